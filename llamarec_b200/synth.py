"""Synthetic inputs for the five BASELINE.json configurations.

Everything here is deterministic given the seed (42 = the reference default, config.py:161) and
uses only numpy/torch on the CPU, so tests, bench.py and the oracle all consume identical tensors.
No file under oracle/ is imported here.

Weight initialisation mirrors the *distributions* LRURec uses (truncated normal sigma=0.02 clipped
to +-0.04, model/lru.py:16-36; ring initialisation r in [0.8, 0.99], model/lru.py:106-119) but is
not the reference's RNG stream; fixtures that must equal the reference's own weights are produced
by oracle/make_golden.py, which constructs the reference model itself.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np
import torch

D_MODEL = 64


@dataclass(frozen=True)
class Config:
    name: str
    num_users: int
    num_items: int
    max_len: int
    batch: int
    n_blocks: int = 2
    k: int = 20
    metric_ks: Tuple[int, ...] = (1, 5, 10, 20, 50)


CONFIGS: Dict[str, Config] = {
    # configs[0]: ML-100k shaped, runs on CPU in the reference (config.py:59, :104 batch 16)
    "c1_ml100k": Config("c1_ml100k", 943, 1682, 200, 16),
    # configs[1]: Beauty shaped (paper table 2: 22,332 users / 12,086 items), max_len 50, batch 64
    "c2_beauty": Config("c2_beauty", 22332, 12086, 50, 64),
    # configs[2]: Games shaped per BASELINE.json wording, batch 2048
    "c3_games": Config("c3_games", 25000, 11000, 50, 2048),
    # configs[3]: scaled catalogue, 10M items, batch 4096
    "c4_10m": Config("c4_10m", 4096, 10_000_000, 50, 4096),
}


def trunc_normal_(t: torch.Tensor, gen: torch.Generator, std: float = 0.02, lo: float = -0.04,
                  hi: float = 0.04) -> torch.Tensor:
    """In-place truncated normal via inverse-CDF sampling (same family as model/lru.py:16-36)."""
    a = (1.0 + math.erf((lo / std) / math.sqrt(2.0))) / 2.0
    b = (1.0 + math.erf((hi / std) / math.sqrt(2.0))) / 2.0
    u = torch.rand(t.shape, generator=gen, dtype=torch.float32) * (2 * (b - a)) + (2 * a - 1)
    t.copy_(torch.erfinv(u) * (std * math.sqrt(2.0)))
    return t


def make_state_dict(num_items: int, n_blocks: int = 2, seed: int = 42, table_scale: float = 1.0,
                    bias_std: float = 0.0) -> Dict[str, torch.Tensor]:
    """A random-init LRURec state_dict with the reference's key names and dtypes."""
    g = torch.Generator().manual_seed(seed)
    d, h = D_MODEL, 2 * D_MODEL
    sd: Dict[str, torch.Tensor] = {}

    def tn(*shape):
        return trunc_normal_(torch.empty(*shape), g)

    sd["embedding.token.weight"] = tn(num_items + 1, d) * table_scale
    sd["embedding.layer_norm.weight"] = torch.ones(d)
    sd["embedding.layer_norm.bias"] = torch.zeros(d)
    sd["model.bias"] = (torch.randn(num_items + 1, generator=g) * bias_std if bias_std > 0
                        else torch.zeros(num_items + 1))
    for i in range(n_blocks):
        p = f"model.lru_blocks.{i}."
        u1 = torch.rand(h, generator=g)
        u2 = torch.rand(h, generator=g)
        r_min, r_max = 0.8, 0.99
        nu_log = torch.log(-0.5 * torch.log(u1 * (r_max ** 2 - r_min ** 2) + r_min ** 2))
        theta_log = torch.log(u2 * (2 * math.pi))
        lam_abs = torch.exp(-torch.exp(nu_log))
        gamma_log = torch.log(torch.sqrt(1 - lam_abs ** 2))
        sd[p + "lru_layer.params_log"] = torch.vstack((nu_log, theta_log, gamma_log))
        sd[p + "lru_layer.in_proj.weight"] = torch.complex(tn(h, d), tn(h, d))
        sd[p + "lru_layer.in_proj.bias"] = torch.complex(tn(h), tn(h))
        sd[p + "lru_layer.out_proj.weight"] = torch.complex(tn(d, h), tn(d, h))
        sd[p + "lru_layer.out_proj.bias"] = torch.complex(tn(d), tn(d))
        sd[p + "lru_layer.layer_norm.weight"] = torch.ones(d)
        sd[p + "lru_layer.layer_norm.bias"] = torch.zeros(d)
        sd[p + "feed_forward.w_1.weight"] = tn(4 * d, d)
        sd[p + "feed_forward.w_1.bias"] = tn(4 * d)
        sd[p + "feed_forward.w_2.weight"] = tn(d, 4 * d)
        sd[p + "feed_forward.w_2.bias"] = tn(d)
        sd[p + "feed_forward.layer_norm.weight"] = torch.ones(d)
        sd[p + "feed_forward.layer_norm.bias"] = torch.zeros(d)
    return sd


def _zipf_items(rng: np.random.Generator, num_items: int, n: int) -> np.ndarray:
    """n distinct item ids in 1..num_items drawn with probability ~ 1/rank (Zipf s=1)."""
    out: list = []
    seen = set()
    log_n = math.log(num_items + 1.0)
    while len(out) < n:
        # inverse-CDF of the continuous 1/x law on [1, N+1)
        draw = np.exp(rng.random(2 * (n - len(out)) + 8) * log_n).astype(np.int64)
        for v in draw:
            v = int(min(max(v, 1), num_items))
            if v not in seen:
                seen.add(v)
                out.append(v)
                if len(out) == n:
                    break
    return np.asarray(out, dtype=np.int64)


def make_sequences(cfg: Config, num_users: int | None = None, seed: int = 42,
                   split: str = "test") -> Tuple[torch.Tensor, torch.Tensor]:
    """Left-padded id matrix [U, max_len] int64 and labels [U] int64 (dataloader/lru.py:129-180).

    Per user a list of distinct items; last = test label, second-last = validation label
    (llamarec_datasets/base.py:105-116).  `split='val'` feeds seq[:-2] -> label seq[-2];
    `split='test'` feeds seq[:-1] -> label seq[-1].
    """
    users = cfg.num_users if num_users is None else num_users
    rng = np.random.default_rng(seed)
    L = cfg.max_len
    ids = np.zeros((users, L), dtype=np.int64)
    labels = np.zeros((users,), dtype=np.int64)
    for u in range(users):
        if L >= 200:
            n = int(min(737, 20 + math.floor(rng.exponential(86.0))))
        else:
            n = int(min(L + 2, 3 + rng.geometric(1.0 / 6.0)))
        n = max(n, 3)
        n = min(n, cfg.num_items)
        seq = _zipf_items(rng, cfg.num_items, n)
        if split == "val":
            hist, lab = seq[:-2], seq[-2]
        else:
            hist, lab = seq[:-1], seq[-1]
        hist = hist[-L:]
        ids[u, L - len(hist):] = hist
        labels[u] = lab
    return torch.from_numpy(ids), torch.from_numpy(labels)


def make_sequences_fast(num_users: int, num_items: int, max_len: int, seed: int = 42,
                        mean_len: float = 9.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Vectorised generator for very large catalogues (config 4): Zipf ids, geometric lengths.

    Items inside one history may repeat with negligible probability at N=10M; duplicates are legal
    inputs for every kernel.
    """
    rng = np.random.default_rng(seed)
    lens = np.minimum(max_len, 2 + rng.geometric(1.0 / max(mean_len - 2.0, 1.0), size=num_users))
    log_n = math.log(num_items + 1.0)
    raw = np.exp(rng.random((num_users, max_len)) * log_n).astype(np.int64)
    raw = np.clip(raw, 1, num_items)
    pos = np.arange(max_len)[None, :]
    keep = pos >= (max_len - lens)[:, None]
    ids = np.where(keep, raw, 0)
    labels = rng.integers(1, num_items + 1, size=num_users, dtype=np.int64)
    return torch.from_numpy(ids), torch.from_numpy(labels)


def make_table_bf16(num_items: int, seed: int = 42, device: str = "cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """Config-4 item table: truncated normal sigma=0.02 (fp32 master) and zero bias."""
    g = torch.Generator(device=device).manual_seed(seed)
    n = num_items + 1
    u = torch.rand((n, D_MODEL), generator=g, dtype=torch.float32, device=device)
    a = (1.0 + math.erf(-2.0 / math.sqrt(2.0))) / 2.0
    b = (1.0 + math.erf(2.0 / math.sqrt(2.0))) / 2.0
    u.mul_(2 * (b - a)).add_(2 * a - 1)
    table = torch.erfinv(u).mul_(0.02 * math.sqrt(2.0))
    bias = torch.zeros(n, dtype=torch.float32, device=device)
    return table, bias


def make_verbalizer_inputs(batch: int = 512, hidden: int = 4096, vocab: int = 32000, classes: int = 20,
                           seed: int = 42) -> Dict[str, torch.Tensor]:
    """Config 5: Llama-2-7B shaped last-position hidden states and lm_head (random init, bf16)."""
    g = torch.Generator().manual_seed(seed)
    h = torch.randn(batch, hidden, generator=g).to(torch.bfloat16)
    w = (torch.randn(vocab, hidden, generator=g) * 0.02).to(torch.bfloat16)
    label_ids = torch.randperm(vocab, generator=g)[:classes].to(torch.int64)
    return {"hidden": h, "lm_head": w, "label_ids": label_ids}
