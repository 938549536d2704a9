"""CPU: host-side logic -- state-dict compatibility, weight packing, the scan algorithm the CUDA kernel
implements, synthetic-data generators, shard arithmetic, and the multi-rank plumbing under gloo."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT
from llamarec_b200 import LRURec, synth
from llamarec_b200.packing import BLOCK_FLOATS, OFF_BLOCKS, pack_encoder_weights
from llamarec_b200.sharded import ShardedRetriever, shard_range
from oracle import lru_oracle as O


def _args(n=400):
    return SimpleNamespace(num_items=n, bert_hidden_units=64, bert_num_blocks=2, bert_dropout=0.2, bert_attn_dropout=0.2)


def test_state_dict_matches_reference_keys(golden_sd):
    m = LRURec(_args())
    own = m.state_dict()
    assert set(own.keys()) == set(golden_sd.keys())
    for k in own:
        assert own[k].shape == golden_sd[k].shape and own[k].dtype == golden_sd[k].dtype, k
    m.load_state_dict(golden_sd)   # a reference checkpoint loads unchanged


def test_no_cpu_fallback(golden_sd):
    m = LRURec(_args())
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.encode(torch.zeros(2, 20, dtype=torch.int64))


def test_packing_layout(golden_sd):
    blob = pack_encoder_weights(golden_sd)
    assert blob.numel() == OFF_BLOCKS + 2 * BLOCK_FLOATS
    p = "model.lru_blocks.1."
    base = OFF_BLOCKS + BLOCK_FLOATS
    lam, gamma = O.lru_constants(golden_sd[p + "lru_layer.params_log"])
    assert torch.allclose(blob[base:base + 128], lam.real.reshape(-1))
    assert torch.allclose(blob[base + 256:base + 384], gamma.reshape(-1))
    win_t = blob[base + 384: base + 384 + 64 * 256].reshape(64, 256)
    w = golden_sd[p + "lru_layer.in_proj.weight"]
    assert torch.equal(win_t[5, 2 * 17], w.real[17, 5]) and torch.equal(win_t[5, 2 * 17 + 1], w.imag[17, 5])
    off = base + 384 + 64 * 256 + 256
    wout_t = blob[off: off + 256 * 64].reshape(256, 64)
    wo = golden_sd[p + "lru_layer.out_proj.weight"]
    assert torch.equal(wout_t[2 * 9, 3], wo.real[3, 9]) and torch.equal(wout_t[2 * 9 + 1, 3], -wo.imag[3, 9])
    # the packed real-arithmetic form reproduces the complex layer on a random vector
    x = torch.randn(64)
    bu_ref = torch.nn.functional.linear(x.to(torch.cfloat), w, golden_sd[p + "lru_layer.in_proj.bias"]) * gamma
    b_in = blob[base + 384 + 64 * 256: base + 384 + 64 * 256 + 256]
    bu = (x @ win_t + b_in).reshape(128, 2) * gamma.reshape(128, 1)
    assert torch.allclose(torch.view_as_complex(bu.contiguous()), bu_ref.reshape(-1), atol=1e-6)


def fenwick_scan(bu, lam, mask, L):
    """numpy transliteration of lru_scan_kernel (llamarec_b200/csrc/encode.cu)."""
    B, Lp, H = bu.shape
    levels = int(np.log2(Lp))
    off = Lp - L
    out = np.zeros_like(bu)
    for b in range(B):
        q = np.zeros((levels, H), dtype=np.complex64)
        for t in range(L):
            pos = t + off
            acc = bu[b, pos].copy()
            snap, z, below = None, -1, True
            for l in range(levels):
                if (pos >> l) & 1:
                    acc = acc + q[l]
                elif below:
                    snap, z, below = acc.copy(), l, False
            out[b, pos] = acc
            m = 1.0 if mask[b, pos] else 0.0
            for l in range(levels):
                x = snap * m if l == z else q[l]
                q[l] = (x * lam).astype(np.complex64)
    return out


@pytest.mark.parametrize("L", [1, 5, 20, 37, 50, 64, 200])
def test_kernel_scan_algorithm_equals_tree_scan_for_any_mask(L):
    rng = np.random.default_rng(L)
    Lp = 1 << int(np.ceil(np.log2(L))) if L > 1 else 1
    B, H = 3, 8
    bu = (rng.standard_normal((B, Lp, H)) + 1j * rng.standard_normal((B, Lp, H))).astype(np.complex64)
    lam = (0.9 * np.exp(1j * rng.uniform(0, 6.28, H))).astype(np.complex64)
    mask = rng.random((B, Lp)) < 0.7
    mask[:, : Lp - L] = False
    if Lp == 1:
        return
    ref = O.tree_scan(torch.from_numpy(bu), torch.from_numpy(lam).reshape(1, H), torch.from_numpy(mask)).numpy()
    got = fenwick_scan(bu, lam, mask, L)
    assert np.abs(ref[:, Lp - L:] - got[:, Lp - L:]).max() < 1e-5


def test_synth_is_deterministic_and_left_padded():
    cfg = synth.CONFIGS["c2_beauty"]
    a, la = synth.make_sequences(cfg, num_users=50, seed=42)
    b, lb = synth.make_sequences(cfg, num_users=50, seed=42)
    assert torch.equal(a, b) and torch.equal(la, lb)
    nz = a > 0
    assert torch.all(nz[:, 1:] >= nz[:, :-1])            # zeros only on the left
    assert a.max() <= cfg.num_items and la.min() >= 1
    ids, lab = synth.make_sequences_fast(64, 10_000_000, 50)
    assert ids.shape == (64, 50) and int(ids.max()) <= 10_000_000
    sd = synth.make_state_dict(100)
    assert sd["model.lru_blocks.0.lru_layer.in_proj.weight"].dtype == torch.complex64


def test_shard_range_covers_everything():
    for n in (1, 7, 1683, 10_000_001):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


class OracleBackend:
    """Test-only backend: the CPU oracle restricted to a row range (stands in for the CUDA kernels)."""

    def __init__(self, sd, rank, world):
        self.sd = sd
        n = sd["embedding.token.weight"].shape[0]
        self.lo, self.hi = shard_range(n, rank, world)

    def encode(self, x):
        return O.encode(x, self.sd)

    def local_topk_packed(self, x, u, k, exclude_history):
        s = u @ self.sd["embedding.token.weight"][self.lo:self.hi].t() + self.sd["model.bias"][self.lo:self.hi]
        if exclude_history:
            for b in range(x.shape[0]):
                for i in x[b].tolist() + [0]:
                    if self.lo <= i < self.hi:
                        s[b, i - self.lo] = -1e9
        ts, ti = O.topk_sorted(s, k)
        return torch.stack((ts.contiguous().view(torch.int32), (ti + self.lo).to(torch.int32)))

    # data-parallel users: opaque exchange records (the retriever only moves them)
    def encode_states(self, x, exclude_history):
        u = O.encode(x, self.sd)
        return {"state": u, "excl": x.to(torch.int32) if exclude_history else None, "bloom": None,
                "excl_stride": x.shape[1], "u": u}

    def local_topk_rows(self, state, excl, bloom, excl_stride, k):
        p = self.local_topk_packed(excl if excl is not None else torch.zeros(state.shape[0], 0, dtype=torch.int32),
                                   state, k, excl is not None)
        return p.permute(1, 0, 2).contiguous()                           # [B, 2, k]

    def merge_rows(self, recv, k, labels, ks):
        return self.merge_packed(recv.permute(0, 2, 1, 3), k, labels, ks)   # [R, 2, b, k]

    def merge_packed(self, gathered, k, labels, ks):
        s_all = gathered[:, 0].contiguous().view(torch.float32)
        i_all = gathered[:, 1].contiguous()
        R, B, K = s_all.shape
        s = s_all.permute(1, 0, 2).reshape(B, R * K)
        i = i_all.permute(1, 0, 2).reshape(B, R * K).to(torch.int64)
        key = torch.argsort(i, dim=1, stable=True)                      # id asc ...
        s, i = s.gather(1, key), i.gather(1, key)
        order = torch.argsort(-s.double(), dim=1, stable=True)[:, :k]   # ... then score desc (stable)
        return {"scores": s.gather(1, order), "ids": i.gather(1, order).to(torch.int32)}


def _worker(rank, world, port, sd_np, ids_np, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sd = {k: torch.from_numpy(v) for k, v in sd_np.items()}
    x = torch.from_numpy(ids_np)
    r = ShardedRetriever(OracleBackend(sd, rank, world))
    out = r.retrieve(x, k=20, exclude_history=True)
    # data-parallel users: each rank brings its own slice of the batch
    b = x.shape[0] // world
    dp = r.retrieve_dp(x[rank * b:(rank + 1) * b], k=20, exclude_history=True)
    q.put(("dp", rank, dp["scores"].numpy(), dp["ids"].numpy()))
    if rank == 0:
        q.put(("full", out["scores"].numpy(), out["ids"].numpy(), out["u"].numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_retrieval_two_ranks_gloo(golden_sd):
    ids = np.load(os.path.join(ROOT, "tests", "golden", "lru_case_left_l20.npz"))["ids"][:11]   # odd batch
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    sd_np = {k: v.numpy() for k, v in golden_sd.items()}
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sd_np, ids, q)) for r in range(2)]
    for p in procs:
        p.start()
    msgs = [q.get(timeout=120) for _ in range(3)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = torch.from_numpy(ids)
    ref_s, ref_i = O.retrieve(x, golden_sd, 20)
    (_, s, i, u), = [m for m in msgs if m[0] == "full"]
    b = x.shape[0] // 2
    for _, rank, ds, di in [m for m in msgs if m[0] == "dp"]:
        assert np.array_equal(di, ref_i[rank * b:(rank + 1) * b].numpy().astype(np.int32))
        np.testing.assert_allclose(ds, ref_s[rank * b:(rank + 1) * b].numpy(), atol=1e-6)
    np.testing.assert_allclose(u, O.encode(x, golden_sd).numpy(), atol=1e-6)
    assert np.array_equal(i, ref_i.numpy().astype(np.int32))
    np.testing.assert_allclose(s, ref_s.numpy(), atol=1e-6)


# ------------------------------------------------------------------------------------------------
# Invariants behind the scoring kernel's threshold sharing, checked on the CPU:
# whatever the kernel prunes with these bounds can never belong to a user's exact top-K.
# ------------------------------------------------------------------------------------------------
def _kth_best(values, k):
    v = np.sort(np.asarray(values))[::-1]
    return v[k - 1] if len(v) >= k else -np.inf


@pytest.mark.parametrize("seed", range(8))
def test_union_bound_never_exceeds_the_kth_best(seed):
    """Every full-stream thread publishes the c-th best admissible score it has seen so far; with c * S >= K the
    minimum over a user's S streams is a lower bound on the user's K-th best admissible score (ties included,
    so filtering with `>=` keeps every top-K element)."""
    rng = np.random.default_rng(seed)
    for _ in range(200):
        K = int(rng.integers(1, 51))
        S = int(rng.choice([2, 4, 8, 12]))
        c = -(-K // S)
        n = int(rng.integers(S * c, 400))
        scores = rng.integers(-5, 6, size=n).astype(np.float64)          # few distinct values: ties everywhere
        excluded = rng.random(n) < 0.1
        stream = rng.integers(0, S, size=n)
        seen = rng.random(n) < rng.uniform(0.2, 1.0)                       # each stream is somewhere in its sweep
        published = []
        for s in range(S):
            mine = scores[(stream == s) & seen & ~excluded]
            published.append(_kth_best(mine, c))
        bound = min(published)
        if bound == -np.inf:
            continue                                                       # undefined bound: nothing is pruned
        kth = _kth_best(scores[~excluded], K)
        assert kth >= bound
        # hence every element of the exact top-K passes the `>= bound` filter
        order = np.argsort(-scores[~excluded], kind="stable")[:K]
        assert np.all(scores[~excluded][order] >= bound)


@pytest.mark.parametrize("seed", range(8))
def test_scout_bound_accounts_for_excluded_ids(seed):
    """The scout pass only sees the maxima of 16-item groups and does not look at ids; publishing the (c+E)-th
    largest group maximum, E = excluded ids inside the scouted range, still leaves >= c admissible items at or
    above the published value (the c+E maxima belong to c+E distinct items, at most E of them excluded)."""
    rng = np.random.default_rng(100 + seed)
    for _ in range(200):
        c = int(rng.integers(1, 7))
        groups = int(rng.integers(8, 64))
        scores = rng.integers(-3, 4, size=(groups, 16)).astype(np.float64)
        excluded = rng.random((groups, 16)) < 0.05
        E = int(excluded.sum())
        gmax = np.sort(scores.max(axis=1))[::-1]
        if c + E > len(gmax):
            continue
        published = gmax[c + E - 1]
        assert int(((scores >= published) & ~excluded).sum()) >= c


def test_device_eval_set_matches_reference_datasets():
    """DeviceEvalSet against LRUValidDataset / LRUTestDataset of the reference itself (dataloader/lru.py:129-180;
    fixture evalset_case.npz from `python oracle/make_golden.py evalset`): same users in the same order, same
    left-padded histories and labels, same batch boundaries as a shuffle=False DataLoader."""
    import json
    from llamarec_b200.evalset import DeviceEvalSet
    d = np.load(os.path.join(ROOT, "tests", "golden", "evalset_case.npz"))
    dicts = {k: {int(u): v for u, v in m.items()} for k, m in json.loads(str(d["dicts_json"])).items()}
    L = int(d["max_len"])
    val = DeviceEvalSet(dicts["train"], dicts["val"], L, batch_size=16, device="cpu")
    test = DeviceEvalSet(dicts["train"], dicts["test"], L, batch_size=16, u2val=dicts["val"], device="cpu")
    for name, ds in (("val", val), ("test", test)):
        assert ds.users == d[f"{name}_users"].tolist()
        assert ds.seqs.dtype == torch.int32 and ds.labels.dtype == torch.int64
        assert np.array_equal(ds.seqs.numpy().astype(np.int64), d[f"{name}_seqs"])
        assert np.array_equal(ds.labels.numpy(), d[f"{name}_labels"].reshape(-1))
        batches = list(ds)
        assert len(batches) == len(ds) == -(-len(ds.users) // 16)
        assert torch.equal(torch.cat([b[0] for b in batches]), ds.seqs)
        assert all(b[1].shape == (b[0].shape[0], 1) for b in batches)
        assert ds.id_bytes()["int32_device"] * 2 == ds.id_bytes()["int64_reference"]
    sub = DeviceEvalSet(dicts["train"], dicts["test"], L, 8, u2val=dicts["val"], device="cpu", subset_users=test.users[:5])
    assert torch.equal(sub.seqs, test.seqs[:5])


def test_gradient_blob_unpack_is_the_inverse_of_weight_packing():
    """lrb_train_step writes gradients in the layout of the packed weight blob; unpack_encoder_grads must map every slot
    back to the reference's parameter it was packed from (complex parameters as re + i*im)."""
    from llamarec_b200 import synth
    from llamarec_b200.packing import pack_encoder_weights, unpack_encoder_grads
    sd = synth.make_state_dict(50, seed=1)
    g = unpack_encoder_grads(pack_encoder_weights(sd), 2)
    assert len(g) == 2 + 2 * 13
    for k, v in g.items():
        if "params_log" in k:                    # those slots carry lambda / gamma, not params_log
            assert tuple(v.shape) == (3, 128)
            continue
        ref = sd[k]
        assert v.shape == ref.shape and v.dtype == ref.dtype, k
        if k.endswith("out_proj.bias"):
            assert torch.equal(v.real, ref.real) and not v.imag.any(), k      # Im b_out never reaches the output
        else:
            assert torch.equal(v, ref), k


def test_radix_select_algorithm_equals_sorted_topk_with_ties():
    """numpy transliteration of select_topk_kernel (csrc/score.cu: exact-fp32 path of small catalogues): 4 passes of
    8 bits over the order-preserving key find the K-th largest key, everything above it plus the lowest-id entries
    equal to it are collected, -inf / NaN entries are dropped, the survivors are rank-sorted (score desc, id asc).
    Must equal a plain sort for random rows with massive ties, masked entries, NaNs and K > valid entries."""
    def key(v):
        i = v.view(np.int32).astype(np.int64)
        k = np.where(i >= 0, i, i ^ 0x7fffffff)                      # float_to_key
        k = (k & 0xffffffff) ^ 0x80000000
        return np.where(np.isnan(v), 0, k).astype(np.uint64)

    def select(row, K):
        rows = len(row)
        keys = key(row)
        K_eff = min(K, rows)
        prefix, need = 0, K_eff
        for p in range(4):
            shift = 24 - 8 * p
            hi_mask = 0 if p == 0 else (0xffffffff << (shift + 8)) & 0xffffffff
            act = (keys & hi_mask) == prefix
            hist = np.bincount(((keys[act] >> shift) & 255).astype(np.int64), minlength=256)
            acc = 0
            for b in range(255, -1, -1):                             # highest bin first
                if acc < need <= acc + hist[b]:
                    prefix |= b << shift
                    need -= acc
                    break
                acc += hist[b]
        gt = np.nonzero(keys > prefix)[0]
        eq = np.nonzero(keys == prefix)[0][:need]                    # lowest columns first
        assert len(gt) == K_eff - need
        cand = np.concatenate([gt, eq])
        cand = [c for c in cand if row[c] > -np.inf]                 # NaN > -inf is False as well
        cand.sort(key=lambda c: (-row[c], c))
        return cand

    rng = np.random.default_rng(5)
    for trial in range(60):
        rows = int(rng.integers(1, 700))
        K = int(rng.choice([1, 5, 20, 50]))
        row = rng.choice(rng.standard_normal(max(2, rows // 9)).astype(np.float32), size=rows).astype(np.float32)
        row[rng.random(rows) < 0.2] = -np.inf                       # history mask
        if trial % 3 == 0:
            row[rng.random(rows) < 0.05] = np.nan
        if trial % 7 == 0:
            row[:] = -np.inf if trial % 2 else np.float32(0.0)       # everything masked / everything tied (+0.0)
        valid = [c for c in range(rows) if row[c] > -np.inf]
        want = sorted(valid, key=lambda c: (-row[c], c))[:K]
        assert select(row, K) == want, (trial, rows, K)


def test_retriever_export_round_trip(tmp_path):
    """stage2.export_retriever / load_retriever (the demo's retriever artefact, setup_demo.py:46-47,
    demo/inference.py:20-23): every parameter of the reference's state_dict survives the file, keys and dtypes
    included; a foreign file is rejected."""
    from types import SimpleNamespace
    from llamarec_b200 import LRURec, stage2, synth
    sd = synth.make_state_dict(120, seed=4, bias_std=0.02)
    m = LRURec(SimpleNamespace(num_items=120, bert_hidden_units=64, bert_num_blocks=2, bert_dropout=0.2,
                               bert_attn_dropout=0.2))
    m.load_state_dict(sd)
    path = str(tmp_path / "retriever.pth")
    stage2.export_retriever(m, path)
    back = stage2.load_retriever(path, device="cpu")
    assert not back.training and back.num_items == 120
    got = back.state_dict()
    assert set(got) == set(sd)
    for k, v in sd.items():
        assert got[k].dtype == v.dtype and torch.equal(got[k], v), k
    torch.save({"format": "something else"}, path)
    with pytest.raises(ValueError):
        stage2.load_retriever(path, device="cpu")
