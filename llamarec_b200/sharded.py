"""Row-sharded multi-GPU retrieval (north_star: item table sharded by row range across the GPUs of one
box, local top-k per rank, one all-gather over NVLink, k-way merge kernel).

The reference has no multi-device retrieval path; its single-device result is the
oracle: merging exact per-shard top-k lists reproduces the unsharded top-k exactly.

One process per GPU (`torch.distributed`, NCCL).  Per batch:
  1. every rank encodes its slice of the users (the encoder is batch-sharded: replicated encode would
     be the Amdahl term at 8 GPUs) and all-gathers the 64-d user states            [B x 64 fp32]
  2. every rank scores ALL users against ITS item rows and keeps a local top-k     (no communication)
  3. ONE all-gather of the local lists, scores and ids packed in one payload       [R x 2 x B x k x 4 bytes]
  4. every rank merges the R lists (fused with the metrics)                        (no communication)

The numerical work is delegated to a backend object so that the plumbing (ranges, padding, gather
layout) can be exercised on CPU with gloo in tests; the product backend drives the CUDA kernels.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of range(n) into `world` pieces of ceil(n/world) (the last may be short/empty)."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


class CudaBackend:
    """Product backend: llamarec_b200.LRURec restricted to this rank's row shard."""

    def __init__(self, model, rank: int, world: int, precision: str = "auto"):
        self.model = model
        self.precision = precision
        self._packed = {}
        lo, hi = shard_range(model.num_items + 1, rank, world)
        model.set_row_shard(lo, hi)
        self.rows = (lo, hi)

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        return self.model.encode(x)

    def local_topk_packed(self, x: torch.Tensor, u: torch.Tensor, k: int, exclude_history: bool) -> torch.Tensor:
        """Local top-k as ONE gather payload: int32 [2, B, k] = scores (bit pattern), ids."""
        B = x.shape[0]
        key = (B, k, str(u.device))
        buf = self._packed.get(key)
        if buf is None:
            buf = torch.empty(2, B, k, dtype=torch.int32, device=u.device)
            self._packed[key] = buf
        if self.rows[1] <= self.rows[0]:   # empty shard (more ranks than rows): nothing to offer
            buf[0].view(torch.float32).fill_(float("-inf"))
            buf[1].fill_(-1)
            return buf
        self.model.retrieve(x, k=k, exclude_history=exclude_history, precision=self.precision, u=u, packed_out=buf)
        return buf

    def merge_packed(self, gathered: torch.Tensor, k, labels, ks):
        """gathered: int32 [R, 2, B, k] (all-gathered payloads) -> final lists (+ metrics)."""
        from .model import merge_lists
        R, _, B, K = gathered.shape
        scores = gathered[:, 0].view(torch.float32)
        ids = gathered[:, 1]
        # missing entries carry id -1 / score -inf and simply never win
        return merge_lists(scores, ids, None, k_out=k, labels=labels, ks=ks, layout="list_major",
                           strides=(2 * B * K, K))


class ShardedRetriever:
    def __init__(self, backend, group: Optional[dist.ProcessGroup] = None):
        self.backend = backend
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def _all_gather(self, t: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return t.unsqueeze(0)
        t = t.contiguous()
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)   # concatenated along dim 0 (gloo and nccl)
        return out.reshape((self.world,) + tuple(t.shape))

    @torch.no_grad()
    def retrieve(self, x: torch.Tensor, k: int = 20, exclude_history: bool = True,
                 labels: Optional[torch.Tensor] = None, ks: Optional[Sequence[int]] = None) -> Dict[str, torch.Tensor]:
        """x: the FULL batch of id sequences [B, L], identical on every rank."""
        B = x.shape[0]
        per = (B + self.world - 1) // self.world
        lo, hi = shard_range(B, self.rank, self.world)
        # 1. batch-sharded encode, padded to a common slice size for the collective
        u_loc = torch.zeros(per, 64, dtype=torch.float32, device=x.device)
        if hi > lo:
            u_loc[: hi - lo] = self.backend.encode(x[lo:hi])
        u = self._all_gather(u_loc).reshape(self.world * per, 64)[:B].contiguous()
        # 2. local scoring over this rank's rows -> one packed payload [2, B, k] (scores, ids)
        payload = self.backend.local_topk_packed(x, u, k, exclude_history)
        # 3. ONE all-gather of the (score, id) lists
        gathered = self._all_gather(payload)
        # 4. merge (+ metrics)
        out = self.backend.merge_packed(gathered, k, labels, ks)
        out["u"] = u
        return out
