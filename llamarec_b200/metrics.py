"""Ranking metrics with the reference's interface (trainer/utils.py:6-90), computed on the GPU by the
fused merge+metrics kernel instead of a full argsort + one-hot gather.

    absolute_recall_mrr_ndcg_for_ks(scores, labels, ks, num_classes=None, preprocessed=False)
    absolute_metrics_batch_wrapper(scores, labels, ks, num_classes=None, preprocessed=False, batch_size=10000)

Both return {"Recall@k": float, "MRR@k": float, "NDCG@k": float} for every k in ks -- the batch mean, as
the reference does.  One relevant item per user (the only way the reference calls them).
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch

from .model import merge_lists


def metrics_from_sums(sums: torch.Tensor, ks: Sequence[int], denom: float) -> Dict[str, float]:
    """sums [len(ks), 3] (Recall, MRR, NDCG) -> dict keyed like the reference, divided by denom."""
    host = (sums.double() / float(denom)).cpu().tolist()
    out: Dict[str, float] = {}
    for k, (r, m, n) in zip(ks, host):
        out["Recall@%d" % k] = r
        out["MRR@%d" % k] = m
        out["NDCG@%d" % k] = n
    return out


def absolute_recall_mrr_ndcg_for_ks(scores: torch.Tensor, labels: torch.Tensor, ks: Sequence[int],
                                    num_classes=None, preprocessed: bool = False) -> Dict[str, float]:
    """trainer/utils.py:43-90.  `scores` is [B, N+1] (or, with preprocessed=True, ranked ids [B, R])."""
    if not scores.is_cuda:
        raise RuntimeError("llamarec_b200 metrics run on the GPU: pass CUDA tensors (no CPU fallback)")
    ks = list(ks)
    kmax = max(ks)
    B = scores.shape[0]
    if preprocessed:
        ranked = scores.to(torch.int32).contiguous()
        R = ranked.shape[1]
        # ranked ids carry descending pseudo-scores so the merge keeps their order
        pseudo = (-torch.arange(R, device=scores.device, dtype=torch.float32)).expand(B, R).contiguous()
        res = merge_lists(pseudo.unsqueeze(1), ranked.unsqueeze(1), None, k_out=min(kmax, R), labels=labels, ks=ks)
    else:
        N = scores.shape[1]
        ids = torch.arange(N, device=scores.device, dtype=torch.int32).expand(B, N).contiguous()
        res = merge_lists(scores.float().contiguous().unsqueeze(1), ids.unsqueeze(1), None, k_out=min(kmax, N),
                          labels=labels, ks=ks)
    out = metrics_from_sums(res["metric_sums"], ks, B)
    return {key: out[key] for k in sorted(ks, reverse=True) for key in ("Recall@%d" % k, "MRR@%d" % k, "NDCG@%d" % k)}


def absolute_metrics_batch_wrapper(scores, labels, ks, num_classes=None, preprocessed=False,
                                   batch_size: int = 10000) -> Dict[str, float]:
    """trainer/utils.py:6-40 -- size-weighted mean over chunks of `batch_size` rows."""
    total = labels.size(0)
    acc: Dict[str, float] = {}
    for lo in range(0, total, batch_size):
        hi = min(lo + batch_size, scores.size(0))
        part = absolute_recall_mrr_ndcg_for_ks(scores[lo:hi], labels[lo:hi], ks, num_classes=num_classes,
                                               preprocessed=preprocessed)
        for key, v in part.items():
            acc[key] = acc.get(key, 0.0) + v * (hi - lo)
    return {k: v / total for k, v in acc.items()}
