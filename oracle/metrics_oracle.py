"""CPU oracle for ranking metrics and candidate emission -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates trainer/utils.py:43-90 (absolute_recall_mrr_ndcg_for_ks) and trainer/lru.py:44-175
(LRUTrainer.generate_candidates) in plain torch on the CPU.  Parity status: PINNED against the
reference's own outputs (tests/golden/lru_case_*.npz, generate_candidates_ref.pkl) by
oracle/make_golden.py and tests/test_oracle_golden.py.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import torch

from . import lru_oracle as O


def recall_mrr_ndcg(scores: torch.Tensor, labels: torch.Tensor, ks: Sequence[int]) -> Dict[str, float]:
    """trainer/utils.py:43-90 for one relevant item per user.

    rank r = 0-based position of the label in the descending sort of `scores`; then
    Recall@k = [r<k] / min(k, 1), MRR@k = [r<k] / (r+1), NDCG@k = [r<k] / log2(r+2) / idcg with idcg = 1;
    each averaged over the batch in fp32 (the reference's .mean())."""
    order = (-scores).argsort(dim=1)                       # same call as the reference (:56)
    pos = (order == labels.view(-1, 1)).float().argmax(dim=1)
    out: Dict[str, float] = {}
    for k in sorted(ks, reverse=True):
        hit = (pos < k).float()
        out["Recall@%d" % k] = hit.mean().item()
        out["MRR@%d" % k] = (hit / (pos.float() + 1)).mean().item()
        w = 1 / torch.log2(pos.float() + 2)
        out["NDCG@%d" % k] = (hit * w).mean().item()
    return out


def ranked_metrics(ranked: torch.Tensor, labels: torch.Tensor, ks: Sequence[int]) -> Dict[str, float]:
    """preprocessed=True branch (:58): `ranked` holds item ids in rank order; absent label = miss."""
    eq = ranked == labels.view(-1, 1)
    found = eq.any(dim=1)
    pos = torch.where(found, eq.float().argmax(dim=1), torch.full_like(labels, 1 << 30))
    out: Dict[str, float] = {}
    for k in sorted(ks, reverse=True):
        hit = (pos < k).float()
        safe = torch.where(found, pos, torch.zeros_like(pos)).float()
        out["Recall@%d" % k] = hit.mean().item()
        out["MRR@%d" % k] = (hit / (safe + 1)).mean().item()
        out["NDCG@%d" % k] = (hit / torch.log2(safe + 2)).mean().item()
    return out


def generate_candidates(last_scores_fn: Callable[[torch.Tensor], torch.Tensor], val_loader, test_loader, args,
                        ks: Sequence[int]) -> Dict:
    """trainer/lru.py:44-175, batched: same outputs, same aggregation (sum over users / args.num_users),
    same pickle schema.  `last_scores_fn(ids) -> [B, N+1]` is model(ids)[:, -1, :]."""
    k_cand = args.llm_negative_sample_size + 1
    kmax = max(ks)

    def sweep(loader, want_probs):
        sums = {f"{n}@{k}": 0.0 for k in sorted(ks, reverse=True) for n in ("Recall", "MRR", "NDCG")}
        users: List[int] = []
        cands: List[List[int]] = []
        non_users: List[int] = []
        probs: List[List[int]] = []
        labs: List[int] = []
        seen = 0
        for seqs, labels in loader:
            labels = labels.view(-1)
            s = O.mask_history(last_scores_fn(seqs), seqs)                  # :69-74
            for j in range(seqs.shape[0]):                                  # metrics are summed per user (:75-81)
                m = recall_mrr_ndcg(s[j:j + 1], labels[j:j + 1], ks)
                for key in sums:
                    sums[key] += m[key]
                top = torch.topk(s[j:j + 1], k_cand)[1][0].tolist()         # :82-84
                uid = seen + j + 1
                if want_probs:
                    probs.extend((-s[j:j + 1]).argsort(dim=1)[:, :kmax].tolist())   # :113-115
                    labs.append(int(labels[j]))
                if int(labels[j]) in top:
                    users.append(uid)
                    cands.append(top)
                else:
                    non_users.append(uid)
            seen += seqs.shape[0]
        for key in sums:
            sums[key] /= args.num_users                                     # :90-93
        return sums, users, cands, non_users, probs, labs

    val_metrics, val_users, val_cands, _, _, _ = sweep(val_loader, False)
    test_metrics, test_users, test_cands, non_test, probs, labs = sweep(test_loader, True)

    def subset(users):
        idx = torch.tensor(users, dtype=torch.int64) - 1
        return ranked_metrics(torch.tensor(probs)[idx], torch.tensor(labs)[idx], ks)

    return {
        "val_metrics": val_metrics, "val_users": val_users, "val_candidates": val_cands,
        "test_probs": probs, "test_labels": labs, "test_metrics": test_metrics, "test_users": test_users,
        "test_candidates": test_cands, "non_test_users": non_test,
        "test_retrieval": {
            "original_size": len(probs), "retrieval_size": len(test_cands), "original_metrics": test_metrics,
            "retrieval_metrics": subset(test_users) if test_users else {},
            "non_retrieval_metrics": subset(non_test) if non_test else {},
        },
    }
