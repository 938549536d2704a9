"""CPU oracle for the stage-2 verbalizer tail -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates ManualVerbalizer.project / process_logits / normalize / aggregate (trainer/verb.py:524-614,
all three multi_token_handlers, :280-305) and the last-position lm_head of model/llm.py:113-114,131.
Parity status: PINNED against the reference class executed by oracle/make_golden.py
(tests/golden/verbalizer_case.npz)."""
from __future__ import annotations

import torch


def process_logits(logits: torch.Tensor, label_words_ids: torch.Tensor, words_ids_mask: torch.Tensor,
                   label_words_mask: torch.Tensor, post_log_softmax: bool,
                   multi_token_handler: str = "first") -> torch.Tensor:
    """logits [B, V] fp32 -> [B, C].  label_words_ids/words_ids_mask [C, W, T], label_words_mask [C, W]."""
    picked = logits[:, label_words_ids]                               # [B, C, W, T]
    if multi_token_handler == "first":                                # verb.py:280-305
        picked = picked[..., 0]
    elif multi_token_handler == "max":
        picked = (picked - 1000 * (1 - words_ids_mask.unsqueeze(0))).max(dim=-1).values
    elif multi_token_handler == "mean":
        picked = (picked * words_ids_mask.unsqueeze(0)).sum(dim=-1) / (words_ids_mask.unsqueeze(0).sum(dim=-1) + 1e-15)
    else:
        raise ValueError(multi_token_handler)
    picked = picked - 10000 * (1 - label_words_mask)                  # verb.py:543
    if post_log_softmax:
        B = picked.shape[0]
        p = torch.softmax(picked.reshape(B, -1), dim=-1).reshape(picked.shape)   # over ALL label words (:599-600)
        picked = torch.log(p + 1e-15)                                 # :582
    return (picked * label_words_mask).sum(-1) / label_words_mask.sum(-1)       # :611-614


def lm_head_last(hidden_last_bf16: torch.Tensor, lm_head_bf16: torch.Tensor) -> torch.Tensor:
    """model/llm.py:113-114,131 restricted to the last position: a bf16 linear, widened to fp32."""
    return torch.nn.functional.linear(hidden_last_bf16, lm_head_bf16).float()
