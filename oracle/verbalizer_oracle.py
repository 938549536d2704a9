"""CPU oracle for the stage-2 verbalizer tail -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates ManualVerbalizer.project / process_logits / normalize / aggregate (trainer/verb.py:524-614,
multi_token_handler='first', :280-305) and the last-position lm_head of model/llm.py:113-114,131.
Parity status: PINNED against the reference class executed by oracle/make_golden.py
(tests/golden/verbalizer_case.npz)."""
from __future__ import annotations

import torch


def process_logits(logits: torch.Tensor, label_words_ids: torch.Tensor, words_ids_mask: torch.Tensor,
                   label_words_mask: torch.Tensor, post_log_softmax: bool) -> torch.Tensor:
    """logits [B, V] fp32 -> [B, C].  label_words_ids/words_ids_mask [C, W, T], label_words_mask [C, W]."""
    picked = logits[:, label_words_ids][..., 0]                       # first sub-token of every label word
    picked = picked - 10000 * (1 - label_words_mask)                  # verb.py:543
    if post_log_softmax:
        B = picked.shape[0]
        p = torch.softmax(picked.reshape(B, -1), dim=-1).reshape(picked.shape)   # over ALL label words (:599-600)
        picked = torch.log(p + 1e-15)                                 # :582
    return (picked * label_words_mask).sum(-1) / label_words_mask.sum(-1)       # :611-614


def lm_head_last(hidden_last_bf16: torch.Tensor, lm_head_bf16: torch.Tensor) -> torch.Tensor:
    """model/llm.py:113-114,131 restricted to the last position: a bf16 linear, widened to fp32."""
    return torch.nn.functional.linear(hidden_last_bf16, lm_head_bf16).float()
