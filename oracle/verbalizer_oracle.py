"""CPU oracle for the stage-2 verbalizer tail -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates ManualVerbalizer.project / process_logits / normalize / calibrate / aggregate (trainer/verb.py:524-643,
all three multi_token_handlers, :280-305) and the last-position lm_head of model/llm.py:113-114,131.
Parity status: PINNED against the reference class executed by oracle/make_golden.py
(tests/golden/verbalizer_case.npz)."""
from __future__ import annotations

import torch


def project(logits: torch.Tensor, label_words_ids: torch.Tensor, words_ids_mask: torch.Tensor,
            label_words_mask: torch.Tensor, multi_token_handler: str = "first") -> torch.Tensor:
    """verb.py:524-544: logits [B, V] -> label-word logits [B, C, W]."""
    picked = logits[:, label_words_ids]                               # [B, C, W, T]
    if multi_token_handler == "first":                                # verb.py:280-305
        picked = picked[..., 0]
    elif multi_token_handler == "max":
        picked = (picked - 1000 * (1 - words_ids_mask.unsqueeze(0))).max(dim=-1).values
    elif multi_token_handler == "mean":
        picked = (picked * words_ids_mask.unsqueeze(0)).sum(dim=-1) / (words_ids_mask.unsqueeze(0).sum(dim=-1) + 1e-15)
    else:
        raise ValueError(multi_token_handler)
    return picked - 10000 * (1 - label_words_mask)                    # verb.py:543


def process_logits(logits: torch.Tensor, label_words_ids: torch.Tensor, words_ids_mask: torch.Tensor,
                   label_words_mask: torch.Tensor, post_log_softmax: bool,
                   multi_token_handler: str = "first", calibrate_logits: torch.Tensor = None) -> torch.Tensor:
    """logits [B, V] fp32 -> [B, C].  label_words_ids/words_ids_mask [C, W, T], label_words_mask [C, W];
    calibrate_logits: optional [V] (ManualVerbalizer._calibrate_logits, verb.py:202-208)."""
    picked = project(logits, label_words_ids, words_ids_mask, label_words_mask, multi_token_handler)
    if post_log_softmax:
        B = picked.shape[0]
        p = torch.softmax(picked.reshape(B, -1), dim=-1).reshape(picked.shape)   # over ALL label words (:599-600)
        if calibrate_logits is not None:                              # calibrate, verb.py:616-643
            cl = project(calibrate_logits.unsqueeze(0), label_words_ids, words_ids_mask, label_words_mask,
                         multi_token_handler)
            cp = torch.softmax(cl.reshape(1, -1), dim=-1).reshape(cl.shape)
            p = p / (cp + 1e-15)
            p = (p.reshape(B, -1) / p.reshape(B, -1).sum(dim=-1, keepdim=True)).reshape(p.shape)
        picked = torch.log(p + 1e-15)                                 # :582
    return (picked * label_words_mask).sum(-1) / label_words_mask.sum(-1)       # :611-614


def lm_head_last(hidden_last_bf16: torch.Tensor, lm_head_bf16: torch.Tensor) -> torch.Tensor:
    """model/llm.py:113-114,131 restricted to the last position: a bf16 linear, widened to fp32."""
    return torch.nn.functional.linear(hidden_last_bf16, lm_head_bf16).float()
