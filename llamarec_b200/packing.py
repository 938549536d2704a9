"""Packs an LRURec state_dict (reference key names) into the flat fp32 blob
consumed by lrb_encode_fwd.  Layout constants mirror llamarec_b200/csrc/encode.cu.

  [0,64) emb LN weight   [64,128) emb LN bias
  per block (128 + i*BLOCK_FLOATS):
    lam_re[128] lam_im[128] gamma[128]                    model/lru.py:151-152
    W_in^T  [64][256]  column 2c = Re W_in[c,:], 2c+1 = Im W_in[c,:]     (x is real, so the complex
    b_in    [256]      interleaved (re, im)                               in_proj is two real GEMMs)
    W_out^T [256][64]  row 2c = Re W_out[:,c], row 2c+1 = -Im W_out[:,c]  (only Re(out_proj h) is kept,
    b_out   [64]       Re b_out                                            model/lru.py:160)
    LN1 w,b [64]x2 ; W1^T [64][256], b1[256] ; W2^T [256][64], b2[64] ; LN2 w,b [64]x2
"""
from __future__ import annotations

from typing import Dict

import torch

D, H, FF = 64, 128, 256
OFF_BLOCKS = 128
BLOCK_FLOATS = 384 + D * 2 * H + 2 * H + 2 * H * D + D + 2 * D + D * FF + FF + FF * D + D + 2 * D


def n_blocks_of(sd: Dict[str, torch.Tensor]) -> int:
    n = 0
    while f"model.lru_blocks.{n}.lru_layer.params_log" in sd:
        n += 1
    return n


def pack_encoder_weights(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Returns a 1-D fp32 CPU tensor of lrb_encoder_weight_floats(n_blocks) elements."""
    nb = n_blocks_of(sd)
    f32 = lambda t: t.detach().to("cpu", torch.float32).contiguous()
    parts = [f32(sd["embedding.layer_norm.weight"]), f32(sd["embedding.layer_norm.bias"])]
    for i in range(nb):
        p = f"model.lru_blocks.{i}."
        params_log = sd[p + "lru_layer.params_log"].detach().cpu().float()
        nu, theta, gamma = torch.exp(params_log).split((1, 1, 1))
        lam = torch.exp(torch.complex(-nu, theta)).reshape(-1)      # same ops as model/lru.py:151-152
        w_in = sd[p + "lru_layer.in_proj.weight"].detach().cpu().to(torch.complex64)    # [128, 64]
        b_in = sd[p + "lru_layer.in_proj.bias"].detach().cpu().to(torch.complex64)      # [128]
        w_out = sd[p + "lru_layer.out_proj.weight"].detach().cpu().to(torch.complex64)  # [64, 128]
        b_out = sd[p + "lru_layer.out_proj.bias"].detach().cpu().to(torch.complex64)    # [64]
        win_t = torch.stack((w_in.real, w_in.imag), dim=1).reshape(2 * H, D).t().contiguous()   # [64][256]
        bin_i = torch.stack((b_in.real, b_in.imag), dim=1).reshape(2 * H)
        wout_t = torch.stack((w_out.real.t(), -w_out.imag.t()), dim=1).reshape(2 * H, D).contiguous()  # [256][64]
        parts += [
            lam.real.contiguous(), lam.imag.contiguous(), gamma.reshape(-1).contiguous(),
            win_t.reshape(-1), bin_i, wout_t.reshape(-1), b_out.real.contiguous(),
            f32(sd[p + "lru_layer.layer_norm.weight"]), f32(sd[p + "lru_layer.layer_norm.bias"]),
            f32(sd[p + "feed_forward.w_1.weight"]).t().contiguous().reshape(-1), f32(sd[p + "feed_forward.w_1.bias"]),
            f32(sd[p + "feed_forward.w_2.weight"]).t().contiguous().reshape(-1), f32(sd[p + "feed_forward.w_2.bias"]),
            f32(sd[p + "feed_forward.layer_norm.weight"]), f32(sd[p + "feed_forward.layer_norm.bias"]),
        ]
    blob = torch.cat([x.reshape(-1).float() for x in parts]).contiguous()
    assert blob.numel() == OFF_BLOCKS + nb * BLOCK_FLOATS, (blob.numel(), nb)
    return blob
