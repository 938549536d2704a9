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


def unpack_encoder_grads(blob: torch.Tensor, n_blocks: int) -> Dict[str, torch.Tensor]:
    """Inverse of pack_encoder_weights for the GRADIENT blob lrb_train_step writes (same layout; the lambda_re /
    lambda_im / gamma slots hold d nu_log / d theta_log / d gamma_log).  Returns tensors keyed by the reference's
    parameter names, complex where the parameter is complex (real part = dL/dRe, imaginary part = dL/dIm)."""
    g: Dict[str, torch.Tensor] = {}
    g["embedding.layer_norm.weight"] = blob[0:D]
    g["embedding.layer_norm.bias"] = blob[D:2 * D]
    for i in range(n_blocks):
        p = f"model.lru_blocks.{i}."
        o = OFF_BLOCKS + i * BLOCK_FLOATS
        cur = [o]

        def nxt(n):
            v = blob[cur[0]:cur[0] + n]
            cur[0] += n
            return v
        g[p + "lru_layer.params_log"] = nxt(3 * H).reshape(3, H)
        win_t = nxt(D * 2 * H).reshape(D, 2 * H)                  # [64][256], column 2c = Re W_in[c,:], 2c+1 = Im
        g[p + "lru_layer.in_proj.weight"] = torch.complex(win_t[:, 0::2].t().contiguous(), win_t[:, 1::2].t().contiguous())
        b_in = nxt(2 * H)
        g[p + "lru_layer.in_proj.bias"] = torch.complex(b_in[0::2].contiguous(), b_in[1::2].contiguous())
        wout_t = nxt(2 * H * D).reshape(2 * H, D)                 # row 2c = Re W_out[:,c], row 2c+1 = -Im W_out[:,c]
        g[p + "lru_layer.out_proj.weight"] = torch.complex(wout_t[0::2].t().contiguous(), (-wout_t[1::2]).t().contiguous())
        b_out = nxt(D)
        g[p + "lru_layer.out_proj.bias"] = torch.complex(b_out.contiguous(), torch.zeros_like(b_out))   # Im b_out is unused
        g[p + "lru_layer.layer_norm.weight"] = nxt(D)
        g[p + "lru_layer.layer_norm.bias"] = nxt(D)
        g[p + "feed_forward.w_1.weight"] = nxt(D * FF).reshape(D, FF).t().contiguous()
        g[p + "feed_forward.w_1.bias"] = nxt(FF)
        g[p + "feed_forward.w_2.weight"] = nxt(FF * D).reshape(FF, D).t().contiguous()
        g[p + "feed_forward.w_2.bias"] = nxt(D)
        g[p + "feed_forward.layer_norm.weight"] = nxt(D)
        g[p + "feed_forward.layer_norm.bias"] = nxt(D)
        assert cur[0] == o + BLOCK_FLOATS
    return g
