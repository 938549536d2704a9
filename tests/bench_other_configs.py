#!/usr/bin/env python
"""Times BASELINE.json configs 0-2 and 4 (the parity-test configurations, not the headline bench line) on one
GPU, with the CPU oracle port beside each, and writes gpurun_out/other_configs.json (copied to
profiles/other_configs_rNN.json).  It lives under tests/ because it uses the oracle as its checker (the oracle is test
infrastructure); it is not collected by pytest.  Run: python tests/bench_other_configs.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root
import torch
from types import SimpleNamespace
from llamarec_b200 import LRURec, LRURetriever, ManualVerbalizer, synth
from oracle import lru_oracle as O, metrics_oracle as MO, verbalizer_oracle as VO

dev = torch.device("cuda")
out = {"cores": os.cpu_count(), "gpu": torch.cuda.get_device_name(0)}
torch.set_num_threads(os.cpu_count())


def gpu_time(fn, n=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def model_for(cfg, bias_std=0.0):
    sd = synth.make_state_dict(cfg.num_items, seed=42, bias_std=bias_std)
    a = SimpleNamespace(num_items=cfg.num_items, bert_hidden_units=64, bert_num_blocks=2, bert_dropout=0.2,
                        bert_attn_dropout=0.2, metric_ks=list(cfg.metric_ks), llm_negative_sample_size=19)
    m = LRURec(a); m.load_state_dict(sd)
    return m.to(dev).eval(), sd, a

# ---- C1: ML-100k shaped, batch 16 and one batch of 943 -------------------------------------------------------
cfg = synth.CONFIGS["c1_ml100k"]
m, sd, a = model_for(cfg)
ids, labels = synth.make_sequences(cfg, seed=42)
for B in (16, 943):
    x, y = ids[:B].to(dev), labels[:B].to(dev)
    t = gpu_time(lambda: m.retrieve(x, k=50, labels=y, ks=list(cfg.metric_ks)))
    t0 = time.perf_counter(); s = O.mask_history(O.last_scores(ids[:B], sd), ids[:B]); MO.recall_mrr_ndcg(s, labels[:B], cfg.metric_ks); torch.topk(s, 20); tc = time.perf_counter() - t0
    res = m.retrieve(x, k=20, labels=y, ks=[1, 5, 10, 20])
    ref_s, ref_i = O.retrieve(ids[:B], sd, 20)
    same = (res["ids"].cpu().long() == ref_i).all(1).float().mean().item()
    out[f"c1_ml100k_B{B}"] = {"gpu_users_per_s": B / t, "gpu_ms": t * 1e3, "cpu_oracle_users_per_s": B / tc,
                              "top20_lists_identical_to_oracle": same, "what": "encode + exact fp32 score + mask + top-50 + Recall/MRR/NDCG"}

# ---- C2: Beauty shaped, batch 64: forward at all positions + top-20 ------------------------------------------
cfg = synth.CONFIGS["c2_beauty"]
m, sd, a = model_for(cfg)
ids, labels = synth.make_sequences(cfg, num_users=4096, seed=42)
x = ids[:64].to(dev)
t_fwd = gpu_time(lambda: m(x))
t_top = gpu_time(lambda: m.retrieve(x, k=20))
t0 = time.perf_counter(); O.forward_scores(ids[:64], sd); tc = time.perf_counter() - t0
out["c2_beauty_B64"] = {"forward_all_positions_ms": t_fwd * 1e3, "forward_users_per_s": 64 / t_fwd,
                        "retrieve_top20_ms": t_top * 1e3, "retrieve_users_per_s": 64 / t_top,
                        "cpu_oracle_forward_users_per_s": 64 / tc, "logits_bytes": 64 * 50 * (cfg.num_items + 1) * 4}
# train-step forward loss (trainer/lru.py:20-28) without the logits tensor, against the oracle's full-logits CE
lab = torch.zeros_like(ids[:64]); lab[:, :-1] = ids[:64, 1:]; lab[:, -1] = labels[:64].view(-1); lab[ids[:64] == 0] = 0
yl = lab.to(dev)
def fwd_loss(mm, xx, yy):
    with torch.no_grad():                  # value only (lrb_ce_loss_fwd); with autograd enabled ce_loss is the train step
        return mm.ce_loss(xx, yy)


def train_step(mm, xx, yy):                # loss + the gradient of every parameter (lrb_train_step) through autograd
    mm.zero_grad(set_to_none=True)
    mm.ce_loss(xx, yy).backward()


t_ce = gpu_time(lambda: fwd_loss(m, x, yl))
t_ts = gpu_time(lambda: train_step(m, x, yl), n=10)
t0 = time.perf_counter(); ref_loss = O.ce_loss(ids[:64], lab, sd).item(); tcl = time.perf_counter() - t0
out["c2_beauty_B64"].update({"ce_loss_fwd_ms": t_ce * 1e3, "train_step_fwd_bwd_ms": t_ts * 1e3,
                             "ce_loss": fwd_loss(m, x, yl).item(), "cpu_oracle_ce_loss": ref_loss,
                             "cpu_oracle_ce_loss_ms": tcl * 1e3})
xb = ids.to(dev)
t_big = gpu_time(lambda: m.retrieve(xb, k=20), n=10)
out["c2_beauty_B4096"] = {"retrieve_top20_ms": t_big * 1e3, "retrieve_users_per_s": 4096 / t_big}

# ---- C3: Games shaped, batch 2048, fused score + top-k + metrics ---------------------------------------------
cfg = synth.CONFIGS["c3_games"]
m, sd, a = model_for(cfg, bias_std=0.01)
ids, labels = synth.make_sequences(cfg, num_users=2048, seed=42)
x, y = ids.to(dev), labels.to(dev)
tr = LRURetriever(a, m)
t32 = gpu_time(lambda: tr.calculate_metrics((x, y.view(-1, 1))), n=10)
t16 = gpu_time(lambda: m.retrieve(x, k=50, labels=y, ks=list(cfg.metric_ks), precision="bf16"), n=10)
t0 = time.perf_counter(); s = O.mask_history(O.last_scores(ids, sd), ids); MO.recall_mrr_ndcg(s, labels, cfg.metric_ks); tc = time.perf_counter() - t0
lab = torch.zeros_like(ids); lab[:, :-1] = ids[:, 1:]; lab[:, -1] = labels.view(-1); lab[ids == 0] = 0
yl = lab.to(dev)
t_ce3 = gpu_time(lambda: fwd_loss(m, x, yl), n=10)
t_ts3 = gpu_time(lambda: train_step(m, x, yl), n=5, warm=2)
out["c3_games_B2048_train_step_loss"] = {"ce_loss_fwd_ms": t_ce3 * 1e3, "train_step_fwd_bwd_ms": t_ts3 * 1e3,
                                         "ce_loss": fwd_loss(m, x, yl).item(),
                                         "logits_bytes_never_materialised": 2048 * 50 * (cfg.num_items + 1) * 4}
out["c3_games_B2048"] = {"calculate_metrics_fp32_ms": t32 * 1e3, "fp32_users_per_s": 2048 / t32,
                         "bf16_k50_ms": t16 * 1e3, "bf16_users_per_s": 2048 / t16, "cpu_oracle_users_per_s": 2048 / tc}

# ---- C5: verbalizer, Llama-2-7B shaped ----------------------------------------------------------------------
v = synth.make_verbalizer_inputs()
h, w = v["hidden"].to(dev), v["lm_head"].to(dev)
class Tok:
    def __init__(self, ids): self.ids = ids
    def encode(self, word, add_special_tokens=False): return [int(self.ids[ord(word[-1]) - ord("A")])]
for pls in (False, True):
    vb = ManualVerbalizer(Tok(v["label_ids"]), classes=list(range(20)), label_words={i: chr(ord("A") + i) for i in range(20)},
                          prefix="", post_log_softmax=pls)
    t = gpu_time(lambda: vb.score_hidden(h, w), n=100)
    hf, wf = v["hidden"].float(), v["lm_head"].float()
    t0 = time.perf_counter(); lg = torch.nn.functional.linear(hf, wf); VO.process_logits(lg, vb.label_words_ids, vb.words_ids_mask, vb.label_words_mask, pls); tc = time.perf_counter() - t0
    got = vb.score_hidden(h, w, round_logits_to_bf16=False).cpu()
    ref = VO.process_logits(lg, vb.label_words_ids, vb.words_ids_mask, vb.label_words_mask, pls)
    algo_bytes = 512 * 4096 * 2 + 20 * 4096 * 2 + 512 * 20 * 4
    out[f"c5_verbalizer_post_log_softmax_{int(pls)}"] = {
        "gpu_us": t * 1e6, "users_per_s": 512 / t, "achieved_GBps_algorithmic": algo_bytes / t / 1e9,
        "cpu_full_vocab_lm_head_users_per_s": 512 / tc, "max_abs_err_vs_fp32_oracle": (got - ref).abs().max().item()}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/other_configs.json", "w"), indent=1)
