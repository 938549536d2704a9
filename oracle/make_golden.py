"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE (GarciaLnk/LlamaRec at /root/reference).

Run in the build container only (the reference is not present on the GPU box):
    python oracle/make_golden.py
The reference has no tests or golden vectors of its own; these fixtures are the
pin for oracle/ and, through it, for the CUDA kernels.  The script also asserts that the oracle
restatement agrees with the reference on every fixture before writing anything.

Import shim: `trainer/__init__` pulls in yacs/peft and `config.py` parses
sys.argv at import, so the package is registered empty and argv is neutralised first.
"""
from __future__ import annotations

import os
import pickle
import sys
import tempfile
import types
from types import SimpleNamespace

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)


def import_reference():
    sys.argv = ["x"]
    sys.path.insert(0, REF)
    pkg = types.ModuleType("trainer")
    pkg.__path__ = [os.path.join(REF, "trainer")]
    sys.modules["trainer"] = pkg
    # trainer/base.py imports loggers -> tensorboard/wandb are present; pytorch_lightning is not needed
    import config  # noqa: F401  (parses the neutralised argv)
    from model.lru import LRURec
    import trainer.utils as tutils
    import trainer.lru as tlru
    # demo/verb.py holds the same ManualVerbalizer as trainer/verb.py without the yacs import
    sys.path.insert(0, os.path.join(REF, "demo"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_demo_verb", os.path.join(REF, "demo", "verb.py"))
    verb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(verb)
    return config, LRURec, tutils, tlru, verb


def make_ids(rng, B, L, N, kind):
    ids = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        if kind == "left":
            n = int(rng.integers(0, L + 1)) if b > 1 else (0 if b == 0 else L)   # includes empty and full rows
            ids[b, L - n:] = rng.integers(1, N + 1, size=n)
        else:   # zeros scattered anywhere: exercises the tree scan's boundary-only masking
            ids[b] = rng.integers(1, N + 1, size=L)
            ids[b, rng.random(L) < 0.3] = 0
    return ids


def main():
    os.makedirs(OUT, exist_ok=True)
    config, RefLRURec, tutils, tlru, verb = import_reference()
    from oracle import lru_oracle as O
    from oracle import metrics_oracle as MO
    from oracle import verbalizer_oracle as VO

    torch.manual_seed(42)
    N = 400
    args = SimpleNamespace(num_items=N, bert_hidden_units=64, bert_num_blocks=2, bert_dropout=0.2,
                           bert_attn_dropout=0.2)
    ref = RefLRURec(args).eval()
    with torch.no_grad():
        # non-trivial LayerNorm affine and bias so every parameter is exercised
        for n, p in ref.named_parameters():
            if "layer_norm.weight" in n:
                p.add_(0.1 * torch.randn_like(p))
            if "layer_norm.bias" in n:
                p.add_(0.05 * torch.randn_like(p))
        ref.model.bias.copy_(0.02 * torch.randn(N + 1))
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    np.savez(os.path.join(OUT, "lru_weights_n400.npz"), **{k: v.numpy() for k, v in sd.items()})

    rng = np.random.default_rng(7)
    ks = [1, 5, 10, 20, 50]
    for name, B, L, kind in [("left_l20", 12, 20, "left"), ("left_l50", 16, 50, "left"), ("left_l200", 6, 200, "left"),
                             ("holes_l50", 8, 50, "holes"), ("holes_l37", 8, 37, "holes")]:
        ids = torch.from_numpy(make_ids(rng, B, L, N, kind))
        labels = torch.from_numpy(rng.integers(1, N + 1, size=B).astype(np.int64))
        with torch.no_grad():
            scores_all = ref(ids)                                  # [B, L, N+1]
            x, mask = ref.embedding(ids)
            # hidden states: re-run the model body without the scoring matmul
            import torch.nn.functional as F
            Lp = 1 << int(np.ceil(np.log2(L)))
            xx = F.pad(x, (0, 0, Lp - L, 0))
            mm = F.pad(mask, (Lp - L, 0))
            for blk in ref.model.lru_blocks:
                xx = blk.forward(xx, mm)
            hidden = xx[:, -L:]
            last = scores_all[:, -1, :].clone()
            m_raw = tutils.absolute_recall_mrr_ndcg_for_ks(last.clone(), labels, ks)
            masked = last.clone()
            for i in range(L):
                masked[torch.arange(B), ids[:, i]] = -1e9
            masked[:, 0] = -1e9
            m_masked = tutils.absolute_recall_mrr_ndcg_for_ks(masked.clone(), labels, ks)
            top_s, top_i = torch.topk(masked, 20)
        # --- the oracle must reproduce the reference before the fixture is trusted ---
        assert torch.allclose(O.hidden_states(ids, sd), hidden, atol=2e-6, rtol=1e-5), name
        assert torch.allclose(O.forward_scores(ids, sd), scores_all, atol=2e-6, rtol=1e-5), name
        assert torch.allclose(O.last_scores(ids, sd), last, atol=2e-6, rtol=1e-5), name
        om = MO.recall_mrr_ndcg(O.mask_history(O.last_scores(ids, sd), ids), labels, ks)
        for k_, v_ in m_masked.items():
            assert abs(om[k_] - v_) < 1e-6, (name, k_, om[k_], v_)
        np.savez(os.path.join(OUT, f"lru_case_{name}.npz"), ids=ids.numpy(), labels=labels.numpy(),
                 hidden=hidden.numpy(), last_scores=last.numpy(), top_scores=top_s.numpy(), top_ids=top_i.numpy(),
                 metrics_raw=np.array([m_raw[f"{n}@{k}"] for k in ks for n in ("Recall", "MRR", "NDCG")]),
                 metrics_masked=np.array([m_masked[f"{n}@{k}"] for k in ks for n in ("Recall", "MRR", "NDCG")]),
                 ks=np.array(ks))
        print("case", name, "ok")

    # ---- generate_candidates: the reference's per-user loop on a tiny val/test split ----
    U, L = 37, 20
    ids_val = torch.from_numpy(make_ids(rng, U, L, N, "left"))
    ids_test = torch.from_numpy(make_ids(rng, U, L, N, "left"))
    # make labels plausible: a third of the users get their top-scored unseen item as label
    with torch.no_grad():
        def plausible(ids):
            s = ref(ids)[:, -1, :].clone()
            for i in range(L):
                s[torch.arange(U), ids[:, i]] = -1e9
            s[:, 0] = -1e9
            order = (-s).argsort(1)
            lab = torch.from_numpy(rng.integers(1, N + 1, size=U).astype(np.int64))
            for u in range(U):
                if u % 3 == 0:
                    lab[u] = order[u, int(rng.integers(0, 30))]
            return lab
        lab_val, lab_test = plausible(ids_val), plausible(ids_test)
    bs = 16
    val_loader = [(ids_val[i:i + bs], lab_val[i:i + bs].unsqueeze(1)) for i in range(0, U, bs)]
    test_loader = [(ids_test[i:i + bs], lab_test[i:i + bs].unsqueeze(1)) for i in range(0, U, bs)]
    targs = SimpleNamespace(num_users=U + 3, num_items=N, llm_negative_sample_size=19, metric_ks=ks)
    config.args.metric_ks = ks
    config.args.num_items = N
    fake = SimpleNamespace(model=ref, metric_ks=ks, val_loader=val_loader, test_loader=test_loader, args=targs,
                           to_device=lambda b: b)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "retrieved.pkl")
        tlru.LRUTrainer.generate_candidates(fake, path)
        with open(path, "rb") as f:
            ref_pkl = pickle.load(f)
    ours = MO.generate_candidates(lambda x: O.last_scores(x, sd), val_loader, test_loader, targs, ks)
    for key in ("val_users", "val_candidates", "test_users", "test_candidates", "non_test_users", "test_labels"):
        assert ours[key] == ref_pkl[key], key
    for key in ("val_metrics", "test_metrics"):
        for k_, v_ in ref_pkl[key].items():
            assert abs(ours[key][k_] - v_) < 1e-6, (key, k_)
    with open(os.path.join(OUT, "generate_candidates_ref.pkl"), "wb") as f:
        pickle.dump({"ref": ref_pkl, "ids_val": ids_val.numpy(), "ids_test": ids_test.numpy(),
                     "lab_val": lab_val.numpy(), "lab_test": lab_test.numpy(), "batch": bs,
                     "num_users": U + 3, "num_items": N, "ks": ks}, f)
    print("generate_candidates ok")

    # ---- verbalizer ----
    class Tok:
        def __init__(self):
            self.vocab = {}
        def encode(self, word, add_special_tokens=False):
            # deterministic fake tokenizer: single-letter words -> one id; longer words -> one id per char
            return [17 + (ord(c) * 7) % 250 for c in word]
    tok = Tok()
    classes = list(range(20))
    label_words = {i: chr(ord("A") + i) for i in range(20)}
    V, Bv, Hd = 300, 9, 512
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(Bv, V, generator=g) * 3
    hid = torch.randn(Bv, Hd, generator=g).to(torch.bfloat16)
    wgt = (torch.randn(V, Hd, generator=g) * 0.05).to(torch.bfloat16)
    out = {}
    for pls in (False, True):
        rv = verb.ManualVerbalizer(tokenizer=tok, prefix="", post_log_softmax=pls, classes=classes,
                                   label_words=label_words)
        r = rv.process_logits(logits.clone())
        o = VO.process_logits(logits, rv.label_words_ids.data, rv.words_ids_mask.data, rv.label_words_mask.data,
                              post_log_softmax=pls)
        assert torch.allclose(o, r, atol=1e-6), pls
        out[f"out_pls{int(pls)}"] = r.detach().numpy()
        # the stage-2 tail as the reference computes it: bf16 lm_head on the last position, .float()
        lg = torch.nn.functional.linear(hid, wgt).float()
        out[f"hidden_out_pls{int(pls)}"] = rv.process_logits(lg.clone()).detach().numpy()
        out["label_words_ids"] = rv.label_words_ids.data.numpy()
        out["words_ids_mask"] = rv.words_ids_mask.data.numpy()
        out["label_words_mask"] = rv.label_words_mask.data.numpy()
    # two label words per class, one of them multi-token, one class with a single word (mask path)
    lw2 = {i: ([chr(ord("A") + i), "x" + chr(ord("a") + i)] if i % 4 else [chr(ord("A") + i)]) for i in range(20)}
    rv2 = verb.ManualVerbalizer(tokenizer=tok, prefix="", post_log_softmax=True, classes=classes, label_words=lw2)
    out["multi_out"] = rv2.process_logits(logits.clone()).detach().numpy()
    o2 = VO.process_logits(logits, rv2.label_words_ids.data, rv2.words_ids_mask.data, rv2.label_words_mask.data, True)
    assert torch.allclose(o2, rv2.process_logits(logits.clone()), atol=1e-6)
    out["multi_label_words_ids"] = rv2.label_words_ids.data.numpy()
    out["multi_words_ids_mask"] = rv2.words_ids_mask.data.numpy()
    out["multi_label_words_mask"] = rv2.label_words_mask.data.numpy()
    np.savez(os.path.join(OUT, "verbalizer_case.npz"), logits=logits.numpy(), hidden=hid.float().numpy(),
             lm_head=wgt.float().numpy(), **out)
    print("verbalizer ok")


def make_ce_fixture():
    """Train-step loss (trainer/lru.py:20-28) of the REFERENCE on the committed weights and id matrices:
    `python oracle/make_golden.py ce` writes tests/golden/ce_case.npz without touching the other fixtures."""
    config, RefLRURec, tutils, tlru, verb = import_reference()
    from oracle import lru_oracle as O
    N = 400
    args = SimpleNamespace(num_items=N, bert_hidden_units=64, bert_num_blocks=2, bert_dropout=0.2,
                           bert_attn_dropout=0.2)
    d = np.load(os.path.join(OUT, "lru_weights_n400.npz"))
    sd = {k: torch.from_numpy(d[k]) for k in d.files}
    ref = RefLRURec(args)
    ref.load_state_dict(sd)
    ref.eval()                                                    # dropout off: the RNG cannot be parity-matched
    fake = SimpleNamespace(model=ref, ce=torch.nn.CrossEntropyLoss(ignore_index=0))
    out = {}
    g = torch.Generator().manual_seed(1)
    for name in ("left_l20", "holes_l37", "left_l200"):
        ids = torch.from_numpy(np.load(os.path.join(OUT, f"lru_case_{name}.npz"))["ids"])
        labels = torch.zeros_like(ids)
        labels[:, :-1] = ids[:, 1:]                               # next-item targets (dataloader/lru.py:98-118)
        labels[:, -1] = torch.randint(1, N + 1, (ids.shape[0],), generator=g)
        labels[ids == 0] = 0                                      # padding positions are ignored
        with torch.no_grad():
            loss = tlru.LRUTrainer.calculate_loss(fake, (ids, labels))
            rows = torch.nn.functional.cross_entropy(ref(ids).view(-1, N + 1), labels.view(-1), ignore_index=0,
                                                     reduction="none").view(ids.shape)
        assert abs(O.ce_loss(ids, labels, sd).item() - loss.item()) < 1e-6, name
        out[f"{name}_labels"] = labels.numpy()
        out[f"{name}_loss"] = np.array(loss.item(), dtype=np.float64)
        out[f"{name}_row_loss"] = rows.numpy()
        print("ce", name, loss.item())
    np.savez(os.path.join(OUT, "ce_case.npz"), **out)


def make_verbalizer_handler_fixture():
    """ManualVerbalizer.process_logits of the REFERENCE with multi_token_handler = max / mean on multi-token label
    words: `python oracle/make_golden.py verb` writes tests/golden/verbalizer_handlers.npz."""
    config, RefLRURec, tutils, tlru, verb = import_reference()
    from oracle import verbalizer_oracle as VO

    class Tok:
        def encode(self, word, add_special_tokens=False):
            return [17 + (ord(c) * 7) % 250 for c in word]
    d = np.load(os.path.join(OUT, "verbalizer_case.npz"))
    logits = torch.from_numpy(d["logits"])
    classes = list(range(16))     # 16 classes x 2 words = 32 label words (the kernel's lanes)
    lw = {i: ([chr(ord("A") + i), "xy" + chr(ord("a") + i)] if i % 4 else [chr(ord("A") + i) + "q"]) for i in range(16)}
    out = {}
    for handler in ("first", "max", "mean"):
        for pls in (False, True):
            rv = verb.ManualVerbalizer(tokenizer=Tok(), prefix="", post_log_softmax=pls, classes=classes,
                                       label_words=lw, multi_token_handler=handler)
            r = rv.process_logits(logits.clone())
            o = VO.process_logits(logits, rv.label_words_ids.data, rv.words_ids_mask.data, rv.label_words_mask.data,
                                  pls, handler)
            assert torch.allclose(o, r, atol=1e-6), (handler, pls)
            out[f"{handler}_pls{int(pls)}"] = r.detach().numpy()
    # calibration (register_calibrate_logits + ManualVerbalizer.calibrate, trainer/verb.py:202-208,616-643)
    gcal = torch.Generator().manual_seed(11)
    cal = torch.randn(logits.shape[1], generator=gcal) * 2.0
    out["calibrate_logits"] = cal.numpy()
    for handler in ("first", "mean"):
        rv = verb.ManualVerbalizer(tokenizer=Tok(), prefix="", post_log_softmax=True, classes=classes,
                                   label_words=lw, multi_token_handler=handler)
        rv.register_calibrate_logits(cal.clone())
        r = rv.process_logits(logits.clone())
        o = VO.process_logits(logits, rv.label_words_ids.data, rv.words_ids_mask.data, rv.label_words_mask.data,
                              True, handler, calibrate_logits=cal)
        assert torch.allclose(o, r, atol=1e-5), (handler, float((o - r).abs().max()))
        out[f"{handler}_calibrated"] = r.detach().numpy()
    np.savez(os.path.join(OUT, "verbalizer_handlers.npz"), **out)
    print("verbalizer handlers ok")


def make_c1_fixture():
    """BASELINE.json configs[0] at its stated shape, computed by the REFERENCE itself: 943 users x 1,682 items,
    max_len 200, eval batch 16 (config.py:104).  The model is the reference's own `LRURec(args)` under
    torch.manual_seed(42) -- including its truncated-normal model.bias (model/lru.py:16-36 initialises every
    parameter whose name has neither 'layer_norm' nor 'params_log').  Records, for the val and test splits of
    synth.make_sequences (the same generator the GPU tests call): per-batch `calculate_metrics` dicts
    (trainer/lru.py:30-42), the masked top-20 of every user (trainer/lru.py:82-84) and the complete
    `generate_candidates` pickle (trainer/lru.py:44-175).
    `python oracle/make_golden.py c1` writes tests/golden/c1_ml100k_ref.pkl + c1_ml100k_weights.npz."""
    config, RefLRURec, tutils, tlru, verb = import_reference()
    from oracle import lru_oracle as O
    from oracle import metrics_oracle as MO
    from llamarec_b200 import synth
    cfg = synth.CONFIGS["c1_ml100k"]
    N, L, bs, ks = cfg.num_items, cfg.max_len, cfg.batch, list(cfg.metric_ks)
    torch.manual_seed(42)
    args = SimpleNamespace(num_items=N, bert_hidden_units=64, bert_num_blocks=2, bert_dropout=0.2,
                           bert_attn_dropout=0.2)
    ref = RefLRURec(args).eval()
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    assert float(sd["model.bias"].abs().max()) > 0          # the reference does initialise the bias
    splits = {name: synth.make_sequences(cfg, seed=42, split=name) for name in ("val", "test")}
    out = {"ks": ks, "batch": bs, "num_items": N, "num_users": cfg.num_users, "max_len": L}
    with torch.no_grad():
        for name, (ids, labels) in splits.items():
            fake = SimpleNamespace(model=ref, metric_ks=ks)
            per_batch, top_i, top_s = [], [], []
            for i in range(0, ids.shape[0], bs):
                x, y = ids[i:i + bs], labels[i:i + bs].unsqueeze(1)
                m = tlru.LRUTrainer.calculate_metrics(fake, (x, y))
                per_batch.append([m[f"{n}@{k}"] for k in ks for n in ("Recall", "MRR", "NDCG")])
                scores = ref(x)[:, -1, :]
                for j in range(L):
                    scores[torch.arange(scores.size(0)), x[:, j]] = -1e9
                scores[:, 0] = -1e9
                s, t = torch.topk(scores, 20)
                top_i.append(t)
                top_s.append(s)
                # the oracle agrees with the reference on this batch (scores to fp32 round-off, metrics exactly)
                o = O.mask_history(O.last_scores(x, sd), x)
                assert torch.allclose(o, scores, atol=2e-6, rtol=1e-5), (name, i)
                om = MO.recall_mrr_ndcg(o, y.view(-1), ks)
                for k_, v_ in m.items():
                    assert abs(om[k_] - v_) < 1e-6, (name, i, k_)
            out[f"{name}_batch_metrics"] = np.array(per_batch, dtype=np.float64)
            out[f"{name}_top_ids"] = torch.cat(top_i).numpy().astype(np.int16)
            out[f"{name}_top_scores"] = torch.cat(top_s).numpy()
            out[f"{name}_ids_sha"] = int(ids.sum().item())
    mk = lambda ids, lab: [(ids[i:i + bs], lab[i:i + bs].unsqueeze(1)) for i in range(0, ids.shape[0], bs)]
    targs = SimpleNamespace(num_users=cfg.num_users, num_items=N, llm_negative_sample_size=19, metric_ks=ks)
    config.args.metric_ks = ks
    config.args.num_items = N
    fake = SimpleNamespace(model=ref, metric_ks=ks, val_loader=mk(*splits["val"]), test_loader=mk(*splits["test"]),
                           args=targs, to_device=lambda b: b)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "retrieved.pkl")
        tlru.LRUTrainer.generate_candidates(fake, path)
        with open(path, "rb") as f:
            out["retrieved"] = pickle.load(f)
    np.savez(os.path.join(OUT, "c1_ml100k_weights.npz"), **{k: v.numpy() for k, v in sd.items()})
    with open(os.path.join(OUT, "c1_ml100k_ref.pkl"), "wb") as f:
        pickle.dump(out, f)
    print("c1 ok:", {k: round(v, 4) for k, v in out["retrieved"]["test_metrics"].items() if "@10" in k},
          "retrieval_size", out["retrieved"]["test_retrieval"]["retrieval_size"])


def make_grad_fixture():
    """Gradient of the train-step loss w.r.t. EVERY parameter, from the REFERENCE's own autograd
    (trainer/lru.py:20-28 + loss.backward(), trainer/base.py:107-111) on the committed weights, eval mode (dropout
    off: the RNG cannot be parity-matched), left-padded batches: `python oracle/make_golden.py grad` writes
    tests/golden/grad_case.npz."""
    config, RefLRURec, tutils, tlru, verb = import_reference()
    N = 400
    args = SimpleNamespace(num_items=N, bert_hidden_units=64, bert_num_blocks=2, bert_dropout=0.2,
                           bert_attn_dropout=0.2)
    d = np.load(os.path.join(OUT, "lru_weights_n400.npz"))
    sd = {k: torch.from_numpy(d[k]) for k in d.files}
    out = {}
    g = torch.Generator().manual_seed(2)
    for name in ("left_l50", "left_l20"):
        ref = RefLRURec(args)
        ref.load_state_dict(sd)
        ref.eval()
        fake = SimpleNamespace(model=ref, ce=torch.nn.CrossEntropyLoss(ignore_index=0))
        ids = torch.from_numpy(np.load(os.path.join(OUT, f"lru_case_{name}.npz"))["ids"])
        labels = torch.zeros_like(ids)
        labels[:, :-1] = ids[:, 1:]
        labels[:, -1] = torch.randint(1, N + 1, (ids.shape[0],), generator=g)
        labels[ids == 0] = 0
        loss = tlru.LRUTrainer.calculate_loss(fake, (ids, labels))
        loss.backward()
        out[f"{name}_labels"] = labels.numpy()
        out[f"{name}_loss"] = np.array(loss.item(), dtype=np.float64)
        for k, p_ in ref.named_parameters():
            gr = p_.grad
            out[f"{name}:{k}"] = gr.detach().numpy()
        print("grad", name, loss.item(), {k: float(p_.grad.abs().max()) for k, p_ in list(ref.named_parameters())[:3]})
    np.savez_compressed(os.path.join(OUT, "grad_case.npz"), **out)


def make_evalset_fixture():
    """LRUValidDataset / LRUTestDataset of the REFERENCE (dataloader/lru.py:129-180) on random user histories:
    `python oracle/make_golden.py evalset` writes tests/golden/evalset_case.npz."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("ref_dataloader_lru", os.path.join(REF, "dataloader", "lru.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(11)
    max_len = 12
    train, val, test = {}, {}, {}
    for u in rng.permutation(np.arange(1, 41)).tolist():              # unsorted insertion order on purpose
        n = int(rng.integers(0, 30))
        train[u] = rng.integers(1, 500, size=n).tolist()
        val[u] = [int(rng.integers(1, 500))] if rng.random() < 0.85 else []
        test[u] = [int(rng.integers(1, 500))] if rng.random() < 0.85 else []
    dv = mod.LRUValidDataset(None, train, val, max_len, None)
    dt = mod.LRUTestDataset(None, train, val, test, max_len, None)
    out = {"max_len": np.array(max_len)}
    for name, ds in (("val", dv), ("test", dt)):
        items = [ds[i] for i in range(len(ds))]
        out[f"{name}_seqs"] = torch.stack([a for a, _ in items]).numpy()
        out[f"{name}_labels"] = torch.stack([b for _, b in items]).numpy()
        out[f"{name}_users"] = np.array(ds.users)
    out["dicts_json"] = np.array(json.dumps({"train": train, "val": val, "test": test}))
    np.savez(os.path.join(OUT, "evalset_case.npz"), **out)
    print("evalset ok", len(dv), len(dt))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "evalset":
        make_evalset_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "verb":
        make_verbalizer_handler_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "ce":
        make_ce_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "grad":
        make_grad_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "c1":
        make_c1_fixture()
    else:
        main()
