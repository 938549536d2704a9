"""CPU: the oracle restatement against the fixtures produced by executing the reference
(oracle/make_golden.py).  This is what pins parity for everything downstream."""
import os
import pickle
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import CASES, GOLDEN, load_case, metrics_vector
from oracle import lru_oracle as O
from oracle import metrics_oracle as MO
from oracle import verbalizer_oracle as VO


@pytest.mark.parametrize("name", CASES)
def test_hidden_and_scores(golden_sd, name):
    c = load_case(name)
    ids = torch.from_numpy(c["ids"])
    hid = O.hidden_states(ids, golden_sd)
    np.testing.assert_allclose(hid.numpy(), c["hidden"], atol=2e-6, rtol=1e-5)
    last = O.last_scores(ids, golden_sd)
    np.testing.assert_allclose(last.numpy(), c["last_scores"], atol=2e-6, rtol=1e-5)
    # scoring only the last position is the same arithmetic as slicing the full forward (probe P4)
    full = O.forward_scores(ids, golden_sd)[:, -1, :]
    np.testing.assert_allclose(full.numpy(), last.numpy(), atol=1e-6, rtol=1e-6)


@pytest.mark.parametrize("name", CASES)
def test_metrics_and_topk(golden_sd, name):
    c = load_case(name)
    ids, labels = torch.from_numpy(c["ids"]), torch.from_numpy(c["labels"])
    ks = c["ks"].tolist()
    last = torch.from_numpy(c["last_scores"])
    raw = MO.recall_mrr_ndcg(last, labels, ks)
    np.testing.assert_allclose(metrics_vector(raw, ks), c["metrics_raw"], atol=1e-6)
    masked = O.mask_history(last, ids)
    mm = MO.recall_mrr_ndcg(masked, labels, ks)
    np.testing.assert_allclose(metrics_vector(mm, ks), c["metrics_masked"], atol=1e-6)
    s, i = O.topk_sorted(masked, 20)
    np.testing.assert_allclose(s.numpy(), c["top_scores"], atol=0, rtol=0)
    # ids equal wherever scores are not tied (torch.topk's tie order is implementation-defined)
    diff = i.numpy() != c["top_ids"]
    if diff.any():
        assert np.all(np.isclose(s.numpy()[diff], c["top_scores"][diff]))
    # chunked running top-k == one-shot top-k
    s2, i2 = O.retrieve(ids, golden_sd, 20, chunk=97)
    assert torch.equal(i2, i) and torch.equal(s2, s)


def test_tree_scan_equals_recurrence_for_left_padding():
    g = torch.Generator().manual_seed(0)
    B, Lp, H = 4, 64, 16
    bu = torch.complex(torch.randn(B, Lp, H, generator=g), torch.randn(B, Lp, H, generator=g))
    lam = 0.9 * torch.exp(1j * torch.rand(1, H, generator=g) * 6.28)
    mask = torch.zeros(B, Lp, dtype=torch.bool)
    for b, start in enumerate([0, 10, 33, 63]):
        mask[b, start:] = True
    a = O.tree_scan(bu, lam.to(torch.complex64), mask)
    r = O.sequential_scan_reference(bu, lam.to(torch.complex64), mask)
    for b, start in enumerate([0, 10, 33, 63]):
        assert (a[b, start:] - r[b, start:]).abs().max() < 2e-5


def test_generate_candidates_matches_reference_pickle(golden_sd):
    with open(os.path.join(GOLDEN, "generate_candidates_ref.pkl"), "rb") as f:
        g = pickle.load(f)
    bs, ks = g["batch"], g["ks"]
    mk = lambda ids, lab: [(torch.from_numpy(ids[i:i + bs]), torch.from_numpy(lab[i:i + bs]).unsqueeze(1))
                           for i in range(0, len(ids), bs)]
    args = SimpleNamespace(num_users=g["num_users"], num_items=g["num_items"], llm_negative_sample_size=19)
    ours = MO.generate_candidates(lambda x: O.last_scores(x, golden_sd), mk(g["ids_val"], g["lab_val"]),
                                  mk(g["ids_test"], g["lab_test"]), args, ks)
    ref = g["ref"]
    for key in ("val_users", "val_candidates", "test_users", "test_candidates", "non_test_users", "test_labels",
                "test_probs"):
        assert ours[key] == ref[key], key
    for key in ("val_metrics", "test_metrics"):
        for k, v in ref[key].items():
            assert abs(ours[key][k] - v) < 1e-6
    for key in ("retrieval_metrics", "non_retrieval_metrics"):
        for k, v in ref["test_retrieval"][key].items():
            assert abs(ours["test_retrieval"][key][k] - v) < 1e-6
    assert ours["test_retrieval"]["original_size"] == ref["test_retrieval"]["original_size"]
    assert ours["test_retrieval"]["retrieval_size"] == ref["test_retrieval"]["retrieval_size"]


def test_verbalizer_oracle():
    d = np.load(os.path.join(GOLDEN, "verbalizer_case.npz"))
    logits = torch.from_numpy(d["logits"])
    ids, tm, wm = (torch.from_numpy(d[k]) for k in ("label_words_ids", "words_ids_mask", "label_words_mask"))
    for pls in (0, 1):
        o = VO.process_logits(logits, ids, tm, wm, bool(pls))
        np.testing.assert_allclose(o.numpy(), d[f"out_pls{pls}"], atol=1e-6)
        lg = VO.lm_head_last(torch.from_numpy(d["hidden"]).to(torch.bfloat16),
                             torch.from_numpy(d["lm_head"]).to(torch.bfloat16))
        o2 = VO.process_logits(lg, ids, tm, wm, bool(pls))
        np.testing.assert_allclose(o2.numpy(), d[f"hidden_out_pls{pls}"], atol=1e-5)
    ids2, tm2, wm2 = (torch.from_numpy(d[k]) for k in ("multi_label_words_ids", "multi_words_ids_mask",
                                                       "multi_label_words_mask"))
    np.testing.assert_allclose(VO.process_logits(logits, ids2, tm2, wm2, True).numpy(), d["multi_out"], atol=1e-6)


@pytest.mark.parametrize("name", ["left_l20", "holes_l37", "left_l200"])
def test_oracle_train_step_loss_matches_reference(golden_sd, name):
    """oracle.ce_loss against LRUTrainer.calculate_loss of the reference itself (tests/golden/ce_case.npz,
    written by `python oracle/make_golden.py ce`)."""
    ce = np.load(os.path.join(GOLDEN, "ce_case.npz"))
    ids = torch.from_numpy(load_case(name)["ids"])
    labels = torch.from_numpy(ce[f"{name}_labels"])
    assert abs(O.ce_loss(ids, labels, golden_sd).item() - float(ce[f"{name}_loss"])) < 1e-6


def test_oracle_multi_token_handlers_match_reference():
    """first / max / mean handlers on multi-token label words against the reference class itself
    (tests/golden/verbalizer_handlers.npz, written by `python oracle/make_golden.py verb`)."""
    from llamarec_b200 import ManualVerbalizer

    class Tok:
        def encode(self, word, add_special_tokens=False):
            return [17 + (ord(c) * 7) % 250 for c in word]
    d = np.load(os.path.join(GOLDEN, "verbalizer_case.npz"))
    h = np.load(os.path.join(GOLDEN, "verbalizer_handlers.npz"))
    logits = torch.from_numpy(d["logits"])
    lw = {i: ([chr(ord("A") + i), "xy" + chr(ord("a") + i)] if i % 4 else [chr(ord("A") + i) + "q"]) for i in range(16)}
    for handler in ("first", "max", "mean"):
        for pls in (0, 1):
            v = ManualVerbalizer(Tok(), classes=list(range(16)), label_words=lw, prefix="",
                                 post_log_softmax=bool(pls), multi_token_handler=handler)   # label-word tables only
            o = VO.process_logits(logits, v.label_words_ids, v.words_ids_mask, v.label_words_mask, bool(pls), handler)
            np.testing.assert_allclose(o.numpy(), h[f"{handler}_pls{pls}"], atol=1e-6)
    # ManualVerbalizer.calibrate (trainer/verb.py:616-643) with registered calibration logits
    cal = torch.from_numpy(h["calibrate_logits"])
    for handler in ("first", "mean"):
        v = ManualVerbalizer(Tok(), classes=list(range(16)), label_words=lw, prefix="", post_log_softmax=True,
                             multi_token_handler=handler)
        o = VO.process_logits(logits, v.label_words_ids, v.words_ids_mask, v.label_words_mask, True, handler,
                              calibrate_logits=cal)
        np.testing.assert_allclose(o.numpy(), h[f"{handler}_calibrated"], atol=1e-5)
