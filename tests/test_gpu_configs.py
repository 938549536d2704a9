"""GPU (-m gpu): BASELINE.json configs 0, 1 and 4 at their STATED shapes (the headline config 3 is
test_full_size_10m_catalogue_properties and bench.py, config 2 is test_games_shaped_batch_2048_against_oracle).

C1  ML-100k shaped: 943 users x 1,682 items, max_len 200, eval batch 16 -- against outputs of the REFERENCE ITSELF
    (tests/golden/c1_ml100k_ref.pkl, written by `python oracle/make_golden.py c1`): top-20 lists of every user
    (trainer/lru.py:82-84), per-batch calculate_metrics (trainer/lru.py:30-42) and the generate_candidates pickle
    (trainer/lru.py:44-175).  north_star: "bit-identical top-20 candidate lists versus the reference on
    ML-100k-shaped data".
C2  Beauty shaped: 12,086 items, max_len 50, batch 64 -- forward at all positions, train-step loss, top-20,
    against the oracle on the same seeded inputs.
C5  Stage-2 verbalizer: 512 x 4096 hidden states, 32,000 x 4096 lm_head, 20 labels, both post_log_softmax modes.
Tolerances (north_star): scores 1e-3 relative, lists identical except ties inside the tolerance, metrics 4 decimals."""
import os
import pickle
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from llamarec_b200 import DeviceEvalSet, LRURec, LRURetriever, ManualVerbalizer, synth
from oracle import lru_oracle as O
from oracle import metrics_oracle as MO
from oracle import verbalizer_oracle as VO
from test_gpu_parity import assert_topk_equivalent

pytestmark = pytest.mark.gpu
RTOL = 1e-3
KEYS = ("Recall", "MRR", "NDCG")


def _args(n, ks=(1, 5, 10, 20, 50), num_users=None):
    return SimpleNamespace(num_items=n, bert_hidden_units=64, bert_num_blocks=2, bert_dropout=0.2,
                           bert_attn_dropout=0.2, metric_ks=list(ks), llm_negative_sample_size=19, num_users=num_users)


# ------------------------------------------------------------------------------------------------ C1
@pytest.fixture(scope="module")
def c1():
    with open(os.path.join(GOLDEN, "c1_ml100k_ref.pkl"), "rb") as f:
        ref = pickle.load(f)
    d = np.load(os.path.join(GOLDEN, "c1_ml100k_weights.npz"))
    sd = {k: torch.from_numpy(d[k]) for k in d.files}
    cfg = synth.CONFIGS["c1_ml100k"]
    assert (ref["num_items"], ref["num_users"], ref["max_len"], ref["batch"]) == (1682, 943, 200, 16)
    m = LRURec(_args(cfg.num_items))
    m.load_state_dict(sd)
    splits = {}
    for name in ("val", "test"):
        ids, labels = synth.make_sequences(cfg, seed=42, split=name)
        assert int(ids.sum().item()) == ref[f"{name}_ids_sha"], "synthetic sequences differ from the fixture's"
        splits[name] = (ids, labels)
    return m.cuda().eval(), sd, ref, splits, cfg


@pytest.mark.parametrize("split", ["val", "test"])
def test_c1_top20_lists_identical_to_the_reference(c1, split):
    m, sd, ref, splits, cfg = c1
    ids, _ = splits[split]
    ref_i = ref[f"{split}_top_ids"].astype(np.int64)
    ref_s = ref[f"{split}_top_scores"]
    got_i, got_s = [], []
    for i in range(0, ids.shape[0], cfg.batch):                  # the reference's eval batch (16), last one ragged
        res = m.retrieve(ids[i:i + cfg.batch].cuda(), k=20, exclude_history=True)     # precision 'auto' -> exact fp32
        got_i.append(res["ids"].cpu().numpy().astype(np.int64))
        got_s.append(res["scores"].cpu().numpy())
    got_i, got_s = np.concatenate(got_i), np.concatenate(got_s)
    np.testing.assert_allclose(got_s, ref_s, atol=1e-5, rtol=1e-4)
    identical = (got_i == ref_i).all(axis=1)
    # a list may differ from the reference's only where the reference's own scores tie to fp32 round-off
    # (torch.topk's tie order is implementation-defined, and two fp32 summation orders differ by ~1e-7)
    assert_topk_equivalent(got_i, got_s, ref_i, ref_s, rtol=2e-6)
    assert identical.mean() >= 0.995, f"only {identical.mean():.4f} of the {len(identical)} lists are bit-identical"
    # one batch of all 943 users gives the same lists as 59 batches of 16
    big = m.retrieve(ids.cuda(), k=20, exclude_history=True)
    assert np.array_equal(big["ids"].cpu().numpy().astype(np.int64), got_i)


@pytest.mark.parametrize("split", ["val", "test"])
def test_c1_calculate_metrics_per_batch_to_4_decimals(c1, split):
    m, sd, ref, splits, cfg = c1
    ids, labels = splits[split]
    ks = ref["ks"]
    tr = LRURetriever(_args(cfg.num_items, ks), m)
    want = ref[f"{split}_batch_metrics"]
    for bi, i in enumerate(range(0, ids.shape[0], cfg.batch)):
        got = tr.calculate_metrics((ids[i:i + cfg.batch].cuda(), labels[i:i + cfg.batch].cuda().view(-1, 1)))
        vec = np.array([got[f"{n}@{k}"] for k in ks for n in KEYS])
        np.testing.assert_allclose(vec, want[bi], atol=5e-5, err_msg=f"batch {bi}")


def test_c1_generate_candidates_pickle_matches_the_reference(c1, tmp_path):
    m, sd, ref, splits, cfg = c1
    ks, bs = ref["ks"], ref["batch"]
    want = ref["retrieved"]
    # the eval splits as pre-padded int32 device tensors (DeviceEvalSet) feed the batched sweep
    loaders = {name: DeviceEvalSet.from_tensors(ids, labels, batch_size=bs, device="cuda")
               for name, (ids, labels) in splits.items()}
    tr = LRURetriever(_args(cfg.num_items, ks, num_users=cfg.num_users), m, loaders["val"], loaders["test"])
    path = str(tmp_path / "retrieved.pkl")
    tr.generate_candidates(path)
    with open(path, "rb") as f:
        ours = pickle.load(f)
    assert set(ours.keys()) == set(want.keys())
    for key in ("val_users", "test_users", "non_test_users", "test_labels"):
        assert ours[key] == want[key], key
    # the reference's scores of the test users decide whether a differing list is a tie
    ref_scores = O.mask_history(O.last_scores(splits["test"][0], sd), splits["test"][0]).numpy()
    for key in ("val_candidates", "test_candidates", "test_probs"):
        assert len(ours[key]) == len(want[key]), key
        n_diff = 0
        for idx, (a, b) in enumerate(zip(ours[key], want[key])):
            assert len(a) == len(b)
            if a == b:
                continue
            n_diff += 1
            if key == "test_probs":
                sc = ref_scores[idx]
                for pa, pb in zip(a, b):
                    assert pa == pb or abs(sc[pa] - sc[pb]) <= 2e-6 * max(1.0, abs(sc[pb])), (key, idx, pa, pb)
        assert n_diff <= max(1, len(want[key]) // 200), (key, n_diff)
    for key in ("val_metrics", "test_metrics"):
        for k, v in want[key].items():
            assert abs(ours[key][k] - v) < 5e-5, (key, k, ours[key][k], v)
    for key in ("retrieval_metrics", "non_retrieval_metrics"):
        for k, v in want["test_retrieval"][key].items():
            assert abs(ours["test_retrieval"][key][k] - v) < 5e-5, (key, k)
    assert ours["test_retrieval"]["retrieval_size"] == want["test_retrieval"]["retrieval_size"]
    assert ours["test_retrieval"]["original_size"] == want["test_retrieval"]["original_size"] == 943


# ------------------------------------------------------------------------------------------------ C2
def test_c2_beauty_forward_loss_and_top20_batch_64():
    cfg = synth.CONFIGS["c2_beauty"]
    assert (cfg.num_items, cfg.max_len, cfg.batch) == (12086, 50, 64)
    sd = synth.make_state_dict(cfg.num_items, seed=42, bias_std=0.02)
    m = LRURec(_args(cfg.num_items))
    m.load_state_dict(sd)
    m = m.cuda().eval()
    ids, labels = synth.make_sequences(cfg, num_users=64, seed=42)
    x = ids.cuda()
    # train-step forward: logits at all 50 positions, [64, 50, 12087]  (model/lru.py:85)
    out = m(x)
    assert out.shape == (64, 50, cfg.num_items + 1)
    ref = O.forward_scores(ids, sd)
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), atol=1e-4, rtol=RTOL)
    # train-step loss without the logits tensor (trainer/lru.py:20-28)
    lab = torch.zeros_like(ids)
    lab[:, :-1] = ids[:, 1:]
    lab[:, -1] = labels
    lab[ids == 0] = 0
    loss = m.ce_loss(x, lab.cuda()).item()
    ref_loss = torch.nn.functional.cross_entropy(ref.view(-1, ref.size(-1)), lab.view(-1), ignore_index=0).item()
    assert abs(loss - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss)), (loss, ref_loss)
    # full-catalogue top-20 at the last position (trainer/lru.py:82-84), exact fp32 ('auto' at 12k items)
    res = m.retrieve(x, k=20, exclude_history=True, labels=labels.cuda(), ks=[1, 5, 10, 20])
    ref_s, ref_i = O.retrieve(ids, sd, 20)
    got_i = res["ids"].cpu().numpy().astype(np.int64)
    assert_topk_equivalent(got_i, res["scores"].cpu().numpy(), ref_i.numpy(), ref_s.numpy(), rtol=2e-6)
    assert (got_i == ref_i.numpy()).all(axis=1).mean() >= 0.98
    mm = MO.recall_mrr_ndcg(O.mask_history(ref[:, -1].clone(), ids), labels, [1, 5, 10, 20])
    sums = res["metric_sums"].cpu().numpy() / 64.0
    for ki, k in enumerate([1, 5, 10, 20]):
        for ni, n in enumerate(KEYS):
            assert abs(sums[ki, ni] - mm[f"{n}@{k}"]) < 5e-5, (n, k)


# ------------------------------------------------------------------------------------------------ C5
@pytest.mark.parametrize("post_log_softmax", [False, True])
def test_c5_verbalizer_llama2_7b_shape(post_log_softmax):
    v = synth.make_verbalizer_inputs()                                   # 512 x 4096, 32000 x 4096, 20 label ids
    assert tuple(v["hidden"].shape) == (512, 4096) and tuple(v["lm_head"].shape) == (32000, 4096)

    class Tok:
        def encode(self, word, add_special_tokens=False):
            return [int(v["label_ids"][ord(word[-1]) - ord("A")])]
    vb = ManualVerbalizer(Tok(), classes=list(range(20)), label_words={i: chr(ord("A") + i) for i in range(20)},
                          prefix="", post_log_softmax=post_log_softmax)
    h, w = v["hidden"].cuda(), v["lm_head"].cuda()
    # the reference chain: lm_head over the whole vocabulary (fp32 here), then process_logits (trainer/verb.py:546-586)
    logits = torch.nn.functional.linear(v["hidden"].float(), v["lm_head"].float())
    ref = VO.process_logits(logits, vb.label_words_ids, vb.words_ids_mask, vb.label_words_mask, post_log_softmax)
    exact = vb.score_hidden(h, w, round_logits_to_bf16=False).cpu()
    np.testing.assert_allclose(exact.numpy(), ref.numpy(), atol=5e-4, rtol=RTOL)
    assert torch.equal(exact.argmax(1), ref.argmax(1))
    # default mode rounds the 20 logits to bf16 like a bf16 lm_head followed by .float() (model/llm.py:113-131)
    ref16 = VO.process_logits(logits.to(torch.bfloat16).float(), vb.label_words_ids, vb.words_ids_mask,
                              vb.label_words_mask, post_log_softmax)
    got16 = vb.score_hidden(h, w).cpu()
    close = np.isclose(got16.numpy(), ref16.numpy(), atol=1e-3, rtol=1e-3).mean()
    assert close > 0.98, close                                           # a different summation order may flip a bf16 rounding
    np.testing.assert_allclose(got16.numpy(), ref16.numpy(), atol=0.1, rtol=2e-2)
