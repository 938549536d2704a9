"""GPU, world_size 2 under NCCL (skipped unless launched with >= 2 visible GPUs): the row-sharded
retriever reproduces the single-GPU result exactly."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["LRB_ROOT"])
from types import SimpleNamespace
from llamarec_b200 import LRURec, synth
from llamarec_b200.sharded import CudaBackend, ShardedRetriever
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
n_items = 200_000
args = SimpleNamespace(num_items=n_items, bert_hidden_units=64, bert_num_blocks=2, bert_dropout=0.2, bert_attn_dropout=0.2)
sd = synth.make_state_dict(n_items, seed=3, bias_std=0.01)
ids, labels = synth.make_sequences_fast(777, n_items, 50, seed=5)
ref_model = LRURec(args); ref_model.load_state_dict(sd); ref_model = ref_model.cuda().eval()
ref = ref_model.retrieve(ids.cuda(), k=20, labels=labels.cuda(), ks=[1, 5, 10, 20], precision="bf16")
m = LRURec(args); m.load_state_dict(sd); m = m.cuda().eval()
sh = ShardedRetriever(CudaBackend(m, rank, world, precision="bf16"))
out = sh.retrieve(ids.cuda(), k=20, labels=labels.cuda(), ks=[1, 5, 10, 20])
assert torch.equal(out["ids"], ref["ids"]), "sharded ids differ"
assert torch.equal(out["scores"], ref["scores"]), "sharded scores differ"
assert torch.equal(out["label_rank"], ref["label_rank"])
assert torch.allclose(out["metric_sums"], ref["metric_sums"], rtol=1e-5)
# data-parallel users: each rank brings its own slice, gets its own users' lists against the whole catalogue
b = 384
lo = rank * b
dp = sh.retrieve_dp(ids[lo:lo + b].cuda(), k=20, labels=labels[lo:lo + b].cuda(), ks=[1, 5, 10, 20])
assert torch.equal(dp["ids"], ref["ids"][lo:lo + b]), "dp ids differ"
assert torch.equal(dp["scores"], ref["scores"][lo:lo + b]), "dp scores differ"
assert torch.equal(dp["label_rank"], ref["label_rank"][lo:lo + b])
assert sh._peer and all(v is not None for v in sh._peer.values()), "peer-memory exchange was not used"
# same step on torch.distributed collectives, and the peer path again (double-buffer parity, 3 more steps)
shc = ShardedRetriever(CudaBackend(m, rank, world, precision="bf16"), exchange="collective")
dc = shc.retrieve_dp(ids[lo:lo + b].cuda(), k=20, labels=labels[lo:lo + b].cuda(), ks=[1, 5, 10, 20])
assert torch.equal(dc["ids"], dp["ids"]) and torch.equal(dc["scores"], dp["scores"])
for it in range(3):
    sl = slice((lo + 7 * it) % 300, (lo + 7 * it) % 300 + b)
    d2 = sh.retrieve_dp(ids[sl].cuda(), k=20, labels=labels[sl].cuda(), ks=[1, 5, 10, 20])
    ref2 = ref_model.retrieve(ids[sl].cuda(), k=20, precision="bf16")
    assert torch.equal(d2["ids"], ref2["ids"]) and torch.equal(d2["scores"], ref2["scores"]), f"peer step {it}"
# one user per rank (2 users in total) over the collective path: the [B, 2, k] payload must not be mistaken for [2, B, k]
d1 = shc.retrieve_dp(ids[rank:rank + 1].cuda(), k=20, labels=labels[rank:rank + 1].cuda(), ks=[1, 5, 10, 20])
assert torch.equal(d1["ids"], ref["ids"][rank:rank + 1]) and torch.equal(d1["scores"], ref["scores"][rank:rank + 1]), "b=1"
d1p = sh.retrieve_dp(ids[rank:rank + 1].cuda(), k=20, labels=labels[rank:rank + 1].cuda(), ks=[1, 5, 10, 20])
assert torch.equal(d1p["ids"], ref["ids"][rank:rank + 1]), "b=1 peer"
# more gathered users than one scoring launch takes (2 x 5000 > 9472): consecutive chunk launches (programmatic
# dependent launch) inside the data-parallel step
big_ids, big_lab = synth.make_sequences_fast(10000, n_items, 50, seed=9)
bb = 5000
mine = slice(rank * bb, (rank + 1) * bb)
dbig = sh.retrieve_dp(big_ids[mine].cuda(), k=20, labels=big_lab[mine].cuda(), ks=[1, 5, 10, 20])
rbig = ref_model.retrieve(big_ids[mine].cuda(), k=20, labels=big_lab[mine].cuda(), ks=[1, 5, 10, 20], precision="bf16")
assert torch.equal(dbig["ids"], rbig["ids"]) and torch.equal(dbig["scores"], rbig["scores"]), "chunked dp"
assert torch.equal(dbig["label_rank"], rbig["label_rank"])
# 'auto' precision is resolved from the whole catalogue, identically on every rank
auto = CudaBackend(m, rank, world, precision="auto")
assert auto.precision == "bf16"
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_retrieval_matches_single_gpu_nccl(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, LRB_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
