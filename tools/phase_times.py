"""Per-phase GPU times of one rank's share of an R-way sharded step, emulated on one GPU (developer tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
from llamarec_b200 import LRURec, synth, merge_lists
from llamarec_b200.sharded import shard_range

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N, B, L, K = 10_000_000, 4096, 50, 20
dev = torch.device("cuda")
args = SimpleNamespace(num_items=N, bert_hidden_units=64, bert_num_blocks=2, bert_dropout=0.2, bert_attn_dropout=0.2)
sd = synth.make_state_dict(1000, seed=42)
table, bias = synth.make_table_bf16(N, seed=42, device=dev)
m = LRURec(args)
m.load_state_dict({k: v for k, v in sd.items() if k not in ("embedding.token.weight", "model.bias")}, strict=False)
m = m.to(dev).eval()
with torch.no_grad():
    m.embedding.token.weight.copy_(table); m.model.bias.copy_(bias)
del table
lo, hi = shard_range(N + 1, 0, R)
m.set_row_shard(lo, hi)
ids, labels = synth.make_sequences_fast(B, N, L, seed=42)
ids, labels = ids.to(dev), labels.to(dev)
per = (B + R - 1) // R
payload = torch.empty(2, B, K, dtype=torch.int32, device=dev)
gathered = torch.empty(R, 2, B, K, dtype=torch.int32, device=dev)

def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

u_full = m.encode(ids)
print(f"R={R}: rows/rank={hi-lo}, users/rank (encode)={per}")
print("encode slice        us:", round(timeit(lambda: m.encode(ids[:per]))))
print("encode full batch   us:", round(timeit(lambda: m.encode(ids))))
print("prep full (excl)    us:", round(timeit(lambda: m._prepare_sequences(ids, False, True))))
print("score+local merge   us:", round(timeit(lambda: m.retrieve(ids, k=K, u=u_full, precision="bf16", packed_out=payload))))
m.profile_events = []
for _ in range(10): m.retrieve(ids, k=K, u=u_full, precision="bf16", packed_out=payload)
torch.cuda.synchronize()
print("  score kernel only us:", round(sum(a.elapsed_time(b) for a, b in m.profile_events) / len(m.profile_events) * 1e3))
m.profile_events = None
for r in range(R): gathered[r].copy_(payload)
fm = lambda: merge_lists(gathered[:, 0].view(torch.float32), gathered[:, 1], None, k_out=K, labels=labels, ks=[1, 5, 10, 20],
                         layout="list_major", strides=(2 * B * K, K))
print("final merge+metrics us:", round(timeit(fm)))
t0 = time.perf_counter()
for _ in range(50): m.retrieve(ids, k=K, u=u_full, precision="bf16", packed_out=payload)
print("python launch cost of retrieve() us:", round((time.perf_counter() - t0) / 50 * 1e6)); torch.cuda.synchronize()

# ---- weak scaling (data-parallel users): one rank scores R*B users against its 1/R of the rows ----
from llamarec_b200.sharded import CudaBackend
be = CudaBackend(m, 0, R, precision="bf16")
sts = []
for r in range(R):
    xr, _ = synth.make_sequences_fast(B, N, L, seed=42 + r)
    st = be.encode_states(xr.to(dev), True)
    sts.append({k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()})
state = torch.cat([s["state"] for s in sts]); excl = torch.cat([s["excl"] for s in sts]); bloom = torch.cat([s["bloom"] for s in sts])
stride = sts[0]["excl_stride"]
print(f"weak R={R}: users scored per rank={state.shape[0]}")
print("encode_states (own 4096) us:", round(timeit(lambda: be.encode_states(ids, True))))
print("score R*B + local merge  us:", round(timeit(lambda: be.local_topk_rows(state, excl, bloom, stride, K), n=10)))
m.profile_events = []
for _ in range(5): be.local_topk_rows(state, excl, bloom, stride, K)
torch.cuda.synchronize()
print("  score call only        us:", round(sum(a.elapsed_time(b) for a, b in m.profile_events) / len(m.profile_events) * 1e3))
m.profile_events = None
recv = torch.empty(R, B, 2, K, dtype=torch.int32, device=dev)
rows = be.local_topk_rows(state, excl, bloom, stride, K)
for r in range(R): recv[r].copy_(rows[:B])
print("final merge (own 4096, R lists) us:", round(timeit(lambda: be.merge_rows(recv, K, labels, [1, 5, 10, 20]))))
