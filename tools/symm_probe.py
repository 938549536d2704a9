"""Developer probe: does torch's symmetric memory (peer-mapped buffers over NVLink) work in this container?
Run under torchrun with >= 2 GPUs.  Prints what works; never raises."""
import os, sys, time, traceback
import torch, torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.int32, device=f"cuda:{rank}")
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in h.buffer_ptrs][:4], "signal_pad", hex(h.signal_pad_ptrs[0]), flush=True)
    t.fill_(rank + 1)
    h.barrier(channel=0)
    peer = h.get_buffer((rank + 1) % world, (1 << 20,), torch.int32)
    v = int(peer[12345].item())
    print(rank, "peer read ->", v, "(expect", (rank + 1) % world + 1, ")", flush=True)
    # peer write + timing of a 5 MB pull
    h.barrier(channel=0)
    mine = torch.empty(1 << 20, dtype=torch.int32, device=f"cuda:{rank}")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        mine.copy_(peer)
    e1.record(); torch.cuda.synchronize()
    print(rank, "4 MB peer pull: %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3), flush=True)
    e0.record()
    for _ in range(20):
        h.barrier(channel=0)
    e1.record(); torch.cuda.synchronize()
    print(rank, "symm barrier: %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3), flush=True)
except Exception:
    print(rank, "symmetric memory probe FAILED:\n" + traceback.format_exc(), flush=True)
# NCCL latencies for comparison
x = torch.empty(4096, 2, 20, dtype=torch.int32, device=f"cuda:{rank}")
y = torch.empty(world * 4096, 2, 20, dtype=torch.int32, device=f"cuda:{rank}")
z = torch.empty(4096 * world, 2, 20, dtype=torch.int32, device=f"cuda:{rank}")
z2 = torch.empty_like(z)
for name, fn in (("all_gather 640KB/rank", lambda: dist.all_gather_into_tensor(y, x)),
                 ("all_to_all %dKB" % (z.numel() * 4 // 1024), lambda: dist.all_to_all_single(z2, z))):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(name, "%.1f us" % (e0.elapsed_time(e1) / 20 * 1e3), flush=True)
dist.barrier(); dist.destroy_process_group()
