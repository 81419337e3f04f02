"""All-reduce bandwidth with few CTAs (what parallel.DataParallel allows NCCL), and whether torch's symmetric memory
rendezvous works on this box (needed for a hand-written NVLink peer-memory all-reduce).  Run under torchrun."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def bw(nbytes, op, iters=30):
    t = torch.ones(nbytes // 4, device=dev)
    for _ in range(5):
        dist.all_reduce(t, op=op)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dist.all_reduce(t, op=op)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, nbytes / ms / 1e6


for mb in (1, 4, 16, 64):
    for name, op in (("SUM", dist.ReduceOp.SUM), ("AVG", dist.ReduceOp.AVG)):
        ms, gbs = bw(mb << 20, op)
        if rank == 0:
            print(f"NCCL_MAX_CTAS={os.environ.get('NCCL_MAX_CTAS')} all_reduce {name} {mb:3d} MB: {ms * 1e3:8.1f} us  algbw {gbs:7.1f} GB/s",
                  flush=True)
if os.environ.get("PROBE_SYMM") == "1":
    try:
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(16 << 18, device=dev, dtype=torch.float32)          # 16 MB
        hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
        t.fill_(rank + 1)
        hdl.barrier()
        peer = hdl.get_buffer((rank + 1) % world, t.shape, t.dtype)
        torch.cuda.synchronize()
        if rank == 0:
            print("symmetric memory rendezvous OK; peer buffer first element:", float(peer[0]), "buffer_ptrs", len(hdl.buffer_ptrs),
                  "signal_pad_ptrs", len(hdl.signal_pad_ptrs), flush=True)
        for fn_name in ("one_shot_all_reduce", "two_shot_all_reduce_"):
            fn = getattr(torch.ops.symm_mem, fn_name)
            for _ in range(3):
                fn(t, "sum", dist.group.WORLD.group_name)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn(t, "sum", dist.group.WORLD.group_name)
            e1.record()
            torch.cuda.synchronize()
            if rank == 0:
                ms = e0.elapsed_time(e1) / 20
                print(f"torch symm_mem {fn_name} 16 MB: {ms * 1e3:.1f} us  algbw {16 * 1.048576 / ms:.1f} GB/s", flush=True)
    except Exception as e:
        if rank == 0:
            print("symmetric memory probe failed:", type(e).__name__, str(e)[:300], flush=True)
dist.destroy_process_group()
