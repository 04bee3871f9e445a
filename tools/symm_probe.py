"""Probe torch symmetric memory / NVLS multicast on this box and time all-reduce variants for the 352 MB gradient arena.
torchrun --nproc-per-node N tools/symm_probe.py"""
import os, sys, time
import torch, torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", init_method="env://", world_size=world, rank=rank)
dev = torch.device("cuda", local)
n = 88_080_384  # ~ arena size (multiple of 1024)
g = torch.randn(n, device=dev)

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

t = timeit(lambda: dist.all_reduce(g))
if rank == 0: print(f"NCCL fp32 all_reduce {n*4/1e6:.0f} MB: {t:.3f} ms ({n*4/t/1e6:.0f} GB/s alg)", flush=True)
gb = g.bfloat16()
t = timeit(lambda: dist.all_reduce(gb))
if rank == 0: print(f"NCCL bf16 all_reduce: {t:.3f} ms", flush=True)
try:
    import torch.distributed._symmetric_memory as symm
    buf = symm.empty(n, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(buf, dist.group.WORLD.group_name)
    if rank == 0:
        print("symm_mem ok: world", hdl.world_size, "multicast_ptr", hex(hdl.multicast_ptr) if hasattr(hdl, "multicast_ptr") else None,
              "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:2], "signal pads", len(hdl.signal_pad_ptrs), flush=True)
    buf.copy_(g)
    for name in ("multimem_all_reduce_", "one_shot_all_reduce", "two_shot_all_reduce_"):
        try:
            op = getattr(torch.ops.symm_mem, name)
            if name == "one_shot_all_reduce":
                fn = lambda: op(buf, "sum", dist.group.WORLD.group_name)
            else:
                fn = lambda: op(buf, "sum", dist.group.WORLD.group_name)
            t = timeit(fn, 5)
            if rank == 0: print(f"symm_mem.{name}: {t:.3f} ms ({n*4/t/1e6:.0f} GB/s alg)", flush=True)
        except Exception as e:
            if rank == 0: print(f"symm_mem.{name}: FAILED {type(e).__name__}: {str(e)[:200]}", flush=True)
except Exception as e:
    if rank == 0: print("symm_mem unavailable:", type(e).__name__, str(e)[:300], flush=True)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
