"""Diagnose e2e variance: repeat the train_one_epoch loop and time raw pinned H2D copies between repeats."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from unite_b200.engine import Stage1Engine
from unite_b200.engine_for_pretraining import train_one_epoch
from unite_b200 import engine_for_pretraining as efp
from unite_b200.synthetic import SyntheticStage1Loader
from unite_b200.ddp import DataParallel

dev = torch.device("cuda", 0)
B = 32
student, teacher = bench.build_models(seed=0)
student, teacher = student.to(dev).train(), teacher.to(dev).eval()
model = DataParallel(student); model.grad_sync = None
eng = Stage1Engine(student, teacher, mask_ratio=0.8, use_graph=True)
efp.register_engine(student, teacher, eng)
loader = SyntheticStage1Loader(B, steps=20, seed=0, rank=0, n_distinct=2)

class A: log_freq = 1; use_cuda_graph = True

def h2d_bw():
    v = loader.batches[0][0]
    d = torch.empty_like(v, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        d.copy_(v, non_blocking=True)
    torch.cuda.synchronize()
    return 3 * v.numel() * 4 / (time.perf_counter() - t0) / 1e9

print("pinned:", loader.batches[0][0].is_pinned(), "H2D GB/s:", round(h2d_bw(), 1), flush=True)
class Timed:
    def __init__(self, inner): self.inner, self.t = inner, []
    def __len__(self): return len(self.inner)
    def __iter__(self):
        for b in self.inner:
            self.t.append(time.perf_counter())
            yield b

for rep in range(6):
    tl = Timed(loader)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    train_one_epoch(model, tl, None, eng.optimizer, dev, 0, None, teacher_model=teacher, mask_type="attention", mask_ratio=0.8, args=A)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / 20
    dts = [round((b - a) * 1e3, 1) for a, b in zip(tl.t[:-1], tl.t[1:])]
    print("   per-step CPU intervals (ms):", dts, flush=True)
    print(f"rep {rep}: {e0.elapsed_time(e1) / 20:.2f} ms/step (wall {wall:.2f})  H2D {h2d_bw():.1f} GB/s  mem reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB", flush=True)
