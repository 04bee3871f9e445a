"""Does an HBM-bound pass (AdamW over 88 M parameters) hide beside the teacher's GEMMs?  Times N GEMMs alone, AdamW alone, and
both on two streams.   python tools/overlap_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unite_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
M, N, K = 50432, 3072, 768
a = torch.randn(M, K, device=dev, generator=g).bfloat16()
w = torch.randn(N, K, device=dev, generator=g).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
bias = torch.randn(N, device=dev, generator=g)
n = 88_015_104 // 4 * 4
p = torch.randn(n, device=dev, generator=g); gr = torch.randn(n, device=dev, generator=g) * 0.01
m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev); w16 = torch.empty(n, device=dev, dtype=torch.bfloat16)
seg_end4 = torch.tensor([n // 4], dtype=torch.int32, device=dev)
hyper = torch.tensor([1e-4, 0.05, 0.9, 0.95, 1e-8, 0.1, 0.2236, 1.0, 1e-4, 0.05], device=dev)
gn = torch.zeros(1, device=dev)
side = torch.cuda.Stream()
NG = 6


def gemms():
    for _ in range(NG):
        ops.gemm(a, w, out, bias=bias, act=ops.UB_ACT_QUICKGELU)


def adamw():
    ops.adamw_seg(p, gr, m, v, w16, seg_end4, hyper, gn)


def both():
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        adamw()
    gemms()
    cur.wait_stream(side)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


tg, ta, tb = timeit(gemms), timeit(adamw), timeit(both)
print(f"{NG} teacher c_fc GEMMs alone {tg:.3f} ms | AdamW (88 M) alone {ta:.3f} ms | sum {tg + ta:.3f} ms | concurrent {tb:.3f} ms "
      f"-> {100 * (tg + ta - tb) / ta:.0f} % of the AdamW time hidden")
