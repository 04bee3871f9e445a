"""Experiment: one GEMM split by rows into a cluster-of-4 launch (33 clusters = 132 SMs) and a CTA-pair launch limited to the 16
SMs no cluster of 4 can use, on two streams.  python tools/gemm_hybrid.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unite_b200 import ops

dev = "cuda"
side = torch.cuda.Stream()


def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def run(M, N, K, out_fp32, resid):
    g = torch.Generator(device=dev).manual_seed(1)
    A = torch.randn(M, K, device=dev, generator=g).bfloat16()
    W = torch.randn(N, K, device=dev, generator=g).bfloat16() * 0.05
    bias = torch.randn(N, device=dev, generator=g)
    R = torch.randn(M, N, device=dev, generator=g) if resid else None
    C0 = torch.empty(M, N, device=dev, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    C1 = torch.empty_like(C0)
    kw = dict(bias=bias)

    def plain():
        ops.gemm(A, W, C0, residual=R, **kw)

    def hybrid(m1, max_ctas):
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)
        ops.gemm(A[:m1], W, C1[:m1], residual=None if R is None else R[:m1], tile_ctas=4, **kw)
        with torch.cuda.stream(side):
            ops.gemm(A[m1:], W, C1[m1:], residual=None if R is None else R[m1:], tile_ctas=2, max_ctas=max_ctas, **kw)
        main.wait_stream(side)

    t0 = timeit(plain)
    print(f"M={M} N={N} K={K} fp32={out_fp32} res={resid}: pairs {t0:.1f} us ({2*M*N*K/t0/1e6:.0f} TF/s)", flush=True)
    t4 = timeit(lambda: ops.gemm(A, W, C1, residual=R, tile_ctas=4, **kw))
    print(f"    cluster-4 only: {t4:.1f} us")
    for frac in (0.88, 0.90, 0.92, 0.94):
        m1 = int(M * frac) // 512 * 512
        for mc in (16, 12):
            t = timeit(lambda: hybrid(m1, mc))
            print(f"    hybrid m1={m1} ({m1/M:.3f}) side max_ctas={mc}: {t:.1f} us ({2*M*N*K/t/1e6:.0f} TF/s)  x{t0/t:.3f}", flush=True)
    hybrid(int(M * 0.9) // 512 * 512, 16)
    plain()
    torch.cuda.synchronize()
    print("    max |hybrid - plain| =", (C1.float() - C0.float()).abs().max().item())


run(50432, 3072, 768, 0, False)
run(50432, 768, 3072, 1, True)
run(50432, 2304, 768, 0, False)
