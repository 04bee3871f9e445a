"""Times the operand-reading GEMM epilogues (fp32 residual, fp16 residual + row statistics, DGELU) at the step's shapes.
    python tools/gemm_epi_time.py            (UB_LIB_VARIANT=<suffix> for another build)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unite_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()                                  # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


def case(name, M, N, K, kind):
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = torch.randn(N, K, device=dev, generator=g).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    if kind == "res32":
        res = torch.randn(M, N, device=dev, generator=g); out = torch.empty(M, N, device=dev)
        fn = lambda: ops.gemm(a, w, out, bias=bias, residual=res)
    elif kind == "res16":
        a16, w16 = a.half(), w.half()
        res = torch.randn(M, N, device=dev, generator=g).half(); out = torch.empty(M, N, device=dev, dtype=torch.float16)
        st = torch.zeros(M, 2, device=dev)
        fn = lambda: ops.gemm(a16, w16, out, bias=bias, residual=res, stats_out=st)
    elif kind == "dgelu":
        wt = torch.randn(K, N, device=dev, generator=g).bfloat16()
        pre = torch.randn(M, N, device=dev, generator=g).bfloat16(); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fn = lambda: ops.gemm(a, wt, out, b_t=True, act=ops.UB_ACT_DGELU, aux_in=pre)
    else:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fn = lambda: ops.gemm(a, w, out, bias=bias)
    us = timeit(fn)
    print(f"{name:34s} {M}x{N}x{K} {kind:6s} {us:7.1f} us  {2.0 * M * N * K / us / 1e6:7.0f} TF/s", flush=True)


case("student proj fwd", 10240, 768, 768, "res32")
case("student fc2 fwd", 10240, 768, 3072, "res32")
case("teacher out_proj", 50432, 768, 768, "res16")
case("teacher c_proj", 50432, 768, 3072, "res16")
case("student dgrad fc2 (DGELU)", 10240, 3072, 768, "dgelu")
case("student qkv fwd (plain)", 10240, 2304, 768, "plain")
