"""Cost of the LayerNorm-fold pieces on the teacher shapes (python tools/gemm_fold_time.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unite_b200 import ops
dev = "cuda"
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
M = 50432
g = torch.Generator(device=dev).manual_seed(0)
x16 = torch.randn(M, 768, device=dev, generator=g).half()
h = torch.randn(M, 768, device=dev, generator=g).bfloat16()
o = torch.randn(M, 768, device=dev, generator=g).bfloat16()
u = torch.randn(M, 3072, device=dev, generator=g).bfloat16()
stats = torch.zeros(M, 2, device=dev)
for N, act in ((2304, 0), (3072, 1)):
    Wb = (torch.randn(N, 768, device=dev, generator=g) * 0.03).bfloat16()
    Wh = Wb.half()
    bias = torch.randn(N, device=dev, generator=g)
    c = Wh.float().sum(1).contiguous()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    stats.copy_(torch.stack([x16.float().sum(1), (x16.float() ** 2).sum(1)], 1))
    t0 = timeit(lambda: ops.gemm(h, Wb, out, bias=bias, act=act))
    t1 = timeit(lambda: ops.gemm(x16, Wh, out, bias=bias, act=act))
    t2 = timeit(lambda: ops.gemm(x16, Wh, out, bias=bias, act=act, ln_stats=stats, ln_c=c, ln_eps=1e-5))
    print(f"N={N} act={act}: bf16 plain {t0:.1f} us | fp16 operands {t1:.1f} | fp16 + LN fold {t2:.1f}", flush=True)
W1 = (torch.randn(768, 768, device=dev, generator=g) * 0.03).bfloat16()
W2 = (torch.randn(768, 3072, device=dev, generator=g) * 0.03).bfloat16()
b = torch.randn(768, device=dev, generator=g)
y = torch.empty(M, 768, device=dev, dtype=torch.float16)
for name, A, W in (("out_proj", o, W1), ("c_proj", u, W2)):
    t0 = timeit(lambda: ops.gemm(A, W, y, bias=b, residual=x16))
    t1 = timeit(lambda: ops.gemm(A, W, y, bias=b, residual=x16, stats_out=stats))
    print(f"{name}: fp16 residual {t0:.1f} us | + row stats {t1:.1f}", flush=True)
