"""Timing of the HBM-bound kernels at the stage-1 B=32 shapes (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unite_b200 import ops
dev = "cuda"

def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def report(name, us, bytes_):
    print(f"{name:40s} {us:8.1f} us   {bytes_/us/1e3:7.0f} GB/s", flush=True)

R, D = 50432, 768
x = torch.randn(R, D, device=dev); g = torch.ones(D, device=dev); b = torch.zeros(D, device=dev)
h = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
report("ln_fwd teacher [50432x768] f32->bf16", timeit(lambda: ops.layernorm_fwd(x, g, b, 1e-5, h)), R * D * 6)
M = 10240
xs = torch.randn(M, D, device=dev); hs = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
report("ln_fwd student [10240x768]", timeit(lambda: ops.layernorm_fwd(xs, g, b, 1e-6, hs)), M * D * 6)
dy = torch.randn(M, D, device=dev).bfloat16(); dx = torch.randn(M, D, device=dev); dxs = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev); ds = torch.zeros(D, device=dev)
report("ln_bwd student [10240x768] (+dsum)", timeit(lambda: ops.layernorm_bwd(dy, xs, g, 1e-6, dx, dx, dxs, None, 0, dg, db, ds)), M * D * (2 + 4 + 4 + 4 + 2))
dp = torch.randn(M, 3072, device=dev).bfloat16(); cs = torch.zeros(3072, device=dev)
report("colsum [10240x3072] bf16", timeit(lambda: ops.colsum_bf16(dp, cs)), M * 3072 * 2)
dq = torch.randn(M, 2304, device=dev).bfloat16(); cq = torch.zeros(768, device=dev)
report("colsum dqkv[:, :768] (strided)", timeit(lambda: ops.colsum_bf16(dq[:, :768], cq)), M * 768 * 2)
cqa = torch.zeros(2304, device=dev)
report("colsum dqkv all 2304", timeit(lambda: ops.colsum_bf16(dq, cqa)), M * 2304 * 2)
v = torch.randn(32, 3, 8, 224, 224, device=dev); pt = torch.empty(32 * 1568, 768, device=dev, dtype=torch.bfloat16)
report("patchify B=32", timeit(lambda: ops.patchify(v, pt, 1)), v.numel() * 4 + pt.numel() * 2)
y = torch.randn(M, 512, device=dev); g5 = torch.ones(512, device=dev); b5 = torch.zeros(512, device=dev); o5 = torch.empty(M, 512, device=dev)
tg = torch.nn.functional.normalize(torch.randn(M, 512, device=dev), dim=-1); la = torch.zeros(1, device=dev)
report("dec_tail_fwd [10240x512] +loss", timeit(lambda: ops.dec_tail_fwd(y, g5, b5, 1e-6, o5, tg, la, 1.0)), M * 512 * 12)
dy5 = torch.empty(M, 512, device=dev, dtype=torch.bfloat16); dg5 = torch.zeros(512, device=dev); db5 = torch.zeros(512, device=dev)
report("dec_tail_bwd [10240x512]", timeit(lambda: ops.dec_tail_bwd(y, g5, b5, 1e-6, tg, -1e-4, dy5, dg5, db5)), M * 512 * 10)
