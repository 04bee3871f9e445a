"""Small launch sets for `ncu --set full` captures of the kernels the round-1 verdict asked evidence for.

    python tools/ncu_targets.py gemm      # student shapes 10240x768x3072 (fc2 fwd: fp32 out + residual) and 10240x2304x768 (qkv fwd)
    python tools/ncu_targets.py teacher_attn   # the teacher's 197-token forward, 256 frames x 12 heads
    python tools/ncu_targets.py attn      # long-sequence attention fwd + two-pass bwd, n_seq=4 S=1568 H=12
    python tools/ncu_targets.py membound  # patchify, LN fwd/bwd, dec_tail, gather, AdamW at the step's shapes
Each mode warms up once, then runs the launches of interest between cudaProfilerStart/Stop (use --profile-from-start off).
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unite_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
mode = sys.argv[1] if len(sys.argv) > 1 else "gemm"
g = torch.Generator(device=dev).manual_seed(0)


def profiled(fn):
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if mode == "gemm":
    M = 10240
    a1 = torch.randn(M, 3072, device=dev, generator=g).bfloat16(); w1 = torch.randn(768, 3072, device=dev, generator=g).bfloat16()
    res = torch.randn(M, 768, device=dev, generator=g); o1 = torch.empty(M, 768, device=dev); b1 = torch.randn(768, device=dev, generator=g)
    a2 = torch.randn(M, 768, device=dev, generator=g).bfloat16(); w2 = torch.randn(2304, 768, device=dev, generator=g).bfloat16()
    o2 = torch.empty(M, 2304, device=dev, dtype=torch.bfloat16); b2 = torch.randn(2304, device=dev, generator=g)

    def run():
        ops.gemm(a1, w1, o1, bias=b1, residual=res)        # fc2 forward: 10240 x 768 x 3072
        ops.gemm(a2, w2, o2, bias=b2)                      # qkv forward: 10240 x 2304 x 768
    profiled(run)
elif mode == "teacher_attn":
    n_seq, S, H = 256, 197, 12                              # one teacher layer of the B = 32 step: 3072 (frame, head) items
    qkv = (torch.randn(n_seq * S, 3 * H * 64, device=dev, generator=g) * 0.7).bfloat16()
    o = torch.empty(n_seq * S, H * 64, device=dev, dtype=torch.bfloat16)
    profiled(lambda: ops.attn_fwd(qkv, o, None, n_seq, S, H, 0.125))
elif mode == "attn":
    n_seq, S, H = 4, 1568, 12
    qkv = (torch.randn(n_seq * S, 3 * H * 64, device=dev, generator=g) * 0.7).bfloat16()
    o = torch.empty(n_seq * S, H * 64, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(n_seq, H, S, device=dev)
    d_o = torch.randn(n_seq * S, H * 64, device=dev, generator=g).bfloat16()
    dqkv = torch.empty_like(qkv); dws = torch.empty(n_seq, H, S, device=dev)

    def run():
        ops.attn_fwd(qkv, o, lse, n_seq, S, H, 0.125)
        ops.attn_bwd(qkv, o, d_o, lse, dws, dqkv, n_seq, S, H, 0.125)
    profiled(run)
else:
    B, M, D = 32, 10240, 768
    clip = torch.randn(B, 3, 8, 224, 224, device=dev, generator=g)
    patches = torch.empty(B * 1568, 768, device=dev, dtype=torch.bfloat16)
    x = torch.randn(M, D, device=dev, generator=g); gam = torch.ones(D, device=dev); bet = torch.zeros(D, device=dev)
    h = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    xt = torch.randn(B * 8 * 197, D, device=dev, generator=g).half(); ht = torch.empty(B * 8 * 197, D, device=dev, dtype=torch.bfloat16)
    dy = torch.randn(M, D, device=dev, generator=g).bfloat16(); dx = torch.randn(M, D, device=dev, generator=g)
    dxs = torch.empty(M, D, device=dev, dtype=torch.bfloat16); dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev); ds = torch.zeros(D, device=dev)
    y = torch.randn(M, 512, device=dev, generator=g); tgt = torch.nn.functional.normalize(torch.randn(M, 512, device=dev, generator=g), dim=-1)
    out = torch.empty(M, 512, device=dev); g5 = torch.ones(512, device=dev); b5 = torch.zeros(512, device=dev); lacc = torch.zeros(1, device=dev)
    dy5 = torch.empty(M, 512, device=dev, dtype=torch.bfloat16); dg5 = torch.zeros(512, device=dev); db5 = torch.zeros(512, device=dev)
    idx = torch.randint(0, 1568, (M,), device=dev, generator=g).int(); pv = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
    n = 88_015_104
    p = torch.randn(n, device=dev, generator=g) * 0.02; gr = torch.randn(n, device=dev, generator=g) * 1e-3
    m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev); w16 = torch.empty(n, device=dev, dtype=torch.bfloat16)
    d4 = torch.randn(M, 3072, device=dev, generator=g).bfloat16(); cs = torch.zeros(3072, device=dev)

    def run():
        ops.patchify(clip, patches, 1)
        ops.layernorm_fwd(x, gam, bet, 1e-6, h)                                   # student LN: fp32 rows -> bf16
        ops.layernorm_fwd(xt, gam, bet, 1e-5, ht)                                 # teacher-sized LN: fp16 rows -> bf16
        ops.layernorm_bwd(dy, x, gam, 1e-6, dx, dx, dxs, None, 320, dg, db, dsum=ds)
        ops.dec_tail_fwd(y, g5, b5, 1e-6, out, tgt, lacc, 1.0 / M)
        ops.dec_tail_bwd(y, g5, b5, 1e-6, tgt, -2.0 / M, dy5, dg5, db5)
        ops.gather_rows(patches, idx, pv, rows_per_group=320, group_stride_rows=1568)
        ops.colsum_bf16(d4, cs)
        ops.adamw(p, gr, m, v, w16, n - 131072, 1e-3, 0.05, 0.9, 0.95, 1e-8, 1)
    profiled(run)
print("done", mode)
