"""Parity + timing of the long-sequence tcgen05 attention (csrc/attention_long_tc.cu) against an fp32 torch reference on the GPU.

    python tools/attn_long_check.py [--time]        exit code 1 on mismatch
Shapes: ragged and exact multiples of the 128-row tiles / 64- and 128-row chunks, an odd tile count (the duplicated last tile of
warpgroup 1), more units than SMs, a score range that forces the forward's reference-moving slow path, and the stage-2 shape.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unite_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
FAIL = []


def report(name, got, ref, tol):
    got, ref = got.float(), ref.float()
    err = ((got - ref).norm() / (ref.norm() + 1e-30)).item()
    bad = not (err <= tol) or not bool(torch.isfinite(got).all())
    print(f"{'FAIL' if bad else 'ok  '} {name:44s} rel-L2 {err:.3e} (tol {tol:.1e}) max-abs {float((got - ref).abs().max()):.3e}")
    if bad:
        FAIL.append(name)


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def check(n_seq, S, H, amp=0.7, spike=False, time_it=False, no_lse=False):
    g = torch.Generator(device=dev).manual_seed(S * 31 + H)
    qkv = torch.randn(n_seq * S, 3 * H * 64, device=dev, generator=g) * amp
    if spike:
        # keys late in the sequence score ~2^60 above the first chunk's maximum for some rows: the slow path must move the reference
        v3 = qkv.view(n_seq, S, 3, H, 64)
        v3[:, S // 2 + 5, 1] *= 0.0
        v3[:, S // 2 + 5, 1, :, :8] = 24.0
        v3[:, ::3, 0, :, :8] = 6.0
    qkv = qkv.bfloat16()
    o = torch.full((n_seq * S, H * 64), float("nan"), device=dev, dtype=torch.bfloat16)
    lse = torch.full((n_seq, H, S), float("nan"), device=dev)
    scale = 0.125
    ops.attn_fwd(qkv, o, None if no_lse else lse, n_seq, S, H, scale)
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(n_seq, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    q = q.detach().requires_grad_(); k = k.detach().requires_grad_(); v = v.detach().requires_grad_()
    s = (q @ k.transpose(-1, -2)) * scale
    p = s.softmax(-1)
    oref = p @ v
    tag = f"n={n_seq} S={S} H={H}{' spike' if spike else ''}{' no-lse' if no_lse else ''}"
    report(f"fwd o   {tag}", o.view(n_seq, S, H, 64).permute(0, 2, 1, 3), oref, 6e-3)
    if no_lse:
        return
    report(f"fwd lse {tag}", lse, torch.logsumexp(s, -1), 5e-4)
    d_o = torch.randn(n_seq * S, H * 64, device=dev, generator=g).bfloat16()
    oref.backward(d_o.float().view(n_seq, S, H, 64).permute(0, 2, 1, 3))
    dqkv = torch.full_like(qkv, float("nan"))
    dws = torch.empty(n_seq, H, S, device=dev)
    dbias = torch.zeros(3 * H * 64, device=dev)
    ops.attn_bwd(qkv, o, d_o, lse, dws, dqkv, n_seq, S, H, scale, dbias=dbias)
    torch.cuda.synchronize()
    dq, dk, dv = dqkv.float().view(n_seq, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    report(f"bwd dq  {tag}", dq, q.grad, 1.2e-2)
    report(f"bwd dk  {tag}", dk, k.grad, 1.2e-2)
    report(f"bwd dv  {tag}", dv, v.grad, 1.2e-2)
    want = dqkv.float().sum(0)                      # bias gradients = column sums of the bf16 dqkv as stored; key third untouched
    want[H * 64:2 * H * 64] = 0
    report(f"bwd dbias {tag}", dbias, want, 1e-4)
    if time_it:
        fl = 4.0 * S * S * 64 * H * n_seq
        t = timeit(lambda: ops.attn_fwd(qkv, o, lse, n_seq, S, H, scale))
        t2 = timeit(lambda: ops.attn_bwd(qkv, o, d_o, lse, dws, dqkv, n_seq, S, H, scale))
        print(f"   fwd {t*1e3:8.0f} us ({fl/t/1e9:6.0f} TF/s)   bwd {t2*1e3:8.0f} us ({2.5*fl/t2/1e9:6.0f} TF/s of the 10 S^2 d algorithmic count)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--small", action="store_true", help="the two smallest shapes only (compute-sanitizer runs)")
    args = ap.parse_args()
    if args.small:
        check(1, 333, 2)
        print("ATTN LONG " + ("MISMATCH: " + ", ".join(FAIL) if FAIL else "PARITY OK"))
        sys.exit(1 if FAIL else 0)
    check(2, 384, 2)                       # 3 tiles (odd), exact chunks
    check(1, 333, 3)                       # ragged: 3 tiles, partial last chunk of both sizes
    check(2, 1000, 4)                      # 8 tiles, ragged
    check(3, 1568, 12)                     # stage-2 sequence length: 13 tiles (odd), 252 units > 148 SMs
    check(2, 512, 2, spike=True)           # the forward's slow path
    check(2, 400, 2, no_lse=True)          # inference form (no LSE output)
    if args.time:
        check(32, 1568, 12, time_it=True)  # BASELINE configs[0] / [3] shape: stage-2 / stage-3 all-token attention, B = 32
    print("ATTN LONG " + ("MISMATCH: " + ", ".join(FAIL) if FAIL else "PARITY OK"))
    sys.exit(1 if FAIL else 0)


if __name__ == "__main__":
    main()
