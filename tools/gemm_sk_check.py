"""Stream-K tail of the GEMM: parity (split vs whole-tile schedule vs fp32 torch) over every epilogue the step uses, ragged shapes,
repeated launches (the counters must return to zero), and per-shape timing with a flushed L2.
    UB_LIB_VARIANT=sk python tools/gemm_sk_check.py [--time | --quick]        (run under gpurun; writes gpurun_out/gemm_sk_check.json)
The schedule is a build option (-DUB_GEMM_STREAMK -> libunite_b200_sk.so, built by __graft_entry__.build())."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unite_b200 import ops, _cabi  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
results = []


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def build(kind, M, N, K):
    """returns (fn(stream_k) -> out tensor, fp32 reference)"""
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = torch.randn(N, K, device=dev, generator=g).bfloat16() * (K ** -0.5)
    bias = torch.randn(N, device=dev, generator=g)
    acc = a.float() @ w.float().t()
    if kind == "plain":
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        return (lambda sk: ops.gemm(a, w, out, bias=bias, stream_k=sk)), acc + bias
    if kind == "nn":                       # dgrad: B stored [K, N]
        wt = w.t().contiguous()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        return (lambda sk: ops.gemm(a, wt, out, b_t=True, stream_k=sk)), acc
    if kind == "res32":
        res = torch.randn(M, N, device=dev, generator=g)
        scale = torch.rand(max(1, M // 320 + 1), device=dev, generator=g) + 0.5
        out = torch.empty(M, N, device=dev)
        ref = (acc + bias) * scale[torch.arange(M, device=dev) // 320, None] + res
        return (lambda sk: ops.gemm(a, w, out, bias=bias, residual=res, row_scale=scale, rows_per_scale=320, stream_k=sk)), ref
    if kind == "gelu_aux":
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ref = torch.nn.functional.gelu(acc + bias)
        return (lambda sk: (ops.gemm(a, w, out, bias=bias, act=ops.UB_ACT_GELU, aux_out=pre, stream_k=sk), pre)[0]), ref
    if kind == "dgelu":
        wt = w.t().contiguous()
        pre = torch.randn(M, N, device=dev, generator=g).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        x = pre.float()
        gp = 0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * torch.pi) ** 0.5
        return (lambda sk: ops.gemm(a, wt, out, b_t=True, act=ops.UB_ACT_DGELU, aux_in=pre, stream_k=sk)), acc * gp
    if kind == "res16":
        a16, w16 = a.half(), w.half()
        res = torch.randn(M, N, device=dev, generator=g).half()
        out = torch.empty(M, N, device=dev, dtype=torch.float16)
        st = torch.zeros(M, 2, device=dev)
        ref = a16.float() @ w16.float().t() + bias + res.float()
        return (lambda sk: (st.zero_(), ops.gemm(a16, w16, out, bias=bias, residual=res, stats_out=st, stream_k=sk))[1]), ref
    if kind == "acc32":                    # fp32 accumulate into a running output
        out = torch.empty(M, N, device=dev)
        base = torch.randn(M, N, device=dev, generator=g)
        return (lambda sk: (out.copy_(base), ops.gemm(a, w, out, accumulate=True, stream_k=sk))[1]), acc + base
    raise ValueError(kind)


def timeit(fn, iters=15):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


def case(kind, M, N, K, time_it=False, name=""):
    fn, ref = build(kind, M, N, K)
    whole = fn(False).clone()
    torch.cuda.synchronize()
    outs = []
    n0 = _cabi.lib.ub_gemm_sk_launches()
    for _ in range(3):                      # repeated launches: the arrival counters must come back to zero each time
        outs.append(fn(True).clone())
    torch.cuda.synchronize()
    split = _cabi.lib.ub_gemm_sk_launches() - n0 == 3
    e_whole, e_sk = rel(whole, ref), rel(outs[0], ref)
    same_runs = all(torch.equal(outs[0], o) for o in outs[1:])
    d = rel(outs[0], whole)
    tol = 2e-3 if ref.dtype == torch.float32 and kind in ("res32", "acc32") else 8e-3
    ok = e_sk < tol and e_whole < tol and same_runs and d < tol and torch.isfinite(outs[0].float()).all().item()
    row = dict(kind=kind, M=M, N=N, K=K, name=name, rel_whole=e_whole, rel_sk=e_sk, rel_sk_vs_whole=d, reruns_identical=same_runs, split=split, ok=ok)
    msg = f"{kind:9s} {M:6d}x{N:5d}x{K:5d}  whole {e_whole:.2e}  sk {e_sk:.2e}  sk-vs-whole {d:.2e}  reruns identical {same_runs}  split {split}  {'OK' if ok else 'FAIL'}"
    if time_it:
        t0, t1 = timeit(lambda: fn(False)), timeit(lambda: fn(True))
        t0b, t1b = timeit(lambda: fn(False)), timeit(lambda: fn(True))
        row.update(us_whole=min(t0, t0b), us_sk=min(t1, t1b))
        msg += f"   {min(t0, t0b):7.1f} us -> {min(t1, t1b):7.1f} us  ({2.0 * M * N * K / min(t1, t1b) / 1e6:5.0f} TF/s)  {name}"
    print(msg, flush=True)
    results.append(row)
    return ok


QUICK = (("res32", 10240, 768, 3072), ("nn", 10240, 768, 2304), ("gelu_aux", 10240, 3072, 768), ("dgelu", 10240, 3072, 768),
         ("res32", 10100, 776, 3000), ("res16", 10240, 768, 3072), ("acc32", 4096, 1024, 4096), ("nn", 20480, 1024, 1024),
         ("plain", 3200, 768, 1024))

if __name__ == "__main__":
    if not _cabi.lib.ub_gemm_sk_compiled():
        print("this build of the library has no stream-K tail (default build): run with UB_LIB_VARIANT=sk")
        sys.exit(2)
    if "--quick" in sys.argv:            # the pytest subset: every epilogue, ragged shapes, >= 8 of 9 cases must really split
        ok = all([case(*c) for c in QUICK])
        n_split = sum(r["split"] for r in results)
        print(f"{n_split} of {len(results)} cases took the split")
        print("ALL OK" if ok and n_split >= 8 else "FAILURES", flush=True)
        sys.exit(0 if ok and n_split >= 8 else 1)
    timing = "--time" in sys.argv
    ok = True
    # the step's shapes (B = 32: M = 10 240), every epilogue that can meet a split tail
    ok &= case("res32", 10240, 768, 3072, timing, "student fc2 fwd")
    ok &= case("nn", 10240, 768, 3072, timing, "student fc1 dgrad")
    ok &= case("nn", 10240, 768, 2304, timing, "student qkv dgrad")
    ok &= case("gelu_aux", 10240, 3072, 768, timing, "student fc1 fwd")
    ok &= case("dgelu", 10240, 3072, 768, timing, "student fc2 dgrad")
    ok &= case("res32", 10240, 768, 768, timing, "student proj fwd")
    ok &= case("nn", 10240, 768, 768, timing, "student proj dgrad")
    ok &= case("plain", 10240, 2304, 768, timing, "student qkv fwd")
    # tails of other sizes: fewer tiles than pairs (every tile halved), ragged M / N / K, three-piece tiles, fp16 stream, accumulate
    ok &= case("plain", 2560, 768, 1024)
    ok &= case("res32", 10100, 776, 3000)
    ok &= case("res16", 10240, 768, 3072)
    ok &= case("acc32", 4096, 1024, 4096)
    ok &= case("plain", 300, 264, 4096)
    ok &= case("nn", 20480, 1024, 1024)
    ok &= case("plain", 19200, 768, 512)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gemm_sk_check.json"), "w") as f:
        json.dump(dict(device=torch.cuda.get_device_name(0), env={k: v for k, v in os.environ.items() if k.startswith("UB_")}, cases=results, ok=bool(ok)), f, indent=1)
    print("ALL OK" if ok else "FAILURES", flush=True)
    sys.exit(0 if ok else 1)
