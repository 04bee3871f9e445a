"""GPU check + timing of ub_gemm_bf16 against torch.matmul (run under gpurun)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unite_b200 import _cabi as cabi


def run(M, N, K, a_mn=0, b_mn=0, out_fp32=0, bias=False, act=0, resid=False, split_k=1, accumulate=0, time_it=False, tile_ctas=0):
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device=dev, generator=g).bfloat16()
    B = torch.randn(N, K, device=dev, generator=g).bfloat16()
    A_st = A.t().contiguous() if a_mn else A
    B_st = B.t().contiguous() if b_mn else B
    ref = A.float() @ B.float().t()
    ep = cabi.GemmEpilogue()
    keep = []
    if bias:
        bv = torch.randn(N, device=dev, generator=g)
        ep.bias = bv.data_ptr(); ref = ref + bv; keep.append(bv)
    if act == 1:
        ref = ref * torch.sigmoid(1.702 * ref)
    elif act == 2:
        ref = torch.nn.functional.gelu(ref)
    if resid:
        rv = torch.randn(M, N, device=dev, generator=g)
        ep.residual = rv.data_ptr(); ep.ldr = N; ref = ref + rv; keep.append(rv)
    ep.act = act; ep.out_fp32 = out_fp32; ep.accumulate = accumulate; ep.tile_ctas = tile_ctas
    Cout = torch.zeros(M, N, device=dev, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    def call():
        rc = cabi.lib.ub_gemm_bf16(A_st.data_ptr(), A_st.stride(0), a_mn, B_st.data_ptr(), B_st.stride(0), b_mn,
                                   Cout.data_ptr(), N, M, N, K, C.byref(ep), split_k, st)
        cabi.check(rc, "gemm")
    call()
    torch.cuda.synchronize()
    err = (Cout.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    rel = ((Cout.float() - ref).norm() / ref.norm()).item()
    ok = rel < (2e-3 if out_fp32 else 6e-3)
    msg = f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} fp32={out_fp32} bias={bias} act={act} res={resid} split={split_k}: max_err={err:.4g} (scale {scale:.3g}) rel={rel:.3g} {'OK' if ok else 'FAIL'}"
    if time_it:
        for _ in range(3): call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        msg += f"  {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s"
        Cr = torch.empty_like(Cout, dtype=torch.bfloat16)
        for _ in range(3): torch.matmul(A, B.t(), out=Cr)
        e0.record()
        for _ in range(n): torch.matmul(A, B.t(), out=Cr)
        e1.record(); torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / n
        msg += f"  | cublas {ms2*1e3:.1f} us {2*M*N*K/ms2/1e9:.0f} TFLOP/s"
    print(msg, flush=True)
    return ok


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "SMs", cabi.lib.ub_sm_count(), "cluster-of-4 capacity", cabi.lib.ub_gemm_cluster4_capacity())
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        M, N, K = map(int, sys.argv[2:5])
        kw = dict(a.split("=") for a in sys.argv[5:])
        kw = {k: (v == "True") if v in ("True", "False") else int(v) for k, v in kw.items()}
        sys.exit(0 if run(M, N, K, **kw) else 1)
    ok = True
    ok &= run(128, 128, 64)
    ok &= run(128, 256, 64)
    ok &= run(128, 256, 256)
    ok &= run(256, 512, 768, out_fp32=1)
    ok &= run(200, 264, 200, out_fp32=1)           # ragged everything
    ok &= run(10240, 768, 768, bias=True, resid=True, out_fp32=1, time_it=True)
    ok &= run(10240, 2304, 768, bias=True, time_it=True)
    ok &= run(10240, 3072, 768, bias=True, act=2, time_it=True)
    ok &= run(10240, 768, 3072, bias=True, resid=True, out_fp32=1, time_it=True)
    ok &= run(50432, 2304, 768, bias=True, time_it=True)
    ok &= run(50432, 3072, 768, bias=True, act=1, time_it=True)
    ok &= run(50432, 768, 3072, bias=True, resid=True, out_fp32=1, time_it=True)
    ok &= run(8192, 8192, 8192, time_it=True)
    # dgrad form: B MN-major
    ok &= run(128, 128, 64, b_mn=1)
    ok &= run(256, 256, 128, b_mn=1)
    ok &= run(10240, 768, 3072, b_mn=1, time_it=True)
    # wgrad form: both MN-major, split-K with atomics
    ok &= run(128, 128, 64, a_mn=1, b_mn=1, out_fp32=1)
    ok &= run(256, 256, 256, a_mn=1, b_mn=1, out_fp32=1)
    ok &= run(768, 768, 10240, a_mn=1, b_mn=1, out_fp32=1, split_k=4, accumulate=1)
    ok &= run(3072, 768, 10240, a_mn=1, b_mn=1, out_fp32=1, time_it=True)
    print("ALL OK" if ok else "SOME FAILED")
    sys.exit(0 if ok else 1)


def run_lnfold(M, N, K, act=0):
    """LayerNorm folded into the GEMM (fp16 operands) + fp16-residual epilogue producing the row statistics, vs torch."""
    import torch.nn.functional as F
    from unite_b200 import ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(M + N + K)
    x0 = (torch.randn(M, K, device=dev, generator=g) * 2 + 0.3)
    r0 = torch.randn(M, K, device=dev, generator=g).half()
    W0 = (torch.randn(K, K, device=dev, generator=g) * 0.03).bfloat16()
    # producer: x = r0 + x0 @ W0^T in fp16 with row statistics
    stats = torch.zeros(M, 2, device=dev)
    x = torch.empty(M, K, device=dev, dtype=torch.float16)
    ops.gemm(x0.bfloat16(), W0, x, residual=r0, stats_out=stats)
    xf = x.float()
    ok = True
    e1 = (stats[:, 0] - xf.sum(1)).abs().max().item() / xf.sum(1).abs().max().item()
    e2 = (stats[:, 1] - (xf * xf).sum(1)).abs().max().item() / (xf * xf).sum(1).abs().max().item()
    ok &= e1 < 1e-5 and e2 < 1e-5
    gamma = torch.rand(K, device=dev, generator=g) + 0.5
    beta = torch.randn(K, device=dev, generator=g) * 0.1
    W = torch.randn(N, K, device=dev, generator=g) * 0.05
    b = torch.randn(N, device=dev, generator=g) * 0.1
    ref = F.layer_norm(xf, (K,), gamma, beta, 1e-5) @ W.t() + b
    if act == 1:
        ref = ref * torch.sigmoid(1.702 * ref)
    wg = (W * gamma[None, :]).half().contiguous()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, wg, out, bias=(W @ beta + b).contiguous(), act=act, ln_stats=stats, ln_c=wg.float().sum(1).contiguous(), ln_eps=1e-5)
    torch.cuda.synchronize()
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    ok &= rel < 4e-3
    print(f"LN-fold GEMM M={M} N={N} K={K} act={act}: stats err {e1:.1e}/{e2:.1e}, out rel {rel:.3e} {'OK' if ok else 'FAIL'}", flush=True)
    return ok
