"""GPU unit checks of the non-GEMM kernels against torch (run under gpurun)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from unite_b200 import ops

dev = "cuda"
OK = True


def report(name, got, ref, tol):
    global OK
    rel = ((got.float() - ref.float()).norm() / (ref.float().norm() + 1e-30)).item()
    mx = (got.float() - ref.float()).abs().max().item()
    ok = rel < tol and math.isfinite(rel)
    OK &= ok
    print(f"{name:34s} rel={rel:.3e} max={mx:.3e} {'OK' if ok else 'FAIL'}", flush=True)


def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def check_attention(n_seq, S, H, time_it=False):
    g = torch.Generator(device=dev).manual_seed(S * 31 + H)
    qkv = (torch.randn(n_seq * S, 3 * H * 64, device=dev, generator=g) * 0.7).bfloat16()
    o = torch.empty(n_seq * S, H * 64, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(n_seq, H, S, device=dev)
    scale = 0.125
    ops.attn_fwd(qkv, o, lse, n_seq, S, H, scale)
    q, k, v = qkv.float().view(n_seq, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    q = q.detach().requires_grad_(); k = k.detach().requires_grad_(); v = v.detach().requires_grad_()
    s = (q @ k.transpose(-1, -2)) * scale
    p = s.softmax(-1)
    oref = (p @ v)
    report(f"attn_fwd o   S={S} H={H}", o.view(n_seq, S, H, 64).permute(0, 2, 1, 3), oref, 6e-3)
    # the tcgen05 forward sums the bf16 P it multiplies with V (row sum on the tensor core): |d lse| <= 2^-9, so O stays
    # exactly normalised w.r.t. the P the backward recomputes from this lse
    report(f"attn_fwd lse S={S} H={H}", lse, torch.logsumexp(s, -1), 5e-4)
    d_o = (torch.randn(n_seq * S, H * 64, device=dev, generator=g)).bfloat16()
    oref.backward(d_o.float().view(n_seq, S, H, 64).permute(0, 2, 1, 3))
    dqkv = torch.zeros_like(qkv)
    dws = torch.empty(n_seq, H, S, device=dev)
    dbias = torch.zeros(3 * H * 64, device=dev)
    ops.attn_bwd(qkv, o, d_o, lse, dws, dqkv, n_seq, S, H, scale, dbias=dbias)
    dq, dk, dv = dqkv.float().view(n_seq, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    report(f"attn_bwd dq  S={S} H={H}", dq, q.grad, 1.2e-2)
    report(f"attn_bwd dk  S={S} H={H}", dk, k.grad, 1.2e-2)
    report(f"attn_bwd dv  S={S} H={H}", dv, v.grad, 1.2e-2)
    want = dqkv.float().sum(0)
    want[H * 64:2 * H * 64] = 0
    report(f"attn_bwd dbias S={S} H={H}", dbias, want, 1e-4)
    # cls attention
    att = torch.empty(n_seq, S - 1, device=dev)
    ops.cls_attn(qkv, att, n_seq, S, H, scale)
    report(f"cls_attn     S={S} H={H}", att, p.mean(1)[:, 0, 1:], 3e-3)
    if S <= 240:
        o2 = torch.full_like(o, float("nan"))
        ops.attn_fwd(qkv, o2, None, n_seq, S, H, scale)       # tcgen05 / TMEM inference kernel (no LSE)
        torch.cuda.synchronize()
        report(f"attn_fwd_tc  S={S} H={H}", o2.view(n_seq, S, H, 64).permute(0, 2, 1, 3), oref.detach(), 6e-3)
        if time_it:
            fl = 4 * S * S * 64 * H * n_seq
            t = timeit(lambda: ops.attn_fwd(qkv, o2, None, n_seq, S, H, scale))
            print(f"   attn fwd (tcgen05) {t*1e3:.0f} us ({fl/t/1e9:.0f} TF/s)")
    if time_it:
        fl = 4 * S * S * 64 * H * n_seq
        t = timeit(lambda: ops.attn_fwd(qkv, o, lse, n_seq, S, H, scale))
        t2 = timeit(lambda: ops.attn_bwd(qkv, o, d_o, lse, dws, dqkv, n_seq, S, H, scale))
        print(f"   attn fwd {t*1e3:.0f} us ({fl/t/1e9:.0f} TF/s)   bwd {t2*1e3:.0f} us ({2.5*fl/t2/1e9:.0f} TF/s alg)")


def check_attention_tc_late_maximum(n_seq=6, S=197, H=3):
    """The register-resident teacher kernel (no LSE, S = 197) takes the maximum of the FIRST 96 keys as its softmax reference and
    only redoes that half when a later key exceeds it by more than 2^60: rows whose maximum sits in the second half by a little, by
    a lot (redo path), and rows dominated by the first half."""
    g = torch.Generator(device=dev).manual_seed(1234)
    qkv = (torch.randn(n_seq, S, 3, H, 64, device=dev, generator=g) * 0.7)
    qkv[0, :, 1, :, :] *= 6.0                       # large scores everywhere
    qkv[1, 120:, 1, :, :] *= 90.0                   # second-half keys far above the first half: the redo path (diff >> 60 / (scale log2 e))
    qkv[2, :96, 1, :, :] *= 90.0                    # first-half keys dominate
    qkv[3, 150, 1, :, :] *= 400.0                   # one huge late key
    qkv[4, 196, 1, :, :] *= 400.0                   # ... the very last valid key (masked tail block)
    qkv = qkv.bfloat16().view(n_seq * S, 3 * H * 64).contiguous()
    o = torch.full((n_seq * S, H * 64), float("nan"), device=dev, dtype=torch.bfloat16)
    scale = 0.125
    ops.attn_fwd(qkv, o, None, n_seq, S, H, scale)
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(n_seq, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * scale
    oref = s.softmax(-1) @ v
    late = ((s[..., 96:].max(-1).values - s[..., :96].max(-1).values) * 1.4427 > 60).float().mean().item()
    print(f"   rows on the redo path: {100 * late:.1f} %")
    global OK
    OK &= late > 0.05
    got = o.view(n_seq, S, H, 64).permute(0, 2, 1, 3)
    OK &= bool(torch.isfinite(got.float()).all())
    report(f"attn_fwd_tc late maximum S={S} H={H}", got, oref, 8e-3)


def check_gemm_dot_aux(B=5, N=320, H=12, time_it=False):
    """ub_gemm_epilogue.dot_out: the proj dgrad GEMM that writes dO also leaves D = rowsum(dO o O) per (token, head) in ub_attn_bwd's
    [n_seq, H, S] layout; dO must be bit-identical to the plain GEMM's, D within fp32 summation order of the torch value."""
    global OK
    g = torch.Generator(device=dev).manual_seed(B * 7 + N)
    M, D = B * N, H * 64
    dy = (torch.randn(M, D, device=dev, generator=g) * 0.3).bfloat16()
    w = (torch.randn(D, D, device=dev, generator=g) * D ** -0.5).bfloat16()          # [out, in] as stored; dgrad contracts over `out`
    o = torch.randn(M, D, device=dev, generator=g).bfloat16()
    d_o_plain = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    ops.gemm(dy, w, d_o_plain, b_t=True)
    d_o = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    dws = torch.full((B, H, N), float("nan"), device=dev)
    ops.gemm(dy, w, d_o, b_t=True, act=ops.UB_ACT_DOT_AUX, aux_in=o, dot_out=dws, dot_seq_len=N)
    torch.cuda.synchronize()
    same = torch.equal(d_o, d_o_plain)
    OK &= same
    print(f"gemm dot_aux dO bit-identical to the plain GEMM: {same}")
    want = (d_o.float() * o.float()).view(B, N, H, 64).sum(-1).permute(0, 2, 1)
    report(f"gemm dot_aux D B={B} N={N} H={H}", dws, want, 2e-6)
    if time_it:
        t0 = timeit(lambda: ops.gemm(dy, w, d_o_plain, b_t=True))
        t1 = timeit(lambda: ops.gemm(dy, w, d_o, b_t=True, act=ops.UB_ACT_DOT_AUX, aux_in=o, dot_out=dws, dot_seq_len=N))
        print(f"   proj dgrad {M}x{D}x{D}: plain {t0*1e3:.1f} us, with D {t1*1e3:.1f} us")


def check_ln(rows, D):
    g = torch.Generator(device=dev).manual_seed(D + rows)
    x = torch.randn(rows, D, device=dev, generator=g) * 2 + 0.3
    gam = 1 + 0.1 * torch.randn(D, device=dev, generator=g); bet = 0.1 * torch.randn(D, device=dev, generator=g)
    out = torch.empty(rows, D, device=dev, dtype=torch.bfloat16)
    ops.layernorm_fwd(x, gam, bet, 1e-6, out)
    report(f"ln_fwd bf16 D={D}", out, F.layer_norm(x, (D,), gam, bet, 1e-6), 4e-3)
    outf = torch.empty(rows, D, device=dev)
    ops.layernorm_fwd(x, gam, bet, 1e-5, outf)
    report(f"ln_fwd fp32 D={D}", outf, F.layer_norm(x, (D,), gam, bet, 1e-5), 1e-5)
    # gather + post add
    src = torch.randperm(rows, device=dev, generator=g)[: rows // 2].int().contiguous()
    tab = torch.randn(37, D, device=dev, generator=g); pidx = torch.randint(0, 37, (rows // 2,), device=dev, generator=g).int()
    outg = torch.empty(rows // 2, D, device=dev)
    ops.layernorm_fwd(x, gam, bet, 1e-6, outg, src_rows=src, post_add=tab, post_idx=pidx)
    report(f"ln_fwd gather+add D={D}", outg, F.layer_norm(x[src.long()], (D,), gam, bet, 1e-6) + tab[pidx.long()], 1e-5)
    # backward
    xr = x.clone().requires_grad_(); gr = gam.clone().requires_grad_(); br = bet.clone().requires_grad_()
    dy = torch.randn(rows, D, device=dev, generator=g).bfloat16()
    dx_in = torch.randn(rows, D, device=dev, generator=g)
    F.layer_norm(xr, (D,), gr, br, 1e-6).backward(dy.float())
    dx_out = torch.empty(rows, D, device=dev); dxs = torch.empty(rows, D, device=dev, dtype=torch.bfloat16)
    dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev)
    rs = torch.rand(rows // 16, device=dev, generator=g) + 0.5
    dsum = torch.zeros(D, device=dev)
    ops.layernorm_bwd(dy, x, gam, 1e-6, dx_in, dx_out, dxs, rs, 16, dg, db, dsum)
    report(f"ln_bwd dsum D={D}", dsum, ((xr.grad + dx_in) * rs.repeat_interleave(16)[:, None]).sum(0), 1e-4)
    report(f"ln_bwd dx D={D}", dx_out, xr.grad + dx_in, 1e-5)
    report(f"ln_bwd dxs D={D}", dxs, (xr.grad + dx_in) * rs.repeat_interleave(16)[:, None], 4e-3)
    report(f"ln_bwd dgamma D={D}", dg, gr.grad, 1e-4)
    report(f"ln_bwd dbeta D={D}", db, br.grad, 1e-4)


def check_dec_tail(rows, D):
    g = torch.Generator(device=dev).manual_seed(D * 3 + rows)
    y = torch.randn(rows, D, device=dev, generator=g)
    gam = 1 + 0.1 * torch.randn(D, device=dev, generator=g); bet = 0.1 * torch.randn(D, device=dev, generator=g)
    tgt = F.normalize(torch.randn(rows, D, device=dev, generator=g), dim=-1)
    yr = y.clone().requires_grad_(); gr = gam.clone().requires_grad_(); br = bet.clone().requires_grad_()
    u = F.layer_norm(yr, (D,), gr, br, 1e-6)
    oref = u / u.norm(dim=-1, keepdim=True)
    lref = (2 - 2 * (oref * tgt).sum(-1)).mean()
    lref.backward()
    out = torch.empty(rows, D, device=dev); lacc = torch.zeros(1, device=dev)
    ops.dec_tail_fwd(y, gam, bet, 1e-6, out, tgt, lacc, 1.0 / rows)
    report(f"dec_tail_fwd out D={D}", out, oref, 1e-5)
    report(f"dec_tail_fwd loss D={D}", lacc, lref.detach().view(1), 1e-5)
    dy = torch.empty(rows, D, device=dev, dtype=torch.bfloat16); dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev)
    ops.dec_tail_bwd(y, gam, bet, 1e-6, tgt, -2.0 / rows, dy, dg, db)
    report(f"dec_tail_bwd dy D={D}", dy, yr.grad, 4e-3)
    report(f"dec_tail_bwd dgamma D={D}", dg, gr.grad, 1e-4)
    report(f"dec_tail_bwd dbeta D={D}", db, br.grad, 1e-4)
    z = torch.randn(rows, D, device=dev, generator=g); zz = z.clone()
    ops.l2norm_rows(zz)
    report(f"l2norm_rows D={D}", zz, z / z.norm(dim=-1, keepdim=True), 1e-6)


def check_tokens():
    global OK
    g = torch.Generator(device=dev).manual_seed(7)
    B, T, H, W = 2, 4, 64, 96
    for tub in (1, 2):
        x = torch.randn(B, 3, T, H, W, device=dev, generator=g)
        n_tok = B * (T // tub) * (H // 16) * (W // 16)
        out = torch.empty(n_tok, 3 * tub * 256, device=dev, dtype=torch.bfloat16)
        ops.patchify(x, out, tub)
        ref = x.reshape(B, 3, T // tub, tub, H // 16, 16, W // 16, 16).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(n_tok, -1)
        ok = torch.equal(out, ref.bfloat16()); OK &= ok
        print(f"patchify tub={tub}: {'bit-exact OK' if ok else 'FAIL'}")
    # mask select vs torch
    frames, P, Tm = 16, 196, 8
    attn = torch.rand(frames, P, device=dev, generator=g) + 1e-3
    q = torch.empty(frames, P, device=dev).exponential_(1, generator=g)
    n_vis = P - int(P * 0.8)
    mask = torch.empty(1, frames * P, device=dev, dtype=torch.uint8); vis = torch.empty(1, frames // Tm, Tm * n_vis, device=dev, dtype=torch.int32)
    tea = torch.empty_like(vis)
    ops.mask_select(attn, q, mask, vis, tea, Tm, 1, n_vis)
    order = torch.topk(attn / q, P, dim=-1).indices
    m = torch.ones(frames, P, device=dev); m[torch.arange(frames, device=dev).view(-1, 1).repeat(1, n_vis), order[:, :n_vis]] = 0
    mref = m.view(frames // Tm, -1).bool()
    ok = torch.equal(mask.view(frames // Tm, -1).bool(), mref); OK &= ok
    print(f"mask_select (multinomial form): {'bit-exact OK' if ok else 'FAIL'}")
    vref = (~mref).nonzero()[:, 1].view(frames // Tm, -1).int()
    ok = torch.equal(vis[0], vref); OK &= ok
    print(f"mask_select vis_idx: {'bit-exact OK' if ok else 'FAIL'}")
    b_ = torch.arange(frames // Tm, device=dev).view(-1, 1); t_ = vref // P; p_ = vref % P
    ok = torch.equal(tea[0], ((b_ * Tm + t_) * (P + 1) + 1 + p_).int()); OK &= ok
    print(f"mask_select tea_rows: {'bit-exact OK' if ok else 'FAIL'}")
    # greedy
    k = 2
    maskk = torch.empty(k, frames * P, device=dev, dtype=torch.uint8); visk = torch.empty(k, frames // Tm, Tm * n_vis, device=dev, dtype=torch.int32)
    ops.mask_select(attn, None, maskk, visk, None, Tm, k, n_vis)
    order = attn.sort(dim=1, descending=True).indices
    gm = torch.ones(k, frames, P, dtype=torch.bool, device=dev)
    for i in range(k):
        gm[i].scatter_(1, order[:, i::k][:, :n_vis], False)
    ok = torch.equal(maskk.view(k, frames, P).bool(), gm); OK &= ok
    print(f"mask_select (greedy k=2): {'bit-exact OK' if ok else 'FAIL'}")
    # gather rows
    src = torch.randn(1000, 768, device=dev, generator=g).bfloat16()
    idx = torch.randint(0, 100, (4, 50), device=dev, generator=g).int()
    out = torch.empty(200, 768, device=dev, dtype=torch.bfloat16)
    ops.gather_rows(src, idx.view(-1), out, rows_per_group=50, group_stride_rows=250)
    ref = src[(idx.long() + torch.arange(4, device=dev).view(-1, 1) * 250).view(-1)]
    ok = torch.equal(out, ref); OK &= ok
    print(f"gather_rows: {'bit-exact OK' if ok else 'FAIL'}")
    x = torch.randn(1000, 2304, device=dev, generator=g).bfloat16(); cs = torch.zeros(2304, device=dev)
    ops.colsum_bf16(x, cs)
    report("colsum_bf16", cs, x.float().sum(0), 1e-5)
    cs2 = torch.full((2304,), 7.0, device=dev)
    ops.colsum_bf16(x, cs2, skip=(768, 1536))             # the key third of a q|k|v bias gradient is left untouched
    ref2 = x.float().sum(0) + 7.0; ref2[768:1536] = 7.0
    report("colsum_bf16 skip range", cs2, ref2, 1e-5)
    xf = torch.randn(640, 768, device=dev, generator=g); o16 = torch.empty(640, 768, device=dev, dtype=torch.bfloat16)
    rs = torch.rand(4, device=dev, generator=g)
    ops.cast_scale_bf16(xf, o16, rs, 160)
    ok = torch.equal(o16, (xf * rs.repeat_interleave(160)[:, None]).bfloat16()); OK &= ok
    print(f"cast_scale_bf16: {'bit-exact OK' if ok else 'FAIL'}")


def check_optim():
    g = torch.Generator(device=dev).manual_seed(3)
    n, nd = 4096 * 4, 4096 * 3
    p = torch.randn(n, device=dev, generator=g); gr = torch.randn(n, device=dev, generator=g) * 0.01
    pr1 = p[:nd].clone().requires_grad_(); pr2 = p[nd:].clone().requires_grad_()
    opt = torch.optim.AdamW([dict(params=[pr1], weight_decay=0.05), dict(params=[pr2], weight_decay=0.0)], lr=1e-3, betas=(0.9, 0.95), eps=1e-8)
    m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev); w16 = torch.empty(n, device=dev, dtype=torch.bfloat16)
    for step in (1, 2, 3):
        pr1.grad = gr[:nd].clone(); pr2.grad = gr[nd:].clone(); opt.step()
        ops.adamw(p, gr, m, v, w16, nd, 1e-3, 0.05, 0.9, 0.95, 1e-8, step)
    report("adamw params (3 steps)", p, torch.cat([pr1.detach(), pr2.detach()]), 1e-6)
    report("adamw bf16 shadow", w16, p.bfloat16(), 1e-7)
    s = torch.zeros(1, device=dev); ops.sumsq(gr, s)
    report("sumsq", s, (gr * gr).sum().view(1), 1e-5)


if __name__ == "__main__":
    check_tokens()
    check_ln(4096, 768); check_ln(1024, 128); check_ln(512, 1024)
    check_dec_tail(4096, 512); check_dec_tail(300, 128)
    check_optim()
    check_attention(2, 17, 2)
    check_attention(4, 197, 12)
    check_attention(3, 128, 4)
    check_attention(5, 240, 3)
    check_attention(2, 256, 3)
    check_attention(3, 320, 12)
    check_attention(2, 64, 2)
    check_attention(1, 1568, 4)
    check_attention_tc_late_maximum()
    check_gemm_dot_aux(5, 320, 12); check_gemm_dot_aux(3, 197, 4); check_gemm_dot_aux(2, 1568, 12)
    check_gemm_dot_aux(32, 320, 12, time_it=True)
    check_attention(256, 197, 12, time_it=True)
    check_attention(32, 320, 12, time_it=True)
    print("ALL OK" if OK else "SOME FAILED")
    sys.exit(0 if OK else 1)
