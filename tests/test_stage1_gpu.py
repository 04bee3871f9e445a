"""Stage-1 parity on the GPU: unite_b200 (CUDA kernels through the C ABI) vs the CPU oracle / golden fixtures.

Tolerances (BASELINE.json north_star): mask and gather indices bit-exact; per-token features within 1e-2
relative (bf16 operands vs the fp32 reference); loss within 1e-3 relative.  Gradients (our addition, SURVEY.md
§8(c)): per-parameter cosine >= 0.999 and relative L2 <= 2e-2 on the large tensors.
"""
import pytest
import torch

from tests.util import (build_student, build_teacher, cosine, load_golden, oracle_cfgs, per_token_rel, rel_l2, seeded_states)

pytestmark = pytest.mark.gpu
FEAT_TOL, LOSS_TOL = 1e-2, 1e-3


def _models(scfg, tcfg, ssd, tsd):
    student = build_student(scfg)
    teacher = build_teacher(tcfg)
    student.load_state_dict(ssd, strict=True)
    teacher.load_state_dict(tsd, strict=True)
    return student.cuda().train(), teacher.cuda().eval()


def _check_step(engine, ref, videos, q, big_grad_only=False):
    loss = engine.forward_backward(videos.cuda(), q.cuda(), attn_override=ref["attn"].cuda().contiguous())
    torch.cuda.synchronize()
    last = engine.last
    # teacher attention (fp) and the mask it implies given the ORACLE's attn (bit-exact)
    assert rel_l2(last["attn"], ref["attn"]) < FEAT_TOL
    assert torch.equal(last["mask"].cpu(), ref["mask"]), "mask differs from the reference"
    from oracle.unite_oracle import visible_indices
    assert torch.equal(last["vis_idx"].cpu().long(), visible_indices(ref["mask"])), "visible index list differs"
    K, B, Nv, C = ref["targets"].shape
    t_err = per_token_rel(last["targets"].view(K, B, Nv, C), ref["targets"])
    o_err = per_token_rel(last["outputs"], ref["outputs"])
    print(f"targets per-token rel: mean {t_err.mean():.2e} max {t_err.max():.2e}; outputs: mean {o_err.mean():.2e} max {o_err.max():.2e}")
    assert t_err.max() < FEAT_TOL, f"teacher target features off by {t_err.max():.3e}"
    assert o_err.max() < FEAT_TOL, f"student output features off by {o_err.max():.3e}"
    l_rel = abs(loss.item() - ref["loss"].item()) / abs(ref["loss"].item())
    print(f"loss {loss.item():.6f} vs {ref['loss'].item():.6f} (rel {l_rel:.2e})")
    assert l_rel < LOSS_TOL
    arena = engine.core.arena
    worst_cos, worst_rel = 1.0, 0.0
    for k, g_ref in ref["grads"].items():
        g = arena.g32(k)
        if big_grad_only and g_ref.numel() < 4096:
            continue
        c, r = cosine(g, g_ref), rel_l2(g, g_ref)
        worst_cos, worst_rel = min(worst_cos, c), max(worst_rel, r)
        assert c >= 0.999, f"grad {k}: cosine {c:.5f}"
        assert r <= 2e-2, f"grad {k}: rel L2 {r:.3e}"
    print(f"grads: worst cosine {worst_cos:.6f}, worst rel L2 {worst_rel:.2e}")


def test_tiny_stage1_against_golden_fixture():
    from unite_b200.engine import Stage1Engine
    fix = load_golden("tiny_stage12.pt")
    scfg, tcfg = oracle_cfgs(fix)
    ssd, tsd, _ = seeded_states(fix)
    student, teacher = _models(scfg, tcfg, ssd, tsd)
    eng = Stage1Engine(student, teacher, mask_ratio=fix["cfg"]["mask_ratio"])
    ref = dict(attn=fix["attn"], mask=fix["mask"], targets=fix["targets"], outputs=fix["outputs"], loss=fix["loss"], grads=fix["grads"])
    _check_step(eng, ref, fix["videos"], fix["q"])
    for k, n in fix["grad_norms"].items():
        got = eng.core.arena.g32(k).norm().item()
        assert abs(got - n.item()) <= 3e-2 * n.item() + 1e-6, f"|grad {k}| = {got} vs {n.item()}"


def test_full_vitb16_stage1_against_oracle():
    """ViT-B/16 student + CLIP-B/16 teacher, 8x224^2, B=2: the CUDA step against the oracle run here on the CPU."""
    from oracle import unite_oracle as O
    from oracle.weights import seeded_state
    from unite_b200.engine import Stage1Engine
    scfg, tcfg = O.StudentCfg(), O.TeacherCfg()
    student, teacher = build_student(scfg), build_teacher(tcfg)
    ssd = seeded_state({k: tuple(v.shape) for k, v in student.state_dict().items()}, 0)
    tsd = seeded_state({k: tuple(v.shape) for k, v in teacher.state_dict().items()}, 1)
    student.load_state_dict(ssd); teacher.load_state_dict(tsd)
    g = torch.Generator().manual_seed(21)
    B = 2
    videos = torch.randn(B, 3, 8, 224, 224, generator=g)
    q = torch.empty(B * 8, 196).exponential_(1, generator=g)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref = O.stage1_step(ssd, tsd, videos, q, scfg, tcfg, mask_ratio=0.8)
    assert ref["vis_idx"].shape == (B, 320)
    eng = Stage1Engine(student.cuda().train(), teacher.cuda().eval(), mask_ratio=0.8)
    _check_step(eng, ref, videos, q, big_grad_only=True)


@pytest.mark.parametrize("B,mask_ratio,n_vis", [(1, 0.8, 320), (3, 0.9, 160), (1, 0.75, 392), (3, 0.5, 784)])
def test_full_vitb16_stage1_odd_batches_and_other_mask_ratios(B, mask_ratio, n_vis):
    """Edge shapes of the same step: a single clip, an odd number of clips, and mask ratios other than the shipped 0.8 — 0.9 leaves
    160 visible tokens (half a 128-row tile in every student GEMM and attention item), 0.75 and 0.5 leave 392 / 784, which takes the
    student's attention (forward with LSE and backward) off the resident S <= 320 kernels onto the streamed long-sequence ones.
    run_stage1.py:378-393 (the mask ratio only enters through N_vis = P - int(P * ratio) per frame)."""
    from oracle import unite_oracle as O
    from oracle.weights import seeded_state
    from unite_b200.engine import Stage1Engine
    scfg, tcfg = O.StudentCfg(), O.TeacherCfg()
    student, teacher = build_student(scfg), build_teacher(tcfg)
    ssd = seeded_state({k: tuple(v.shape) for k, v in student.state_dict().items()}, 0)
    tsd = seeded_state({k: tuple(v.shape) for k, v in teacher.state_dict().items()}, 1)
    student.load_state_dict(ssd); teacher.load_state_dict(tsd)
    g = torch.Generator().manual_seed(100 + B + n_vis)
    videos = torch.randn(B, 3, 8, 224, 224, generator=g)
    q = torch.empty(B * 8, 196).exponential_(1, generator=g)
    ref = O.stage1_step(ssd, tsd, videos, q, scfg, tcfg, mask_ratio=mask_ratio)
    assert ref["vis_idx"].shape == (B, n_vis)
    eng = Stage1Engine(student.cuda().train(), teacher.cuda().eval(), mask_ratio=mask_ratio)
    _check_step(eng, ref, videos, q, big_grad_only=True)


def test_module_api_matches_engine_and_autograd_accumulates():
    """The reference-facing call `model(videos, mask, clip_only=True)` + loss.backward() gives the same numbers as
    the fused engine, and a second backward accumulates into .grad like torch does."""
    from unite_b200.engine import Stage1Engine
    fix = load_golden("tiny_stage12.pt")
    scfg, tcfg = oracle_cfgs(fix)
    ssd, tsd, _ = seeded_states(fix)
    student, teacher = _models(scfg, tcfg, ssd, tsd)
    videos, mask = fix["videos"].cuda(), fix["mask"].cuda()
    with torch.no_grad():
        feat, attn = teacher(videos)
    K, C = feat.shape[0], feat.shape[-1]
    targets = feat[~mask.unsqueeze(0).repeat(K, 1, 1)].reshape(K, videos.shape[0], -1, C)
    assert per_token_rel(targets, fix["targets"]).max() < FEAT_TOL
    out = student(videos, mask, clip_only=True)
    assert per_token_rel(out, fix["outputs"]).max() < FEAT_TOL
    loss = (2 - 2 * (out * targets).sum(dim=-1)).mean()
    loss.backward()
    g1 = {k: p.grad.clone() for k, p in student.named_parameters()}
    for k, gr in fix["grads"].items():
        assert cosine(g1[k], gr) >= 0.999, k
    out2 = student(videos, mask, clip_only=True)
    ((2 - 2 * (out2 * targets).sum(dim=-1)).mean()).backward()
    for k, p in student.named_parameters():
        assert rel_l2(p.grad, 2 * g1[k]) < 1e-3, f"{k}: second backward did not accumulate"
    # eval / no-grad path and the (x_vis, x_clip) return form
    student.eval()
    with torch.no_grad():
        x_vis, x_clip = student(videos, mask, clip_only=False)
    assert per_token_rel(x_vis, fix["x_vis"]).max() < FEAT_TOL
    assert x_clip.shape == fix["outputs"].shape


def test_optimizer_step_matches_torch_adamw():
    from unite_b200.engine import Stage1Engine
    fix = load_golden("tiny_stage12.pt")
    scfg, tcfg = oracle_cfgs(fix)
    ssd, tsd, _ = seeded_states(fix)
    student, teacher = _models(scfg, tcfg, ssd, tsd)
    eng = Stage1Engine(student, teacher, mask_ratio=fix["cfg"]["mask_ratio"], lr=1e-3, weight_decay=0.05)
    videos, q = fix["videos"].cuda(), fix["q"].cuda()
    before = {k: v.detach().clone() for k, v in student.state_dict().items()}
    eng.optimizer.zero_grad()
    eng.forward_backward(videos, q)
    grads = {k: eng.core.arena.g32(k).clone() for k in before}
    eng.optimizer.step()
    torch.cuda.synchronize()
    # replay with torch.optim.AdamW on clones (decay on >=2-D weights only: optim_factory.py:83-88)
    ps = {k: before[k].clone().requires_grad_() for k in before}
    dec = [p for k, p in ps.items() if not (p.ndim == 1 or k.endswith(".bias"))]
    nod = [p for k, p in ps.items() if (p.ndim == 1 or k.endswith(".bias"))]
    opt = torch.optim.AdamW([dict(params=dec, weight_decay=0.05), dict(params=nod, weight_decay=0.0)], lr=1e-3, betas=(0.9, 0.95), eps=1e-8)
    for k, p in ps.items():
        p.grad = grads[k]
    opt.step()
    for k, p in ps.items():
        assert rel_l2(student.state_dict()[k], p) < 1e-6, k
    assert abs(eng.optimizer.grad_norm().item() - torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()) < 1e-3
    # graph-friendly form (scalars in device memory, gradient norm accumulated by the AdamW pass itself): same update
    eng.optimizer.zero_grad()
    eng.forward_backward(videos, q)
    g2 = eng.core.arena.grads.clone()
    p_before = eng.core.arena.params.clone()
    m_before, v_before = eng.optimizer.exp_avg.clone(), eng.optimizer.exp_avg_sq.clone()
    eng.optimizer.prepare_step()
    eng.optimizer.step_dev()
    torch.cuda.synchronize()
    p_dev = eng.core.arena.params.clone()
    assert abs(eng.optimizer.grad_norm().item() - g2.double().norm().item()) < 1e-3 * g2.double().norm().item() + 1e-6
    eng.core.arena.params.copy_(p_before); eng.optimizer.exp_avg.copy_(m_before); eng.optimizer.exp_avg_sq.copy_(v_before)
    eng.optimizer.step_count -= 1
    eng.optimizer.step()
    torch.cuda.synchronize()
    assert rel_l2(p_dev, eng.core.arena.params) < 1e-6


def test_clip_grad_matches_torch_clip_grad_norm_then_adamw():
    """utils.py:613-615: clip_grad_norm_(parameters, clip_grad) between backward and optimizer.step(); here the coefficient is
    computed on the device and folded into the AdamW pass.  Checked in the clipping regime (max_norm well below the norm) and
    in the no-op regime (max_norm above it)."""
    from unite_b200.engine import Stage1Engine
    fix = load_golden("tiny_stage12.pt")
    scfg, tcfg = oracle_cfgs(fix)
    ssd, tsd, _ = seeded_states(fix)
    student, teacher = _models(scfg, tcfg, ssd, tsd)
    eng = Stage1Engine(student, teacher, mask_ratio=fix["cfg"]["mask_ratio"], lr=1e-3, weight_decay=0.05)
    videos, q = fix["videos"].cuda(), fix["q"].cuda()
    for regime in ("clip", "noop"):
        before = {k: v.detach().clone() for k, v in student.state_dict().items()}
        m0, v0, t0 = eng.optimizer.exp_avg.clone(), eng.optimizer.exp_avg_sq.clone(), eng.optimizer.step_count
        eng.optimizer.zero_grad()
        eng.forward_backward(videos, q)
        grads = {k: eng.core.arena.g32(k).clone() for k in before}
        total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
        max_norm = total * (0.25 if regime == "clip" else 4.0)
        eng.optimizer.step(max_norm=max_norm)
        torch.cuda.synchronize()
        assert abs(eng.optimizer.grad_norm().item() - total) < 1e-4 * total              # the norm BEFORE clipping is what is reported
        ps = {k: before[k].clone().requires_grad_() for k in before}
        dec = [p for k, p in ps.items() if not (p.ndim == 1 or k.endswith(".bias"))]
        nod = [p for k, p in ps.items() if (p.ndim == 1 or k.endswith(".bias"))]
        opt = torch.optim.AdamW([dict(params=dec, weight_decay=0.05), dict(params=nod, weight_decay=0.0)], lr=1e-3, betas=(0.9, 0.95), eps=1e-8)
        # bring torch's moments to the engine's state before this step
        for k, p in ps.items():
            o, n = eng.core.arena.offsets[k]
            opt.state[p] = dict(step=torch.tensor(float(t0)), exp_avg=m0[o:o + n].view_as(p).clone(), exp_avg_sq=v0[o:o + n].view_as(p).clone())
            p.grad = grads[k].clone()
        torch.nn.utils.clip_grad_norm_(list(ps.values()), max_norm)
        opt.step()
        for k, p in ps.items():
            assert rel_l2(student.state_dict()[k], p) < 1e-6, (regime, k)


def test_vitl_student_tubelet2_teacher_kernel2_against_oracle():
    """BASELINE configs[4] shapes: ViT-L/16 student (D=1024, 24 layers, 16 heads), 16 frames, tubelet 2, CLIP-B/16 teacher built
    with kernel_size=2 -> 1568 tokens, 320 visible.  B=1, against the oracle on the CPU."""
    from oracle import unite_oracle as O
    from oracle.weights import seeded_state
    from unite_b200.engine import Stage1Engine
    scfg = O.StudentCfg(embed_dim=1024, depth=24, num_heads=16, num_frames=16, tubelet_size=2)
    tcfg = O.TeacherCfg(kernel_size=2)
    student, teacher = build_student(scfg), build_teacher(tcfg)
    ssd = seeded_state({k: tuple(v.shape) for k, v in student.state_dict().items()}, 5)
    tsd = seeded_state({k: tuple(v.shape) for k, v in teacher.state_dict().items()}, 6)
    student.load_state_dict(ssd); teacher.load_state_dict(tsd)
    g = torch.Generator().manual_seed(77)
    videos = torch.randn(1, 3, 16, 224, 224, generator=g)
    q = torch.empty(8, 196).exponential_(1, generator=g)
    ref = O.stage1_step(ssd, tsd, videos, q, scfg, tcfg, mask_ratio=0.8)
    assert ref["vis_idx"].shape == (1, 320)
    eng = Stage1Engine(student.cuda().train(), teacher.cuda().eval(), mask_ratio=0.8)
    _check_step(eng, ref, videos, q, big_grad_only=True)


def test_stage1_step_from_uint8_frames_equals_step_from_normalised_clip():
    """§8 row f2: feeding decoded uint8 frames (normalised on the device inside the patchify kernel) gives the same mask and
    the same loss (up to the order of the fp32 atomic adds) as feeding the fp32 clip the reference's data pipeline would have produced on the host."""
    from oracle.unite_oracle import normalize_frames_u8
    from unite_b200.engine import Stage1Engine
    fix = load_golden("tiny_stage12.pt")
    scfg, tcfg = oracle_cfgs(fix)
    ssd, tsd, _ = seeded_states(fix)
    student, teacher = _models(scfg, tcfg, ssd, tsd)
    eng = Stage1Engine(student, teacher, mask_ratio=fix["cfg"]["mask_ratio"])
    B, _, T, H, W = fix["videos"].shape
    frames = torch.randint(0, 256, (B, T, H, W, 3), generator=torch.Generator().manual_seed(11), dtype=torch.uint8)
    q = fix["q"].cuda()
    eng.optimizer.zero_grad()
    l_u8 = eng.forward_backward(frames.cuda(), q).clone()
    mask_u8 = eng.last["mask"].clone()
    g_u8 = eng.core.arena.grads.clone()
    eng.optimizer.zero_grad()
    l_f32 = eng.forward_backward(normalize_frames_u8(frames).cuda(), q).clone()
    torch.cuda.synchronize()
    assert torch.equal(mask_u8, eng.last["mask"])
    assert abs(l_u8.item() - l_f32.item()) <= 2e-6 * abs(l_f32.item()), (l_u8.item(), l_f32.item())   # atomic summation order only
    assert rel_l2(g_u8, eng.core.arena.grads) < 1e-5      # fp32 red.add ordering only


@pytest.mark.parametrize("kind", ["mse", "smooth_l1", "l1"])
def test_alternative_alignment_losses_against_oracle(kind):
    """clip_loss_type in {'mse','smooth_l1','l1'} (run_stage1.py:403-408,432-433), tiny configuration,
    against the oracle's autograd on the CPU.  Same tolerances as the shipped 'l2' form ('l1' gradients are sign(d)/n: an
    element of d within bf16 noise of zero can flip, so its gradient check is on the big tensors' cosine only)."""
    from oracle import unite_oracle as O
    from unite_b200.engine import Stage1Engine
    fix = load_golden("tiny_stage12.pt")
    scfg, tcfg = oracle_cfgs(fix)
    ssd, tsd, _ = seeded_states(fix)
    student, teacher = _models(scfg, tcfg, ssd, tsd)
    eng = Stage1Engine(student, teacher, mask_ratio=fix["cfg"]["mask_ratio"], clip_loss_type=kind)
    ref = O.stage1_step(ssd, tsd, fix["videos"], fix["q"], scfg, tcfg, mask_ratio=fix["cfg"]["mask_ratio"], clip_loss_type=kind)
    eng.optimizer.zero_grad()
    loss = eng.forward_backward(fix["videos"].cuda(), fix["q"].cuda(), attn_override=ref["attn"].cuda().contiguous())
    torch.cuda.synchronize()
    assert torch.equal(eng.last["mask"].cpu(), ref["mask"])
    l_rel = abs(loss.item() - ref["loss"].item()) / abs(ref["loss"].item())
    print(f"{kind}: loss {loss.item():.6e} vs {ref['loss'].item():.6e} (rel {l_rel:.2e})")
    assert l_rel < (3e-3 if kind == "l1" else LOSS_TOL * 2)
    worst = 1.0
    for k, g_ref in ref["grads"].items():
        if g_ref.numel() < 4096:
            continue
        c = cosine(eng.core.arena.g32(k), g_ref)
        worst = min(worst, c)
        assert c >= (0.99 if kind == "l1" else 0.999), f"{kind} grad {k}: cosine {c:.5f}"
    print(f"{kind}: worst big-tensor gradient cosine {worst:.6f}")
