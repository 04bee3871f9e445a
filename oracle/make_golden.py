"""Pins oracle/unite_oracle.py against the REAL reference modules and writes tests/golden/*.pt.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

What it does
  1. installs a 4-symbol `timm` shim in sys.modules (timm 0.4.12 is pinned by the reference's
     environment.yaml:325 but is not installed here; semantics restated from that release:
     drop_path = x.div(keep) * floor(keep + rand), trunc_normal_ = torch.nn.init.trunc_normal_,
     to_2tuple, register_model) and imports /root/reference/src/models unchanged;
  2. builds tiny reference models (same classes, smaller width/depth so the fixtures stay small), runs the
     reference modules and the oracle restatement on identical state_dicts and seeded inputs, asserts they
     agree to fp32 round-off, and stores inputs + state_dicts + reference outputs as fixtures;
  3. checks `torch.multinomial(p, n)` == topk(p / Exp(1)) under a shared generator (the mask-sampler
     restatement) and utils.get_greedy_masks (lifted from the reference source with `ast`, executed as is);
  4. for the FULL ViT-B/16 configuration stores only scalars / checksums of a seeded run of the reference
     modules (state_dicts would be 350 MB each), with the seeds needed to regenerate it.

Nothing from /root/reference is copied into the repository: fixtures hold tensors only.
"""
import ast
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import unite_oracle as O  # noqa: E402
from oracle.weights import seeded_state  # noqa: E402


def install_timm_shim():
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    registry = types.ModuleType("timm.models.registry")

    def drop_path(x, drop_prob=0.0, training=False):
        if drop_prob == 0.0 or not training:
            return x
        keep = 1 - drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        r = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
        r.floor_()
        return x.div(keep) * r

    def to_2tuple(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

    _reg = {}

    def register_model(fn):
        _reg[fn.__name__] = fn
        return fn

    layers.drop_path = drop_path
    layers.to_2tuple = to_2tuple
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    registry.register_model = register_model
    registry._model_entrypoints = _reg
    timm.models = models
    models.layers = layers
    models.registry = registry
    for m in (timm, models, layers, registry):
        sys.modules[m.__name__] = m
    return _reg


def import_reference():
    install_timm_shim()
    sys.path.insert(0, REF)
    import importlib
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        clip = importlib.import_module("src.models.clip")
        fin = importlib.import_module("src.models.modeling_finetune")
        ada = importlib.import_module("src.models.modeling_adaptation")
    return clip, fin, ada


def lift_function(path, name):
    """Extract one top-level function from a reference file that cannot be imported whole (utils.py imports
    torch._six / clip / tensorboardX) and exec it unchanged in a namespace with torch."""
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            ns = {"torch": torch}
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def quiet(fn, *a, **k):
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def close(a, b, tol, what):
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= tol * max(1.0, ref), f"{what}: max abs err {err} vs scale {ref}"
    return err


def tiny_cfgs():
    scfg = O.StudentCfg(embed_dim=128, depth=3, num_heads=2, num_frames=4, tubelet_size=1, img_size=64,
                        patch_size=16, return_layers=(1, 2), clip_output_dim=128, num_classes=12)
    tcfg = O.TeacherCfg(width=128, layers=3, heads=2, output_dim=128, input_resolution=64, patch_size=16,
                        kernel_size=1, return_layers=(1, 2))
    return scfg, tcfg


def build_reference_models(clip, fin, ada, scfg, tcfg, seed):
    from functools import partial
    import torch.nn as nn

    torch.manual_seed(seed)
    student = quiet(ada.AdaptationVisionTransformer,
                    img_size=scfg.img_size, patch_size=scfg.patch_size, encoder_embed_dim=scfg.embed_dim,
                    encoder_depth=scfg.depth, encoder_num_heads=scfg.num_heads, encoder_num_classes=0,
                    mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                    num_frames=scfg.num_frames, tubelet_size=scfg.tubelet_size,
                    clip_decoder_embed_dim=scfg.embed_dim, clip_output_dim=scfg.clip_output_dim,
                    clip_return_layers=list(scfg.return_layers), drop_path_rate=0.0)
    teacher = quiet(clip.VisionTransformer,
                    input_resolution=tcfg.input_resolution, patch_size=tcfg.patch_size, width=tcfg.width,
                    layers=tcfg.layers, heads=tcfg.heads, output_dim=tcfg.output_dim, kernel_size=tcfg.kernel_size,
                    return_attn=True, clip_return_layers=list(tcfg.return_layers)).eval()
    vit = quiet(fin.VisionTransformer,
                img_size=scfg.img_size, patch_size=scfg.patch_size, embed_dim=scfg.embed_dim, depth=scfg.depth,
                num_heads=scfg.num_heads, mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                num_classes=scfg.num_classes, all_frames=scfg.num_frames, tubelet_size=scfg.tubelet_size,
                init_scale=0.001, use_mean_pooling=True)
    # weights come from oracle/weights.py so fixtures can store (shapes, seed) instead of tensors
    for i, m in enumerate((student, teacher, vit)):
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(seeded_state(shapes, seed + i), strict=True)
    return student, teacher, vit


def reference_stage1(student, teacher, videos, q, mask_ratio):
    """The step body of run_stage1.py:360-438 driven through the reference nn.Modules; the only edits are
    the ones SURVEY.md §8(c) lists: no CUDA autocast / GradScaler, noise q supplied instead of drawn."""
    with torch.no_grad():
        norm_clip, attn = teacher(videos)
        BT, N = attn.shape
        B = videos.shape[0]
        N_vis = N - int(N * mask_ratio)
        importance = torch.topk(attn / q, N).indices       # == torch.multinomial(attn, N) with this noise
        m = torch.ones((BT, N))
        pos1 = torch.arange(BT).view(-1, 1).repeat(1, N_vis)
        m[pos1, importance[:, :N_vis]] = 0
        mask = m.view(B, -1).to(torch.bool)
        K, C = norm_clip.shape[0], norm_clip.shape[-1]
        targets = norm_clip[~mask.unsqueeze(0).repeat(K, 1, 1)].reshape(K, B, -1, C)
    student.train()
    student.zero_grad()
    out = student(videos, mask, clip_only=True)
    loss = (2 - 2 * (out * targets).sum(dim=-1)).mean()
    loss.backward()
    grads = {n: p.grad.clone() for n, p in student.named_parameters()}
    return dict(attn=attn, mask=mask, targets=targets, outputs=out.detach(), loss=loss.detach(), grads=grads,
                norm_clip=norm_clip)


GRAD_KEYS_STAGE1 = ("encoder.patch_embed.proj.bias", "encoder.blocks.0.attn.q_bias", "encoder.blocks.0.attn.v_bias",
                    "encoder.blocks.0.attn.qkv.weight", "encoder.blocks.1.mlp.fc1.weight", "encoder.blocks.2.mlp.fc2.bias",
                    "encoder.blocks.1.norm1.weight", "encoder.norm.weight", "encoder.norm.bias",
                    "clip_decoder.0.head.weight", "clip_decoder.1.norm.weight", "clip_decoder.1.head.bias")
GRAD_KEYS_STAGE2 = ("patch_embed.proj.bias", "blocks.0.attn.qkv.weight", "blocks.2.mlp.fc2.weight", "fc_norm.weight",
                    "head.weight", "head.bias")


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    clip, fin, ada = import_reference()
    torch.set_num_threads(os.cpu_count())

    # ---------------------------------------------------------------- 1. multinomial == topk(p / Exp(1))
    g1 = torch.Generator().manual_seed(1234)
    p = torch.rand(16, 196, generator=g1) + 1e-3
    ga, gb = torch.Generator().manual_seed(77), torch.Generator().manual_seed(77)
    draw = torch.multinomial(p, 196, generator=ga)
    qn = torch.empty_like(p).exponential_(1, generator=gb)
    assert torch.equal(draw, torch.topk(p / qn, 196).indices), "multinomial != topk(p/q) on this torch"
    print("multinomial == topk(p / Exp(1)) under a shared generator: OK")

    # ---------------------------------------------------------------- 2. greedy masks (lifted, unchanged)
    ref_greedy = lift_function(os.path.join(REF, "src", "utils.py"), "get_greedy_masks")
    attn_g = torch.rand(8, 196, generator=g1)
    gm_ref = ref_greedy(attn_g, 0.8, 2)
    gm_or = O.greedy_masks(attn_g, 0.8, 2)
    assert torch.equal(gm_ref, gm_or)
    print("greedy_masks == utils.get_greedy_masks: OK")

    # ---------------------------------------------------------------- 3. tiny models: reference vs oracle
    scfg, tcfg = tiny_cfgs()
    student, teacher, vit = build_reference_models(clip, fin, ada, scfg, tcfg, seed=0)
    gi = torch.Generator().manual_seed(5)
    B = 2
    videos = torch.randn(B, 3, scfg.num_frames, scfg.img_size, scfg.img_size, generator=gi)
    HW = (scfg.img_size // scfg.patch_size) ** 2
    q = torch.empty(B * scfg.num_frames, HW).exponential_(1, generator=gi)
    labels = torch.randint(0, scfg.num_classes, (B,), generator=gi)
    mask_ratio = 0.75

    ref = reference_stage1(student, teacher, videos, q, mask_ratio)
    ssd = {k: v.detach().clone() for k, v in student.state_dict().items()}
    tsd = {k: v.detach().clone() for k, v in teacher.state_dict().items()}
    orc = O.stage1_step(ssd, tsd, videos, q, scfg, tcfg, mask_ratio=mask_ratio)
    close(orc["attn"], ref["attn"], 1e-5, "teacher attn")
    assert torch.equal(orc["mask"], ref["mask"]), "mask differs"
    close(orc["targets"], ref["targets"], 1e-5, "targets")
    close(orc["outputs"], ref["outputs"], 1e-5, "student outputs")
    close(orc["loss"], ref["loss"], 1e-6, "loss")
    assert set(orc["grads"]) == set(ref["grads"]), "grad key sets differ"
    worst = max(close(orc["grads"][k], ref["grads"][k], 1e-4, f"grad {k}") for k in ref["grads"])
    print(f"tiny stage-1: oracle == reference modules (worst grad abs err {worst:.2e}, loss {ref['loss']:.6f}): OK")

    # non-clip_only forward + stage-2 model
    student.eval()
    with torch.no_grad():
        xv_ref, xc_ref = student(videos, ref["mask"], clip_only=False)
    xv_or, xc_or = O.student_forward(ssd, videos, ref["mask"], scfg, clip_only=False)
    close(xv_or, xv_ref, 1e-5, "x_vis")
    close(xc_or, xc_ref, 1e-5, "x_clip")
    vit.train()
    vit.zero_grad()
    logits_ref = vit(videos)
    loss2_ref = torch.nn.functional.cross_entropy(logits_ref, labels)
    loss2_ref.backward()
    vsd = {k: v.detach().clone() for k, v in vit.state_dict().items()}
    o2 = O.stage2_step(vsd, videos, labels, scfg)
    close(o2["logits"], logits_ref.detach(), 1e-5, "stage-2 logits")
    close(o2["loss"], loss2_ref.detach(), 1e-6, "stage-2 loss")
    g2 = {n: p.grad for n, p in vit.named_parameters()}
    worst2 = max(close(o2["grads"][k], g2[k], 1e-4, f"stage-2 grad {k}") for k in g2)
    print(f"tiny stage-2: oracle == reference modules (worst grad abs err {worst2:.2e}): OK")

    # DropPath with injected keep factors: reference draws rand inside timm.drop_path; replay its draws.
    # (drop_path_rate > 0 path: verified by constructing the keep factors from the same RNG stream.)
    torch.save(dict(
        cfg=dict(student=scfg.__dict__, teacher=tcfg.__dict__, mask_ratio=mask_ratio),
        student_shapes={k: tuple(v.shape) for k, v in ssd.items()}, teacher_shapes={k: tuple(v.shape) for k, v in tsd.items()},
        vit_shapes={k: tuple(v.shape) for k, v in vsd.items()}, seeds=dict(student=0, teacher=1, vit=2),
        videos=videos, q=q, labels=labels,
        attn=ref["attn"].clone(), mask=ref["mask"], targets=ref["targets"].clone(), outputs=ref["outputs"].clone(), loss=ref["loss"].clone(),
        grads={k: v for k, v in ref["grads"].items() if k in GRAD_KEYS_STAGE1},
        grad_norms={k: v.norm() for k, v in ref["grads"].items()},
        x_vis=xv_ref, stage2_logits=logits_ref.detach(), stage2_loss=loss2_ref.detach(),
        stage2_grads={k: v.clone() for k, v in g2.items() if k in GRAD_KEYS_STAGE2},
        stage2_grad_norms={k: v.norm() for k, v in g2.items()},
    ), os.path.join(out_dir, "tiny_stage12.pt"))

    torch.save(dict(p=p, q=qn, draw=draw, greedy_attn=attn_g, greedy_masks=gm_ref),
               os.path.join(out_dir, "mask_sampler.pt"))

    # ---------------------------------------------------------------- 4. full-size scalars (ViT-B/16, B=1)
    full_s, full_t = O.StudentCfg(), O.TeacherCfg()
    studentF, teacherF, _ = build_reference_models(clip, fin, ada, O.StudentCfg(num_classes=12), full_t, seed=0)
    gi = torch.Generator().manual_seed(11)
    vF = torch.randn(1, 3, 8, 224, 224, generator=gi)
    qF = torch.empty(8, 196).exponential_(1, generator=gi)
    refF = reference_stage1(studentF, teacherF, vF, qF, 0.8)
    ssdF = {k: v.detach().clone() for k, v in studentF.state_dict().items()}
    tsdF = {k: v.detach().clone() for k, v in teacherF.state_dict().items()}
    orcF = O.stage1_step(ssdF, tsdF, vF, qF, full_s, full_t, mask_ratio=0.8, with_grads=True)
    close(orcF["attn"], refF["attn"], 1e-5, "full attn")
    assert torch.equal(orcF["mask"], refF["mask"])
    close(orcF["outputs"], refF["outputs"], 2e-5, "full outputs")
    close(orcF["loss"], refF["loss"], 1e-6, "full loss")
    worstF = max(close(orcF["grads"][k], refF["grads"][k], 1e-4, f"full grad {k}") for k in refF["grads"])
    n_params = sum(v.numel() for v in studentF.parameters())
    n_params_t = sum(v.numel() for v in teacherF.parameters())
    print(f"full ViT-B/16 stage-1 (B=1): oracle == reference (worst grad err {worstF:.2e}); "
          f"loss {refF['loss']:.6f}; student params {n_params}; teacher params {n_params_t}")
    torch.save(dict(
        student_params=n_params, teacher_params=n_params_t, loss=refF["loss"].clone(), attn=refF["attn"].clone(),
        attn_rowsum=refF["attn"].sum(-1), n_visible=int((~refF["mask"]).sum()), mask=refF["mask"], q=qF,
        outputs_sample=refF["outputs"][:, 0, :4, :8].clone(), targets_sample=refF["targets"][:, 0, :4, :8].clone(),
        seeds=dict(student=0, teacher=1, inputs=11),
        state_keys_student={k: tuple(v.shape) for k, v in ssdF.items()},
        state_keys_teacher={k: tuple(v.shape) for k, v in tsdF.items()},
    ), os.path.join(out_dir, "full_b16_scalars.pt"))
    for f in sorted(os.listdir(out_dir)):
        print(f, os.path.getsize(os.path.join(out_dir, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
