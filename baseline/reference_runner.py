"""Runs the UNMODIFIED reference model modules (reddyav1/unite src/models) as a baseline — CPU fp32 and GPU eager bf16.

Not product code: only bench.py (`--impl reference`, `cpu_baseline`, `gpu_eager_baseline`) and tools/ call it.  Nothing from
`unite_b200` is imported here, so a process that only runs this file never loads libunite_b200.so.

The reference has no setup.py / pyproject (it is a script tree), so `pip install --target baseline/_ref /root/reference`
cannot work; `install()` instead copies the model package `src/models/*.py` (+ `src/__init__.py`) unchanged into the
git-ignored `baseline/_ref/` so that it travels to the GPU box (where /root/reference does not exist).  `timm` 0.4.12
(environment.yaml:325) is absent from the image: the four symbols the model files import are provided by a shim
(semantics restated from that release).

The reference's step LOOP (run_stage1.py:294-505) cannot be imported (src/utils.py needs torch._six, OpenAI clip,
tensorboardX; src/knn.py is missing), so `stage1_step_reference` restates the loop body line by line around the
unmodified modules:
    teacher forward under autocast, no grad            run_stage1.py:360-377
    multinomial attention mask                         :379-387   (mask built on attn.device: the CPU tensor of :383
                                                                   cannot be indexed with CUDA indices in torch >= 2)
    boolean-mask target gather                         :389-397
    student forward under autocast, l2 loss            :410-438
    loss.item() x2                                     :440-441
    zero_grad, backward, grad-norm, optimizer step     :451-456, src/utils.py:608-622, :631-643
    torch.cuda.synchronize()                           :458
The optimizer is torch.optim.AdamW with the decay / no-decay groups of src/optim_factory.py:76-118.
"""
import contextlib
import io
import os
import shutil
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
MODEL_FILES = ["__init__.py", "clip.py", "modeling_finetune.py", "modeling_adaptation.py", "modeling_pretrain.py",
               "modeling_pretrain_umt.py"]


def install(force: bool = False) -> bool:
    """Copy the reference's model package into baseline/_ref (git-ignored).  Returns True when baseline/_ref is usable."""
    dst = os.path.join(REF_DST, "src", "models")
    if os.path.isdir(os.path.join(REF_SRC, "src", "models")) and (force or not os.path.exists(os.path.join(dst, "clip.py"))):
        os.makedirs(dst, exist_ok=True)
        for f in MODEL_FILES:
            shutil.copyfile(os.path.join(REF_SRC, "src", "models", f), os.path.join(dst, f))
        open(os.path.join(REF_DST, "src", "__init__.py"), "w").close()
    return available()


def available() -> bool:
    return os.path.exists(os.path.join(REF_DST, "src", "models", "modeling_adaptation.py"))


def install_timm_shim():
    """timm.models.layers.{drop_path,to_2tuple,trunc_normal_} + timm.models.registry.register_model (timm 0.4.12)."""
    import torch
    if "timm" in sys.modules and hasattr(sys.modules["timm"], "_ub_shim"):
        return sys.modules["timm.models.registry"]._model_entrypoints
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    registry = types.ModuleType("timm.models.registry")

    def drop_path(x, drop_prob: float = 0.0, training: bool = False):
        if drop_prob == 0.0 or not training:
            return x
        keep = 1 - drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        r = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
        r.floor_()
        return x.div(keep) * r

    def to_2tuple(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

    reg = {}

    def register_model(fn):
        reg[fn.__name__] = fn
        return fn

    layers.drop_path, layers.to_2tuple, layers.trunc_normal_ = drop_path, to_2tuple, torch.nn.init.trunc_normal_
    registry.register_model, registry._model_entrypoints = register_model, reg
    timm.models, models.layers, models.registry = models, layers, registry
    timm._ub_shim = True
    for m in (timm, models, layers, registry):
        sys.modules[m.__name__] = m
    return reg


def import_reference():
    """-> (clip module, modeling_adaptation module, registry dict) of the unmodified reference model package."""
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python -c 'import __graft_entry__ as g; g.build()'` in the build "
                           "container (it copies /root/reference/src/models there)")
    reg = install_timm_shim()
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    import importlib
    with contextlib.redirect_stdout(io.StringIO()):
        clip = importlib.import_module("src.models.clip")
        ada = importlib.import_module("src.models.modeling_adaptation")
    return clip, ada, reg


def build_reference_models(seed: int = 0, drop_path: float = 0.1):
    """Student / teacher exactly as run_stage1.py:273-291 and :782-788 build them from configs/stage1_config.yaml; random
    init by the reference's own initialisers under torch.manual_seed(seed)."""
    import torch
    clip, ada, reg = import_reference()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        student = reg["adaptation_umt_base_patch16_224"](
            pretrained=False, use_learnable_pos_emb=False, drop_path_rate=drop_path, use_checkpoint=False, checkpoint_num=0,
            clip_decoder_embed_dim=768, clip_output_dim=512, clip_norm_type="l2", num_frames=8, tubelet_size=1,
            clip_return_layers=[6, 7, 8, 9, 10, 11], clip_student_return_interval=1, use_cls_token=False)
        teacher = clip.clip_b16(pretrained=False, clip_norm_type="l2", input_resolution=224, return_attn=True,
                                clip_return_layers=[6, 7, 8, 9, 10, 11], clip_return_interval=1)
    return student, teacher


def param_groups(model, weight_decay, skip_list=()):
    """src/optim_factory.py:76-118 without layer decay (layer_decay: 1.0 in configs/stage1_config.yaml)."""
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if (p.ndim == 1 or name.endswith(".bias") or name in skip_list) else decay).append(p)
    return [dict(params=decay, weight_decay=weight_decay, lr_scale=1.0), dict(params=no_decay, weight_decay=0.0, lr_scale=1.0)]


def get_grad_norm_(parameters):
    """src/utils.py:631-643."""
    import torch
    ps = [p for p in parameters if p.grad is not None]
    return torch.norm(torch.stack([torch.norm(p.grad.detach(), 2.0) for p in ps]), 2.0)


def stage1_step_reference(student, teacher, optimizer, videos, mask_ratio=0.8, autocast=None):
    """The body of run_stage1.py:360-458 (mask_type='attention', clip_loss_type='l2', clip_loss_data='mixed')."""
    import torch
    ac = autocast if autocast is not None else contextlib.nullcontext
    with torch.no_grad():
        B, C, T, H, W = videos.shape
        with ac():
            norm_clip, attn = teacher(videos)                                         # :375
        BT, N = attn.shape
        N_vis = N - int(N * mask_ratio)
        importance = torch.multinomial(attn.float(), N)                                # :382
        bool_masked_pos = torch.ones((BT, N), device=attn.device)                      # :383 (device: see module docstring)
        pos1 = torch.arange(BT, device=attn.device).view(-1, 1).repeat(1, N_vis)
        pos2 = importance[:, :N_vis]
        bool_masked_pos[pos1, pos2] = 0
        bool_masked_pos = bool_masked_pos.view(B, -1).to(torch.bool)
        C_CLIP = norm_clip.shape[-1]
        K = norm_clip.shape[0]
        clip_bool_masked_pos = bool_masked_pos.unsqueeze(0).repeat(K, 1, 1)
        targets_clip = norm_clip[~clip_bool_masked_pos].reshape(K, B, -1, C_CLIP)      # :392
    with ac():
        outputs_clip = student(videos, bool_masked_pos, clip_only=True)               # :415
        loss_clip = (2 - 2 * (outputs_clip * targets_clip).sum(dim=-1)).mean()         # :431
        loss = loss_clip
    loss_clip_value = loss_clip.item()                                                 # :440
    loss_value = loss.item()                                                           # :441
    optimizer.zero_grad()
    loss.backward()                                                                    # utils.py:609
    norm = get_grad_norm_(student.parameters())                                        # utils.py:618
    optimizer.step()                                                                   # utils.py:619
    if videos.is_cuda:
        torch.cuda.synchronize()                                                       # :458
    return loss_value, norm


def run_stage1(device: str, batch: int, steps: int, warmup: int, seed: int = 0, drop_path: float = 0.1, world_batch=None,
               threads=None):
    """Times `steps` reference steps after `warmup` untimed ones.  device 'cuda' -> eager under
    torch.autocast('cuda', bfloat16) (the reference uses fp16 + GradScaler, run_stage1.py:372,410; bf16 needs no scaler),
    fused AdamW, CUDA-event timed; device 'cpu' -> fp32, wall clock, all host threads."""
    import torch
    cuda = device.startswith("cuda")
    if not cuda:
        torch.set_num_threads(threads or os.cpu_count() or 1)
    student, teacher = build_reference_models(seed, drop_path)
    student, teacher = student.to(device).train(), teacher.to(device).eval()
    for p in teacher.parameters():
        p.requires_grad_(False)
    lr = 1.5e-4 * (world_batch or batch) / 256                                          # run_stage1.py:796-798
    groups = param_groups(student, 0.05, student.no_weight_decay())
    optimizer = torch.optim.AdamW(groups, lr=lr, betas=(0.9, 0.95), eps=1e-8, **(dict(fused=True) if cuda else {}))
    g = torch.Generator().manual_seed(1000 + seed)
    vids = [torch.randn(batch, 3, 8, 224, 224, generator=g) for _ in range(2)]
    if cuda:
        vids = [v.to(device) for v in vids]
    ac = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if cuda else None
    loss = float("nan")
    for i in range(warmup):
        loss, _ = stage1_step_reference(student, teacher, optimizer, vids[i % 2], 0.8, ac)
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.perf_counter()
    for i in range(steps):
        loss, _ = stage1_step_reference(student, teacher, optimizer, vids[i % 2], 0.8, ac)
    if cuda:
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / max(1, steps)
        peak = torch.cuda.max_memory_allocated() / 2 ** 30
    else:
        ms = (time.perf_counter() - t0) * 1e3 / max(1, steps)
        peak = None
    return dict(value=batch / (ms * 1e-3), unit="clips/s", ms_per_step=ms, loss=loss, batch=batch, steps=steps, warmup=warmup,
                peak_mem_gib=peak, torch=torch.__version__)


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--drop-path", type=float, default=0.1)
    a = ap.parse_args()
    print(json.dumps(run_stage1(a.device, a.batch, a.steps, a.warmup, drop_path=a.drop_path)))
