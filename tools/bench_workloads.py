"""bench.py --workload stage2 | stage3 | vitl: the other BASELINE.json configs as their own JSON lines (same timing rules as
the headline: W >= 3 warm-up steps, CUDA events, barrier + synchronize on both sides, max over ranks, inputs larger than L2).

  stage2  configs[0] shape on the GPU: ViT-B/16 all 1568 tokens, supervised CE fwd/bwd + layer-decay AdamW (0.65), B=32/GPU   F_alg 1074.7 GF/clip
  stage3  configs[3]: collaborative self-training step, B_s = B_t per GPU, dual-view target batch, k=2 committee              F_alg 2229 GF/pair
  vitl    configs[4]: stage-1 step with the ViT-L/16 student, 16x224^2, tubelet 2, teacher kernel_size 2 (1568 -> 320 tokens)  F_alg 902.4 GF/clip
"""
import json
import os
import sys

import torch
import torch.distributed as dist


def _timed(step, n_warm, n_steps, world, dev):
    for _ in range(n_warm):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() / n_steps, out


def run_workload(args, rank, local, world, dev, ClockSampler, measured_peaks, F_ALG, NOMINAL):
    from unite_b200 import ops
    from unite_b200.ddp import GradSync
    import bench
    B, wl = args.batch, args.workload
    gs = GradSync() if world > 1 else None
    g = torch.Generator().manual_seed(1000 + rank)
    unit = "clips/s"
    extra = {}
    use_graph = os.environ.get("UB_NO_GRAPH", "0") != "1"
    if wl == "stage2":
        from unite_b200.registry import create_model
        from unite_b200 import modeling_finetune  # noqa: F401
        from unite_b200.optim_factory import LayerDecayValueAssigner, create_optimizer
        torch.manual_seed(0)
        # run_stage2.py:328-346 with configs/stage2_config.yaml
        model = create_model("vit_base_patch16_224", pretrained=False, num_classes=12, all_frames=8, tubelet_size=1, drop_rate=0.0,
                             drop_path_rate=bench.DROP_PATH, attn_drop_rate=0.0, drop_block_rate=None, use_mean_pooling=True,
                             init_scale=0.001, fc_drop_rate=0.0, use_checkpoint=False, checkpoint_num=0).to(dev).train()
        L = model.get_num_layers()
        asg = LayerDecayValueAssigner([0.65 ** (L + 1 - i) for i in range(L + 2)])

        class A:
            opt, lr, weight_decay, opt_betas, opt_eps = "adamw", 1e-3, 0.05, (0.9, 0.999), 1e-8
        opt = create_optimizer(A, model, get_num_layer=asg.get_layer_id, get_layer_scale=asg.get_scale)
        if gs is not None:
            gs.arena = model.core().arena
        batches = [(torch.randn(B, 3, 8, 224, 224, generator=g).to(dev), torch.randint(0, 12, (B,), generator=g).to(dev)) for _ in range(2)]
        from unite_b200.engine_for_finetuning import Stage2Engine
        eng = Stage2Engine(model, opt, grad_sync=gs, use_graph=use_graph)
        i = [0]

        def step():
            v, y = batches[i[0] % 2]
            i[0] += 1
            return eng.step(v, y)
        metric = "clips/sec (ViT-B/16 8x224^2 stage-2 supervised step, 1568 tokens)"
        workload = ("BASELINE configs[0] shape on the GPU: stage-2 supervised fine-tune, ViT-B/16 on all 1568 tokens, CE, drop_path 0.1, "
                    "AdamW with layer decay 0.65 (14 x 2 groups)")
        extra["optimizer_groups"] = len(opt.param_groups)
        extra["cuda_graph"] = use_graph
    elif wl == "stage3":
        from unite_b200.engine_stage3 import Stage3Engine
        student, teacher = bench.build_models(seed=0)
        student, teacher = student.to(dev).train(), teacher.to(dev).eval()
        gw = torch.Generator().manual_seed(5)
        eng = Stage3Engine(student, teacher, torch.randn(12, 768, generator=gw) * 0.5, torch.zeros(12), torch.randn(12, 512, generator=gw),
                           mask_ratio=0.8, k=2, grad_sync=gs, use_graph=use_graph)
        if gs is not None:
            gs.arena = eng.core.arena
        mk = lambda: torch.randn(B, 3, 8, 224, 224, generator=g)
        batches = []
        for _ in range(2):
            vt = mk()
            batches.append((mk().to(dev), torch.randint(0, 12, (B,), generator=g).to(dev), vt.to(dev), (vt + 0.1 * mk()).to(dev)))
        i = [0]

        def step():
            b = batches[i[0] % 2]
            i[0] += 1
            return eng.step(*b)
        metric, unit = "clip pairs/sec (stage-3 collaborative self-training step, ViT-B/16 8x224^2)", "pairs/s"
        workload = ("BASELINE configs[3]: stage-3 step per (source, target) pair — teacher attention on vid_aug + zero-shot CLS on vid, source "
                    "full pass (grad), target full pass, k=2 masked committee (last member trains), MatchOrConf fusion, AdamW; drop_path 0.1")
        extra["per_gpu_pairs"] = B
        extra["cuda_graph"] = use_graph
        extra["ddp"] = "fused NVLink step" if getattr(eng, "nvls", None) is not None else ("NCCL all-reduce" if world > 1 else "n/a")
    else:
        from unite_b200.engine import Stage1Engine
        student, teacher = bench.build_models(seed=0, large=True)
        student, teacher = student.to(dev).train(), teacher.to(dev).eval()
        eng = Stage1Engine(student, teacher, mask_ratio=0.8, lr=1.5e-4 * B * world / 256, grad_sync=gs, use_graph=use_graph)
        extra["cuda_graph"] = use_graph
        if gs is not None:
            gs.arena = eng.core.arena
        batches = [(v.to(dev), q.to(dev)) for v, q in bench.host_batches(B, rank, frames=16, tokens_per_frame=2)]
        i = [0]

        def step():
            b = batches[i[0] % 2]
            i[0] += 1
            return eng.step(*b)
        metric = "clips/sec (ViT-L/16 student 16x224^2 stage-1 step)"
        workload = ("BASELINE configs[4]: stage-1 UMT step, ViT-L/16 student (24 layers, D=1024; 16 frames, tubelet 2: 1568 tokens, 320 visible, "
                    "drop_path 0.1) + CLIP ViT-B/16 teacher with kernel_size 2")
        extra["ddp"] = "fused NVLink step" if getattr(eng, "nvls", None) is not None else ("NCCL all-reduce" if world > 1 else "n/a")
    n_warm = max(args.warmup, 5)
    sampler = ClockSampler(local)
    l0 = ops.LAUNCHES
    # warm-up outside the sampler, then the timed region under it
    ms0, _ = _timed(step, n_warm, 1, world, dev)
    if rank == 0:
        sampler.start()
    l0 = ops.LAUNCHES
    ms, loss = _timed(step, 0, args.steps, world, dev)
    launches = ops.LAUNCHES - l0
    clocks = sampler.stop() if rank == 0 else None
    value = world * B / (ms * 1e-3)
    if rank == 0:
        f = F_ALG[wl]
        per_gpu = value / world
        line = dict(metric=metric, value=round(value, 2), unit=unit, n_gpus=world, steps=args.steps, warmup=n_warm, ms_per_step=round(ms, 3),
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                    config=dict(workload=workload, per_gpu_batch=B, global_batch=B * world, parallelism=f"dp{world}",
                                l2_policy="inputs_exceed_l2 (two resident input batches alternate; activations are multi-GB)", **extra),
                    clocks=clocks, gpu_launches=launches,
                    mfu=dict(alg_gflop_per_unit=f, tflops_per_gpu=round(per_gpu * f / 1e3, 1), of_nominal_2250=round(per_gpu * f / 1e3 / NOMINAL, 4),
                             of_measured_sustained=round(per_gpu * f / 1e3 / measured_peaks()["bf16"], 4)),
                    loss=round(float(loss.item()), 5))
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    if world > 1:
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        os._exit(0)
