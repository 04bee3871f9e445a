#!/usr/bin/env python
"""Headline benchmark: clips/s of one full stage-1 UMT distillation step (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 32]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

A "step" = teacher forward + attention-guided mask + student forward/backward + gradient all-reduce + AdamW on one
batch of synthetic 8x224^2 clips (ViT-B/16 student, CLIP ViT-B/16 teacher, 80 % mask, per-GPU batch 32, bf16
operands / fp32 accumulate).  Prints ONE JSON line (rank 0).  Keys follow the driver contract; in addition:
  roofline      tensor-bound GEMM kernel family: algorithmic FLOPs / CUDA-event time of every GEMM launch of one
                instrumented step (events on the launching stream), against the measured sustained bf16 peak
  cpu_baseline  the CPU oracle port of the same step on this box's host cores (bounded sample, B=2)
  e2e           the same metric through the public train_one_epoch() API with pinned-host inputs (H2D inside)
--impl reference times the oracle port (the reference's algorithm, fp32, torch CPU) — /root/reference itself
cannot travel to the GPU box and its own step loop does not run anywhere.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_ALG_GFLOP_PER_CLIP = 462.2          # teacher 282.5 + student fwd 60.0 + bwd 119.7 (no padding / recompute)
NOMINAL_BF16_TFLOPS = 2250.0
METRIC = "clips/sec (ViT-B/16 8x224^2 stage-1 step)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d.get("bf16_tflops_sustained", 1341.6), bf16_burst=d.get("bf16_tflops", 1640.4), hbm=d.get("hbm_gbs", 6529.1),
                    source="measured (MEASURED_PEAKS.json, sustained)")
    return dict(bf16=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def build_models(seed=0):
    import torch
    from unite_b200.registry import create_model
    from unite_b200 import modeling_adaptation  # noqa: F401  (registers the factories)
    from unite_b200.clip import clip_b16
    torch.manual_seed(seed)
    # kwargs exactly as run_stage1.py:275-291 passes them with configs/stage1_config.yaml
    student = create_model("adaptation_umt_base_patch16_224", pretrained=False, drop_path_rate=0.0, drop_block_rate=None,
                           use_learnable_pos_emb=False, use_checkpoint=False, checkpoint_num=0, clip_decoder_embed_dim=768,
                           clip_output_dim=512, clip_norm_type="l2", num_frames=8, tubelet_size=1,
                           clip_return_layers=[6, 7, 8, 9, 10, 11], clip_student_return_interval=1, use_cls_token=False)
    teacher = clip_b16(pretrained=False, clip_norm_type="l2", input_resolution=224, return_attn=True,
                       clip_return_layers=[6, 7, 8, 9, 10, 11], clip_return_interval=1)
    return student, teacher


def cpu_reference_run(steps, warmup, batch=2, seed=0):
    """The oracle port of the stage-1 step (fp32, torch CPU, all host threads) on a bounded sample: B=2 clips/step."""
    import torch
    from oracle import unite_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    student, teacher = build_models(seed)
    ssd = {k: v.detach() for k, v in student.state_dict().items()}
    tsd = {k: v.detach() for k, v in teacher.state_dict().items()}
    g = torch.Generator().manual_seed(seed)
    videos = torch.randn(batch, 3, 8, 224, 224, generator=g)
    q = torch.empty(batch * 8, 196).exponential_(1, generator=g)
    scfg, tcfg = O.StudentCfg(), O.TeacherCfg()
    for _ in range(warmup):
        O.stage1_step(ssd, tsd, videos, q, scfg, tcfg, mask_ratio=0.8)
    t0 = time.perf_counter()
    for _ in range(steps):
        r = O.stage1_step(ssd, tsd, videos, q, scfg, tcfg, mask_ratio=0.8)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return dict(value=batch / dt, unit="clips/s", cores=cores, kind="port",
                sample=f"oracle/unite_oracle.stage1_step (teacher fwd + mask + student fwd/bwd, fp32 torch CPU), B={batch} clips/step, "
                       f"{warmup} warm-up + {steps} timed steps; optimizer not included",
                ms_per_step=dt * 1e3, loss=float(r["loss"]))


WORKLOAD = ("BASELINE configs[1]: stage-1 UMT masked distillation, ViT-B/16 student (80% CLIP-attn mask, 320 of 1568 tokens) + "
            "frozen CLIP ViT-B/16 teacher, 8x224^2, tubelet 1, K=6 aligned layers, AdamW")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    cb = cpu_reference_run(steps, warm)
    line = dict(metric=METRIC, value=cb["value"], unit="clips/s", n_gpus=args.gpus, steps=steps, warmup=warm, ms_per_step=cb["ms_per_step"],
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD, per_gpu_batch=args.batch, global_batch=args.batch * args.gpus, parallelism=f"dp{args.gpus}",
                            cpu_sample="each step = B=2 clips of that workload (teacher fwd + mask + student fwd/bwd), fp32 torch CPU, all host "
                                       "threads; clips/s does not depend on the batch the sample is cut from", l2_policy="n/a (CPU)"),
                cpu_baseline=dict(kind=cb["kind"], cores=cb["cores"], sample=cb["sample"], value=cb["value"], unit="clips/s"),
                e2e=dict(value=cb["value"], unit="clips/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-op breakdown of one instrumented step here (json)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from unite_b200 import ops
    from unite_b200.ddp import GradSync, DataParallel, init_distributed_from_env
    from unite_b200.engine import Stage1Engine
    from unite_b200.engine_for_pretraining import train_one_epoch
    from unite_b200.synthetic import SyntheticStage1Loader

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path in unite_b200); use --impl reference for the CPU oracle")
    # NCCL prints its version banner on stdout when the communicator is created; the driver wants ONE JSON line there
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        rank, local, world = init_distributed_from_env()
        dev = torch.device("cuda", local)
        if world > 1:
            warm = torch.ones(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)
    B = args.batch
    student, teacher = build_models(seed=0)                       # identical weights on every rank (DDP broadcast equivalent)
    student, teacher = student.to(dev).train(), teacher.to(dev).eval()
    model = DataParallel(student)
    model.grad_sync = GradSync() if world > 1 else None
    use_graph = os.environ.get("UB_NO_GRAPH", "0") != "1"
    eng = Stage1Engine(student, teacher, mask_ratio=0.8, lr=1.5e-4 * B * world / 256, weight_decay=0.05, grad_sync=model.grad_sync,
                       use_graph=use_graph)
    from unite_b200 import engine_for_pretraining as efp
    efp._ENGINES[(id(student), id(teacher))] = eng          # train_one_epoch (e2e) drives the same engine / optimizer state

    # ---- device-resident inputs (value) ------------------------------------------------------------------------
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    dev_batches = [(torch.randn(B, 3, 8, 224, 224, device=dev, generator=g),
                    torch.empty(B * 8, 196, device=dev).exponential_(1, generator=g)) for _ in range(2)]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # untimed warm-up: W steps, and at least 2 eager steps + one CUDA-graph capture per resident input buffer (the engine keeps
    # one graph per buffer pair), so that no capture can fall inside the timed region whatever W the caller asked for
    n_warm = max(args.warmup, 2 + len(dev_batches))
    for i in range(n_warm):
        eng.step(*dev_batches[i % 2])
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = eng.step(*dev_batches[i % 2])
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = ops.LAUNCHES - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = t.item() / args.steps
    value = world * B / (ms_per_step * 1e-3)
    loss_value = loss.item()

    # ---- end to end through the public API: pinned host batches, H2D inside the timed region, loss read back ------
    loader = SyntheticStage1Loader(B, steps=args.steps, seed=0, rank=rank, n_distinct=2)
    # warm-up through the SAME pinned buffers, staging allocations and graphs as the timed loop: 2 eager steps + one graph
    # capture per slot of the 3-slot staging ring must all happen before the timed region (first-touch effects of a
    # fresh process / box otherwise land inside the timed region: 1516 vs 1704 clips/s measured back to back)
    warm_loader = SyntheticStage1Loader(B, steps=max(6, args.warmup), seed=0, rank=rank, n_distinct=2)
    warm_loader.batches = loader.batches

    class _Args:
        log_freq = 1          # read the loss back every step: the D2H of the step's result is inside the timed region
        use_cuda_graph = use_graph
    train_one_epoch(model, warm_loader, None, eng.optimizer, dev, 0, None, teacher_model=teacher, mask_type="attention",
                    mask_ratio=0.8, args=_Args)
    sync_all()
    e0.record()
    stats = train_one_epoch(model, loader, None, eng.optimizer, dev, 0, None, teacher_model=teacher, mask_type="attention", mask_ratio=0.8,
                            args=_Args)
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item() / args.steps
    e2e_value = world * B / (e2e_ms * 1e-3)
    h2d = B * 3 * 8 * 224 * 224 * 4 + B * 8 * 196 * 4
    d2h = 4

    # ---- the same, fed with DECODED uint8 frames [B,T,H,W,3] (what decord / NVDEC deliver; normalised inside the patchify
    #      kernel): 4x fewer H2D bytes per step.  Reported beside e2e, not instead of it.
    e2e_u8 = None
    if os.environ.get("UB_BENCH_U8", "1") != "0":
        loader8 = SyntheticStage1Loader(B, steps=args.steps, seed=0, rank=rank, n_distinct=2, uint8=True)
        warm8 = SyntheticStage1Loader(B, steps=max(6, args.warmup), seed=0, rank=rank, n_distinct=1, uint8=True)
        warm8.batches = loader8.batches
        train_one_epoch(model, warm8, None, eng.optimizer, dev, 0, None, teacher_model=teacher, mask_type="attention", mask_ratio=0.8,
                        args=_Args)
        sync_all()
        e0.record()
        train_one_epoch(model, loader8, None, eng.optimizer, dev, 0, None, teacher_model=teacher, mask_type="attention", mask_ratio=0.8,
                        args=_Args)
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        u8_ms = t.item() / args.steps
        e2e_u8 = dict(value=round(world * B / (u8_ms * 1e-3), 2), unit="clips/s", ms_per_step=round(u8_ms, 3),
                      h2d_bytes_per_step=B * 8 * 224 * 224 * 3 + B * 8 * 196 * 4, d2h_bytes_per_step=4,
                      input="uint8 frames [B,T,H,W,3], ImageNet-normalised on the device inside the patchify kernel")

    # ---- one instrumented step: per-op CUDA-event durations on the launching stream (roofline of the GEMM kernel) --
    roof, breakdown = None, None
    if rank == 0:
        ops.PROFILE = []
    eng.use_graph = False              # the instrumented step runs eagerly (events around every launch)
    # Eagerly, the host needs ~10 us per launch and the GPU would idle INSIDE the event pairs waiting for the next command
    # (that inflated the per-kernel times of earlier rounds' JSON by 3-5 us per launch).  A spin kernel ahead of the step lets
    # the host queue the whole step first, so every event pair brackets device execution only.
    torch.cuda._sleep(int(0.15 * 1.9e9))
    eng.step(*dev_batches[0])          # every rank runs it (the step contains the all-reduce); only rank 0 records
    sync_all()
    if rank == 0:
        peaks = measured_peaks()
        prof, ops.PROFILE = ops.PROFILE, None
        by = {}
        shapes = {}
        gemm_flops = gemm_ms = 0.0
        n_gemm = 0
        for name, info, a, b in prof:
            d = a.elapsed_time(b)
            if name == "gemm":
                sh = shapes.setdefault(str(info), [0, 0.0, 2.0 * info[0] * info[1] * info[2]])
                sh[0] += 1
                sh[1] += d
            by.setdefault(name, [0, 0.0])
            by[name][0] += 1
            by[name][1] += d
            if name == "gemm":
                M, N, K = info[0], info[1], info[2]
                gemm_flops += 2.0 * M * N * K
                gemm_ms += d
                n_gemm += 1
        total_ms = sum(v[1] for v in by.values())
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
        top = None
        tp = os.path.join(ROOT, "profiles", "gemm_ncu_r01c.json")
        if os.path.exists(tp):
            t = json.load(open(tp))
            key = str((t["shape_M_N_K"][0], t["shape_M_N_K"][1], t["shape_M_N_K"][2], False, False, False))
            if key in shapes:
                us = shapes[key][1] / shapes[key][0] * 1e3
                top = dict(shape_M_N_K=t["shape_M_N_K"], launches_per_step=shapes[key][0], avg_launch_us=round(us, 1),
                           achieved=round(t["algorithmic_flops"] / (us * 1e-6) / 1e12, 1), unit="TFLOP/s",
                           frac=round(t["algorithmic_flops"] / (us * 1e-6) / 1e12 / peaks["bf16"], 4),
                           traffic=t["dram_bytes_read"] + t["dram_bytes_write"], algorithmic_bytes=t["algorithmic_bytes"],
                           traffic_source=t["source"])
        roof = dict(bound="tensor", kernel="ub::gemm_kernel (tcgen05/TMEM/TMA bf16 GEMM family, all launches of one step)",
                    achieved=round(achieved, 1), peak=peaks["bf16"], unit="TFLOP/s", frac=round(achieved / peaks["bf16"], 4),
                    traffic=None, top_shape=top, peak_source=peaks["source"], launches_per_step=n_gemm,
                    alg_flops_per_launch=gemm_flops / n_gemm, avg_launch_us=round(gemm_ms / n_gemm * 1e3, 2),
                    share_of_step=round(gemm_ms / total_ms, 4))
        breakdown = {k: dict(launches=v[0], ms=round(v[1], 3)) for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])}
        if args.profile_out:
            gs = {k: dict(launches=v[0], ms=round(v[1], 3), us_each=round(v[1] / v[0] * 1e3, 1), tflops=round(v[2] * v[0] / (v[1] * 1e-3) / 1e12, 1))
                  for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][1])}
            json.dump(dict(ms_per_step_sum_of_ops=total_ms, ops=breakdown, gemm_tflops=achieved,
                           gemm_shapes_M_N_K_aT_bT_fp32out=gs), open(args.profile_out, "w"), indent=1)

    def finish():
        # NCCL communicators referenced by captured CUDA graphs can hang in destroy_process_group(): leave together, hard
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        finish()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_run(steps=2, warmup=1)
        cpu = dict(value=round(cb["value"], 3), unit="clips/s", cores=cb["cores"], kind=cb["kind"], sample=cb["sample"])
    per_gpu = value / world
    line = dict(
        metric=METRIC, value=round(value, 2), unit="clips/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=round(ms_per_step, 3), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
        config=dict(workload=WORKLOAD, per_gpu_batch=B, global_batch=B * world, parallelism=f"dp{world}", cuda_graph=use_graph,
                    warmup_steps_run=n_warm, ddp=("fused NVLink step (ub_adamw_nvls)" if getattr(eng, "nvls", None) is not None else ("NCCL all-reduce" if world > 1 else "n/a")),
                    l2_policy="inputs_exceed_l2 (154 MB clip batch + "
                    "multi-GB activations per step vs 126 MB L2; two input batches alternate)", init="random (reference initialisers), seed 0"),
        clocks=clocks,
        e2e=dict(value=round(e2e_value, 2), unit="clips/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=round(e2e_ms, 3),
                 api="unite_b200.engine_for_pretraining.train_one_epoch (reference run_stage1.py:294 signature), pinned host batches"),
        e2e_uint8_frames=e2e_u8, gpu_launches=launches, roofline=roof, cpu_baseline=cpu,
        mfu=dict(alg_gflop_per_clip=F_ALG_GFLOP_PER_CLIP, tflops_per_gpu=round(per_gpu * F_ALG_GFLOP_PER_CLIP / 1e3, 1),
                 of_nominal_2250=round(per_gpu * F_ALG_GFLOP_PER_CLIP / 1e3 / NOMINAL_BF16_TFLOPS, 4),
                 of_measured_sustained=round(per_gpu * F_ALG_GFLOP_PER_CLIP / 1e3 / measured_peaks()["bf16"], 4)),
        loss=round(loss_value, 5), e2e_loss=round(stats["loss"], 5), op_breakdown_ms=breakdown)
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    main()
