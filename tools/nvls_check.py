"""Parity + timing of the fused NVLS optimizer step (csrc/ddp_nvls.cu) against NCCL all-reduce -> ub_adamw_dev.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/nvls_check.py [--numel 88080384]

Every rank draws its own gradients; both paths start from identical p / m / v and run 3 steps.  After consolidate() the
sharded fp32 state must equal the replicated reference to 1e-5 relative (the two kernels contract FMAs differently; at
N>2 the switch's summation order also differs from NCCL's), the bf16 shadow must be identical on every rank, and the gradient norm must agree.  Exit code 1 on mismatch.
"""
import argparse
import math
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--numel", dest="n", type=int, default=88_080_384)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    from unite_b200 import ops
    from unite_b200.ddp import NvlsShardedStep, init_distributed_from_env
    rank, local, world = init_distributed_from_env()
    dev = torch.device("cuda", local)
    n, n_decay = args.n, args.n - 131072 - 8 * 5          # ragged shard sizes on purpose
    g = torch.Generator(device=dev).manual_seed(7)
    p0 = torch.randn(n, device=dev, generator=g) * 0.02

    def make():
        arena = types.SimpleNamespace(device=dev, numel=n, n_decay=n_decay, params=p0.clone(), grads=torch.zeros(n, device=dev),
                                      w16=p0.bfloat16(), _params={})
        opt = types.SimpleNamespace(exp_avg=torch.zeros(n, device=dev), exp_avg_sq=torch.zeros(n, device=dev),
                                    gnorm_sq=torch.zeros(1, device=dev), _hyper_dev=torch.zeros(8, device=dev))
        return arena, opt

    ref_a, ref_o = make()
    new_a, new_o = make()
    nv = NvlsShardedStep(new_a, new_o)
    gr = torch.Generator(device=dev).manual_seed(100 + rank)
    ok = True
    b1, b2 = 0.9, 0.95
    for step in range(1, args.steps + 1):
        grads = torch.randn(n, device=dev, generator=gr) * 1e-3
        hyper = torch.tensor([1e-3, 0.05, b1, b2, 1e-8, 1 - b1 ** step, math.sqrt(1 - b2 ** step), 1.0 / world], device=dev)
        ref_o._hyper_dev.copy_(hyper); new_o._hyper_dev.copy_(hyper)
        # reference: NCCL sum, replicated AdamW
        ref_a.grads.copy_(grads)
        dist.all_reduce(ref_a.grads)
        ref_o.gnorm_sq.zero_()
        ops.adamw_dev(ref_a.params, ref_a.grads, ref_o.exp_avg, ref_o.exp_avg_sq, ref_a.w16, n_decay, ref_o._hyper_dev, ref_o.gnorm_sq)
        # fused
        new_a.grads.copy_(grads)
        if step % 2 == 0 and nv.early_push:
            # the overlapped form: prefixes of the decay segment pushed to their owners (copy engines) before the kernel starts
            for frac in (0.21, 0.5, 0.87):
                nv.range_ready(new_a.grads, int(n_decay * frac))
        nv.step_dev()
        torch.cuda.synchronize()
        nv.check()
        gn_ref, gn_new = ref_o.gnorm_sq.item(), new_o.gnorm_sq.item()
        if abs(gn_ref - gn_new) > 1e-4 * abs(gn_ref):
            ok = False
            print(f"[rank {rank}] step {step}: gnorm_sq {gn_new} vs {gn_ref}", flush=True)
        if not torch.equal(new_a.w16.view(torch.int16), ref_a.w16.view(torch.int16)):
            d = (new_a.w16.float() - ref_a.w16.float()).abs()
            bad = int((d > 0).sum())
            # a bf16 rounding boundary can flip where the two kernels' fp32 results differ by an ulp (FMA contraction; at N>2
            # also the summation order inside the switch)
            if bad > n * 1e-5 or float(d.max()) > 1e-3:
                ok = False
            print(f"[rank {rank}] step {step}: shadow differs in {bad} of {n} elements, max {float(d.max()):.3e}", flush=True)
    nv.consolidate()
    torch.cuda.synchronize()
    for name, a, b in (("p", new_a.params, ref_a.params), ("m", new_o.exp_avg, ref_o.exp_avg), ("v", new_o.exp_avg_sq, ref_o.exp_avg_sq)):
        # p: where the summed gradient cancels to ~eps the Adam ratio m / (sqrt(v) + eps) is sensitive to the summation order
        # (NCCL's vs rank order), bounded by a fraction of lr = 1e-3
        same = torch.allclose(a, b, rtol=1e-5, atol=2e-5 if name == "p" else 1e-9)
        if not same:
            ok = False
            print(f"[rank {rank}] {name}: max abs diff {float((a - b).abs().max()):.3e}", flush=True)

    # ---- clip_grad with the fused step (NvlsShardedStep.step_dev_clipped) against all-reduce -> clip coefficient -> AdamW ----
    # both paths continue from the consolidated state above (identical on every rank)
    ref_a.params.copy_(new_a.params); ref_o.exp_avg.copy_(new_o.exp_avg); ref_o.exp_avg_sq.copy_(new_o.exp_avg_sq)
    grads = torch.randn(n, device=dev, generator=gr) * 1e-3
    step = args.steps + 1
    hyper = torch.tensor([1e-3, 0.05, b1, b2, 1e-8, 1 - b1 ** step, math.sqrt(1 - b2 ** step), 1.0 / world], device=dev)
    ref_o._hyper_dev.copy_(hyper); new_o._hyper_dev.copy_(hyper)
    ref_a.grads.copy_(grads)
    dist.all_reduce(ref_a.grads)
    total = ref_a.grads.double().norm().item() / world                       # norm of the averaged gradient
    max_norm = 0.3 * total
    coef = min(1.0, max_norm / (total + 1e-6))
    h2 = hyper.clone(); h2[7] = coef / world
    ops.adamw_dev(ref_a.params, ref_a.grads, ref_o.exp_avg, ref_o.exp_avg_sq, ref_a.w16, n_decay, h2, None)
    new_a.grads.copy_(grads)
    nv.step_dev_clipped(max_norm)
    torch.cuda.synchronize()
    nv.check()
    nv.consolidate()
    torch.cuda.synchronize()
    gn = new_o.gnorm_sq.sqrt().item() / world
    if abs(gn - total) > 1e-4 * total:
        ok = False
        print(f"[rank {rank}] clipped step: reported grad norm {gn} vs {total}", flush=True)
    if not torch.allclose(new_a.params, ref_a.params, rtol=1e-5, atol=2e-5):
        ok = False
        print(f"[rank {rank}] clipped step: p max abs diff {float((new_a.params - ref_a.params).abs().max()):.3e}", flush=True)

    # ---- timing -------------------------------------------------------------------------------------------------
    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def ref_step():
        dist.all_reduce(ref_a.grads)
        ref_o.gnorm_sq.zero_()
        ops.adamw_dev(ref_a.params, ref_a.grads, ref_o.exp_avg, ref_o.exp_avg_sq, ref_a.w16, n_decay, ref_o._hyper_dev, ref_o.gnorm_sq)

    ref_a.grads.zero_(); new_a.grads.zero_()
    t_ref = timeit(ref_step)
    t_new = timeit(nv.step_dev)

    def pushed_step():
        nv.range_ready(new_a.grads, n_decay)
        nv.step_dev()
    t_pre = timeit(pushed_step) if nv.early_push else float("nan")
    nv._pushed_hi = 0
    nv.check()
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print(f"nvls_check world={world} n={n}: {'PARITY OK' if flag.item() == 0 else 'MISMATCH'}; "
              f"NCCL all-reduce + AdamW {t_ref:.3f} ms, fused NVLS step {t_new:.3f} ms, copy-engine pushes + update-only kernel {t_pre:.3f} ms "
              f"(pushes not overlapped here; {nv.pushes} copies)", flush=True)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
