"""Data-parallel gradient path: bucketed all-reduce of the flat gradient arena, overlapped with backward.

Replaces torch.nn.parallel.DistributedDataParallel at run_stage1.py:809 / run_stage2.py:641 / run_stage3.py:1246
(25 MiB autograd-hook buckets of per-tensor fp32 grads).  Here the gradients already live in ONE buffer ordered
by backward completion (arena.py), so a "bucket" is simply a contiguous range: as soon as backward has finished
a block, its range is all-reduced (SUM) on a side stream while the next block's backward runs; the 1/world
scale is folded into the fused AdamW kernel (grad_scale) instead of a separate division pass.
One process per GPU (torchrun env contract, src/utils.py:532-548); NCCL over NVLink/NVSwitch on GPUs, gloo on
CPU tensors (used by the world_size-2 tests).  The path has exactly this one collective per step.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


import os as _os

# UB_DDP_OVERLAP=0: one all-reduce of the whole arena after backward instead of per-block ranges overlapped with it
_OVERLAP = _os.environ.get("UB_DDP_OVERLAP", "1") != "0"


class GradSync:
    def __init__(self, arena=None, process_group=None):
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.arena = arena
        self._stream = None
        self._pending = []
        self._done_hi_decay = 0      # decay-segment prefix already reduced this step
        self.calls = 0

    # ---- plain (non-overlapped) form ------------------------------------------------------------
    def all_reduce(self, flat: torch.Tensor) -> float:
        """SUM all-reduce of whatever has not been reduced yet; returns the scale (1/world) the optimizer must apply."""
        if self.world > 1:
            lo = self._done_hi_decay
            self._launch(flat[lo:])
            self.finish()
        self._done_hi_decay = 0
        return 1.0 / self.world

    # ---- overlapped form: called by backward as ranges become final ------------------------------
    def range_ready(self, flat: torch.Tensor, hi: int):
        """The decay-segment prefix [done, hi) of `flat` holds final gradients: reduce it now, asynchronously."""
        if self.world == 1 or hi <= self._done_hi_decay or not _OVERLAP:
            return
        self._launch(flat[self._done_hi_decay:hi])
        self._done_hi_decay = hi

    def _launch(self, t: torch.Tensor):
        if t.numel() == 0:
            return
        self.calls += 1
        if t.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=t.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(t.device))
            self._stream.wait_event(ev)
            with torch.cuda.stream(self._stream):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
        else:
            self._pending.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def finish(self):
        for w in self._pending:
            w.wait()
        self._pending = []
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)


class DataParallel(torch.nn.Module):
    """Thin wrapper with DDP's surface (`.module`, forward passthrough); gradient averaging is done by GradSync
    inside the engines, not by autograd hooks."""

    def __init__(self, module, device_ids=None, find_unused_parameters=False, process_group=None):
        super().__init__()
        self.module = module
        self.grad_sync = GradSync(process_group=process_group)

    def forward(self, *a, **k):
        return self.module(*a, **k)


def init_distributed_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """torchrun env:// contract (src/utils.py:532-548).  Returns (rank, local_rank, world_size)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # SM partition between compute and the overlapped all-reduce: NCCL gets `reserve` CTAs (NCCL_MAX_CTAS), every persistent
    # grid of the library is sized for the remaining SMs (ub_set_sm_limit).  Without it the collective's CTAs land on SMs a
    # persistent GEMM assumed it owned and that GEMM's static schedule waits for them (measured: 17.21 -> 18.07 ms per step
    # from 1 to 2 GPUs).  The experiment did not pay (see below); UB_DDP_RESERVED_SMS=n turns the partition on.
    reserve = int(os.environ.get("UB_DDP_RESERVED_SMS", "0"))   # measured at 2 GPUs: 18.88 ms without, 19.11 (4) / 19.41 (8) with -> off
    if world > 1 and reserve > 0 and torch.cuda.is_available():
        os.environ.setdefault("NCCL_MAX_CTAS", str(reserve))
        os.environ.setdefault("NCCL_MIN_CTAS", str(min(reserve, 4)))
        from . import _cabi
        _cabi.lib.ub_set_sm_limit(0)
        _cabi.lib.ub_set_sm_limit(_cabi.lib.ub_sm_count() - reserve)
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, init_method="env://", world_size=world, rank=rank)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world
