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


class NvlsShardedStep:
    """Gradient reduce-scatter + AdamW + bf16-shadow all-gather in ONE kernel over NVSwitch multicast (csrc/ddp_nvls.cu),
    instead of NCCL all-reduce -> AdamW.  The gradient arena and the bf16 shadow are re-homed into symmetric memory
    (torch.distributed._symmetric_memory: one VMM allocation per rank bound to a multicast object); the kernel reads the
    SUM of all ranks' gradients with multimem.ld_reduce and writes the refreshed shadow to all ranks with multimem.st.
    fp32 master weights / Adam moments of the decay segment are sharded by rank (ZeRO-1): call consolidate() before
    reading them (state_dict, checkpoint); the no-decay segment and the bf16 shadow are always replicated.

    Raises RuntimeError when the box has no multicast support; the caller then stays on the NCCL path (GradSync)."""

    def __init__(self, arena, optimizer, process_group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _cabi
        if not getattr(optimizer, "plain_two_groups", True):
            raise NotImplementedError("ub_adamw_nvls updates the plain [decay | no-decay] arena layout; layer-decay groups or frozen "
                                      "parameters need the NCCL path (UB_DDP_NVLS=0)")
        self.arena, self.opt = arena, optimizer
        self.pg = process_group if process_group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.pg), dist.get_world_size(self.pg)
        dev, n = arena.device, arena.numel
        if n % 8 or arena.n_decay % 8:
            raise RuntimeError("arena size / decay boundary must be multiples of 8 elements")
        slots = int(_cabi.lib.ub_nvls_slots())
        self._grads = symm.empty(n, dtype=torch.float32, device=dev)
        self._w16 = symm.empty(n, dtype=torch.bfloat16, device=dev)
        self._sync = symm.empty(slots + 8, dtype=torch.int32, device=dev)     # [gnorm_sq f32 | 7 pad | flags u32[slots]]
        # push form: slot s of my staging buffer receives rank s's gradients of MY slice of the decay segment
        lo, hi = self.shard_of(arena.n_decay, 0, self.world)
        self._stage = symm.empty(max(8, self.world * (hi - lo)), dtype=torch.float32, device=dev)
        self._hdl = [symm.rendezvous(t, self.pg) for t in (self._grads, self._w16, self._sync, self._stage)]
        if any(int(h.multicast_ptr) == 0 for h in self._hdl[:3]):
            raise RuntimeError("symmetric memory has no multicast mapping on this box (NVLS unavailable)")
        self._grads.copy_(arena.grads)
        self._w16.copy_(arena.w16)
        self._sync.zero_()
        arena.grads, arena.w16 = self._grads, self._w16          # p.grad views are re-pointed by arena.attach_grads()
        for p in arena._params.values():
            p.grad = None
        optimizer.gnorm_sq = self._sync[:1].view(torch.float32)
        optimizer._sharded = self                                  # FusedAdamW.consolidate() / checkpoint.save_model() find it here
        self._epoch = torch.zeros(slots // 2, device=dev, dtype=torch.int32)
        self._err = torch.zeros(1, device=dev, dtype=torch.int32)
        self.g_mc, self.w16_mc, sync_mc = (int(h.multicast_ptr) for h in self._hdl[:3])
        self.gnorm_mc, self.flags_mc = sync_mc, sync_mc + 32
        import ctypes
        self.g_peers = (ctypes.c_void_p * self.world)(*[int(a) for a in self._hdl[0].buffer_ptrs])
        self.w16_peers = (ctypes.c_void_p * self.world)(*[int(a) for a in self._hdl[1].buffer_ptrs])
        self.stage_peers = (ctypes.c_void_p * self.world)(*[int(a) for a in self._hdl[3].buffer_ptrs])
        self.flags = self._sync.data_ptr() + 32
        self.calls = 0
        # ---- gradient pushes overlapped with backward (the role of DDP's bucketed all-reduce, run_stage1.py:809): as soon as a
        # prefix of the decay segment is final, its pieces go to the owners' staging slots as peer-to-peer copies on the copy
        # engines (no SM is taken from the backward GEMMs); the fused kernel then starts at the reduce + AdamW phase
        self.early_push = _os.environ.get("UB_NVLS_EARLY_PUSH", "1") != "0" and _os.environ.get("UB_NVLS_MODE", "push") == "push" \
            and self.world in (2, 4, 8)
        self._push_stream = torch.cuda.Stream(device=dev) if self.early_push else None
        self._pushed_hi = 0
        self._shard = self.shard_of(arena.n_decay, 0, self.world)[1]            # elements per (full) shard
        self._stage_of = [self._hdl[3].get_buffer(r, (self._stage.numel(),), torch.float32) if r != self.rank else None
                          for r in range(self.world)] if self.early_push else None
        self.pushes = 0
        torch.cuda.synchronize(dev)
        dist.barrier(self.pg)                                      # every rank's buffers are initialised before anyone's kernel
        torch.cuda.synchronize(dev)

    @staticmethod
    def shard_of(n_decay: int, rank: int, world: int) -> Tuple[int, int]:
        """[lo, hi) elements of the decay segment owned by `rank`: contiguous slices of ceil(n_decay/8 / world) 8-element
        units (the arithmetic of adamw_nvls_kernel)."""
        n8 = n_decay // 8
        shard = (n8 + world - 1) // world
        lo = min(rank * shard, n8)
        return lo * 8, min(lo + shard, n8) * 8

    def shard_range(self, rank=None):
        return self.shard_of(self.arena.n_decay, self.rank if rank is None else rank, self.world)

    def range_ready(self, flat: torch.Tensor, hi: int):
        """GradSync.range_ready's twin for the fused step: the decay-segment prefix [pushed, hi) of the gradient arena is final
        (backward has finished those blocks) — push its pieces to the ranks that own them, asynchronously, on the copy engines."""
        if not self.early_push:
            return
        hi = min(hi, self.arena.n_decay)
        if hi <= self._pushed_hi:
            if hi < self._pushed_hi:
                self._pushed_hi = 0                   # a new backward without an optimizer step in between: start over
            else:
                return
        lo, self._pushed_hi = self._pushed_hi, hi
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.arena.device))
        self._push_stream.wait_event(ev)
        g = self.arena.grads
        with torch.cuda.stream(self._push_stream):
            for j in range(1, self.world):
                q = (self.rank + j) % self.world          # staggered: the N ranks address N different peers at any time
                qlo, qhi = self.shard_range(q)
                a, b = max(lo, qlo), min(hi, qhi)
                if b > a:
                    off = self.rank * self._shard + (a - qlo)
                    self._stage_of[q][off:off + (b - a)].copy_(g[a:b], non_blocking=True)
                    self.pushes += 1

    def step_dev(self):
        """Device side of the step (graph-capturable); the optimizer's prepare_step() uploaded hyper[] with grad_scale = 1/world."""
        from . import ops
        a, o = self.arena, self.opt
        prepushed = False
        if self.early_push and self._pushed_hi > 0:
            self.range_ready(a.grads, a.n_decay)          # whatever backward finished last (the patch embedding)
            torch.cuda.current_stream(a.device).wait_stream(self._push_stream)
            prepushed = True
        self._pushed_hi = 0
        o.gnorm_sq.zero_()
        ops.adamw_nvls(a.params, self.g_mc, o.exp_avg, o.exp_avg_sq, a.w16, self.w16_mc, a.n_decay, self.rank, self.world,
                       o._hyper_dev, self.gnorm_mc, self.flags, self.flags_mc, self._epoch, self._err, self.g_peers, self.w16_peers, self.stage_peers,
                       prepushed=prepushed)
        self.calls += 1

    def step_dev_clipped(self, max_norm: float):
        """clip_grad with the fused step (loss_scaler(..., clip_grad=max_norm), utils.py:613-615): the clip coefficient needs the
        norm of the SUMMED gradient before any update, which the one-pass kernel cannot know.  So the sum is formed first (one
        NCCL all-reduce of the symmetric gradient arena, in place), its norm measured, and the fused kernel then runs on N
        identical copies: it adds them up again (exactly N x for N a power of two), which the 1/N^2 in grad_scale undoes.
        Slower than the unclipped step by one all-reduce; clip_grad is null in every shipped config."""
        from . import ops
        a, o = self.arena, self.opt
        if self._pushed_hi > 0:                           # pushes of unreduced gradients are of no use here: let them finish, ignore them
            torch.cuda.current_stream(a.device).wait_stream(self._push_stream)
            self._pushed_hi = 0
        dist.all_reduce(a.grads, op=dist.ReduceOp.SUM, group=self.pg)
        if not hasattr(self, "_clip_sq"):
            self._clip_sq = torch.zeros(1, device=a.device, dtype=torch.float32)
            self._hyper_clip = torch.zeros_like(o._hyper_dev)
        self._clip_sq.zero_()
        ops.sumsq(a.grads, self._clip_sq)
        h = self._hyper_clip
        h.copy_(o._hyper_dev)
        norm = self._clip_sq.sqrt() / self.world                              # norm of the rank-averaged gradient
        h[7:8] = (max_norm / (norm + 1e-6)).clamp(max=1.0) / float(self.world * self.world)
        ops.adamw_nvls(a.params, self.g_mc, o.exp_avg, o.exp_avg_sq, a.w16, self.w16_mc, a.n_decay, self.rank, self.world,
                       h, None, self.flags, self.flags_mc, self._epoch, self._err, self.g_peers, self.w16_peers, self.stage_peers)
        o.gnorm_sq.copy_(self._clip_sq)                                        # grad_norm(1/world) then reports the pre-clip norm
        self.calls += 1

    _ERR_NAMES = {1: "entry", 2: "exit", 3: "mid"}

    def raise_if(self, code: int):
        if code:
            raise RuntimeError(f"ub_adamw_nvls (rank {self.rank}): a peer never reached the {self._ERR_NAMES.get(code, code)} barrier "
                               "within UB_NVLS_SPIN_S seconds; the step was declared void and no later step updates anything — "
                               "the run must stop (the other ranks time out on their next step and raise the same error)")

    def check(self):
        """Host-side health check (syncs): raises if a peer ever missed a barrier inside the kernel."""
        self.raise_if(int(self._err.item()))

    def poll_error_async(self, pinned_int32: torch.Tensor):
        """Enqueue a 4-byte D2H copy of the error word (the training loops do it next to their loss read-back, then call
        raise_if() on the value once the copy's event has completed — no extra synchronisation)."""
        pinned_int32.copy_(self._err, non_blocking=True)

    def consolidate(self):
        """Gather the sharded fp32 master weights and Adam moments so that every rank holds all of them."""
        self.check()
        for r in range(self.world):
            lo, hi = self.shard_range(r)
            if hi > lo:
                for t in (self.arena.params, self.opt.exp_avg, self.opt.exp_avg_sq):
                    dist.broadcast(t[lo:hi], src=dist.get_global_rank(self.pg, r), group=self.pg)


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
