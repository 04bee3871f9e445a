"""Stage-2 supervised fine-tuning step with the reference's `train_one_epoch` signature
(src/engines/engine_for_finetuning.py:48-54; non-DeepSpeed branch :116-127, no mixup).

    loss = CE(model(samples), targets) / update_freq          engine_for_finetuning.py:37-40, :119
    backward every micro-step, optimizer every update_freq     :120-126
The loss and its gradient come from one fused kernel (ub_softmax_ce); the model's backward is the explicit one in
finetune_core.py; gradients accumulate in the parameter arena across micro-steps like `.grad` does in torch.
"""
import math
import sys
from typing import Iterable, Optional

import torch

from . import ops
from .engine import FusedAdamW

_OPT = {}


def train_class_batch(model, samples, target, criterion=None):
    """engine_for_finetuning.py:37-40 on the fused path: returns (device loss [1], logits) with gradients already
    accumulated in the arena (scale folded in by the caller through `loss_scale`)."""
    raise NotImplementedError("use train_one_epoch / finetune_step; the fused path computes loss and gradient together")


def finetune_step(model, samples, targets, loss_acc, scale=1.0):
    """One micro-step: forward, CE, backward.  Adds scale * mean CE to loss_acc (fp32 [1]).  Returns logits."""
    net = model.module if hasattr(model, "module") else model
    core = net.core()
    dp = None
    if net.training:
        from .modeling_finetune import drop_path_factors
        dp = drop_path_factors(net.drop_path_rates, samples.shape[0], samples.device)
    logits, state = core.run_forward(samples, dp, save=True)
    B = logits.shape[0]
    dlogits = torch.empty_like(logits)
    ops.softmax_ce(logits, targets.to(torch.int32), None, scale / B, loss_acc, dlogits)
    core.run_backward(state, dlogits)
    return logits


def train_one_epoch(model: torch.nn.Module, criterion=None, data_loader: Iterable = (), optimizer=None, device=None, epoch: int = 0,
                    loss_scaler=None, max_norm: float = 0, model_ema=None, mixup_fn=None, log_writer=None, start_steps=None,
                    lr_schedule_values=None, wd_schedule_values=None, num_training_steps_per_epoch=None, update_freq=None,
                    num_epochs=None, train_head_only=False, wandb_run=None, args=None):
    if mixup_fn is not None or model_ema is not None or train_head_only:
        raise NotImplementedError("mixup / EMA / head-only training are off in the shipped stage-2 config")
    if max_norm:
        raise NotImplementedError("clip_grad is null in every shipped config")
    model.train(True)
    net = model.module if hasattr(model, "module") else model
    core = net.core()
    dev = core.arena.device
    update_freq = update_freq or 1
    start_steps = start_steps or 0
    if optimizer is None or not hasattr(optimizer, "arena"):
        optimizer = _OPT.setdefault(id(net), FusedAdamW(core.arena, lr=1e-3, weight_decay=0.05, betas=(0.9, 0.999)))
    gs = getattr(model, "grad_sync", None)
    loss_sum = torch.zeros(1, device=dev)
    loss_acc = torch.zeros(1, device=dev)
    correct = torch.zeros(1, device=dev)
    seen = 0
    optimizer.zero_grad()
    for data_iter_step, batch in enumerate(data_loader):
        samples, targets = batch[0], batch[1]
        step = data_iter_step // update_freq
        it = start_steps + step
        if data_iter_step % update_freq == 0:
            for group in optimizer.param_groups:                                   # engine_for_finetuning.py:76-81
                if lr_schedule_values is not None:
                    group["lr"] = lr_schedule_values[min(it, len(lr_schedule_values) - 1)] * group.get("lr_scale", 1.0)
                if wd_schedule_values is not None and group["weight_decay"] > 0:
                    group["weight_decay"] = wd_schedule_values[min(it, len(wd_schedule_values) - 1)]
        samples = samples.to(dev, non_blocking=True)
        targets = targets.to(dev, non_blocking=True)
        loss_acc.zero_()
        logits = finetune_step(model, samples, targets, loss_acc, scale=1.0 / update_freq)
        loss_sum += loss_acc * update_freq
        correct += (logits.argmax(-1) == targets).sum()
        seen += samples.shape[0]
        if (data_iter_step + 1) % update_freq == 0:
            scale = gs.all_reduce(core.arena.grads) if gs is not None else 1.0
            optimizer.step(grad_scale=scale)
            optimizer.zero_grad()
    n = max(1, data_iter_step + 1 if seen else 1)
    loss_avg = (loss_sum / n).item()
    if not math.isfinite(loss_avg):
        print("Loss is {}, stopping training".format(loss_avg))
        sys.exit(1)
    lrs = [g["lr"] for g in optimizer.param_groups]
    return {"loss": loss_avg, "class_acc": (correct / max(seen, 1)).item(), "loss_scale": 1.0, "lr": max(lrs), "min_lr": min(lrs),
            "grad_norm": optimizer.grad_norm().item()}
