"""create_optimizer / LayerDecayValueAssigner with the reference's signatures (src/optim_factory.py:66-74, 120-175) on top of the
fused arena optimizer.

    assigner = LayerDecayValueAssigner([layer_decay ** (L + 1 - i) for i in range(L + 2)])          run_stage2.py:616-617
    optimizer = create_optimizer(args, model_without_ddp, skip_list=model.no_weight_decay(),
                                 get_num_layer=assigner.get_layer_id, get_layer_scale=assigner.get_scale)   run_stage2.py:651-654

Returns a FusedAdamW whose `param_groups` are the groups get_parameter_groups (optim_factory.py:76-118) would build — "decay" /
"no_decay" or "layer_%d_decay" / "layer_%d_no_decay" with their lr_scale, frozen parameters left out — so the training loops'
per-step `param_group["lr"] = schedule[it] * param_group["lr_scale"]` writes carry over unchanged.  Only AdamW exists on the fused
path (every shipped config uses opt: adamw); other optimizer names raise.
"""
from .engine import FusedAdamW, LayerDecayValueAssigner, get_num_layer_for_vit  # noqa: F401


def create_optimizer(args, model, get_num_layer=None, get_layer_scale=None, filter_bias_and_bn=True, skip_list=None):
    opt_lower = str(getattr(args, "opt", "adamw")).lower().split("_")[-1]
    if opt_lower != "adamw":
        raise NotImplementedError(f"opt={args.opt!r}: the fused path implements AdamW (opt: adamw in every shipped config)")
    net = model.module if hasattr(model, "module") else model
    arena = net.core().arena
    weight_decay = float(getattr(args, "weight_decay", 0.05) or 0.0)
    if not (weight_decay and filter_bias_and_bn):
        # optim_factory.py:140-141: all parameters in one group with args.weight_decay (also on biases)
        raise NotImplementedError("weight_decay == 0 / filter_bias_and_bn=False: the arena keeps the decay / no-decay split of "
                                  "get_parameter_groups; pass weight_decay > 0 (shipped configs use 0.05)")
    skip = set(skip_list) if skip_list is not None else (set(net.no_weight_decay()) if hasattr(net, "no_weight_decay") else set())
    model_skip = set(net.no_weight_decay()) if hasattr(net, "no_weight_decay") else set()
    named = dict(net.named_parameters())
    if {n for n in skip if n in named} != {n for n in model_skip if n in named}:
        raise ValueError("skip_list names parameters the model's arena did not place in its no-decay segment")
    betas = tuple(getattr(args, "opt_betas", None) or (0.9, 0.999))
    eps = getattr(args, "opt_eps", None) or 1e-8
    return FusedAdamW(arena, lr=float(args.lr), weight_decay=weight_decay, betas=betas, eps=eps, get_num_layer=get_num_layer,
                      get_layer_scale=get_layer_scale)
