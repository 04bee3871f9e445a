"""Host logic of round 2 that needs no GPU: the Philox restatement against the Random123 known-answer vectors, the parameter
groups FusedAdamW derives from the arena (src/optim_factory.py:45-118 semantics), the per-step scalar upload, and the refusal of
foreign optimizers."""
from functools import partial

import numpy as np
import pytest
import torch
import torch.nn as nn


def test_philox4x32_10_known_answer_vectors():
    """Random123 kat_vectors (philox4x32 10 rounds): zero, all-ones and the pi-digits counter / key."""
    from oracle.philox import philox4x32_10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox4x32_10(*[[c] for c in ctr], *key)
        assert tuple(int(v[0]) for v in got) == want


def test_drop_path_factor_statistics_follow_timm_drop_path():
    """floor(keep + u) / keep: 0 with probability p_l, 1 / keep_l otherwise; layer 0 (p = 0) never drops."""
    from oracle.philox import drop_path_factors
    rates = np.linspace(0, 0.3, 6).astype(np.float32)
    f = np.stack([drop_path_factors(rates, 64, 5, s) for s in range(40)])         # [40, 6, 2, 64]
    assert np.all(f[:, 0] == 1.0)
    for l in range(1, 6):
        vals = np.unique(f[:, l])
        assert set(vals.tolist()) <= {0.0, float(np.float32(1.0) / (np.float32(1.0) - rates[l]))}
        assert abs((f[:, l] == 0).mean() - rates[l]) < 0.03
    assert not np.array_equal(f[0], f[1])


def _tiny_vit():
    from unite_b200.modeling_finetune import VisionTransformer
    return VisionTransformer(img_size=32, patch_size=16, embed_dim=64, depth=3, num_heads=1, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), num_classes=5, all_frames=2, tubelet_size=1, init_scale=0.001)


def _arena(model):
    from unite_b200.arena import ParamArena
    from unite_b200.modeling_adaptation import no_decay_rule, student_order_key
    return ParamArena(model, "cpu", student_order_key(len(model.blocks)), no_decay_rule(model.no_weight_decay()),
                      gap_after=lambda n, p: p.numel() if n.endswith("attn.q_bias") else 0)


def test_fused_adamw_groups_follow_get_parameter_groups():
    from unite_b200.engine import FusedAdamW, LayerDecayValueAssigner
    vit = _tiny_vit()
    vit.blocks[1].attn.proj.weight.requires_grad_(False)
    arena = _arena(vit)
    plain = FusedAdamW(arena, lr=1e-3, weight_decay=0.05)
    assert [g["name"] for g in plain.param_groups] == ["decay", "no_decay"] and not plain.plain_two_groups      # a frozen run splits "decay"
    L = 3
    asg = LayerDecayValueAssigner([0.65 ** (L + 1 - i) for i in range(L + 2)])
    opt = FusedAdamW(arena, lr=1e-3, weight_decay=0.05, get_num_layer=asg.get_layer_id, get_layer_scale=asg.get_scale)
    # restated get_parameter_groups (optim_factory.py:76-118)
    want = {}
    for n, p in vit.named_parameters():
        if not p.requires_grad:
            continue
        nd = p.ndim == 1 or n.endswith(".bias") or n in vit.no_weight_decay()
        lid = asg.get_layer_id(n)
        want.setdefault("layer_%d_%s" % (lid, "no_decay" if nd else "decay"), []).append(n)
    got = {g["name"]: g for g in opt.param_groups}
    assert set(got) == set(want)
    ids = {id(p): n for n, p in vit.named_parameters()}
    for name, names in want.items():
        assert sorted(ids[id(p)] for p in got[name]["params"]) == sorted(names)
        lid = int(name.split("_")[1])
        assert got[name]["lr_scale"] == asg.get_scale(lid) and got[name]["lr"] == 1e-3 * asg.get_scale(lid)
        assert got[name]["weight_decay"] == (0.0 if name.endswith("no_decay") else 0.05)
    # the runs tile the arena in order, frozen runs are marked, every trainable element belongs to its group's run
    ends = (opt._seg_end4 * 4).tolist()
    assert ends == sorted(ends) and ends[-1] == arena.numel and len(ends) == opt.n_seg <= 128
    starts = [0] + ends[:-1]
    for n, p in vit.named_parameters():
        off, k = arena.offsets[n]
        s = next(i for i, (a, b) in enumerate(zip(starts, ends)) if a <= off < b)
        assert off + k <= ends[s]
        gi = opt._seg_group[s]
        assert (gi == -1) == (not p.requires_grad)
        if gi >= 0:
            assert any(q is p for q in opt.param_groups[gi]["params"])
    # per-step upload: lr / wd of every run, frozen runs flagged with wd = -1
    for g in opt.param_groups:
        g["lr"] = 0.5 * g["lr_scale"]
    opt.prepare_step(grad_scale=0.25)
    h = opt._hyper_dev
    S = opt.n_seg
    assert h.numel() == 8 + 2 * S and abs(h[7].item() - 0.25) < 1e-7 and abs(h[5].item() - (1 - 0.9)) < 1e-6
    for s, gi in enumerate(opt._seg_group):
        if gi < 0:
            assert h[8 + S + s].item() == -1.0
        else:
            assert abs(h[8 + s].item() - opt.param_groups[gi]["lr"]) < 1e-7
            assert abs(h[8 + S + s].item() - opt.param_groups[gi]["weight_decay"]) < 1e-7
    sd = opt.state_dict()
    assert "params" not in sd["param_groups"][0]
    opt.load_state_dict(sd)


def test_get_num_layer_for_vit_matches_the_reference_table():
    """src/optim_factory.py:45-63, including its prefix blindness (an `encoder.`-prefixed name falls into the last group)."""
    from unite_b200.engine import get_num_layer_for_vit
    n = 14
    assert get_num_layer_for_vit("pos_embed", n) == 0 and get_num_layer_for_vit("patch_embed.proj.weight", n) == 0
    assert get_num_layer_for_vit("blocks.0.attn.qkv.weight", n) == 1 and get_num_layer_for_vit("blocks.11.mlp.fc2.bias", n) == 12
    assert get_num_layer_for_vit("transformer.resblocks.4.ln_1.weight", n) == 5 and get_num_layer_for_vit("conv1.weight", n) == 0
    assert get_num_layer_for_vit("fc_norm.weight", n) == 13 and get_num_layer_for_vit("head.weight", n) == 13
    assert get_num_layer_for_vit("encoder.blocks.3.norm1.weight", n) == 13 and get_num_layer_for_vit("rel_pos_bias.x", n) == 13


def test_foreign_optimizers_are_refused_not_ignored():
    from unite_b200.engine import FusedAdamW, require_fused_optimizer
    vit = _tiny_vit()
    arena = _arena(vit)
    assert require_fused_optimizer(None, arena, "x") is None
    with pytest.raises(TypeError, match="FusedAdamW"):
        require_fused_optimizer(torch.optim.AdamW(vit.parameters(), lr=1e-3), arena, "train_one_epoch")
    other = FusedAdamW(_arena(_tiny_vit()))
    with pytest.raises(ValueError, match="different parameter arena"):
        require_fused_optimizer(other, arena, "train_one_epoch")
    mine = FusedAdamW(arena)
    assert require_fused_optimizer(mine, arena, "x") is mine and mine.plain_two_groups


def test_bench_parity_fixture_matches_the_current_seed0_initialisation():
    """bench.py refuses to print when its first eager B=32 step disagrees with tests/golden/bench_b32_check.json; a changed
    initialisation order would make that fixture stale — caught here, on the CPU, before any GPU time is spent."""
    import json
    import os
    import bench
    fix = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bench_b32_check.json")))
    student, _ = bench.build_models(seed=0)
    assert bench.weights_digest(student.state_dict()) == fix["weights_sha16"], "rerun oracle/make_bench_fixture.py"
    assert fix["visible_tokens"] == 32 * 320 and len(fix["mask_hex"]) == 32 * 1568 // 8 * 2 and 1.9 < fix["loss"] < 2.1
