"""Oracle (oracle/unite_oracle.py) against the fixtures generated from the REAL reference modules
(oracle/make_golden.py).  CPU only."""
import torch

from oracle import unite_oracle as O
from tests.util import load_golden, oracle_cfgs, seeded_states


def _close(a, b, tol):
    assert (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item())


def test_multinomial_restatement_matches_torch_draws():
    fix = load_golden("mask_sampler.pt")
    order = torch.topk(fix["p"] / fix["q"], fix["p"].shape[1]).indices
    assert torch.equal(order, fix["draw"])                       # torch.multinomial(p, N) with the same generator
    m = O.multinomial_mask(fix["p"], fix["q"], 0.8, clips=2)
    assert m.shape == (2, 8 * 196) and int((~m).sum()) == 16 * 40


def test_greedy_masks_match_reference():
    fix = load_golden("mask_sampler.pt")
    gm = O.greedy_masks(fix["greedy_attn"], 0.8, 2)
    assert torch.equal(gm, fix["greedy_masks"])
    assert int((~gm[0] & ~gm[1]).sum()) == 0 and int((~gm).sum(-1).min()) == 40   # disjoint, 40 visible per row


def test_tiny_stage1_matches_reference_modules():
    fix = load_golden("tiny_stage12.pt")
    scfg, tcfg = oracle_cfgs(fix)
    ssd, tsd, _ = seeded_states(fix)
    r = O.stage1_step(ssd, tsd, fix["videos"], fix["q"], scfg, tcfg, mask_ratio=fix["cfg"]["mask_ratio"])
    _close(r["attn"], fix["attn"], 1e-5)
    assert torch.equal(r["mask"], fix["mask"])
    _close(r["targets"], fix["targets"], 1e-5)
    _close(r["outputs"], fix["outputs"], 1e-5)
    _close(r["loss"], fix["loss"], 1e-6)
    for k, g in fix["grads"].items():
        _close(r["grads"][k], g, 1e-4)
    for k, n in fix["grad_norms"].items():
        assert abs(r["grads"][k].norm().item() - n.item()) <= 1e-4 * max(1.0, n.item())


def test_tiny_stage2_matches_reference_modules():
    fix = load_golden("tiny_stage12.pt")
    scfg, _ = oracle_cfgs(fix)
    _, _, vsd = seeded_states(fix)
    r = O.stage2_step(vsd, fix["videos"], fix["labels"], scfg)
    _close(r["logits"], fix["stage2_logits"], 1e-5)
    _close(r["loss"], fix["stage2_loss"], 1e-6)
    for k, g in fix["stage2_grads"].items():
        _close(r["grads"][k], g, 1e-4)


def test_drop_path_factors_enter_both_branches():
    fix = load_golden("tiny_stage12.pt")
    scfg, _ = oracle_cfgs(fix)
    ssd, _, _ = seeded_states(fix)
    B = fix["videos"].shape[0]
    ones = torch.ones(scfg.depth, 2, B)
    a = O.student_forward(ssd, fix["videos"], fix["mask"], scfg, clip_only=True)
    b = O.student_forward(ssd, fix["videos"], fix["mask"], scfg, clip_only=True, keep_scales=ones)
    assert torch.allclose(a, b)
    ks = ones.clone(); ks[1, 1, 0] = 0.0
    c = O.student_forward(ssd, fix["videos"], fix["mask"], scfg, clip_only=True, keep_scales=ks)
    assert not torch.allclose(a[:, 0], c[:, 0]) and torch.allclose(a[:, 1], c[:, 1])


def test_full_size_anchors():
    fix = load_golden("full_b16_scalars.pt")
    assert fix["student_params"] == 88005888 and fix["teacher_params"] == 86192640
    assert fix["n_visible"] == 320
    assert abs(fix["loss"].item() - 2.0) < 0.1
    assert fix["attn"].shape == (8, 196) and (fix["attn_rowsum"] < 1.0).all() and (fix["attn_rowsum"] > 0.9).all()


def test_pseudo_label_fusion_logic():
    g = torch.Generator().manual_seed(0)
    lf = torch.randn(16, 12, generator=g) * 3
    lm = torch.randn(2, 16, 12, generator=g)
    cp = torch.softmax(torch.randn(16, 12, generator=g) * 3, -1)
    r = O.pseudo_label_fusion(lf, lm, cp, clip_threshold=0.5)
    msp, preds = torch.softmax(lf, -1).max(-1)
    cm, cpred = cp.max(-1)
    match = cpred == preds
    assert torch.equal(r["sel_mask"], match | (((msp >= 0.5) ^ (cm >= 0.5)) & ~match))
    assert torch.equal(r["pseudo"], preds)
