"""Host-side rows f3 / f4 of SURVEY.md §8 against fixtures produced by the REFERENCE's own functions
(oracle/make_golden_interop.py lifts load_student_from_ckpt, load_from_ckpt, merge, tensor_normalize unchanged)."""
import os
import types

import pytest
import torch

from tests.util import load_golden


def _model_stub(fix):
    m = fix["model"]
    return types.SimpleNamespace(patch_embed=types.SimpleNamespace(num_patches=m["num_patches"], tubelet_size=m["tubelet_size"]),
                                 pos_embed=torch.zeros(1, m["pos_tokens"], 16))


def _adapted(monkeypatch, loader, args, model):
    from unite_b200 import checkpoint as ck
    got = {}
    monkeypatch.setattr(ck, "load_state_dict", lambda model, state, prefix="", **k: got.update(state=state, prefix=prefix))
    loader(args, model)
    return got


def _same_state(a, b):
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_student_checkpoint_adapter_matches_reference(tmp_path, monkeypatch):
    from unite_b200 import checkpoint as ck
    fix = load_golden("interop.pt")["checkpoint"]
    raw, wrapped, dec = tmp_path / "raw.pth", tmp_path / "wrapped.pth", tmp_path / "dec.pth"
    torch.save(fix["raw"], raw); torch.save(fix["wrapped"], wrapped); torch.save(fix["dec"], dec)
    for case, path in (("student_raw", raw), ("student_wrapped", wrapped)):
        ref = fix["cases"][case]
        a = dict(ref["args"]); a["student_init"] = str(path)
        if a["clip_decoder_init"]:
            a["clip_decoder_init"] = str(dec)
        got = _adapted(monkeypatch, ck.load_student_from_ckpt, types.SimpleNamespace(**a), _model_stub(fix))
        _same_state(got["state"], ref["state"])
        assert got["prefix"] == ref["prefix"]
    # temporal (4 -> 8 frames, linear) and spatial (6x6 -> 8x8, bicubic) interpolation really happened
    assert fix["cases"]["student_raw"]["state"]["pos_embed"].shape == (1, 8 * 8 * 8, 16)


def test_finetune_checkpoint_adapter_matches_reference(tmp_path, monkeypatch):
    from unite_b200 import checkpoint as ck
    fix = load_golden("interop.pt")["checkpoint"]
    raw = tmp_path / "raw.pth"
    torch.save(fix["raw"], raw)
    for case in ("finetune_k400", "finetune_delete_head"):
        ref = fix["cases"][case]
        a = dict(ref["args"]); a["finetune"] = str(raw)
        got = _adapted(monkeypatch, ck.load_from_ckpt, types.SimpleNamespace(**a), _model_stub(fix))
        _same_state(got["state"], ref["state"])
    assert fix["cases"]["finetune_k400"]["state"]["head.weight"].shape[0] == 400
    assert "head.weight" not in fix["cases"]["finetune_delete_head"]["state"]


def test_load_state_dict_reports_like_the_reference(capsys):
    from unite_b200 import checkpoint as ck
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.LayerNorm(3))
    sd = {"m.0.weight": torch.ones(3, 4), "m.0.bias": torch.zeros(5), "m.extra": torch.zeros(1), "other.0.weight": torch.zeros(3, 4)}
    missing, unexpected, ignored = ck.load_state_dict(net, sd, prefix="m.", ignore_missing="1.bias")
    assert torch.equal(net[0].weight, torch.ones(3, 4))
    assert sorted(missing) == ["0.bias", "1.weight"] and unexpected == ["extra"] and ignored == ["1.bias"]
    out = capsys.readouterr().out
    assert "not initialized from pretrained model" in out and "size mismatch for 0.bias" in out


def test_save_and_auto_resume_roundtrip(tmp_path):
    from unite_b200 import checkpoint as ck
    net = torch.nn.Linear(3, 2)
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9)
    net(torch.ones(1, 3)).sum().backward(); opt.step()
    ck.save_model(str(tmp_path), 3, net, opt)
    ck.save_model(str(tmp_path), 11, net, opt)
    net2 = torch.nn.Linear(3, 2)
    opt2 = torch.optim.SGD(net2.parameters(), lr=0.1, momentum=0.9)
    assert ck.auto_load_model(str(tmp_path), net2, opt2) == 12
    assert torch.equal(net2.weight, net.weight)
    ck.save_model(str(tmp_path), 5, net, opt, tag="latest")
    assert ck.auto_load_model(str(tmp_path), net2, opt2) == 6          # checkpoint-latest.pth wins


def test_merge_matches_reference_merge(tmp_path):
    from unite_b200.engine_for_finetuning import merge
    fix = load_golden("interop.pt")["merge"]
    for r, text in fix["files"].items():
        (tmp_path / f"{r}.txt").write_text(text)
    top1, top5 = merge(str(tmp_path), fix["num_tasks"])
    assert abs(top1 - fix["top1"]) < 1e-9 and abs(top5 - fix["top5"]) < 1e-9


def test_oracle_normalisation_is_pinned_to_reference_tensor_normalize():
    from oracle.unite_oracle import normalize_frames_u8
    fix = load_golden("interop.pt")["normalize"]
    assert torch.equal(normalize_frames_u8(fix["frames"]), fix["clip"])


def test_accuracy_and_ece():
    from unite_b200.engine_for_finetuning import accuracy, compute_ece
    out = torch.tensor([[0.1, 0.7, 0.2], [0.8, 0.1, 0.1], [0.3, 0.3, 0.4]])
    tgt = torch.tensor([1, 2, 2])
    a1, a2 = accuracy(out, tgt, topk=(1, 2))
    assert abs(a1.item() - 200 / 3) < 1e-4 and abs(a2.item() - 200 / 3) < 1e-4
    p = torch.tensor([[0.9, 0.1], [0.9, 0.1], [0.6, 0.4], [0.6, 0.4]])
    y = torch.tensor([0, 0, 0, 1])
    assert abs(compute_ece(p, y) - (0.5 * abs(1.0 - 0.9) + 0.5 * abs(0.5 - 0.6))) < 1e-6
