"""Shared helpers for the parity tests (oracle = checker, unite_b200 = thing under test)."""
import os
from functools import partial

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def oracle_cfgs(fix):
    from oracle import unite_oracle as O
    return O.StudentCfg(**fix["cfg"]["student"]), O.TeacherCfg(**fix["cfg"]["teacher"])


def seeded_states(fix):
    from oracle.weights import seeded_state
    s = fix["seeds"]
    return (seeded_state(fix["student_shapes"], s["student"]), seeded_state(fix["teacher_shapes"], s["teacher"]),
            seeded_state(fix["vit_shapes"], s["vit"]) if "vit_shapes" in fix else None)


def build_student(scfg, drop_path_rate=0.0):
    import torch.nn as nn
    from unite_b200.modeling_adaptation import AdaptationVisionTransformer
    return AdaptationVisionTransformer(
        img_size=scfg.img_size, patch_size=scfg.patch_size, encoder_embed_dim=scfg.embed_dim, encoder_depth=scfg.depth,
        encoder_num_heads=scfg.num_heads, encoder_num_classes=0, mlp_ratio=4, qkv_bias=True,
        norm_layer=partial(nn.LayerNorm, eps=1e-6), num_frames=scfg.num_frames, tubelet_size=scfg.tubelet_size,
        clip_decoder_embed_dim=scfg.embed_dim, clip_output_dim=scfg.clip_output_dim, clip_return_layers=list(scfg.return_layers),
        drop_path_rate=drop_path_rate)


def build_teacher(tcfg):
    from unite_b200.clip import VisionTransformer
    return VisionTransformer(input_resolution=tcfg.input_resolution, patch_size=tcfg.patch_size, width=tcfg.width, layers=tcfg.layers,
                             heads=tcfg.heads, output_dim=tcfg.output_dim, kernel_size=tcfg.kernel_size, return_attn=True,
                             clip_return_layers=list(tcfg.return_layers)).eval()


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def per_token_rel(a, b):
    """max over tokens of ||a_t - b_t|| / ||b_t|| (last dim = feature)."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm(dim=-1) / (b.norm(dim=-1) + 1e-30))


def cosine(a, b):
    a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()
