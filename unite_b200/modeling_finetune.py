"""Student ViT building blocks and the stage-2 classifier — parameter containers with the reference's names.

Mirrors the public surface of reference src/models/modeling_finetune.py (Block :122, Attention :76, Mlp :56,
PatchEmbed :153, VisionTransformer :237, factories :386-415): same constructor kwargs, attribute names and
state_dict keys/shapes, so checkpoints and driver code (`model.patch_embed.patch_size`, `get_num_layers()`,
`no_weight_decay()`) carry over.  The modules hold parameters only — the arithmetic is in vit_core.ViTTrunk
and runs on the CUDA kernels; calling forward without a GPU raises.
"""
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from .registry import register_model


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def get_sinusoid_encoding_table(n_position, d_hid):
    """fp64 table -> fp32 [1, n, d]: angle = pos / 10000^(2*(j//2)/d), sin on even / cos on odd columns
    (reference modeling_finetune.py:225-235, modeling_adaptation.py:41-51)."""
    j = np.arange(d_hid)
    ang = np.arange(n_position, dtype=np.float64)[:, None] / np.power(10000, 2 * (j // 2) / d_hid)[None, :]
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.tensor(ang, dtype=torch.float).unsqueeze(0)


class _ParamsOnly(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(f"{type(self).__name__} only stores parameters; the owning model runs the fused CUDA path")


class DropPath(_ParamsOnly):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def extra_repr(self):
        return "p={}".format(self.drop_prob)


class Mlp(_ParamsOnly):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)


class Attention(_ParamsOnly):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0, attn_head_dim=None):
        super().__init__()
        if attn_drop or proj_drop:
            raise NotImplementedError("attention / projection dropout are 0 in every shipped UNITE config")
        head_dim = attn_head_dim or dim // num_heads
        self.num_heads = num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, head_dim * num_heads * 3, bias=False)
        if not qkv_bias:
            raise NotImplementedError("qkv_bias=False is not used by any UNITE factory")
        self.q_bias = nn.Parameter(torch.zeros(head_dim * num_heads))
        self.v_bias = nn.Parameter(torch.zeros(head_dim * num_heads))
        self.proj = nn.Linear(head_dim * num_heads, dim)


class Block(_ParamsOnly):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 init_values=None, act_layer=nn.GELU, norm_layer=nn.LayerNorm, attn_head_dim=None):
        super().__init__()
        if init_values:
            raise NotImplementedError("layer-scale (init_values > 0) is off in every UNITE config (modeling_finetune.py:137-141)")
        if drop:
            raise NotImplementedError("dropout is 0 in every shipped UNITE config")
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop, attn_head_dim=attn_head_dim)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.gamma_1, self.gamma_2 = None, None


class PatchEmbed(_ParamsOnly):
    """Conv3d(3, D, (tubelet,16,16), stride=same) parameters; applied as patchify + GEMM."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, num_frames=16, tubelet_size=2):
        super().__init__()
        img_size, patch_size = to_2tuple(img_size), to_2tuple(patch_size)
        if patch_size != (16, 16) or in_chans != 3:
            raise NotImplementedError("the patchify kernel is specialised for 3-channel 16x16 patches")
        self.tubelet_size = int(tubelet_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0]) * (num_frames // self.tubelet_size)
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=(self.tubelet_size, 16, 16), stride=(self.tubelet_size, 16, 16))


def drop_path_factors(rates, B, device, generator=None):
    """Per-sample DropPath factors floor(keep + u)/keep (timm 0.4.12 drop_path) for every block and both
    branches: fp32 [depth, 2, B]; None if every rate is 0."""
    if not any(r > 0 for r in rates):
        return None
    keep = 1.0 - torch.tensor(rates, dtype=torch.float32, device=device).view(-1, 1, 1)
    u = torch.rand(len(rates), 2, B, device=device, generator=generator)
    return (torch.floor(keep + u) / keep).contiguous()
