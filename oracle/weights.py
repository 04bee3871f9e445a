"""Deterministic, seed-addressed state_dicts for parity runs — TEST INFRASTRUCTURE (see unite_oracle.py).

Fixtures under tests/golden/ store (key -> shape, seed) instead of hundreds of MB of weights; both the
fixture generator (which loads the tensors into the reference nn.Modules) and the tests (which feed them to
the oracle and to the CUDA path) rebuild the same tensors from this function.  Magnitudes follow the
reference initialisers (xavier-uniform-like 1/sqrt(fan_in) matrices, modeling_adaptation.py:108-115;
width^-0.5 CLIP embeddings, clip.py:130-143) but biases and LayerNorm affines are made non-trivial so a
dropped bias or gamma shows up as a parity failure.
"""
import math
import torch


def seeded_state(shapes, seed: int):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        if len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if key in ("proj", "positional_embedding"):
                fan_in = shape[-1] if key == "positional_embedding" else shape[0]
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        elif key.endswith("norm.weight") or key.endswith("norm1.weight") or key.endswith("norm2.weight") or \
                key.endswith("ln_1.weight") or key.endswith("ln_2.weight") or key.endswith("ln_pre.weight") or \
                key.endswith("ln_post.weight") or key.endswith("fc_norm.weight"):
            t = 1.0 + 0.05 * torch.randn(shape, generator=g)
        elif key == "class_embedding":
            t = torch.randn(shape, generator=g) / math.sqrt(shape[0])
        else:
            t = 0.05 * torch.randn(shape, generator=g)
        sd[key] = t
    return sd
