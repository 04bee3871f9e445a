"""Minimal model registry with timm's calling convention.

The reference registers its factories with `timm.models.registry.register_model` and builds them through
`timm.create_model(name, pretrained=False, **kwargs)` (run_stage1.py:273-292, run_stage2.py:326-347), which
drops kwargs whose value is None before calling the factory.  timm is not a dependency here; this module
reproduces exactly that contract.
"""
_MODELS = {}


def register_model(fn):
    _MODELS[fn.__name__] = fn
    return fn


def create_model(model_name, pretrained=False, **kwargs):
    if model_name not in _MODELS:
        raise RuntimeError(f"Unknown model ({model_name}); registered: {sorted(_MODELS)}")
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    return _MODELS[model_name](pretrained=pretrained, **kwargs)


def list_models():
    return sorted(_MODELS)
