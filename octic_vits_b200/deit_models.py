"""The reference's timm-registered DeiT-III octic configurations (octic_vits/deit_models.py:11-72), same factory
names and keyword handling.  timm itself is not a dependency: `register_model` / `create_model` are a small local
registry with timm's calling convention (None-valued kwargs are dropped, as timm.create_model does)."""
from .layers import Layer_scale_init_Block, Layer_scale_init_BlockD8
from .model import OcticVisionTransformer

_REGISTRY = {}


def register_model(fn):
    _REGISTRY[fn.__name__] = fn
    return fn


def _register_all() -> None:
    """the DINOv2 factories (dinov2_models.py) register themselves on import, as the reference's modules do with timm"""
    from . import dinov2_models  # noqa: F401


def create_model(model_name: str, **kwargs):
    if model_name not in _REGISTRY:
        _register_all()
    if model_name not in _REGISTRY:
        raise RuntimeError(f"Unknown model ({model_name})")
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    return _REGISTRY[model_name](**kwargs)


def list_models():
    _register_all()
    return sorted(_REGISTRY)


def _deit(img_size, patch_size, embed_dim, depth, num_heads, invariant, **kwargs):
    return OcticVisionTransformer(img_size=img_size, patch_size=patch_size, embed_dim=embed_dim, depth=depth,
                                  num_heads=num_heads, mlp_ratio=4, qkv_bias=True, invariant=invariant,
                                  standard_block_layers=Layer_scale_init_Block,
                                  octic_block_layers=Layer_scale_init_BlockD8, **kwargs)


@register_model
def hybrid_deit_large_patch16(img_size=224, **kwargs):
    return _deit(img_size, 16, 1024, 24, 16, False, **kwargs)


@register_model
def hybrid_deit_huge_patch14(img_size=224, **kwargs):
    return _deit(img_size, 14, 1280, 32, 16, False, **kwargs)


@register_model
def d8_inv_early_deit_huge_patch14(img_size=224, **kwargs):
    return _deit(img_size, 14, 1280, 32, 16, True, **kwargs)


@register_model
def d8_inv_early_deit_large_patch16(img_size=224, **kwargs):
    return _deit(img_size, 16, 1024, 24, 16, True, **kwargs)


@register_model
def hybrid_deit_small_patch16(img_size=224, **kwargs):
    """BASELINE.json configs[0]: hybrid octic ViT-S/16 (embed 384, depth 12, heads 6); not in the reference registry,
    built there by calling OcticVisionTransformer directly."""
    return _deit(img_size, 16, 384, 12, 6, False, **kwargs)
