from . import vision_transformer  # noqa: F401

_REGISTRY = {}


def register_model(fn):
    _REGISTRY[fn.__name__] = fn
    return fn


def create_model(name, **kwargs):
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    return _REGISTRY[name](**kwargs)
