"""timm.models stand-ins: a registry + pass-through builders."""
_REGISTRY = {}


def register_model(fn):
    _REGISTRY[fn.__name__] = fn
    return fn


def create_model(name, pretrained=False, **kwargs):
    return _REGISTRY[name](pretrained=pretrained, **kwargs)


def build_model_with_cfg(model_cls, variant, pretrained, feature_cfg=None, **kwargs):
    kwargs.pop("pretrained_cfg", None)
    kwargs.pop("pretrained_cfg_overlay", None)
    return model_cls(**kwargs)


def generate_default_cfgs(cfgs):
    return cfgs
