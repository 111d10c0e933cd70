"""recnext_b200 — B200-native (sm_100a) implementation of RecNeXt's RecConv hot path.

Public surface mirrors the reference (suous/RecNeXt): ``RecConv2d`` (model/recnext.py) and ``RecAttn2d`` (model/recattn.py).  The arithmetic runs in
hand-written CUDA behind the C ABI of include/recnext_b200.h; importing this package does not need a GPU, but
calling it does — there is no CPU fallback.
"""
from .recattn import RecAttn2d, recattn_down_forward, recattn_up_forward  # noqa: F401
from .recconv import RecConv2d, plan_describe, recconv_backward, recconv_forward  # noqa: F401
from .variants import LsRecAttn2d, MllaRecConv2d, PartialChannelOperation  # noqa: F401

__all__ = ["RecConv2d", "recconv_forward", "recconv_backward", "plan_describe", "RecAttn2d", "recattn_down_forward", "recattn_up_forward",
           "MllaRecConv2d", "LsRecAttn2d", "PartialChannelOperation"]
