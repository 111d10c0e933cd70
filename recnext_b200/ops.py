"""``torch.library`` registration of the RecConv kernels: ``torch.ops.recnext.recconv`` / ``recnext.recconv_backward``.

The reference's release path scripts / traces the model (publish.py:32-38, fuse_eval.py:48, export_coreml.py:30) and
``torch.compile`` graphs need an operator the tracer can see: a ``ctypes`` call inside an ``autograd.Function`` is opaque to
both.  These are registered custom operators with fake (meta) implementations and an autograd formula, so

  * ``torch.jit.trace`` records ``recnext::recconv`` as one node,
  * ``torch.compile(fullgraph=True)`` keeps it as one call (no graph break), forward and backward,
  * ``torch.library.opcheck`` validates schema, fake tensors and autograd registration.

``RecConv2d.forward`` dispatches through the operator (``recnext_b200.recconv``); the eager cost is the same C-ABI call.
Parameters travel as tensor lists in the ``state_dict`` order: weights = [down.weight, convs.0.weight, ..., convs.L.weight],
biases = the same order or an empty list.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import recconv as _rc


@torch.library.custom_op("recnext::recconv", mutates_args=(), device_types="cuda")
def recconv(x: torch.Tensor, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], kernel_size: int, level: int, mode: str) -> torch.Tensor:
    return _rc.recconv_forward(x, list(weights), list(biases) if len(biases) else None, kernel_size, level, mode)


@recconv.register_fake
def _(x, weights, biases, kernel_size, level, mode):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


@torch.library.custom_op("recnext::recconv_backward", mutates_args=(), device_types="cuda")
def recconv_backward(x: torch.Tensor, gy: torch.Tensor, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], kernel_size: int, level: int,
                     mode: str) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (gx, gw [(L+2), C, k*k] fp32, gb [(L+2), C] fp32 — an empty tensor when there are no biases)"""
    gx, gw, gb = _rc.recconv_backward(x, gy, list(weights), list(biases) if len(biases) else None, kernel_size, level, mode)
    return gx, gw, gb if gb is not None else gw.new_empty((0,))


@recconv_backward.register_fake
def _(x, gy, weights, biases, kernel_size, level, mode):
    C = x.shape[1]
    gw = x.new_empty((level + 2, C, kernel_size * kernel_size), dtype=torch.float32)
    gb = x.new_empty((level + 2, C), dtype=torch.float32) if len(biases) else x.new_empty((0,), dtype=torch.float32)
    return torch.empty_like(x, memory_format=torch.contiguous_format), gw, gb


def _setup_context(ctx, inputs, output):
    x, weights, biases, kernel_size, level, mode = inputs
    ctx.save_for_backward(x, *weights, *biases)
    ctx.cfg = (kernel_size, level, mode, len(weights), len(biases))


def _backward(ctx, gy):
    k, level, mode, nw, nb = ctx.cfg
    x, *params = ctx.saved_tensors
    weights, biases = params[:nw], params[nw:]
    gx, gw, gb = torch.ops.recnext.recconv_backward(x, gy.contiguous(), weights, biases, k, level, mode)
    C = x.shape[1]
    wdt = weights[0].dtype
    gws: List = [gw[j].view(C, 1, k, k).to(wdt) for j in range(level + 2)]
    gbs: List = [gb[j].to(wdt) for j in range(level + 2)] if nb else []
    if level == 0:   # `down` does not take part in the graph at level 0 (the reference leaves its grad at None)
        gws[0] = None
        if nb:
            gbs[0] = None
    return gx, gws, gbs, None, None, None


recconv.register_autograd(_backward, setup_context=_setup_context)
