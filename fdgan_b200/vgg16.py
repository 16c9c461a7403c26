"""``myutils/vgg16.py`` surface of the reference (Vgg16 perceptual-feature extractor), on fdgan_b200 kernels.

``Vgg16()`` declares all 13 convolutions like the reference (myutils/vgg16.py:9-25; conv5_x are parameters
only) and ``forward`` returns ``[relu1_2, relu2_2, relu3_3, relu4_3]`` (:27-49).  The outputs are
NCHW-shaped views of channels-last memory.  Gradients flow to the input for any subset of the four outputs;
parameter gradients are produced only if the parameters require grad (FD-GAN uses the extractor frozen).
"""
from __future__ import annotations

import torch

from . import engine
from .dehaze1113 import ConvParams, _alloc_grads, _KernelNet

_CFG = (("conv1_1", 3, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128), ("conv2_2", 128, 128),
        ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256),
        ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512),
        ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512))


class _VggFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, *params):
        need = any(ctx.needs_input_grad[1:])
        outs, ectx = engine.vgg_forward(mod, x, need)
        ctx.mod, ctx.ectx = mod, (ectx if need else None)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        mod = ctx.mod
        need_dx = ctx.needs_input_grad[1]
        named = mod._used_named_parameters()
        need_w = any(ctx.needs_input_grad[2:])
        grads = None
        if need_w:
            _flat, grads = _alloc_grads(named)
        dx = engine.vgg_backward(mod, ctx.ectx, list(gouts), grads, need_dx)
        ctx.ectx = None
        if not need_w:
            return (None, dx) + (None,) * len(named)
        return (None, dx) + tuple(grads[n] if ctx.needs_input_grad[2 + i] else None for i, (n, _p) in enumerate(named))


class Vgg16(_KernelNet):
    def __init__(self):
        super().__init__()
        for name, ci, co in _CFG:
            setattr(self, name, ConvParams(ci, co, 3, True))

    def _is_used(self, name):
        return not name.startswith("conv5_")

    def forward(self, X):
        named = self._used_named_parameters()
        if torch.is_grad_enabled() and (X.requires_grad or any(p.requires_grad for _n, p in named)):
            return list(_VggFn.apply(self, X, *[p for _n, p in named]))
        outs, _ = engine.vgg_forward(self, X, False)
        return outs
