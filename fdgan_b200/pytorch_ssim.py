"""``models/pytorch_ssim`` surface of the reference (models/pytorch_ssim/__init__.py:39-73) on the fused SSIM kernel.

``ssim(img1, img2, window_size=11, size_average=True)`` and ``SSIM(window_size, size_average)``: window 11, sigma 1.5, zero padding,
C1 = 0.01^2, C2 = 0.03^2 (``_ssim`` :17-37).  Differentiable w.r.t. either image (SSIM is symmetric in its arguments, so the
gradient w.r.t. img2 is the same kernel with the images swapped).  The reference runs ten 11x11 depth-wise convolutions and a
dozen element-wise passes for this; here it is one separable filter kernel per direction (fdg_ssim_loss_grad, csrc/ssim.cu).
"""
from __future__ import annotations

import torch

from . import ops
from .ops import View


def _check(img1, img2, window_size):
    if window_size != 11:
        raise NotImplementedError("fdgan_b200.pytorch_ssim implements the reference's default window (11, sigma 1.5)")
    if img1.dim() != 4 or tuple(img1.shape) != tuple(img2.shape):
        raise ValueError("ssim expects two [B,C,H,W] tensors of the same shape, got %s and %s" % (tuple(img1.shape), tuple(img2.shape)))
    if not (img1.is_cuda and img2.is_cuda) or img1.dtype != torch.float32 or img2.dtype != torch.float32:
        raise RuntimeError("fdgan_b200 runs on CUDA fp32 tensors only (no CPU fallback)")


class _SsimFn(torch.autograd.Function):
    """mean over (C,H,W) of the SSIM map, one value per image (``size_average=False``); the mean over images is taken outside."""

    @staticmethod
    def forward(ctx, img1, img2):
        B, Cc, H, W = img1.shape
        out = torch.zeros(B, dtype=torch.float64, device=img1.device)
        per = 1.0 / (Cc * H * W)
        for b in range(B):
            ops.ssim_loss_grad(View.from_nchw(img1[b:b + 1]), View.from_nchw(img2[b:b + 1]), per, 0.0, out[b:b + 1], None)
        ctx.save_for_backward(img1, img2)
        return out.float()

    @staticmethod
    def backward(ctx, gout):
        img1, img2 = ctx.saved_tensors
        B, Cc, H, W = img1.shape
        per = 1.0 / (Cc * H * W)
        gh = gout.detach().double().cpu().tolist()      # B scalars: the per-image weights of the gradient
        sink = torch.zeros(1, dtype=torch.float64, device=img1.device)
        grads = []
        for need, a, b_ in ((ctx.needs_input_grad[0], img1, img2), (ctx.needs_input_grad[1], img2, img1)):
            if not need:
                grads.append(None)
                continue
            g = torch.empty_like(a, memory_format=torch.contiguous_format)
            for b in range(B):
                ops.ssim_loss_grad(View.from_nchw(a[b:b + 1]), View.from_nchw(b_[b:b + 1]), 0.0, per * gh[b], sink, View.from_nchw(g[b:b + 1]))
            grads.append(g)
        return tuple(grads)


def ssim(img1, img2, window_size=11, size_average=True):
    """models/pytorch_ssim/__init__.py:65-73."""
    _check(img1, img2, window_size)
    per_image = _SsimFn.apply(img1, img2)
    return per_image.mean() if size_average else per_image


class SSIM(torch.nn.Module):
    """models/pytorch_ssim/__init__.py:39-63."""

    def __init__(self, window_size=11, size_average=True):
        super().__init__()
        self.window_size = window_size
        self.size_average = size_average
        self.channel = 1

    def forward(self, img1, img2):
        self.channel = img1.shape[1]
        return ssim(img1, img2, self.window_size, self.size_average)
