"""``loss.py`` surface of the reference (frequency decomposition), on fdgan_b200 kernels.

The reference's loss.py survives only as ``__pycache__/loss.cpython-36.pyc``:
``Blur(l=15, kernel=None, use_input_norm=True)`` (@L122-151), ``isotropic_gaussian_kernel`` (@L153-159),
``Laplacian(kernel_size)`` (@L245-301) and the singletons ``blur`` (@L161-162) and ``laplace_filter`` (@L304).
One fused kernel produces all nine channels [x, LF, HF] of the Fusion-discriminator input; ``Blur`` and
``Laplacian`` are views of it, and ``freq_concat`` is the fused call the training step uses.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .ops import View


def isotropic_gaussian_kernel(l, sigma, tensor=True):
    """loss.pyc@L153-159."""
    ax = np.arange(-l // 2 + 1.0, l // 2 + 1.0)
    xx, yy = np.meshgrid(ax, ax)
    kernel = np.exp(-(xx ** 2 + yy ** 2) / (2.0 * sigma ** 2))
    kernel = kernel / np.sum(kernel)
    return torch.FloatTensor(kernel) if tensor else kernel


def get_laplacian_kernel2d(kernel_size):
    """loss.pyc@L205-241: ones(k,k) with centre 1 - k^2 (not normalised)."""
    if not isinstance(kernel_size, int) or kernel_size % 2 == 0 or kernel_size <= 0:
        raise TypeError("ksize must be an odd positive integer. Got {}".format(kernel_size))
    k = torch.ones((kernel_size, kernel_size))
    k[kernel_size // 2, kernel_size // 2] = 1 - kernel_size ** 2
    return k


class _FreqFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("frequency decomposition expects a [B,3,H,W] input, got %s" % (tuple(x.shape),))
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("fdgan_b200 runs on CUDA fp32 tensors only (no CPU fallback)")
        B, _, H, W = x.shape
        z = View.alloc(B, H, W, 9, x.device)          # channels-last memory
        ops.freq_concat_fwd(View.from_nchw(x), z)
        ctx.shape = (B, H, W)
        return z.as_nchw()

    @staticmethod
    def backward(ctx, dz):
        B, H, W = ctx.shape
        dx = torch.empty((B, 3, H, W), dtype=torch.float32, device=dz.device)
        scratch = torch.empty(B * 3 * H * W, dtype=torch.float32, device=dz.device)
        ops.freq_concat_bwd(View.from_nchw(dz), View.from_nchw(dx), scratch)
        return dx


def freq_concat(x):
    """[x, Blur(x), Laplacian(x)] along channels: the Fusion-discriminator input (facades/network.png)."""
    return _FreqFn.apply(x)


class Blur(nn.Module):
    """loss.pyc@L122-151.  Only the configuration the reference instantiates is built as a kernel:
    l=15, sigma=3 Gaussian (the ``blur`` singleton), ImageNet input normalisation."""

    def __init__(self, l=15, kernel=None, use_input_norm=True):
        super().__init__()
        self.l = l
        ref = isotropic_gaussian_kernel(15, 3.0)
        if l != 15 or not use_input_norm or (kernel is not None and not torch.allclose(torch.as_tensor(kernel, dtype=torch.float32).view(15, 15), ref, atol=1e-7)):
            raise NotImplementedError("fdgan_b200.Blur implements the reference's blur = Blur(l=15, kernel=isotropic_gaussian_kernel(15, 3.0))")
        self.use_input_norm = use_input_norm

    def forward(self, input):
        return freq_concat(input)[:, 3:6]


class Laplacian(nn.Module):
    """loss.pyc@L245-301 with kernel_size 3 (the ``laplace_filter`` singleton); 3-channel inputs."""

    def __init__(self, kernel_size=3):
        super().__init__()
        self.kernel = get_laplacian_kernel2d(kernel_size)
        if kernel_size != 3:
            raise NotImplementedError("fdgan_b200.Laplacian implements the reference's laplace_filter = Laplacian(kernel_size=3)")
        self._padding = (kernel_size - 1) // 2

    def forward(self, x):
        if not torch.is_tensor(x) or x.dim() != 4:
            raise ValueError("Invalid input shape, we expect BxCxHxW. Got: {}".format(getattr(x, "shape", None)))
        return freq_concat(x)[:, 6:9]


blur_kernel = isotropic_gaussian_kernel(l=15, sigma=3.0)
blur = Blur(l=15, kernel=blur_kernel)
laplace_filter = Laplacian(kernel_size=3)
