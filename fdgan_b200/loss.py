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


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
MAX_FILTER = 31      # fdg_depthwise2d: odd sizes up to 31 x 31


class _DepthwiseFn(torch.autograd.Function):
    """One l x l kernel on every (image, channel) plane (fdg_depthwise2d_fwd / _bwd): the generic form behind Blur and Laplacian."""

    @staticmethod
    def forward(ctx, x, kernel, pad_mode, mean, inv_std):
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("fdgan_b200 runs on CUDA fp32 tensors only (no CPU fallback)")
        B, Cc, H, W = x.shape
        y = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
        ops.depthwise2d(View.from_nchw(x), View.from_nchw(y), kernel, pad_mode, mean, inv_std)
        ctx.cfg = (kernel, pad_mode, inv_std)
        return y

    @staticmethod
    def backward(ctx, dy):
        kernel, pad_mode, inv_std = ctx.cfg
        dx = torch.zeros(dy.shape, dtype=torch.float32, device=dy.device)
        ops.depthwise2d(View.from_nchw(dy), View.from_nchw(dx), kernel, pad_mode, None, inv_std, backward=True)
        return dx, None, None, None, None


def _check_4d(x):
    if not torch.is_tensor(x) or x.dim() != 4:
        raise ValueError("Invalid input shape, we expect BxCxHxW. Got: {}".format(getattr(x, "shape", None)))


class Blur(nn.Module):
    """loss.pyc@L122-151: optional ImageNet normalisation, ReflectionPad2d(l // 2), ONE l x l kernel on every (batch, channel)
    plane.  The configuration FD-GAN instantiates (l=15, sigma-3 Gaussian, normalisation on: the ``blur`` singleton) on 3-channel
    images runs inside the fused frequency-decomposition kernel; every other (l, kernel, use_input_norm) or channel count runs on
    the generic depth-wise kernel (odd l <= 31)."""

    def __init__(self, l=15, kernel=None, use_input_norm=True):
        super().__init__()
        self.l = int(l)
        if self.l % 2 == 0 or not 1 <= self.l <= MAX_FILTER:
            raise NotImplementedError("fdgan_b200.Blur: odd kernel sizes up to %d (ReflectionPad2d(l // 2) keeps the size only for odd l)" % MAX_FILTER)
        ref = isotropic_gaussian_kernel(15, 3.0)
        k = ref if kernel is None and self.l == 15 else kernel
        if k is None:
            raise ValueError("Blur(l=%d): a kernel is required (the reference passes isotropic_gaussian_kernel(l, sigma))" % self.l)
        k = torch.as_tensor(k, dtype=torch.float32).reshape(self.l, self.l).contiguous()
        self.use_input_norm = bool(use_input_norm)
        self._default = self.l == 15 and self.use_input_norm and bool(torch.allclose(k, ref, atol=1e-7))
        self.register_buffer("_kernel", k, persistent=False)
        if self.use_input_norm:      # loss.pyc@L131-136: mean / std buffers of shape [1,3,1,1]
            self.register_buffer("_mean", torch.tensor(IMAGENET_MEAN, dtype=torch.float32), persistent=False)
            self.register_buffer("_inv_std", 1.0 / torch.tensor(IMAGENET_STD, dtype=torch.float32), persistent=False)

    def forward(self, input):
        _check_4d(input)
        if self._default and input.shape[1] == 3:
            return freq_concat(input)[:, 3:6]
        if self.use_input_norm and input.shape[1] != 3:
            raise RuntimeError("Blur(use_input_norm=True) normalises with the 3-channel ImageNet mean / std; got %d channels" % input.shape[1])
        if min(input.shape[2:]) <= self.l // 2:
            raise RuntimeError("Blur: ReflectionPad2d(%d) needs H, W > %d" % (self.l // 2, self.l // 2))
        k = self._kernel.to(input.device)
        mean = self._mean.to(input.device) if self.use_input_norm else None
        istd = self._inv_std.to(input.device) if self.use_input_norm else None
        return _DepthwiseFn.apply(input, k, 1, mean, istd)


class Laplacian(nn.Module):
    """loss.pyc@L245-301: depth-wise ones(k, k) with centre 1 - k^2 on every channel (``kernel.repeat(c,1,1,1)``, ``groups=c``), zero
    padding (k - 1) // 2.  Channel-agnostic like the reference; kernel_size 3 on 3 channels (the ``laplace_filter`` singleton as the
    Fusion-discriminator uses it) runs inside the fused frequency-decomposition kernel."""

    def __init__(self, kernel_size=3):
        super().__init__()
        self.kernel = get_laplacian_kernel2d(kernel_size)
        if kernel_size > MAX_FILTER:
            raise NotImplementedError("fdgan_b200.Laplacian: kernel sizes up to %d" % MAX_FILTER)
        self.kernel_size = kernel_size
        self._padding = (kernel_size - 1) // 2
        self.register_buffer("_kernel", self.kernel.clone().float().contiguous(), persistent=False)

    def forward(self, x):
        _check_4d(x)
        if self.kernel_size == 3 and x.shape[1] == 3:
            return freq_concat(x)[:, 6:9]
        return _DepthwiseFn.apply(x, self._kernel.to(x.device), 0, None, None)


blur_kernel = isotropic_gaussian_kernel(l=15, sigma=3.0)
blur = Blur(l=15, kernel=blur_kernel)
laplace_filter = Laplacian(kernel_size=3)
