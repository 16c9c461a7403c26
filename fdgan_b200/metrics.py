"""GPU side of the reference's output path and metric script.

``save_image_u8`` = the bytes ``torchvision.utils.save_image(x, normalize=True)`` writes at demo.py:151 (min-max over the
whole tensor, *255 + 0.5, truncation to uint8, HWC); ``psnr_ssim`` = PSNRSSIM.py:201-240 (1-pixel border crop, PSNR on
/255 values, Gaussian-weighted SSIM per channel with a 5-pixel crop).  Both run as fdgan_b200 kernels on CUDA tensors, so
image quality can be evaluated inside a validation or benchmark loop without PNG files.
"""
from __future__ import annotations

import math

import torch

from . import _lib as L
from .ops import View, _byref, _stream


def save_image_u8(x: torch.Tensor) -> torch.Tensor:
    """x: [3,H,W] or [N,3,H,W] fp32 CUDA -> uint8 [H,W,3] / [N,H,W,3] (normalised over the WHOLE tensor, like save_image)."""
    if not x.is_cuda or x.dtype != torch.float32:
        raise RuntimeError("fdgan_b200.metrics runs on CUDA fp32 tensors only (no CPU fallback)")
    squeeze = x.dim() == 3
    if squeeze:
        x = x.unsqueeze(0)
    if x.dim() != 4:
        raise ValueError("save_image_u8 expects [3,H,W] or [N,3,H,W], got %s" % (tuple(x.shape),))
    N, C, H, W = x.shape
    v = View.from_nchw(x)
    vt = v.ft()
    mm = torch.empty(2, dtype=torch.float32, device=x.device)
    out = torch.empty((N, H, W, C), dtype=torch.uint8, device=x.device)
    L.check(L.lib.fdg_image_minmax(_byref(vt), N, H, W, C, mm.data_ptr(), _stream()), "image_minmax")
    L.check(L.lib.fdg_image_pack_u8(_byref(vt), N, H, W, C, mm.data_ptr(), out.data_ptr(), _stream()), "image_pack_u8")
    return out[0] if squeeze else out


def psnr_ssim(ref_u8: torch.Tensor, res_u8: torch.Tensor) -> tuple[float, float]:
    """ref_u8, res_u8: uint8 [H,W,3] CUDA tensors -> (PSNR, SSIM) exactly as PSNRSSIM.py computes them for one image pair."""
    if not (ref_u8.is_cuda and res_u8.is_cuda and ref_u8.dtype == torch.uint8 and res_u8.dtype == torch.uint8):
        raise RuntimeError("psnr_ssim expects uint8 CUDA tensors")
    if ref_u8.shape != res_u8.shape or ref_u8.dim() != 3 or ref_u8.shape[2] != 3:
        raise ValueError("psnr_ssim expects two [H,W,3] images of the same size")
    H, W, _ = ref_u8.shape
    if H < 13 or W < 13:
        raise ValueError("images must be at least 13x13")
    sums = torch.empty(4, dtype=torch.float64, device=ref_u8.device)
    L.check(L.lib.fdg_psnr_ssim_u8(ref_u8.contiguous().data_ptr(), res_u8.contiguous().data_ptr(), H, W, sums.data_ptr(), _stream()), "psnr_ssim_u8")
    s = sums.tolist()
    h, w = H - 2, W - 2
    mse = s[0] / (3.0 * h * w)
    psnr = float("inf") if mse == 0.0 else 10.0 * math.log10(1.0 / mse)
    ssim = (s[1] + s[2] + s[3]) / (3.0 * (h - 10) * (w - 10))
    return psnr, ssim
