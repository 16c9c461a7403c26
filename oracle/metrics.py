"""CPU restatement of the reference's image pipeline and metric script (TEST INFRASTRUCTURE ONLY).

  * read_sample      datasets/pix2pix.py:62-77 semantics on the contiguous-float64 HDF5 samples (SURVEY Appendix C;
                     h5py is not installed, the two datasets sit at fixed byte offsets)
  * save_image_u8    torchvision.utils.save_image(normalize=True, scale_each=False) as called at demo.py:151:
                     min-max over the whole tensor, *255 + 0.5, clamp, truncate to uint8, HWC
  * psnr / mssim     PSNRSSIM.py:201-240: 1-pixel border crop, PSNR on /255 floats, SSIM per channel with
                     gaussian_weights=True (scipy gaussian_filter sigma 1.5), use_sample_covariance=False,
                     data_range 255, K1 .01, K2 .03, 5-pixel crop before the mean (PSNRSSIM.py:46-194)
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.ndimage import gaussian_filter


def read_sample(path):
    buf = open(path, "rb").read()
    assert buf[:8] == b"\x89HDF\r\n\x1a\n"
    gt = np.frombuffer(buf, "<f8", 384 * 512 * 3, 2144).reshape(384, 512, 3)
    haze = np.frombuffer(buf, "<f8", 384 * 512 * 3, 4720736).reshape(384, 512, 3)
    # HWC -> CHW by the two swapaxes of datasets/pix2pix.py:69-77
    to_chw = lambda a: np.swapaxes(np.swapaxes(a, 0, 2), 1, 2)
    return to_chw(haze), to_chw(gt)


def save_image_u8(x: torch.Tensor) -> np.ndarray:
    """x: [3,H,W] float -> uint8 [H,W,3] exactly as save_image(normalize=True) would write it."""
    t = x.detach().float().cpu().clone()
    lo, hi = float(t.min()), float(t.max())
    t.clamp_(min=lo, max=hi)
    t.sub_(lo).div_(max(hi - lo, 1e-5))
    return t.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8).numpy()


def _ssim_channel(X, Y, data_range=255.0):
    X, Y = X.astype(np.float64), Y.astype(np.float64)
    f = lambda a: gaussian_filter(a, sigma=1.5)
    ux, uy = f(X), f(Y)
    uxx, uyy, uxy = f(X * X), f(Y * Y), f(X * Y)
    vx, vy, vxy = uxx - ux * ux, uyy - uy * uy, uxy - ux * uy
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = 5
    return S[pad:-pad, pad:-pad].mean()


def psnr(ref_u8: np.ndarray, res_u8: np.ndarray) -> float:
    a = ref_u8.astype(float)[1:-1, 1:-1, :] / 255.0
    b = res_u8.astype(float)[1:-1, 1:-1, :] / 255.0
    return float(10 * np.log10(1.0 / np.mean(np.square(a - b))))


def mssim(ref_u8: np.ndarray, res_u8: np.ndarray) -> float:
    a, b = ref_u8[1:-1, 1:-1, :], res_u8[1:-1, 1:-1, :]
    return float(np.mean([_ssim_channel(a[:, :, i], b[:, :, i]) for i in range(3)]))
