"""CPU checks of the metric / image-pipeline restatement (oracle/metrics.py) against independent formulations."""
import numpy as np
import torch

from oracle import metrics as M


def test_save_image_matches_torchvision_semantics():
    import torchvision.utils as vutils
    g = torch.Generator().manual_seed(3)
    x = torch.randn((3, 20, 24), generator=g)
    grid = vutils.make_grid(x.unsqueeze(0), normalize=True, scale_each=False, padding=0)
    want = grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8).numpy()
    assert np.array_equal(M.save_image_u8(x), want)


def test_psnr_ssim_properties():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (40, 48, 3), dtype=np.uint8)
    assert M.mssim(a, a) == 1.0
    b = a.copy(); b[10:20, 10:20] ^= 4
    assert 0.0 < M.mssim(a, b) < 1.0
    mse = np.mean(np.square(a[1:-1, 1:-1].astype(float) / 255 - b[1:-1, 1:-1].astype(float) / 255))
    assert abs(M.psnr(a, b) - 10 * np.log10(1 / mse)) < 1e-12
    # symmetric in its arguments, as PSNRSSIM.py's swapped directory names rely on (PSNRSSIM.py:245-246)
    assert abs(M.mssim(a, b) - M.mssim(b, a)) < 1e-12
