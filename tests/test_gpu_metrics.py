"""GPU output path / metric kernels against the CPU restatement of torchvision.save_image and
PSNRSSIM.py (oracle/metrics.py, itself checked in tests/test_oracle_metrics.py)."""
import numpy as np
import pytest
import torch

from oracle import metrics as M
from tests.util import seeded

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(3, 40, 56), (3, 17, 13), (3, 256, 256)])
def test_save_image_u8_bytes_identical(shape):
    from fdgan_b200 import metrics
    x = seeded(shape, 1, -1.0, 1.0)
    want = M.save_image_u8(x)                                  # uint8 [H,W,3]
    got = metrics.save_image_u8(x.cuda()).cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got, want)
    xcl = x.unsqueeze(0).contiguous(memory_format=torch.channels_last)[0]
    assert np.array_equal(metrics.save_image_u8(xcl.cuda()).cpu().numpy(), want)      # layout independent


@pytest.mark.parametrize("hw", [(40, 56), (64, 33), (256, 256)])
def test_psnr_ssim_matches_metric_script(hw):
    from fdgan_b200 import metrics
    g = np.random.default_rng(3)
    ref = g.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
    noise = g.integers(-20, 21, size=ref.shape)
    res = np.clip(ref.astype(int) + noise, 0, 255).astype(np.uint8)
    # smooth content as well (SSIM on noise is near zero)
    yy, xx = np.mgrid[0:hw[0], 0:hw[1]]
    base = (127 + 100 * np.sin(yy / 7.0)[..., None] * np.cos(xx / 5.0)[..., None] * np.ones(3)).astype(np.uint8)
    for a, b in ((ref, res), (base, np.clip(base.astype(int) + noise // 4, 0, 255).astype(np.uint8))):
        p, s = metrics.psnr_ssim(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
        assert abs(p - M.psnr(a, b)) <= 1e-9 * max(1.0, abs(p))
        assert abs(s - M.mssim(a, b)) <= 1e-9
    p, s = metrics.psnr_ssim(torch.from_numpy(ref).cuda(), torch.from_numpy(ref).cuda())
    assert p == float("inf") and abs(s - 1.0) <= 1e-12


def test_run_demo_loop(tmp_path, monkeypatch):
    """fdgan_b200.io.run_demo (demo.py:118-151) on the committed 256x256 crop of testsample1/3.h5 (the HDF5 reader itself is
    a CPU test against the reference's files): generator forward, GPU byte conversion, PNG on disk == oracle pipeline."""
    import os
    import fdgan_b200
    from fdgan_b200 import io as fio
    from oracle import fdgan_oracle as O
    from tests.util import GOLDEN
    Image = pytest.importorskip("PIL.Image")
    haze = torch.from_numpy(np.load(os.path.join(GOLDEN, "testsample1_3_crop256.npz"))["haze"])          # [3,256,256]
    monkeypatch.setattr(fio, "read_h5_pair", lambda path: (haze.clone(), haze.clone()))
    net = fdgan_b200.FDGAN()
    net.load_state_dict(O.make_fdgan_state(0))
    net = net.cuda().train()
    pairs = fio.run_demo(net, str(tmp_path), str(tmp_path / "out"), 1)
    out_u8, gt_u8 = pairs[0]
    png = np.asarray(Image.open(str(tmp_path / "out" / "0.png")).convert("RGB"))
    assert np.array_equal(png, out_u8.cpu().numpy())
    with torch.no_grad():
        y_ref = O.fdgan_forward(O.make_fdgan_state(0), haze.unsqueeze(0), True, False)
    ref_u8 = M.save_image_u8(y_ref[0])
    assert M.psnr(ref_u8, png) > 55.0 or np.array_equal(ref_u8, png)
    assert np.array_equal(gt_u8.cpu().numpy(), M.save_image_u8(haze))
