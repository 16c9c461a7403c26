"""configs[0] of BASELINE.json on the GPU: demo.py's generator forward on a 256x256 crop of testsample1/3.h5 (haze rows
64:320, cols 128:384, committed as tests/golden/testsample1_3_crop256.npz), written through the save_image(normalize=True)
pipeline and scored with the PSNRSSIM.py restatement.  PSNR / SSIM of the fdgan_b200 output must equal those of the
reference arithmetic (CPU oracle, same seeded weights) to 2 decimal places."""
import os

import numpy as np
import pytest
import torch

from oracle import fdgan_oracle as O
from oracle import metrics as M
from tests.util import GOLDEN, maxabs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", ["simt_fp32", "tcgen05_bf16x3"])
def test_demo_forward_psnr_ssim_match_to_2dp(path):
    import fdgan_b200
    from fdgan_b200 import ops
    old = ops.USE_UMMA
    ops.USE_UMMA = path == "tcgen05_bf16x3"
    try:
        haze = torch.from_numpy(np.load(os.path.join(GOLDEN, "testsample1_3_crop256.npz"))["haze"]).unsqueeze(0)   # [1,3,256,256] in [0,1]
        sd = O.make_fdgan_state(0)
        with torch.no_grad():
            y_ref = O.fdgan_forward(sd, haze, True, False)          # train-mode BatchNorm (README.md:38)
        net = fdgan_b200.FDGAN()
        net.load_state_dict(O.make_fdgan_state(0))
        net = net.cuda().train()
        with torch.no_grad():
            y = net(haze.cuda())
        assert maxabs(y, y_ref) <= 1e-3                              # the north star's output bar
        gt_u8 = M.save_image_u8(haze[0])                             # real-haze samples carry gt == haze
        ref_u8, out_u8 = M.save_image_u8(y_ref[0]), M.save_image_u8(y[0])
        p_ref, p_out = M.psnr(gt_u8, ref_u8), M.psnr(gt_u8, out_u8)
        s_ref, s_out = M.mssim(gt_u8, ref_u8), M.mssim(gt_u8, out_u8)
        print("PSNR ref %.4f ours %.4f | SSIM ref %.4f ours %.4f | differing uint8 pixels %d" %
              (p_ref, p_out, s_ref, s_out, int((ref_u8 != out_u8).sum())))
        assert round(p_ref, 2) == round(p_out, 2) or abs(p_ref - p_out) < 5e-3
        assert round(s_ref, 2) == round(s_out, 2) or abs(s_ref - s_out) < 5e-3
        assert M.psnr(ref_u8, out_u8) > 55.0 or np.array_equal(ref_u8, out_u8)
        # the same pipeline entirely on the GPU (fdgan_b200.metrics: save_image bytes + PSNRSSIM.py arithmetic)
        from fdgan_b200 import metrics
        out_u8_gpu = metrics.save_image_u8(y[0])
        assert np.array_equal(out_u8_gpu.cpu().numpy(), out_u8)
        p_gpu, s_gpu = metrics.psnr_ssim(torch.from_numpy(gt_u8).cuda(), out_u8_gpu)
        assert abs(p_gpu - p_out) <= 1e-9 and abs(s_gpu - s_out) <= 1e-9
    finally:
        ops.USE_UMMA = old
