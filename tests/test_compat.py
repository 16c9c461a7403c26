"""The reference's own import lines resolve to fdgan_b200 after compat.install(): demo.py:18
`import models.dehaze1113  as net`, loss.pyc@L7 `from myutils.vgg16 import Vgg16`, and the training-side `models.pytorch_ssim`."""
import subprocess
import sys
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_lines_work_unchanged():
    import fdgan_b200
    from fdgan_b200 import compat
    compat.install(loss=True)
    try:
        ns = {}
        exec("import models.dehaze1113  as net", ns)                 # demo.py:18, verbatim (two spaces included)
        exec("from myutils.vgg16 import Vgg16", ns)                  # loss.pyc@L7
        exec("import models.pytorch_ssim as pytorch_ssim", ns)
        exec("from loss import blur, laplace_filter, Blur, Laplacian, isotropic_gaussian_kernel", ns)
        assert ns["net"].FDGAN is fdgan_b200.FDGAN and ns["net"].D is fdgan_b200.D and ns["Vgg16"] is fdgan_b200.Vgg16
        assert ns["pytorch_ssim"].ssim is fdgan_b200.pytorch_ssim.ssim and hasattr(ns["pytorch_ssim"], "SSIM")
        assert ns["blur"] is fdgan_b200.blur and ns["laplace_filter"] is fdgan_b200.laplace_filter
        netG = ns["net"].FDGAN()                                      # demo.py:73
        assert len(netG.state_dict()) == 786
    finally:
        compat.uninstall()
    assert "models" not in sys.modules and "myutils" not in sys.modules and "loss" not in sys.modules


def test_install_refuses_to_shadow_a_real_package():
    code = ("import sys, types; sys.modules['models'] = types.ModuleType('models'); import fdgan_b200.compat as c\n"
            "try:\n    c.install()\nexcept RuntimeError as e:\n    print('refused'); \nelse:\n    print('shadowed')\n"
            "c.install(force=True); import models.dehaze1113 as net; print(net.FDGAN.__module__)")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    assert r.stdout.split() == ["refused", "fdgan_b200.dehaze1113"]


def test_demo_style_checkpoint_roundtrip_with_module_prefix():
    """demo.py:78-86 strips the first 7 characters ('module.') of every key of a DataParallel checkpoint; the helper writes such a
    checkpoint and the reference's own stripping loop + load_state_dict accepts it."""
    import torch
    import fdgan_b200
    from fdgan_b200.train import dataparallel_state_dict
    net = fdgan_b200.FDGAN()
    sd = dataparallel_state_dict(net)
    assert all(k.startswith("module.") for k in sd) and len(sd) == 786
    from collections import OrderedDict
    new_state_dict = OrderedDict()
    for k, v in sd.items():
        name = k[7:]      # remove `module.`  (demo.py:83)
        new_state_dict[name] = v
    other = fdgan_b200.FDGAN()
    other.load_state_dict(new_state_dict)
    for (ka, a), (kb, b) in zip(net.state_dict().items(), other.state_dict().items()):
        assert ka == kb and torch.equal(a, b)
