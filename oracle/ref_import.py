"""Import the UNMODIFIED reference modules from /root/reference (authoring container) or from the byte-for-byte
copy under baseline/_ref (GPU box; oracle/vendor_ref.py).

TEST / BASELINE INFRASTRUCTURE ONLY: tests, oracle/make_golden.py and bench.py's baseline legs.  Shims, none of which edit the reference:
  * torchvision.models.densenet121(pretrained=True) would download weights
    (models/dehaze1113.py:707) -> patched to build the architecture with weights=None;
  * class D registers sub-modules with dotted names ('layer1.conv',
    models/dehaze1113.py:196) which torch >= 1.x rejects -> nn.Module.add_module is
    wrapped to accept them while the reference constructors run.
"""
from __future__ import annotations

import contextlib
import os
import sys

import torch
import torch.nn as nn

_VENDORED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _find_root() -> str:
    """/root/reference in the authoring container; on the GPU box the byte-for-byte copy oracle/vendor_ref.py placed under
    baseline/_ref (git-ignored, travels with the snapshot)."""
    env = os.environ.get("FDGAN_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile(os.path.join("/root/reference", "models", "dehaze1113.py")):
        return "/root/reference"
    return _VENDORED


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "dehaze1113.py"))


@contextlib.contextmanager
def _shims():
    import torchvision.models as tvm

    orig_densenet = tvm.densenet121
    orig_add = nn.Module.add_module

    def densenet121(pretrained=False, **kw):
        return orig_densenet(weights=None, **kw)

    def add_module(self, name, module):
        if "." in name:
            # what torch 0.3 did: plain registration, no name validation
            self._modules[name] = module
            return
        return orig_add(self, name, module)

    tvm.densenet121 = densenet121
    nn.Module.add_module = add_module
    sys.path.insert(0, REF_ROOT)
    try:
        yield
    finally:
        sys.path.remove(REF_ROOT)
        tvm.densenet121 = orig_densenet
        nn.Module.add_module = orig_add


def _import(name):
    import importlib

    with _shims():
        return importlib.import_module(name)


def ref_fdgan():
    with _shims():
        import models.dehaze1113 as net  # noqa: the reference module
        return net.FDGAN()


def ref_d(nc, nf):
    with _shims():
        import models.dehaze1113 as net
        return net.D(nc, nf)


def ref_vgg16():
    with _shims():
        from myutils.vgg16 import Vgg16
        return Vgg16()


def ref_ssim():
    with _shims():
        import models.pytorch_ssim as pytorch_ssim
        return pytorch_ssim.ssim


def load_state(module, sd):
    """strict load of a reference-keyed state dict (demo.py:86 semantics).  torch 2.x's
    strict key check mis-reports the dotted sub-module names of D as unexpected, so key
    equality is asserted here and the copy itself runs non-strict."""
    own = module.state_dict()
    assert set(own.keys()) == set(sd.keys()), (set(own) ^ set(sd))
    module.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=False)
    for k, v in module.state_dict().items():
        assert torch.equal(v, sd[k]), k
    return module
