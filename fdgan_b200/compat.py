"""Make the reference's own import lines resolve to fdgan_b200 without editing the caller.

    import fdgan_b200.compat; fdgan_b200.compat.install()
    import models.dehaze1113 as net            # demo.py:18          -> fdgan_b200.dehaze1113 (net.FDGAN(), net.D(nc, nf))
    from myutils.vgg16 import Vgg16            # loss.pyc@L7         -> fdgan_b200.vgg16
    import models.pytorch_ssim as pytorch_ssim # (training-side SSIM) -> fdgan_b200.pytorch_ssim
    from loss import blur, laplace_filter      # loss.pyc@L161-162, L304 (only with install(loss=True): `loss` is a generic name)

``install()`` registers alias modules in ``sys.modules``; it refuses to shadow a real ``models`` / ``myutils`` package that is
already imported (e.g. the reference itself on sys.path) unless ``force=True``.  ``uninstall()`` removes exactly what it added.
"""
from __future__ import annotations

import sys
import types

_INSTALLED = []


def install(loss: bool = False, force: bool = False) -> None:
    from . import dehaze1113, pytorch_ssim, vgg16
    from . import loss as loss_mod
    wanted = {
        "models": None, "models.dehaze1113": dehaze1113, "models.pytorch_ssim": pytorch_ssim,
        "myutils": None, "myutils.vgg16": vgg16,
    }
    if loss:
        wanted["loss"] = loss_mod
    for name in wanted:
        cur = sys.modules.get(name)
        if cur is not None and name not in _INSTALLED and not force:
            raise RuntimeError("fdgan_b200.compat.install: a module named %r is already imported from %s; pass force=True to shadow it"
                               % (name, getattr(cur, "__file__", "?")))
    for name, target in wanted.items():
        if target is None:
            pkg = types.ModuleType(name)
            pkg.__path__ = []          # a package with no search path: only the aliases registered here resolve
            pkg.__doc__ = "fdgan_b200 alias of the reference's %s package" % name
            sys.modules[name] = pkg
        else:
            sys.modules[name] = target
            parent, _, leaf = name.rpartition(".")
            if parent:
                setattr(sys.modules[parent], leaf, target)
        if name not in _INSTALLED:
            _INSTALLED.append(name)


def uninstall() -> None:
    while _INSTALLED:
        sys.modules.pop(_INSTALLED.pop(), None)
