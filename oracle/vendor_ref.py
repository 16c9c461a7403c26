"""Recipe: place the UNMODIFIED reference modules of the hot path under baseline/_ref/ (git-ignored, NOT gpurun-ignored).

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is a directory of scripts (no setup.py / pyproject.toml), so
`pip install --target baseline/_ref /root/reference` has nothing to build; this recipe is that install: a byte-for-byte copy
of the files the path needs, made in the authoring container where /root/reference exists.  baseline/_ref travels to
the GPU box with the snapshot, where `bench.py --impl reference` and the `gpu_baseline` leg import the modules from it
through oracle/ref_import.py (same shims: no weight download, dotted sub-module names).  Nothing is ever copied into
tracked paths, and the product (fdgan_b200/) never imports from here.

    python -m oracle.vendor_ref            # idempotent; prints what it copied
"""
from __future__ import annotations

import filecmp
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("FDGAN_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

# models/dehaze1113.py: FDGAN, D, BottleneckBlockdy, TransitionBlockdy; myutils/vgg16.py: Vgg16; models/pytorch_ssim: ssim
FILES = ("models/__init__.py", "models/dehaze1113.py", "models/densenet.py", "models/pytorch_ssim/__init__.py",
         "myutils/__init__.py", "myutils/vgg16.py", "PSNRSSIM.py")


def vendor(verbose: bool = True) -> bool:
    if not os.path.isfile(os.path.join(SRC, "models", "dehaze1113.py")):
        if verbose:
            print("vendor_ref: %s not present; keeping whatever baseline/_ref already holds" % SRC)
        return os.path.isfile(os.path.join(DST, "models", "dehaze1113.py"))
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.isfile(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
            if verbose:
                print("vendor_ref: %s -> baseline/_ref/%s" % (s, rel))
    return True


if __name__ == "__main__":
    vendor()
