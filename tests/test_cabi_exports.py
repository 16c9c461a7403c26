"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol the header declares,
rejects bad descriptors without touching the GPU, and the host-side module surface mirrors the reference."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from fdgan_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "fdgan_b200.h")).read()
    declared = set(re.findall(r"\b(fdg_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"fdg_stream_t"}
    assert declared, "no declarations parsed"
    raw = ctypes.CDLL(lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), "libfdgan_b200.so does not export %s" % name
    assert declared == set(lib.EXPORTS), (declared ^ set(lib.EXPORTS))
    assert lib.lib.fdg_version() >= 100


def test_ctypes_structs_match_the_header_layout(lib, tmp_path):
    """The header compiles as plain C (no CUDA / torch types) and every descriptor struct has the size and field offsets
    the ctypes binding assumes -- the check a maintainer of another host binding (cgo, JNI ...) would run."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    names = ["FdgTensor", "FdgConv", "FdgWgrad", "FdgBnFinalize", "FdgEwBwd", "FdgBnBwdFinalize", "FdgDgradStrided", "FdgDepthwise", "FdgPackJob"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fdgan_b200.h"', "int main(void) {"]
    for n in names:
        cls = getattr(lib, n)
        lines.append('  printf("%s size %%zu\\n", sizeof(%s));' % (n, n))
        for f in cls._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (n, f[0], n, f[0]))
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.rsplit(" ", 1) for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for n in names:
        cls = getattr(lib, n)
        assert int(got["%s size" % n]) == ctypes.sizeof(cls), n
        for f in cls._fields_:
            assert int(got["%s.%s" % (n, f[0])]) == getattr(cls, f[0]).offset, (n, f[0])


def test_descriptor_validation_needs_no_gpu(lib):
    d = lib.FdgConv()
    rc = lib.lib.fdg_conv2d(ctypes.byref(d), None)
    assert rc == -1
    assert b"null" in lib.lib.fdg_last_error()
    with pytest.raises(RuntimeError):
        lib.check(rc, "conv2d")
    w = lib.FdgWgrad()
    assert lib.lib.fdg_conv2d_wgrad(ctypes.byref(w), None) == -1
    assert lib.lib.fdg_adam_flat(None, None, None, None, 0, 0.0, 0.0, 0.0, 0.0, 0, 1.0, None) == -1


def test_module_surface_matches_reference_state_dicts(lib):
    import fdgan_b200
    from oracle import fdgan_oracle as O
    g = fdgan_b200.FDGAN()
    assert [(k, tuple(v.shape)) for k, v in g.state_dict().items()] == [(n, tuple(s)) for n, s, _k in O.fdgan_specs()]
    assert sum(p.numel() for p in g.parameters()) == 13980691
    used = [n for n, _p in g._used_named_parameters()]
    assert used == O.fdgan_used_param_names()
    for nf in (36, 64):
        d = fdgan_b200.D(9, nf)
        assert [(k, tuple(v.shape)) for k, v in d.state_dict().items()] == [(n, tuple(s)) for n, s, _k in O.d_specs(9, nf)]
    v = fdgan_b200.Vgg16()
    assert [(k, tuple(t.shape)) for k, t in v.state_dict().items()] == [(n, tuple(s)) for n, s, _k in O.vgg_specs()]


def test_no_cpu_fallback(lib):
    import fdgan_b200
    net = fdgan_b200.FDGAN()
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError):
        fdgan_b200.freq_concat(torch.zeros(1, 3, 32, 32))
    from fdgan_b200 import metrics
    with pytest.raises(RuntimeError):
        metrics.save_image_u8(torch.zeros(3, 32, 32))
    with pytest.raises(RuntimeError):
        metrics.psnr_ssim(torch.zeros(32, 32, 3, dtype=torch.uint8), torch.zeros(32, 32, 3, dtype=torch.uint8))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fdgan_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn
