"""ctypes binding of the fdgan_b200 C ABI (include/fdgan_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
fallback: if the library is missing this module raises at import time, and every entry point raises
``RuntimeError`` with ``fdg_last_error()`` when a call fails.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FDG_LIB") or os.path.join(_HERE, "libfdgan_b200.so")   # FDG_LIB: A/B builds of the same sources

if not os.path.isfile(LIB_PATH):
    raise ImportError(
        "fdgan_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a). There is no CPU or PyTorch fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

GATHER_DIRECT, GATHER_AVGPOOL2, GATHER_UP2 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3
STORE_NORMAL, STORE_UP2, STORE_ACCUM = 0, 1, 2
IMPL_AUTO, IMPL_SIMT, IMPL_UMMA = 0, 1, 2
LOSS_L1, LOSS_MSE, LOSS_BCE = 0, 1, 2

c_float_p = C.c_void_p  # device pointers travel as integers


class FdgTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("sn", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64), ("sc", C.c_int64)]


class FdgConv(C.Structure):
    _fields_ = [
        ("x", FdgTensor), ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int),
        ("gather", C.c_int), ("has_affine", C.c_int), ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("slope", C.c_float), ("w", C.c_void_p), ("w_ld", C.c_int),
        ("R", C.c_int), ("S", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
        ("Cout", C.c_int), ("OH", C.c_int), ("OW", C.c_int), ("bias", C.c_void_p), ("act", C.c_int),
        ("e", FdgTensor), ("eslope", C.c_float), ("y", FdgTensor), ("store", C.c_int),
        ("stats", C.c_void_p), ("stats_ld", C.c_int), ("alpha", C.c_float), ("impl", C.c_int),
        ("w_umma", C.c_void_p), ("e_scale", C.c_void_p), ("e_shift", C.c_void_p), ("w_k1", C.c_void_p), ("x_split", C.c_void_p),
    ]


class FdgWgrad(C.Structure):
    _fields_ = [
        ("x", FdgTensor), ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int),
        ("gather", C.c_int), ("has_affine", C.c_int), ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("slope", C.c_float), ("g", FdgTensor),
        ("R", C.c_int), ("S", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
        ("Cout", C.c_int), ("OH", C.c_int), ("OW", C.c_int),
        ("dw", C.c_void_p), ("transposed", C.c_int), ("dbias", C.c_void_p), ("impl", C.c_int), ("g_split", C.c_void_p),
        ("x_split", C.c_void_p),
    ]


class FdgBnFinalize(C.Structure):
    _fields_ = [
        ("stats", C.c_void_p), ("stats_ld", C.c_int), ("C", C.c_int), ("count", C.c_double),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float), ("momentum", C.c_float),
        ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("training", C.c_int),
        ("scale", C.c_void_p), ("shift", C.c_void_p), ("mean", C.c_void_p), ("invstd", C.c_void_p),
    ]


class FdgEwBwd(C.Structure):
    _fields_ = [
        ("g", FdgTensor), ("g_gather", C.c_int), ("gscale", C.c_float), ("x", FdgTensor),
        ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int),
        ("has_affine", C.c_int), ("scale", C.c_void_p), ("shift", C.c_void_p), ("slope", C.c_float),
        ("coef", C.c_void_p), ("out", FdgTensor), ("accumulate", C.c_int), ("stats", C.c_void_p), ("out_split", C.c_void_p),
    ]


class FdgDepthwise(C.Structure):
    _fields_ = [
        ("x", FdgTensor), ("y", FdgTensor), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
        ("kernel", C.c_void_p), ("l", C.c_int32), ("pad_mode", C.c_int32), ("mean", C.c_void_p), ("inv_std", C.c_void_p),
        ("accumulate", C.c_int32),
    ]


class FdgPackJob(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p), ("kind", C.c_int32), ("cout", C.c_int32), ("cin", C.c_int32), ("r", C.c_int32),
        ("s", C.c_int32), ("ld", C.c_int32), ("first_block", C.c_int32), ("nblocks", C.c_int32), ("total", C.c_int64),
    ]


PACK_UMMA, PACK_K1 = 3, 4


class FdgBnBwdFinalize(C.Structure):
    _fields_ = [
        ("stats", C.c_void_p), ("C", C.c_int), ("count", C.c_double),
        ("gamma", C.c_void_p), ("mean", C.c_void_p), ("invstd", C.c_void_p),
        ("coef", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("accumulate", C.c_int), ("unit_alpha", C.c_int),
        ("acc_beta", C.c_void_p), ("acc_delta", C.c_void_p),
    ]


class FdgDgradStrided(C.Structure):
    _fields_ = [
        ("g", FdgTensor), ("N", C.c_int), ("OH", C.c_int), ("OW", C.c_int), ("Cout", C.c_int),
        ("w", C.c_void_p), ("Cin", C.c_int), ("R", C.c_int), ("S", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
        ("dx", FdgTensor), ("H", C.c_int), ("W", C.c_int), ("accumulate", C.c_int),
    ]


_P = C.POINTER
_SIGS = {
    "fdg_conv2d": ([_P(FdgConv), C.c_void_p], C.c_int),
    "fdg_conv2d_wgrad": ([_P(FdgWgrad), C.c_void_p], C.c_int),
    "fdg_pack_weight": ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p], C.c_int),
    "fdg_umma_weight_bytes": ([C.c_int, C.c_int, C.c_int], C.c_int64),
    "fdg_k1_weight_bytes": ([C.c_int], C.c_int64),
    "fdg_pack_weight_k1": ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p], C.c_int),
    "fdg_pack_weight_umma": ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p], C.c_int),
    "fdg_bn_finalize": ([_P(FdgBnFinalize), C.c_void_p], C.c_int),
    "fdg_ew_bwd": ([_P(FdgEwBwd), C.c_void_p], C.c_int),
    "fdg_bn_bwd_finalize": ([_P(FdgBnBwdFinalize), C.c_void_p], C.c_int),
    "fdg_maxpool2_fwd": ([_P(FdgTensor), _P(FdgTensor), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p], C.c_int),
    "fdg_maxpool2_bwd": ([_P(FdgTensor), _P(FdgTensor), _P(FdgTensor), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p], C.c_int),
    "fdg_copy4d": ([_P(FdgTensor), _P(FdgTensor), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p], C.c_int),
    "fdg_pool2_bn_act": ([_P(FdgTensor), _P(FdgTensor), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p], C.c_int),
    "fdg_tap_sum": ([_P(FdgTensor), _P(FdgTensor), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p], C.c_int),
    "fdg_tap_spread": ([_P(FdgTensor), _P(FdgTensor), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p], C.c_int),
    "fdg_act_bwd": ([C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p], C.c_int),
    "fdg_conv2d_dgrad_strided": ([_P(FdgDgradStrided), C.c_void_p], C.c_int),
    "fdg_affine_accum": ([_P(FdgTensor), _P(FdgTensor), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    "fdg_colsum": ([_P(FdgTensor), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p], C.c_int),
    "fdg_freq_concat_fwd": ([_P(FdgTensor), _P(FdgTensor), C.c_int, C.c_int, C.c_int, C.c_void_p], C.c_int),
    "fdg_freq_concat_bwd": ([_P(FdgTensor), _P(FdgTensor), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p], C.c_int),
    "fdg_event_record": ([C.c_void_p], C.c_int),
    "fdg_stream_wait": ([C.c_void_p, C.c_int], C.c_int),
    "fdg_pack_job_items": ([_P(FdgPackJob)], C.c_int64),
    "fdg_umma_ntile": ([C.c_int, C.c_int], C.c_int),
    "fdg_umma_tile_code": ([C.c_int, C.c_int, C.c_int], C.c_int),
    "fdg_pack_batch": ([C.c_void_p, C.c_int, C.c_int, C.c_void_p], C.c_int),
    "fdg_depthwise2d_fwd": ([_P(FdgDepthwise), C.c_void_p], C.c_int),
    "fdg_depthwise2d_bwd": ([_P(FdgDepthwise), C.c_void_p], C.c_int),
    "fdg_ssim_loss_grad": ([_P(FdgTensor), _P(FdgTensor), C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _P(FdgTensor), C.c_int,
                           C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    "fdg_image_minmax": ([_P(FdgTensor), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p], C.c_int),
    "fdg_image_pack_u8": ([_P(FdgTensor), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    "fdg_psnr_ssim_u8": ([C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p], C.c_int),
    "fdg_adam_flat": ([C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p], C.c_int),
    "fdg_adam_flat_dev": ([C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_float, C.c_void_p], C.c_int),
    "fdg_loss_grad": ([C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int64, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p], C.c_int),
    "fdg_profile_enable": ([C.c_int], C.c_int),
    "fdg_profile_collect": ([C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)], C.c_int),
    "fdg_set_option": ([C.c_char_p, C.c_int], C.c_int),
    "fdg_last_error": ([], C.c_char_p),
    "fdg_version": ([], C.c_int),
    "fdg_launch_count": ([], C.c_int64),
}

EXPORTS = tuple(_SIGS)

for _name, (_args, _res) in _SIGS.items():
    _f = getattr(lib, _name)  # AttributeError here = the .so does not export what the header declares
    _f.argtypes = _args
    _f.restype = _res


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.fdg_last_error()
        raise RuntimeError("fdgan_b200 %s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def launch_count() -> int:
    return int(lib.fdg_launch_count())


PROF_FAMILIES = ("conv_simt_f32", "conv_tcgen05", "wgrad", "ew_bwd", "freq", "other")


def profile_enable(on: bool) -> None:
    lib.fdg_profile_enable(1 if on else 0)


def profile_collect() -> dict:
    n = len(PROF_FAMILIES)
    ms, fl, by, ln = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)(), (C.c_int64 * n)()
    check(lib.fdg_profile_collect(ms, fl, by, ln), "profile_collect")
    return {PROF_FAMILIES[i]: dict(ms=ms[i], flops=fl[i], bytes=by[i], launches=int(ln[i])) for i in range(n)}


def set_option(name: str, value: int) -> None:
    check(lib.fdg_set_option(name.encode(), int(value)), "set_option")
