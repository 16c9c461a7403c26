"""Typed Python wrappers over the C ABI (one function per entry point of include/fdgan_b200.h).

PyTorch is used for device memory and streams only: every function enqueues hand-written sm_100a
kernels on ``torch.cuda.current_stream()``; nothing here computes with torch ops.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import threading
import weakref

import torch

from . import _lib as L
from ._lib import (ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, GATHER_AVGPOOL2, GATHER_DIRECT, GATHER_UP2,  # noqa: F401
                   IMPL_AUTO, IMPL_SIMT, IMPL_UMMA, STORE_ACCUM, STORE_NORMAL, STORE_UP2)

_byref = C.byref


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t) -> int | None:
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


class View:
    """A strided fp32 [N,H,W,C] view of device memory (element strides sn, sh, sw, sc).  ``base`` keeps the
    owning torch tensor alive."""

    __slots__ = ("base", "ptr", "N", "H", "W", "C", "sn", "sh", "sw", "sc")

    def __init__(self, base, ptr, N, H, W, Cc, sn, sh, sw, sc):
        self.base, self.ptr = base, ptr
        self.N, self.H, self.W, self.C = N, H, W, Cc
        self.sn, self.sh, self.sw, self.sc = sn, sh, sw, sc

    @staticmethod
    def nhwc(buf: torch.Tensor, N, H, W, Ctot) -> "View":
        assert buf.dtype == torch.float32 and buf.is_cuda and buf.is_contiguous()
        assert buf.numel() >= N * H * W * Ctot, (buf.numel(), N, H, W, Ctot)
        return View(buf, buf.data_ptr(), N, H, W, Ctot, H * W * Ctot, W * Ctot, Ctot, 1)

    @staticmethod
    def alloc(N, H, W, Ctot, device, zero=False) -> "View":
        f = torch.zeros if zero else torch.empty
        return View.nhwc(f(N * H * W * Ctot, dtype=torch.float32, device=device), N, H, W, Ctot)

    @staticmethod
    def from_nchw(t: torch.Tensor) -> "View":
        """A logical-NCHW torch tensor with arbitrary strides (contiguous NCHW, channels_last, slices ...)."""
        assert t.dim() == 4 and t.dtype == torch.float32 and t.is_cuda
        n, c, h, w = t.shape
        sn, sc, sh, sw = t.stride()
        return View(t, t.data_ptr(), n, h, w, c, sn, sh, sw, sc)

    def ch(self, c0, c1) -> "View":
        assert 0 <= c0 < c1 <= self.C
        return View(self.base, self.ptr + 4 * c0 * self.sc, self.N, self.H, self.W, c1 - c0, self.sn, self.sh, self.sw, self.sc)

    def ft(self) -> L.FdgTensor:
        return L.FdgTensor(self.ptr, self.sn, self.sh, self.sw, self.sc)

    def as_nchw(self) -> torch.Tensor:
        """torch view with logical shape [N,C,H,W] over the same memory (no copy)."""
        off = (self.ptr - self.base.data_ptr()) // 4
        return torch.as_strided(self.base, (self.N, self.C, self.H, self.W), (self.sn, self.sc, self.sh, self.sw),
                                self.base.storage_offset() + off)


_NULLT = L.FdgTensor(None, 0, 0, 0, 0)

# tcgen05 path switch: True = every eligible convolution runs on the tensor cores (bf16x3 split, fp32 accumulate);
# False = fp32 SIMT everywhere (used by tests to compare the two paths).
USE_UMMA = True
USE_K1 = True      # 3x3 / stride 1 / pad 1 / Cout <= 32 convolutions on the filter-row-concatenated kernel (conv_k1.cu)


def umma_eligible(x: View, Cout: int, R: int = 1, S: int = 1, stride: int = 1, gather: int = GATHER_DIRECT) -> bool:
    halo = stride == 1 and 2 <= R <= 4 and 2 <= S <= 4 and gather == GATHER_DIRECT   # conv_halo.cu takes Cin % 4 == 0
    return ((x.C % 8 == 0 or (halo and x.C % 4 == 0)) and x.C >= 16 and Cout >= 1 and x.sc == 1 and x.ptr % 16 == 0 and
            x.sn % 4 == 0 and x.sh % 4 == 0 and x.sw % 4 == 0)


def k1_eligible(x: View, Cout: int, R: int, S: int, stride: int, pad: int, gather: int) -> bool:
    return (USE_K1 and R == 3 and S == 3 and stride == 1 and pad == 1 and gather == GATHER_DIRECT and 1 <= Cout <= 32 and
            x.C % 4 == 0 and x.C >= 16 and x.sc == 1 and x.ptr % 16 == 0 and x.sn % 4 == 0 and x.sh % 4 == 0 and x.sw % 4 == 0)


def pack_weight_k1(w, w_ld: int, Cin: int, Cout: int, device) -> torch.Tensor:
    """fdg_pack_weight_k1: [W(ky=0) | W(ky=1) | W(ky=2)] bf16 hi/lo tiles per (64-channel chunk, filter column)."""
    out = torch.empty(int(L.lib.fdg_k1_weight_bytes(Cin)) // 4, dtype=torch.int32, device=device)
    L.check(L.lib.fdg_pack_weight_k1(_ptr(w), w_ld, Cin, Cout, out.data_ptr(), _stream()), "pack_weight_k1")
    return out


def pack_weight_umma(w, w_ld: int, taps: int, Cin: int, Cout: int, device) -> torch.Tensor:
    """fdg_pack_weight_umma: bf16 hi/lo, pre-swizzled per-stage images of the GEMM-form weight w[K][w_ld]."""
    nbytes = int(L.lib.fdg_umma_weight_bytes(taps, Cin, Cout))
    out = torch.empty(nbytes // 4, dtype=torch.int32, device=device)
    L.check(L.lib.fdg_pack_weight_umma(_ptr(w), w_ld, taps, Cin, Cout, out.data_ptr(), _stream()), "pack_weight_umma")
    return out


def conv2d(x: View, w, w_ld, R, S, stride, pad, Cout, y: View, *, gather=GATHER_DIRECT, scale=None, shift=None,
           slope=1.0, bias=None, act=ACT_NONE, e: View | None = None, eslope=0.0, store=STORE_NORMAL, stats=None,
           stats_ld=0, alpha=1.0, impl=IMPL_AUTO, w_umma=None, e_scale=None, e_shift=None, w_k1=None, x_split=None):
    """fdg_conv2d.  ``w`` is the packed [K][w_ld] operand (tensor or pointer)."""
    if gather == GATHER_AVGPOOL2:
        H, W = x.H // 2, x.W // 2
    elif gather == GATHER_UP2:
        H, W = 2 * x.H, 2 * x.W
    else:
        H, W = x.H, x.W
    OH = (H + 2 * pad - R) // stride + 1
    OW = (W + 2 * pad - S) // stride + 1
    mul = 2 if store == STORE_UP2 else 1
    if (y.N, y.H, y.W, y.C) != (x.N, mul * OH, mul * OW, Cout):
        raise ValueError("conv2d: output view %s does not match %s" % ((y.N, y.H, y.W, y.C), (x.N, mul * OH, mul * OW, Cout)))
    if e is not None and (e.N, e.H, e.W, e.C) != (x.N, OH, OW, Cout):
        raise ValueError("conv2d: mask view shape mismatch")
    if w_umma is None and w_k1 is None and impl != IMPL_SIMT and (USE_UMMA or impl == IMPL_UMMA):
        ikey = None
        if k1_eligible(x, Cout, R, S, stride, pad, gather):
            ikey = ("k1", x.C, Cout)
        elif umma_eligible(x, Cout, R, S, stride, gather):
            ikey = ("umma", R * S, x.C, Cout)
        if ikey is not None:
            images = getattr(w, "_fdg_images", None)      # packed operand of a frozen parameter: its tcgen05 images are cached
            plan = getattr(_TLS, "plan", None) if images is None else None
            img = images.get(ikey) if images is not None else (plan.image(_ptr(w), w_ld, ikey) if plan is not None else None)
            if img is None:
                img = (pack_weight_k1(w, w_ld, x.C, Cout, x.base.device) if ikey[0] == "k1" else
                       pack_weight_umma(w, w_ld, R * S, x.C, Cout, x.base.device))
                if images is not None:
                    images[ikey] = img
                elif plan is not None:
                    plan.add_image(_ptr(w), w_ld, ikey, img)
            if ikey[0] == "k1":
                w_k1 = img
            else:
                w_umma = img
    d = L.FdgConv(
        x.ft(), x.N, H, W, x.C, gather, 1 if scale is not None else 0, _ptr(scale), _ptr(shift), slope,
        _ptr(w), w_ld, R, S, stride, pad, Cout, OH, OW, _ptr(bias), act,
        e.ft() if e is not None else _NULLT, eslope, y.ft(), store, _ptr(stats), stats_ld, alpha, impl, _ptr(w_umma),
        _ptr(e_scale), _ptr(e_shift), _ptr(w_k1), _ptr(x_split))
    L.check(L.lib.fdg_conv2d(_byref(d), _stream()), "conv2d")


def event_record(stream: int) -> int:
    """fdg_event_record: a point on ``stream`` (raw cudaStream_t) as a small integer handle."""
    h = int(L.lib.fdg_event_record(stream))
    if h < 0:
        L.check(h, "event_record")
    return h


def stream_wait(stream: int, handle: int) -> None:
    L.check(L.lib.fdg_stream_wait(stream, handle), "stream_wait")


def wgrad(x: View, g: View, R, S, stride, pad, dw, *, gather=GATHER_DIRECT, scale=None, shift=None, slope=1.0,
          transposed=False, dbias=None, impl=None, g_split=None, x_split=None, stream=None):
    """fdg_conv2d_wgrad: dw (+)= A^T g in the parameter's own layout (dw must be pre-zeroed or accumulating).
    ``dw is None`` (frozen parameter) skips the launch."""
    if dw is None:
        return
    if gather == GATHER_AVGPOOL2:
        H, W = x.H // 2, x.W // 2
    elif gather == GATHER_UP2:
        H, W = 2 * x.H, 2 * x.W
    else:
        H, W = x.H, x.W
    OH = (H + 2 * pad - R) // stride + 1
    OW = (W + 2 * pad - S) // stride + 1
    if (g.N, g.H, g.W) != (x.N, OH, OW):
        raise ValueError("wgrad: gradient view %s does not match output extent %s" % ((g.N, g.H, g.W), (x.N, OH, OW)))
    d = L.FdgWgrad(x.ft(), x.N, H, W, x.C, gather, 1 if scale is not None else 0, _ptr(scale), _ptr(shift), slope,
                   g.ft(), R, S, stride, pad, g.C, OH, OW, _ptr(dw), 1 if transposed else 0, _ptr(dbias),
                   (IMPL_AUTO if USE_UMMA else IMPL_SIMT) if impl is None else impl, _ptr(g_split), _ptr(x_split))
    L.check(L.lib.fdg_conv2d_wgrad(_byref(d), _stream() if stream is None else stream), "conv2d_wgrad")


def split_planes(x: View, slope=1.0) -> torch.Tensor:
    """act(x) as split-bf16 planes [pixels][C] (hi plane, then lo plane; the bytes of an fp32 tensor of the same extent): the operand
    format the tensor-core kernels take through bulk tensor loads.  One element-wise pass (fdg_ew_bwd with the tensor as its own mask:
    x * [x > 0 ? 1 : slope])."""
    out = torch.empty(x.N * x.H * x.W * x.C, dtype=torch.float32, device=x.base.device)
    ew_bwd(x, x, slope=slope, out_split=out)
    return out


def wgrad_planes_ok(x: View, g: View, R, S, stride, pad) -> bool:
    """Shapes the all-planes weight-gradient path takes (FdgWgrad.x_split)."""
    OW = x.W + 2 * pad - S + 1
    lin = lambda v: v.sc == 1 and v.sh == v.W * v.sw and v.sn == v.H * v.sh
    return (USE_UMMA and stride == 1 and R >= 2 and S >= 2 and OW % 32 == 0 and x.C % 8 == 0 and g.C % 8 == 0 and 64 < g.C and
            (g.C <= 256 or g.C % 128 == 0) and lin(x) and lin(g))


# Operand images of FROZEN parameters (requires_grad == False: the Vgg16 feature extractor, D during the generator update
# of an autograd user) are packed once and reused until the parameter changes (torch bumps ``_version`` on in-place
# writes; the training step's own optimiser never touches frozen parameters).
_FROZEN = {}


def _frozen_lookup(w: torch.Tensor, mode: int):
    if w.requires_grad or not isinstance(w, torch.nn.Parameter):
        return None, None
    key = (w.data_ptr(), tuple(w.shape), mode)
    hit = _FROZEN.get(key)
    if hit is not None and hit[0] == w._version and hit[3]() is w:      # same live parameter object, unchanged
        return key, hit
    return key, None


def drop_frozen_in_range(lo: int, hi: int) -> int:
    """Forget cached operand images of parameters stored in [lo, hi) (device addresses).  The fused flat Adam writes
    parameters through raw pointers, which does not bump ``_version``: a parameter that is frozen for one phase and updated
    by ``FlatState.adam`` in another (an autograd user's D during the generator update) must not keep a stale image.
    Writes through ``param.data`` bypass the version counter too: call this (or re-create the parameter) after such writes."""
    stale = [k for k in list(_FROZEN) if lo <= k[0] < hi]      # snapshot: DataParallel worker threads may insert concurrently
    for k in stale:
        _FROZEN.pop(k, None)
    return len(stale)


# ----------------------------------------------------------------------------------------------------------------------
# Pack plans: all operand repacks of one network pass in one launch per dependency level (fdg_pack_batch)
# ----------------------------------------------------------------------------------------------------------------------
USE_PACK_PLAN = True
_TLS = threading.local()      # .plan: the PackPlan of the network pass running on this thread (nn.DataParallel workers are threads)


class PackPlan:
    """Operand images of the TRAINABLE parameters one network pass reads.  Parameters keep the PyTorch layout and change every
    optimiser step, so their GEMM-operand / tcgen05 images are rebuilt per pass.  The first pass under a plan records every repack
    (and runs it as its own launch); later passes refresh all of them with ONE launch per level -- level 0: fp32 GEMM operands made
    from parameters, level 1: bf16 hi/lo images made from level-0 operands or directly from 1x1 parameters -- and the pack calls
    return the plan's persistent tensors.  A lookup that misses (other shapes, re-homed parameters) is served by an individual
    launch and makes the plan re-record on its next pass."""

    ITEMS_PER_BLOCK = 1024      # 256 threads x 4 items

    def __init__(self):
        self.reset()

    def reset(self):
        self.jobs = ([], [])
        self.packed = {}
        self.images = {}
        self.tables = [None, None]
        self.ready = False
        self.broken = False

    # -- level 0
    def packed_operand(self, w: torch.Tensor, mode: int):
        k = (w.data_ptr(), tuple(w.shape), mode)
        hit = self.packed.get(k)
        if hit is not None:
            return hit
        out, ld = _pack_weight(w, mode)
        if self.ready:
            self.broken = True
            return out, ld
        if mode == 2:
            cin, cout, R, S = w.shape
        else:
            cout, cin, R, S = w.shape
        self.jobs[0].append(L.FdgPackJob(w.data_ptr(), out.data_ptr(), mode, cout, cin, R, S, ld, 0, 0, 0))
        self.packed[k] = (out, ld)
        return out, ld

    # -- level 1
    def image(self, wptr: int, w_ld: int, ikey):
        return self.images.get((wptr, w_ld, ikey))

    def add_image(self, wptr: int, w_ld: int, ikey, img: torch.Tensor):
        if self.ready:
            self.broken = True
            return
        if ikey[0] == "k1":
            _, cin, cout = ikey
            job = L.FdgPackJob(wptr, img.data_ptr(), L.PACK_K1, cout, cin, 3, 3, w_ld, 0, 0, 0)
        else:
            _, taps, cin, cout = ikey
            job = L.FdgPackJob(wptr, img.data_ptr(), L.PACK_UMMA, cout, cin, taps, int(L.lib.fdg_umma_tile_code(taps, cin, cout)), w_ld, 0, 0, 0)
        self.jobs[1].append(job)
        self.images[(wptr, w_ld, ikey)] = img

    def finalize(self, device):
        for lvl in (0, 1):
            jobs = self.jobs[lvl]
            if not jobs:
                continue
            arr = (L.FdgPackJob * len(jobs))()
            blk = 0
            for i, j in enumerate(jobs):
                total = int(L.lib.fdg_pack_job_items(C.byref(j)))
                if total <= 0:
                    raise RuntimeError("fdgan_b200: malformed pack job (kind %d, cout %d, cin %d)" % (j.kind, j.cout, j.cin))
                j.total, j.first_block = total, blk
                j.nblocks = (total + self.ITEMS_PER_BLOCK - 1) // self.ITEMS_PER_BLOCK
                blk += j.nblocks
                arr[i] = j
            table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
            self.tables[lvl] = (table, len(jobs), blk)
        self.ready = True

    def refresh(self):
        st = _stream()
        for lvl in (0, 1):
            if self.tables[lvl] is not None:
                table, n, blocks = self.tables[lvl]
                L.check(L.lib.fdg_pack_batch(table.data_ptr(), n, blocks, st), "pack_batch")


@contextlib.contextmanager
def pack_scope(owner, name: str, device):
    """Run one network pass (``name``: 'fwd' / 'bwd') of module ``owner`` under its pack plan."""
    if not USE_PACK_PLAN or getattr(owner, "_is_replica", False):
        yield None
        return
    plans = owner.__dict__.setdefault("_fdg_pack_plans", {})
    key = (name, USE_UMMA, USE_K1, str(device))      # the conv path switches change which images a pass needs
    plan = plans.get(key)
    if plan is None:
        plan = plans[key] = PackPlan()
    prev = getattr(_TLS, "plan", None)
    _TLS.plan = plan
    ok = False
    try:
        if plan.ready:
            plan.refresh()
        yield plan
        ok = True
    finally:
        _TLS.plan = prev
        if not ok or plan.broken:
            plan.reset()
        elif not plan.ready:
            plan.finalize(device)


def invalidate_frozen() -> int:
    """Forget EVERY cached operand image.  Call after writing frozen parameters through ``param.data`` (``p.data.copy_``, EMA
    updates, a custom optimiser): such writes do not bump ``_version``, which is the only staleness signal the cache has."""
    n = len(_FROZEN)
    _FROZEN.clear()
    return n


_CAPTURE_KEEPALIVE = None      # list that collects every cached operand handed out while a CUDA graph is being captured


def pack_weight(w: torch.Tensor, mode: int) -> tuple[torch.Tensor, int]:
    """fdg_pack_weight.  mode 0: OIHW -> [(r,s,ci)][co]; 1: OIHW -> flipped [(r,s,co)][ci]; 2: [Cin][Cout] -> [co][ci]."""
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
    key, hit = _frozen_lookup(w, mode)
    if hit is not None:
        if _CAPTURE_KEEPALIVE is not None:      # a captured graph bakes in the raw pointers of the operand AND its tcgen05 images
            _CAPTURE_KEEPALIVE.append(hit[1])
        return hit[1], hit[2]
    plan = getattr(_TLS, "plan", None) if key is None else None
    if plan is not None:
        return plan.packed_operand(w, mode)
    out, ld = _pack_weight(w, mode)
    if key is not None:
        if len(_FROZEN) > 4096:
            _FROZEN.clear()
        out._fdg_images = {}          # tcgen05 images of this packed operand, filled by conv2d
        _FROZEN[key] = (w._version, out, ld, weakref.ref(w))
        if _CAPTURE_KEEPALIVE is not None:
            _CAPTURE_KEEPALIVE.append(out)
    return out, ld


def _pack_weight(w: torch.Tensor, mode: int) -> tuple[torch.Tensor, int]:
    if mode == 2:
        cin, cout, R, S = w.shape
    else:
        cout, cin, R, S = w.shape
    K = R * S * cin if mode == 0 else (R * S * cout if mode == 1 else cout)
    ncols = cout if mode == 0 else cin
    ld = (ncols + 3) // 4 * 4
    out = torch.empty(K * ld, dtype=torch.float32, device=w.device)
    L.check(L.lib.fdg_pack_weight(w.data_ptr(), cout, cin, R, S, mode, out.data_ptr(), ld, _stream()), "pack_weight")
    return out, ld


def bn_finalize(stats, stats_ld, Cc, count, gamma, beta, eps, momentum, running_mean, running_var, training,
                scale, shift, mean=None, invstd=None):
    d = L.FdgBnFinalize(_ptr(stats), stats_ld, Cc, float(count), _ptr(gamma), _ptr(beta), eps, momentum,
                        _ptr(running_mean), _ptr(running_var), 1 if training else 0, _ptr(scale), _ptr(shift),
                        _ptr(mean), _ptr(invstd))
    L.check(L.lib.fdg_bn_finalize(_byref(d), _stream()), "bn_finalize")


def ew_bwd(g: View, x: View, *, out: View | None = None, stats=None, g_gather=GATHER_DIRECT, gscale=1.0, scale=None,
           shift=None, slope=0.0, coef=None, accumulate=False, out_split=None):
    if g_gather == GATHER_UP2:
        assert (g.N, g.H, g.W, g.C) == (x.N, x.H // 2, x.W // 2, x.C), "ew_bwd: pooled gradient extent mismatch"
    else:
        assert (g.N, g.H, g.W, g.C) == (x.N, x.H, x.W, x.C), "ew_bwd: gradient extent mismatch"
    if out is not None:
        assert (out.N, out.H, out.W, out.C) == (x.N, x.H, x.W, x.C)
    d = L.FdgEwBwd(g.ft(), g_gather, gscale, x.ft(), x.N, x.H, x.W, x.C, 1 if scale is not None else 0,
                   _ptr(scale), _ptr(shift), slope, _ptr(coef), out.ft() if out is not None else _NULLT,
                   1 if accumulate else 0, _ptr(stats), _ptr(out_split))
    L.check(L.lib.fdg_ew_bwd(_byref(d), _stream()), "ew_bwd")


def affine_accum(x: View, out: View, cb, cd):
    """fdg_affine_accum: out += cb[c] * x + cd[c] (deferred affine part of the BatchNorm backward)."""
    assert (x.N, x.H, x.W, x.C) == (out.N, out.H, out.W, out.C)
    xt, ot = x.ft(), out.ft()
    L.check(L.lib.fdg_affine_accum(_byref(xt), _byref(ot), x.N, x.H, x.W, x.C, _ptr(cb), _ptr(cd), _stream()), "affine_accum")


def bn_bwd_finalize(stats, Cc, count, gamma, mean, invstd, coef, dgamma=None, dbeta=None, accumulate=True, unit_alpha=False,
                    acc_beta=None, acc_delta=None):
    d = L.FdgBnBwdFinalize(_ptr(stats), Cc, float(count), _ptr(gamma), _ptr(mean), _ptr(invstd), _ptr(coef),
                           _ptr(dgamma), _ptr(dbeta), 1 if accumulate else 0, 1 if unit_alpha else 0, _ptr(acc_beta), _ptr(acc_delta))
    L.check(L.lib.fdg_bn_bwd_finalize(_byref(d), _stream()), "bn_bwd_finalize")


def maxpool2_fwd(x: View, y: View):
    assert (y.N, y.H, y.W, y.C) == (x.N, x.H // 2, x.W // 2, x.C)
    xt, yt = x.ft(), y.ft()
    L.check(L.lib.fdg_maxpool2_fwd(_byref(xt), _byref(yt), y.N, y.H, y.W, y.C, _stream()), "maxpool2_fwd")


def maxpool2_bwd(x: View, gy: View, gx: View, accumulate=True, relu_mask=False):
    """gx (=|+=) the pooled gradient gy routed to the first maximum of every 2x2 block of x; ``relu_mask``: times [max > 0]
    (x is a post-ReLU tensor: the mask of its ReLU comes for free).  Without ``accumulate`` the 2x2 blocks are written in full,
    i.e. all of gx when H and W are even."""
    assert (gy.N, gy.H, gy.W, gy.C) == (x.N, x.H // 2, x.W // 2, x.C) and (gx.H, gx.W, gx.C) == (x.H, x.W, x.C)
    xt, gyt, gxt = x.ft(), gy.ft(), gx.ft()
    L.check(L.lib.fdg_maxpool2_bwd(_byref(xt), _byref(gyt), _byref(gxt), gy.N, gy.H, gy.W, gy.C,
                                   (1 if accumulate else 0) | (2 if relu_mask else 0), _stream()), "maxpool2_bwd")


def copy4d(x: View, y: View, *, gather=GATHER_DIRECT, slope=1.0, scale=1.0, accumulate=False):
    if gather == GATHER_AVGPOOL2:
        exp = (x.N, x.H // 2, x.W // 2, x.C)
    elif gather == GATHER_UP2:
        exp = (x.N, 2 * x.H, 2 * x.W, x.C)
    else:
        exp = (x.N, x.H, x.W, x.C)
    assert (y.N, y.H, y.W, y.C) == exp, ("copy4d extent mismatch", (y.N, y.H, y.W, y.C), exp)
    xt, yt = x.ft(), y.ft()
    L.check(L.lib.fdg_copy4d(_byref(xt), _byref(yt), y.N, y.H, y.W, y.C, gather, slope, scale,
                             1 if accumulate else 0, _stream()), "copy4d")


def pool2_bn_act(x: View, y: View, scale=None, shift=None, slope=1.0):
    """fdg_pool2_bn_act: y = avg_pool2(leaky(x * scale + shift, slope))."""
    assert (y.N, y.H, y.W, y.C) == (x.N, x.H // 2, x.W // 2, x.C)
    xt, yt = x.ft(), y.ft()
    L.check(L.lib.fdg_pool2_bn_act(_byref(xt), _byref(yt), y.N, y.H, y.W, y.C, _ptr(scale), _ptr(shift), slope, _stream()), "pool2_bn_act")


def tap_sum(s: View, out: View, R, S, pad, act=ACT_NONE):
    """fdg_tap_sum: out(n,oy,ox) = act(sum over taps t of s(n, oy+ky-pad, ox+kx-pad)[t])."""
    assert s.C == R * S and out.C == 1 and (out.N, out.H, out.W) == (s.N, s.H + 2 * pad - R + 1, s.W + 2 * pad - S + 1)
    st, ot = s.ft(), out.ft()
    L.check(L.lib.fdg_tap_sum(_byref(st), _byref(ot), s.N, s.H, s.W, R, S, pad, act, _stream()), "tap_sum")


def tap_spread(g: View, gs: View, R, S, pad):
    """fdg_tap_spread: gs(n,y,x)[t] = g(n, y-ky+pad, x-kx+pad)."""
    assert gs.C == R * S and g.C == 1 and (g.N, g.H, g.W) == (gs.N, gs.H + 2 * pad - R + 1, gs.W + 2 * pad - S + 1)
    gt, st = g.ft(), gs.ft()
    L.check(L.lib.fdg_tap_spread(_byref(gt), _byref(st), gs.N, gs.H, gs.W, R, S, pad, _stream()), "tap_spread")


def act_bwd(g: torch.Tensor, y: torch.Tensor, out: torch.Tensor, act: int):
    assert g.is_contiguous() and y.is_contiguous() and out.is_contiguous() and g.numel() == y.numel() == out.numel()
    L.check(L.lib.fdg_act_bwd(g.data_ptr(), y.data_ptr(), out.data_ptr(), g.numel(), act, _stream()), "act_bwd")


def dgrad_strided(g: View, w: torch.Tensor, stride, pad, dx: View, accumulate=False):
    cout, cin, R, S = w.shape
    assert g.C == cout and dx.C == cin and w.is_contiguous()
    d = L.FdgDgradStrided(g.ft(), g.N, g.H, g.W, cout, w.data_ptr(), cin, R, S, stride, pad, dx.ft(), dx.H, dx.W,
                          1 if accumulate else 0)
    L.check(L.lib.fdg_conv2d_dgrad_strided(_byref(d), _stream()), "conv2d_dgrad_strided")


def colsum(x: View, out: torch.Tensor, accumulate=False):
    xt = x.ft()
    L.check(L.lib.fdg_colsum(_byref(xt), x.N, x.H, x.W, x.C, out.data_ptr(), 1 if accumulate else 0, _stream()), "colsum")


def freq_concat_fwd(x: View, z: View):
    assert x.C == 3 and z.C == 9 and (x.N, x.H, x.W) == (z.N, z.H, z.W)
    xt, zt = x.ft(), z.ft()
    L.check(L.lib.fdg_freq_concat_fwd(_byref(xt), _byref(zt), x.N, x.H, x.W, _stream()), "freq_concat_fwd")


def freq_concat_bwd(dz: View, dx: View, scratch: torch.Tensor):
    assert dz.C == 9 and dx.C == 3 and scratch.numel() >= dx.N * 3 * dx.H * dx.W
    a, b = dz.ft(), dx.ft()
    L.check(L.lib.fdg_freq_concat_bwd(_byref(a), _byref(b), scratch.data_ptr(), dx.N, dx.H, dx.W, _stream()), "freq_concat_bwd")


def depthwise2d(x: View, y: View, kernel: torch.Tensor, pad_mode: int, mean=None, inv_std=None, accumulate=False, backward=False):
    """fdg_depthwise2d_fwd / _bwd: one l x l kernel on every (image, channel) plane; pad_mode 0 zero, 1 reflect.
    backward=True: x is the incoming gradient, y accumulates the gradient w.r.t. the filter input (must be initialised)."""
    assert (x.N, x.H, x.W, x.C) == (y.N, y.H, y.W, y.C)
    assert kernel.is_cuda and kernel.dtype == torch.float32 and kernel.is_contiguous() and kernel.dim() == 2 and kernel.shape[0] == kernel.shape[1]
    d = L.FdgDepthwise(x.ft(), y.ft(), x.N, x.H, x.W, x.C, kernel.data_ptr(), kernel.shape[0], pad_mode, _ptr(mean), _ptr(inv_std),
                       1 if accumulate else 0)
    fn = L.lib.fdg_depthwise2d_bwd if backward else L.lib.fdg_depthwise2d_fwd
    L.check(fn(_byref(d), _stream()), "depthwise2d_bwd" if backward else "depthwise2d_fwd")


def adam_flat(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, grad_scale=1.0):
    n = param.numel()
    assert grad.numel() == n and exp_avg.numel() == n and exp_avg_sq.numel() == n
    L.check(L.lib.fdg_adam_flat(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), n,
                                lr, beta1, beta2, eps, step, grad_scale, _stream()), "adam_flat")


def adam_flat_dev(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, state, grad_scale=1.0):
    """fdg_adam_flat_dev: step counter / bias corrections live in ``state`` (3 floats on the device)."""
    n = param.numel()
    assert grad.numel() == n and exp_avg.numel() == n and exp_avg_sq.numel() == n and state.numel() >= 3
    L.check(L.lib.fdg_adam_flat_dev(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), n,
                                    lr, beta1, beta2, eps, state.data_ptr(), grad_scale, _stream()), "adam_flat_dev")


LOSS_L1, LOSS_MSE, LOSS_BCE = L.LOSS_L1, L.LOSS_MSE, L.LOSS_BCE


def ssim_loss_grad(x: View, y: View, lscale: float, gscale: float, loss: torch.Tensor, grad: View | None = None, accumulate=False):
    """fdg_ssim_loss_grad: loss += lscale * sum(ssim_map(x, y)); grad (=|+=) gscale * d(sum ssim_map)/dx."""
    assert loss.dtype == torch.float64 and loss.numel() == 1
    assert (x.N, x.H, x.W, x.C) == (y.N, y.H, y.W, y.C)
    scratch = torch.empty(3 * x.N * x.C * x.H * x.W, dtype=torch.float32, device=x.base.device)
    xt, yt = x.ft(), y.ft()
    gt = grad.ft() if grad is not None else None
    L.check(L.lib.fdg_ssim_loss_grad(_byref(xt), _byref(yt), x.N, x.H, x.W, x.C, lscale, gscale, _byref(gt) if gt is not None else None,
                                     1 if accumulate else 0, loss.data_ptr(), scratch.data_ptr(), _stream()), "ssim_loss_grad")


def loss_grad(kind, a: torch.Tensor, b, n, scale, loss: torch.Tensor, grad=None, accumulate=False, target=0.0):
    """fdg_loss_grad over the first ``n`` floats of the (identically laid out) buffers a and b."""
    assert loss.dtype == torch.float64 and loss.numel() == 1
    L.check(L.lib.fdg_loss_grad(_ptr(a), _ptr(b), target, kind, n, scale, _ptr(grad), 1 if accumulate else 0,
                                loss.data_ptr(), _stream()), "loss_grad")
