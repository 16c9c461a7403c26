"""Forward / backward executors of the three networks on the hot path.

Each executor walks the network once and enqueues fdgan_b200 kernels (ops.py) on the current CUDA
stream.  Design points (DESIGN.md has the long form):

  * activations are NHWC fp32; every DenseNet block owns ONE pre-allocated concat buffer and the
    producing convolutions write their 32 new channels in place (no torch.cat, models/densenet.py:181,242);
  * BatchNorm never runs as a pass: batch statistics come out of the producer's epilogue (fp64 sums),
    ``bn_finalize`` turns them into per-channel scale/shift, and the consumer applies normalise+ReLU in
    its operand loader.  Statistics of a concat channel are computed once and shared by every later norm1;
  * avg-pool / nearest-upsample are folded into loaders and stores (gather / store modes);
  * backward = data-gradient convolutions (same kernel, flipped weights), weight-gradient GEMMs over the
    pixel dimension, and two light element-wise passes per BatchNorm (reduce, then apply).

Reference semantics followed: FDGAN.forward models/dehaze1113.py:758-801, BottleneckBlockdy :268-275,
TransitionBlockdy :366-370, D :188-230, Vgg16.forward myutils/vgg16.py:27-49, torchvision dense
layer/transition (spec copy models/densenet.py:179-242).
"""
from __future__ import annotations

import torch

from . import ops
from .ops import (ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, GATHER_AVGPOOL2, GATHER_DIRECT, GATHER_UP2,
                  STORE_ACCUM, STORE_NORMAL, STORE_UP2, View)

GROWTH = 32
BOTTLENECK = 128


class BNRun:
    """Per-forward state of one BatchNorm: scale/shift applied by the consumer, mean/invstd for backward."""

    __slots__ = ("mod", "C", "count", "scale", "shift", "mean", "invstd")

    def __init__(self, mod, C, count, buf):
        self.mod, self.C, self.count = mod, C, count
        self.scale, self.shift, self.mean, self.invstd = buf[0:C], buf[C:2 * C], buf[2 * C:3 * C], buf[3 * C:4 * C]


class _Pool:
    """Bump allocator over one zero-initialised torch buffer (small per-forward scratch: BN stats, BN vectors)."""

    def __init__(self, n, dtype, device):
        self.buf = torch.zeros(n, dtype=dtype, device=device)
        self.off = 0

    def take(self, n, align=4):
        self.off = (self.off + align - 1) // align * align
        t = self.buf[self.off:self.off + n]
        assert t.numel() == n, "scratch pool exhausted"
        self.off += n
        return t


def _bn_run(mod, stats, stats_ld, C, count, training, fpool, nbt_list):
    run = BNRun(mod, C, count, fpool.take(4 * C))
    track = training and mod.track_running_stats
    use_running = track or not training
    ops.bn_finalize(stats, stats_ld, C, count, mod.weight, mod.bias, mod.eps, mod.momentum,
                    mod.running_mean if use_running else None, mod.running_var if use_running else None,
                    training, run.scale, run.shift, run.mean, run.invstd)
    if track:
        nbt_list.append(mod.num_batches_tracked)
    return run


def _bn_bwd(g: View, x: View, run: BNRun, out: View, dpool, grads, prefix, *, slope=0.0, accumulate=False,
            g_gather=GATHER_DIRECT, gscale=1.0, out_split=None, both=False):
    """BatchNorm + (Leaky)ReLU backward: out (=|+=) dL/dx given g = dL/d(act(bn(x))).  With ``out_split`` the result is
    written as split-bf16 planes (the tensor-core operand format) instead of ``out`` (``both``: in addition to it)."""
    C = run.C
    st = dpool.take(2 * C)
    ops.ew_bwd(g, x, stats=st, scale=run.scale, shift=run.shift, slope=slope, g_gather=g_gather, gscale=gscale)
    coef = torch.empty(3 * C, dtype=torch.float32, device=x.base.device)
    dgamma = grads.get(prefix + ".weight") if grads is not None else None
    dbeta = grads.get(prefix + ".bias") if grads is not None else None
    ops.bn_bwd_finalize(st, C, run.count, run.mod.weight, run.mean, run.invstd, coef, dgamma, dbeta, accumulate=True)
    ops.ew_bwd(g, x, out=None if (out_split is not None and not both) else out, scale=run.scale, shift=run.shift, slope=slope, coef=coef,
               accumulate=accumulate, g_gather=g_gather, gscale=gscale, out_split=out_split)


# ======================================================================================================
# Generator
# ======================================================================================================


class GCtx:
    pass


def _dense_block_fwd(block, prefix, n_layers, c_in, X: View, S, training, fpool, spool, nbt, saved):
    N, H, W, Ctot = X.N, X.H, X.W, X.C
    count = N * H * W
    layers = []
    for i in range(n_layers):
        lyr = getattr(block, "denselayer%d" % (i + 1))
        cin = c_in + GROWTH * i
        bn1 = _bn_run(lyr.norm1, S, Ctot, cin, count, training, fpool, nbt)
        T = View.alloc(N, H, W, BOTTLENECK, X.base.device)
        ST = spool.take(2 * BOTTLENECK) if training else None
        w1, ld1 = ops.pack_weight(lyr.conv1.weight, 0)
        ops.conv2d(X.ch(0, cin), w1, ld1, 1, 1, 1, 0, BOTTLENECK, T, scale=bn1.scale, shift=bn1.shift, slope=0.0,
                   stats=ST, stats_ld=BOTTLENECK)
        bn2 = _bn_run(lyr.norm2, ST, BOTTLENECK, BOTTLENECK, count, training, fpool, nbt)
        w2, ld2 = ops.pack_weight(lyr.conv2.weight, 0)
        ops.conv2d(T, w2, ld2, 3, 3, 1, 1, GROWTH, X.ch(cin, cin + GROWTH), scale=bn2.scale, shift=bn2.shift,
                   slope=0.0, stats=(S.data_ptr() + 8 * cin) if training else None, stats_ld=Ctot)
        layers.append((T, bn1, bn2))
    saved[prefix] = layers


POOL_FIRST = True      # transitions: materialise avg_pool(relu(bn(x))) once; conv and weight gradient run as plain 1x1 kernels on it


def _transition_fwd(tr, X: View, S, y: View, ystats, ystats_ld, training, fpool, nbt):
    """torchvision transition (BatchNorm, ReLU, 1x1 conv, AvgPool2d(2)) with the pool commuted in front of the convolution.
    Returns (BatchNorm state, pooled activation or None)."""
    bn = _bn_run(tr.norm, S, X.C, X.C, X.N * X.H * X.W, training, fpool, nbt)
    if POOL_FIRST and ops.USE_UMMA and X.H % 2 == 0 and X.W % 2 == 0:
        P = View.alloc(X.N, X.H // 2, X.W // 2, X.C, X.base.device)
        ops.pool2_bn_act(X, P, bn.scale, bn.shift, 0.0)
        w, ld = ops.pack_weight(tr.conv.weight, 0)
        ops.conv2d(P, w, ld, 1, 1, 1, 0, y.C, y, stats=ystats, stats_ld=ystats_ld)
        return bn, P
    w, ld = ops.pack_weight(tr.conv.weight, 0)
    ops.conv2d(X, w, ld, 1, 1, 1, 0, y.C, y, gather=GATHER_AVGPOOL2, scale=bn.scale, shift=bn.shift, slope=0.0,
               stats=ystats, stats_ld=ystats_ld)
    return bn, None


def generator_forward(m, x: torch.Tensor, training: bool, need_ctx: bool):
    """FDGAN.forward (models/dehaze1113.py:758-801).  x: [B,3,H,W] fp32 CUDA (any strides) -> [B,3,H',W']."""
    with ops.pack_scope(m, "fwd", x.device):      # all weight-operand repacks of the pass in one launch per level
        return _generator_forward(m, x, training, need_ctx)


def _generator_forward(m, x: torch.Tensor, training: bool, need_ctx: bool):
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError("FDGAN expects a [B,3,H,W] input, got %s" % (tuple(x.shape),))
    if not x.is_cuda or x.dtype != torch.float32:
        raise RuntimeError("fdgan_b200 runs on CUDA fp32 tensors only (no CPU fallback); got %s %s" % (x.device, x.dtype))
    B, _, H1, W1 = x.shape
    H2, W2 = H1 // 2, W1 // 2
    H3, W3 = H2 // 2, W2 // 2
    H4, W4 = H3 // 2, W3 // 2
    if H4 < 1 or W4 < 1 or 2 * H4 != H3 or 2 * W4 != W3:
        # same condition under which the reference's torch.cat([x4, x2]) (dehaze1113.py:786) succeeds
        raise RuntimeError("FDGAN: input %dx%d is not supported: floor(H/4) must equal 2*floor(H/8)" % (H1, W1))
    dev = x.device
    ctx = GCtx()
    saved = {}
    nbt = []
    fpool = _Pool(4 * 60000, torch.float32, dev)          # BN scale/shift/mean/invstd vectors
    spool = _Pool(2 * (256 + 512 + 1024) + 42 * 256 + 1024, torch.float64, dev)  # BN statistics (zero-initialised)
    xin = View.from_nchw(x)

    # ---- stem: x0 = relu0(conv_refin1(x)) written straight into block 1's concat buffer
    X1 = View.alloc(B, H1, W1, 256, dev)
    S1 = spool.take(2 * 256)
    w, ld = ops.pack_weight(m.conv_refin1.weight, 0)
    ops.conv2d(xin, w, ld, 3, 3, 1, 1, 64, X1.ch(0, 64), bias=m.conv_refin1.bias, act=ACT_RELU,
               stats=S1 if training else None, stats_ld=256)
    # ---- x01 = conv_refin2(avg_pool2d(x0, 2)) -> first 32 channels of the conv_refine4 input
    C4 = View.alloc(B, H2, W2, 160, dev)
    w, ld = ops.pack_weight(m.conv_refin2.weight, 0)
    pool_first = POOL_FIRST and ops.USE_UMMA and H1 % 2 == 0 and W1 % 2 == 0
    P01 = P22 = None
    if pool_first:      # avg_pool2d(x0, 2) materialised once (dehaze1113.py:763): plain 1x1 convolution / weight gradient on it
        P01 = View.alloc(B, H2, W2, 64, dev)
        ops.pool2_bn_act(X1.ch(0, 64), P01)
        ops.conv2d(P01, w, ld, 1, 1, 1, 0, 32, C4.ch(0, 32), bias=m.conv_refin2.bias)
    else:
        ops.conv2d(X1.ch(0, 64), w, ld, 1, 1, 1, 0, 32, C4.ch(0, 32), gather=GATHER_AVGPOOL2, bias=m.conv_refin2.bias)
    # ---- dense_block1 + trans_block1
    _dense_block_fwd(m.dense_block1, "dense_block1", 6, 64, X1, S1, training, fpool, spool, nbt, saved)
    bn_t1, P_t1 = _transition_fwd(m.trans_block1, X1, S1, C4.ch(32, 160), None, 0, training, fpool, nbt)
    # ---- x10 = conv_refine4(cat[x01, x1])
    X2 = View.alloc(B, H2, W2, 512, dev)
    S2 = spool.take(2 * 512)
    w, ld = ops.pack_weight(m.conv_refine4.weight, 0)
    ops.conv2d(C4, w, ld, 3, 3, 1, 1, 128, X2.ch(0, 128), bias=m.conv_refine4.bias,
               stats=S2 if training else None, stats_ld=512)
    _dense_block_fwd(m.dense_block2, "dense_block2", 12, 128, X2, S2, training, fpool, spool, nbt, saved)
    X3 = View.alloc(B, H3, W3, 1024, dev)
    S3 = spool.take(2 * 1024)
    bn_t2, P_t2 = _transition_fwd(m.trans_block2, X2, S2, X3.ch(0, 256), S3 if training else None, 1024, training, fpool, nbt)
    _dense_block_fwd(m.dense_block3, "dense_block3", 24, 256, X3, S3, training, fpool, spool, nbt, saved)
    C6 = View.alloc(B, H4, W4, 640, dev)
    bn_t3, P_t3 = _transition_fwd(m.trans_block3, X3, S3, C6.ch(0, 512), None, 0, training, fpool, nbt)
    # ---- x22 = conv_refin5(avg_pool2d(x2, 2))
    w, ld = ops.pack_weight(m.conv_refin5.weight, 0)
    if pool_first and H3 % 2 == 0 and W3 % 2 == 0:
        P22 = View.alloc(B, H4, W4, 256, dev)
        ops.pool2_bn_act(X3.ch(0, 256), P22)
        ops.conv2d(P22, w, ld, 1, 1, 1, 0, 128, C6.ch(512, 640), bias=m.conv_refin5.bias)
    else:
        ops.conv2d(X3.ch(0, 256), w, ld, 1, 1, 1, 0, 128, C6.ch(512, 640), gather=GATHER_AVGPOOL2, bias=m.conv_refin5.bias)
    # ---- decoder level 4: conv_refin6 -> BottleneckBlockdy(512,256) -> TransitionBlockdy(768,128)
    D4 = View.alloc(B, H4, W4, 768, dev)
    w, ld = ops.pack_weight(m.conv_refin6.weight, 0)
    ops.conv2d(C6, w, ld, 3, 3, 1, 1, 512, D4.ch(0, 512), bias=m.conv_refin6.bias, act=ACT_RELU)  # in-place ReLU of Bdy
    T4 = View.alloc(B, H4, W4, 1024, dev)
    w, ld = ops.pack_weight(m.dense_block4.conv1.weight, 0)
    ops.conv2d(D4.ch(0, 512), w, ld, 1, 1, 1, 0, 1024, T4)
    w, ld = ops.pack_weight(m.dense_block4.conv2.weight, 0)
    ops.conv2d(T4, w, ld, 3, 3, 1, 1, 256, D4.ch(512, 768), slope=0.0)
    X42 = View.alloc(B, H3, W3, 512, dev)  # [relu(x4) | relu(x2) | block-5 growth]
    ops.conv2d(D4, m.trans_block4.conv1.weight, 128, 1, 1, 1, 0, 128, X42.ch(0, 128), slope=0.0, act=ACT_RELU,
               store=STORE_UP2)
    ops.copy4d(X3.ch(0, 256), X42.ch(128, 384), slope=0.0)
    # ---- decoder level 5
    T5 = View.alloc(B, H3, W3, 512, dev)
    w, ld = ops.pack_weight(m.dense_block5.conv1.weight, 0)
    ops.conv2d(X42.ch(0, 384), w, ld, 1, 1, 1, 0, 512, T5)
    w, ld = ops.pack_weight(m.dense_block5.conv2.weight, 0)
    ops.conv2d(T5, w, ld, 3, 3, 1, 1, 128, X42.ch(384, 512), slope=0.0)
    Hd2, Wd2 = 2 * H3, 2 * W3
    D6 = View.alloc(B, Hd2, Wd2, 96, dev)   # [relu(x5) | block-6 growth]
    ops.conv2d(X42, m.trans_block5.conv1.weight, 64, 1, 1, 1, 0, 64, D6.ch(0, 64), slope=0.0, act=ACT_RELU,
               store=STORE_UP2)
    # ---- decoder level 6
    T6 = View.alloc(B, Hd2, Wd2, 128, dev)
    w, ld = ops.pack_weight(m.dense_block6.conv1.weight, 0)
    ops.conv2d(D6.ch(0, 64), w, ld, 1, 1, 1, 0, 128, T6)
    w, ld = ops.pack_weight(m.dense_block6.conv2.weight, 0)
    ops.conv2d(T6, w, ld, 3, 3, 1, 1, 32, D6.ch(64, 96), slope=0.0)
    Hd1, Wd1 = 2 * Hd2, 2 * Wd2
    X6 = View.alloc(B, Hd1, Wd1, 16, dev)
    ops.conv2d(D6, m.trans_block6.conv1.weight, 16, 1, 1, 1, 0, 16, X6, slope=0.0, store=STORE_UP2)
    # ---- head: tanh(conv_refin3(x6)), NCHW output
    out = torch.empty((B, 3, Hd1, Wd1), dtype=torch.float32, device=dev)
    w, ld = ops.pack_weight(m.conv_refin3.weight, 0)
    ops.conv2d(X6, w, ld, 3, 3, 1, 1, 3, View.from_nchw(out), bias=m.conv_refin3.bias, act=ACT_TANH)
    if nbt:
        torch._foreach_add_(nbt, 1)
    if not need_ctx:
        return out, None
    ctx.x, ctx.out = x, out
    ctx.X1, ctx.C4, ctx.X2, ctx.X3, ctx.C6, ctx.D4, ctx.T4 = X1, C4, X2, X3, C6, D4, T4
    ctx.X42, ctx.T5, ctx.D6, ctx.T6, ctx.X6 = X42, T5, D6, T6, X6
    ctx.bn_t1, ctx.bn_t2, ctx.bn_t3 = bn_t1, bn_t2, bn_t3
    ctx.P_t = (P_t1, P_t2, P_t3)      # pooled transition inputs (saved for the weight gradients)
    ctx.P01, ctx.P22 = P01, P22
    ctx.saved = saved
    ctx.keep = (fpool, spool)
    ctx.training = training
    return out, ctx


def _conv_dgrad_w(weight):
    """[K=(r,s,co)][N=ci] operand for the data gradient of a stride-1 conv; 1x1 weights are used as they are."""
    co, ci, R, S = weight.shape
    if R == 1 and S == 1:
        return weight, ci
    return ops.pack_weight(weight, 1)


TAP_DECOMPOSE_L5 = True   # Fusion-D layer 5 (one output channel) as a 1x1 convolution over taps + shifted sum (fdg_tap_sum / fdg_tap_spread)
FUSED_BN1_BWD = True   # norm1 backward inside the conv1 data-gradient epilogue + deferred per-channel affine term
# norm2 backward with the ReLU mask, alpha and the two reductions inside the conv2 data-gradient epilogue (halo kernel, FdgConv.e_scale with
# a normal store) instead of the reduce pass over dA2 and T.  Round-2 history: through the generic epilogue it was SLOWER (76.1 -> 77.9
# ms/step: mask rows fetched inside the coalesced phase, no bulk-tensor store); conv_halo<128,3,BN2> prefetches the mask rows of the next
# group (level with the reduce pass: the 32 -> 128 data gradient is bound by its four epilogue warps) and runs TWO epilogue sets on the
# warps a 32-channel halo does not need for loading: 65.27 -> 64.90 ms/step at batch 16, 7.61 -> 7.46 ms graphed at batch 1, 42 launches
# and 2.7 GB of DRAM traffic per step less.  FDG_FUSED_BN2_BWD=0 restores the reduce pass.
FUSED_BN2_BWD = bool(int(__import__("os").environ.get("FDG_FUSED_BN2_BWD", "1")))
SPLIT_GRADS = True     # the bottleneck gradient travels as split-bf16 planes: its two consumers are fed by bulk tensor loads


# Weight gradients are leaves of the backward graph: nothing downstream reads them before the optimiser.  On small problems (batch 1,
# deep 64x64 / 32x32 maps) every kernel is latency-bound and leaves most SMs idle, so the dense-block weight gradients are enqueued on a
# side stream and overlap the data-gradient chain (fork / join through events: legal under CUDA-graph capture, where they become
# parallel branches).  ``ASYNC_WGRAD_MAX_PIXELS``: largest N*H*W of a block for which this is done (0 = never).
ASYNC_WGRAD_MAX_PIXELS = int(__import__("os").environ.get("FDG_ASYNC_WGRAD_PIXELS", str(4 * 65536)))
_SIDE_STREAMS = {}


def _side_stream(dev):
    key = (dev.index if dev.index is not None else torch.cuda.current_device())
    s = _SIDE_STREAMS.get(key)
    if s is None:
        s = _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return s


class _SideQueue:
    """Leaf weight gradients of one backward pass on the side stream (see ASYNC_WGRAD_MAX_PIXELS).  ``wgrad`` forks (the side stream waits
    for everything enqueued on the main stream so far), launches there and keeps the operand buffers alive; ``join`` makes the main stream
    wait for the side stream -- called before the pass returns, i.e. before any buffer dies and before the optimiser reads the gradients."""

    def __init__(self, dev, pixels, batch, grads):
        self.on = (0 < pixels <= ASYNC_WGRAD_MAX_PIXELS and grads is not None and not isinstance(grads, _NoGrads) and
                   (batch >= 4 or torch.cuda.is_current_stream_capturing()))
        if self.on:
            self.main = torch.cuda.current_stream(dev).cuda_stream
            self.side = _side_stream(dev).cuda_stream
        self.keep = []
        self.used = False

    def wgrad(self, x, g, *a, **k):
        if not self.on or a[4] is None:      # a[4] = dw: a frozen parameter skips the launch
            return ops.wgrad(x, g, *a, **k)
        ops.stream_wait(self.side, ops.event_record(self.main))
        ops.wgrad(x, g, *a, stream=self.side, **k)
        self.keep.append((x.base, g.base, k.get("x_split"), k.get("g_split")))
        self.used = True

    def join(self):
        if self.on and self.used:
            ops.stream_wait(self.main, ops.event_record(self.side))
        self.keep.clear()
        self.used = False


_TLS_Q = __import__("threading").local()


def _wgrad(*a, **k):
    """ops.wgrad, through the pass's side queue when there is one."""
    q = getattr(_TLS_Q, "q", None)
    return q.wgrad(*a, **k) if q is not None else ops.wgrad(*a, **k)


WIDE_WGRAD_PLANES = True     # wide stride-1 RxS weight gradients: both operands as split-bf16 planes (FdgWgrad.x_split)


def _wgrad_wide(x: View, g: View, R, S, stride, pad, dw, *, slope=1.0, dbias=None):
    """Weight gradient of a wide RxS convolution (decoder / fusion 3x3 layers).  The generic tensor-core kernel converts every input element
    once per filter tap and every gradient element once per 128-row block of dW; here act(x) and g are written once as split-bf16 planes
    and the kernel is fed by bulk tensor loads alone."""
    if dw is not None and WIDE_WGRAD_PLANES and ops.wgrad_planes_ok(x, g, R, S, stride, pad):
        xs = ops.split_planes(x, slope)
        gs = ops.split_planes(g, 1.0)
        return _wgrad(x, g, R, S, stride, pad, dw, slope=slope, dbias=dbias, x_split=xs, g_split=gs)
    return _wgrad(x, g, R, S, stride, pad, dw, slope=slope, dbias=dbias)


def _dense_block_bwd(block, prefix, n_layers, c_in, X: View, dX: View, layers, grads, dpool):
    """Backward of one torchvision dense block.  norm1 of layer j reads concat channels [0, cin_j) with the SAME batch
    statistics as every other consumer of those channels, so its backward dx = alpha dz + (beta x + delta) splits into
    a part the conv1 data-gradient epilogue adds straight into dX (alpha dz, FdgConv.e_scale) and a per-channel affine
    part whose coefficients simply add up over the consumers; that sum is applied ONCE per channel, right before the
    channel's gradient is consumed (ops.affine_accum) -- instead of a reduce pass and a 4-stream apply pass over
    cin_j channels for every layer."""
    N, H, W = X.N, X.H, X.W
    dev = X.base.device
    Ctot = X.C
    cmax = c_in + GROWTH * (n_layers - 1)
    fused = FUSED_BN1_BWD and ops.USE_UMMA
    split = fused and SPLIT_GRADS
    # weight gradients on a side stream (small problems only); the bottleneck-gradient buffers then alternate between two copies, because
    # the side stream may still read layer i's while the main stream produces layer i-1's
    # (eager steps at batch < 4 are bound by the host's launch rate -- 15 ms of Python for 10 ms of kernels at batch 1 -- so the extra fork /
    # join calls would cost more than the overlap returns: there it is used only while a CUDA graph is being captured)
    async_w = (0 < N * H * W <= ASYNC_WGRAD_MAX_PIXELS and grads is not None and not isinstance(grads, _NoGrads) and
               (N >= 4 or torch.cuda.is_current_stream_capturing()))
    nbuf = 2 if async_w else 1
    main = torch.cuda.current_stream(dev).cuda_stream if async_w else None       # raw handles: fork / join are single C-ABI calls
    side_t = _side_stream(dev) if async_w else None
    side = side_t.cuda_stream if async_w else None
    done1 = {}                                     # layer -> point on the side stream after its conv1 weight gradient
    dA2_b = [View.alloc(N, H, W, BOTTLENECK, dev) for _ in range(nbuf)]
    # dL/d(conv1 output) after the norm2 backward, as bf16 hi / lo planes [pixels][128] (same bytes as fp32)
    dA2s_b = [torch.empty(N * H * W * BOTTLENECK, dtype=torch.float32, device=dev) if split else None for _ in range(nbuf)]
    dA1buf = None if fused else torch.empty(N * H * W * cmax, dtype=torch.float32, device=dev)
    cbd = torch.zeros(2, Ctot, dtype=torch.float32, device=dev) if fused else None   # deferred beta / delta sums
    coef_scratch = torch.empty(3 * Ctot, dtype=torch.float32, device=dev) if fused else None

    def wgrad_leaf(*a, **k):
        if not async_w:
            return ops.wgrad(*a, **k)
        ops.stream_wait(side, ops.event_record(main))      # fork: the side stream sees everything enqueued so far
        ops.wgrad(*a, stream=side, **k)

    for i in reversed(range(n_layers)):
        lyr = getattr(block, "denselayer%d" % (i + 1))
        p = "%s.denselayer%d" % (prefix, i + 1)
        cin = c_in + GROWTH * i
        T, bn1, bn2 = layers[i]
        dA2, dA2s = dA2_b[i % nbuf], dA2s_b[i % nbuf]
        dA2sv = View.nhwc(dA2s, N, H, W, BOTTLENECK) if split else None
        if async_w and (i + 2) in done1:
            ops.stream_wait(main, done1.pop(i + 2))      # the side stream has finished reading this pair of buffers (layer i + 2)
        if fused and i < n_layers - 1:   # the later layers' affine terms for this layer's 32 new channels
            ops.affine_accum(X.ch(cin, cin + GROWTH), dX.ch(cin, cin + GROWTH), cbd[0, cin:cin + GROWTH], cbd[1, cin:cin + GROWTH])
        g2 = dX.ch(cin, cin + GROWTH)
        wgrad_leaf(T, g2, 3, 3, 1, 1, grads[p + ".conv2.weight"], scale=bn2.scale, shift=bn2.shift, slope=0.0)
        wd, ldd = _conv_dgrad_w(lyr.conv2.weight)
        if fused and FUSED_BN2_BWD:
            # the data-gradient epilogue reads the bottleneck tile once: dz = acc * [bn2(T) > 0], stores alpha * dz and reduces
            # sum dz, sum dz * T; what is left of the BatchNorm backward is one pass dA2 + beta * T + delta (no mask, alpha = 1)
            st2 = dpool.take(2 * BOTTLENECK)
            ops.conv2d(g2, wd, ldd, 3, 3, 1, 1, BOTTLENECK, dA2, e=T, eslope=0.0, e_scale=bn2.scale, e_shift=bn2.shift, stats=st2,
                       stats_ld=BOTTLENECK)
            coef2 = torch.empty(3 * BOTTLENECK, dtype=torch.float32, device=dev)
            ops.bn_bwd_finalize(st2, BOTTLENECK, bn2.count, bn2.mod.weight, bn2.mean, bn2.invstd, coef2, grads.get(p + ".norm2.weight"),
                                grads.get(p + ".norm2.bias"), accumulate=True, unit_alpha=True)
            ops.ew_bwd(dA2, T, out=None if split else dA2, coef=coef2, slope=1.0, out_split=dA2s)
        else:
            ops.conv2d(g2, wd, ldd, 3, 3, 1, 1, BOTTLENECK, dA2)
            _bn_bwd(dA2, T, bn2, dA2, dpool, grads, p + ".norm2", out_split=dA2s)      # dA2 <- dL/d(conv1 output) (in place / planes)
        wgrad_leaf(X.ch(0, cin), dA2sv if split else dA2, 1, 1, 1, 0, grads[p + ".conv1.weight"], scale=bn1.scale, shift=bn1.shift,
                   slope=0.0, g_split=dA2s)
        if async_w:
            done1[i] = ops.event_record(side)
        if fused:
            st = dpool.take(2 * cin)
            ops.conv2d(dA2sv if split else dA2, lyr.conv1.weight, cin, 1, 1, 1, 0, cin, dX.ch(0, cin), store=STORE_ACCUM, e=X.ch(0, cin),
                       eslope=0.0, e_scale=bn1.scale, e_shift=bn1.shift, stats=st, stats_ld=cin, x_split=dA2s)
            # beta / delta of this layer go straight into the running sums of the deferred affine term (no torch add, no allocation)
            ops.bn_bwd_finalize(st, cin, bn1.count, bn1.mod.weight, bn1.mean, bn1.invstd, coef_scratch,
                                grads.get(p + ".norm1.weight"), grads.get(p + ".norm1.bias"), accumulate=True, acc_beta=cbd[0], acc_delta=cbd[1])
        else:
            dA1 = View.nhwc(dA1buf, N, H, W, cin)
            ops.conv2d(dA2, lyr.conv1.weight, cin, 1, 1, 1, 0, cin, dA1)
            _bn_bwd(dA1, X.ch(0, cin), bn1, dX.ch(0, cin), dpool, grads, p + ".norm1", accumulate=True)
    if fused:
        ops.affine_accum(X.ch(0, c_in), dX.ch(0, c_in), cbd[0, :c_in], cbd[1, :c_in])
    if async_w:
        ops.stream_wait(main, ops.event_record(side))       # join: the buffers above die with this frame; the optimiser reads the gradients on the main stream


def _transition_bwd(tr, prefix, X: View, dX: View, g: View, bn: BNRun, grads, dpool, P: View | None = None, accumulate=True):
    """torchvision transition backward; g = dL/d(output) at the pooled resolution; accumulates into dX (``accumulate=False``: dX is
    uninitialised and this is its first writer -- every element is written, no zero fill and no read of the old value).  ``P``: the pooled
    activation the forward pass materialised (POOL_FIRST): the weight gradient is then a plain 1x1 one on a quarter of the pixels."""
    if P is not None:
        _wgrad(P, g, 1, 1, 1, 0, grads[prefix + ".conv.weight"])
    else:
        _wgrad(X, g, 1, 1, 1, 0, grads[prefix + ".conv.weight"], gather=GATHER_AVGPOOL2, scale=bn.scale, shift=bn.shift,
               slope=0.0)
    dP = View.alloc(g.N, g.H, g.W, X.C, X.base.device)
    ops.conv2d(g, tr.conv.weight, X.C, 1, 1, 1, 0, X.C, dP)
    if (X.H, X.W) != (2 * g.H, 2 * g.W):
        raise RuntimeError("fdgan_b200 backward needs even feature-map sizes at every transition (got %dx%d)" % (X.H, X.W))
    _bn_bwd(dP, X, bn, dX, dpool, grads, prefix + ".norm", accumulate=accumulate, g_gather=GATHER_UP2, gscale=0.25)


def _bdy_bwd(blk, prefix, Dv: View, dD: View, T: View, c_in, grads):
    """BottleneckBlockdy backward.  Dv = [relu(x) | conv2 out]; dD holds dL/dDv (already masked by Dv > 0);
    on return dD[:, :c_in] is dL/d(pre-ReLU x)."""
    c_out = Dv.C - c_in
    g = dD.ch(c_in, Dv.C)
    _wgrad_wide(T, g, 3, 3, 1, 1, grads[prefix + ".conv2.weight"], slope=0.0)
    dT = View.alloc(T.N, T.H, T.W, T.C, T.base.device)
    wd, ldd = _conv_dgrad_w(blk.conv2.weight)
    ops.conv2d(g, wd, ldd, 3, 3, 1, 1, T.C, dT, e=T, eslope=0.0)
    _wgrad(Dv.ch(0, c_in), dT, 1, 1, 1, 0, grads[prefix + ".conv1.weight"])
    ops.conv2d(dT, blk.conv1.weight, c_in, 1, 1, 1, 0, c_in, dD.ch(0, c_in), e=Dv.ch(0, c_in), eslope=0.0,
               store=STORE_ACCUM)
    del c_out


def _tdy_bwd(tr, prefix, Dv: View, dD: View, gUp: View, grads):
    """TransitionBlockdy backward: gUp = dL/d(upsampled output); writes dD = dL/dDv masked by Dv > 0."""
    c_out = gUp.C
    dU = View.alloc(Dv.N, Dv.H, Dv.W, c_out, Dv.base.device)
    ops.copy4d(gUp, dU, gather=GATHER_AVGPOOL2, scale=4.0)              # adjoint of nearest x2
    _wgrad(Dv, dU, 1, 1, 1, 0, grads[prefix + ".conv1.weight"], slope=0.0, transposed=True)
    wd, ldd = ops.pack_weight(tr.conv1.weight, 2)
    ops.conv2d(dU, wd, ldd, 1, 1, 1, 0, Dv.C, dD, e=Dv, eslope=0.0)


def generator_backward(m, ctx: GCtx, dout: torch.Tensor, grads: dict, need_dx: bool):
    """Backward of generator_forward.  ``grads`` maps parameter names to pre-zeroed fp32 tensors that the
    kernels accumulate into.  Returns dL/dx (NCHW) or None."""
    q = _SideQueue(dout.device, ctx.out.shape[0] * ctx.out.shape[2] * ctx.out.shape[3], ctx.out.shape[0], grads)
    _TLS_Q.q = q if q.on else None
    try:
        with ops.pack_scope(m, "bwd", dout.device):
            return _generator_backward(m, ctx, dout, grads, need_dx)
    finally:
        _TLS_Q.q = None
        q.join()


def _generator_backward(m, ctx: GCtx, dout: torch.Tensor, grads: dict, need_dx: bool):
    if not ctx.training:
        raise RuntimeError("fdgan_b200: backward through FDGAN in eval() mode is not implemented (the reference always "
                           "runs BatchNorm on batch statistics, README.md:38)")
    if grads is None:
        grads = _NoGrads()
    dev = dout.device
    B = ctx.out.shape[0]
    dpool = _Pool(2 * 60000, torch.float64, dev)
    X1, C4, X2, X3, C6, D4, T4 = ctx.X1, ctx.C4, ctx.X2, ctx.X3, ctx.C6, ctx.D4, ctx.T4
    X42, T5, D6, T6, X6 = ctx.X42, ctx.T5, ctx.D6, ctx.T6, ctx.X6
    # ---- head
    dpre = torch.empty_like(ctx.out)
    ops.act_bwd(dout.contiguous(), ctx.out, dpre, ACT_TANH)
    gpre = View.from_nchw(dpre)
    _wgrad(X6, gpre, 3, 3, 1, 1, grads["conv_refin3.weight"], dbias=grads["conv_refin3.bias"])
    dX6 = View.alloc(X6.N, X6.H, X6.W, 16, dev)
    wd, ldd = _conv_dgrad_w(m.conv_refin3.weight)
    ops.conv2d(gpre, wd, ldd, 3, 3, 1, 1, 16, dX6)
    # ---- level 6
    dD6 = View.alloc(D6.N, D6.H, D6.W, 96, dev)
    _tdy_bwd(m.trans_block6, "trans_block6", D6, dD6, dX6, grads)
    _bdy_bwd(m.dense_block6, "dense_block6", D6, dD6, T6, 64, grads)
    # ---- level 5
    dX42 = View.alloc(X42.N, X42.H, X42.W, 512, dev)
    _tdy_bwd(m.trans_block5, "trans_block5", X42, dX42, dD6.ch(0, 64), grads)
    _bdy_bwd(m.dense_block5, "dense_block5", X42, dX42, T5, 384, grads)
    # gradient buffers of the encoder blocks: the transition backward is the first writer of every element (no zero fill), everything
    # else accumulates afterwards
    dX3 = View.alloc(X3.N, X3.H, X3.W, 1024, dev)
    # ---- level 4
    dD4 = View.alloc(D4.N, D4.H, D4.W, 768, dev)
    _tdy_bwd(m.trans_block4, "trans_block4", D4, dD4, dX42.ch(0, 128), grads)
    _bdy_bwd(m.dense_block4, "dense_block4", D4, dD4, T4, 512, grads)
    g6 = dD4.ch(0, 512)                                                   # dL/d(conv_refin6 pre-ReLU output)
    _wgrad_wide(C6, g6, 3, 3, 1, 1, grads["conv_refin6.weight"], dbias=grads["conv_refin6.bias"])
    dC6 = View.alloc(C6.N, C6.H, C6.W, 640, dev)
    wd, ldd = _conv_dgrad_w(m.conv_refin6.weight)
    ops.conv2d(g6, wd, ldd, 3, 3, 1, 1, 640, dC6)
    # ---- conv_refin5 branch (x22)
    g5 = dC6.ch(512, 640)
    if ctx.P22 is not None:
        _wgrad(ctx.P22, g5, 1, 1, 1, 0, grads["conv_refin5.weight"], dbias=grads["conv_refin5.bias"])
    else:
        _wgrad(X3.ch(0, 256), g5, 1, 1, 1, 0, grads["conv_refin5.weight"], gather=GATHER_AVGPOOL2, dbias=grads["conv_refin5.bias"])
    dP = View.alloc(g5.N, g5.H, g5.W, 256, dev)
    ops.conv2d(g5, m.conv_refin5.weight, 256, 1, 1, 1, 0, 256, dP)
    # ---- encoder level 3
    _transition_bwd(m.trans_block3, "trans_block3", X3, dX3, dC6.ch(0, 512), ctx.bn_t3, grads, dpool, ctx.P_t[2], accumulate=False)
    ops.copy4d(dX42.ch(128, 384), dX3.ch(0, 256), accumulate=True)       # x2 skip (already masked by relu(x2) > 0)
    ops.copy4d(dP, dX3.ch(0, 256), gather=GATHER_UP2, scale=0.25, accumulate=True)
    _dense_block_bwd(m.dense_block3, "dense_block3", 24, 256, X3, dX3, ctx.saved["dense_block3"], grads, dpool)
    # ---- encoder level 2
    dX2 = View.alloc(X2.N, X2.H, X2.W, 512, dev)
    _transition_bwd(m.trans_block2, "trans_block2", X2, dX2, dX3.ch(0, 256), ctx.bn_t2, grads, dpool, ctx.P_t[1], accumulate=False)
    _dense_block_bwd(m.dense_block2, "dense_block2", 12, 128, X2, dX2, ctx.saved["dense_block2"], grads, dpool)
    g4 = dX2.ch(0, 128)
    _wgrad_wide(C4, g4, 3, 3, 1, 1, grads["conv_refine4.weight"], dbias=grads["conv_refine4.bias"])
    dC4 = View.alloc(C4.N, C4.H, C4.W, 160, dev)
    wd, ldd = _conv_dgrad_w(m.conv_refine4.weight)
    ops.conv2d(g4, wd, ldd, 3, 3, 1, 1, 160, dC4)
    # ---- encoder level 1
    dX1 = View.alloc(X1.N, X1.H, X1.W, 256, dev)
    _transition_bwd(m.trans_block1, "trans_block1", X1, dX1, dC4.ch(32, 160), ctx.bn_t1, grads, dpool, ctx.P_t[0], accumulate=False)
    g2 = dC4.ch(0, 32)
    if ctx.P01 is not None:
        _wgrad(ctx.P01, g2, 1, 1, 1, 0, grads["conv_refin2.weight"], dbias=grads["conv_refin2.bias"])
    else:
        _wgrad(X1.ch(0, 64), g2, 1, 1, 1, 0, grads["conv_refin2.weight"], gather=GATHER_AVGPOOL2, dbias=grads["conv_refin2.bias"])
    dP = View.alloc(g2.N, g2.H, g2.W, 64, dev)
    ops.conv2d(g2, m.conv_refin2.weight, 64, 1, 1, 1, 0, 64, dP)
    if (X1.H, X1.W) != (2 * dP.H, 2 * dP.W):
        raise RuntimeError("fdgan_b200 backward needs an even input size (got %dx%d)" % (X1.H, X1.W))
    ops.copy4d(dP, dX1.ch(0, 64), gather=GATHER_UP2, scale=0.25, accumulate=True)
    _dense_block_bwd(m.dense_block1, "dense_block1", 6, 64, X1, dX1, ctx.saved["dense_block1"], grads, dpool)
    # ---- stem: x0 = relu(conv_refin1(x)) stored post-ReLU
    g0 = dX1.ch(0, 64)
    ops.ew_bwd(g0, X1.ch(0, 64), out=g0, slope=0.0)
    xin = View.from_nchw(ctx.x)
    _wgrad(xin, g0, 3, 3, 1, 1, grads["conv_refin1.weight"], dbias=grads["conv_refin1.bias"])
    if not need_dx:
        return None
    dx = torch.empty((B, 3, xin.H, xin.W), dtype=torch.float32, device=dev)
    wd, ldd = _conv_dgrad_w(m.conv_refin1.weight)
    ops.conv2d(g0, wd, ldd, 3, 3, 1, 1, 3, View.from_nchw(dx))
    return dx


# ======================================================================================================
# Fusion-discriminator
# ======================================================================================================


class DCtx:
    pass


def discriminator_forward(m, z: torch.Tensor, training: bool, need_ctx: bool):
    """D.forward (models/dehaze1113.py:188-230): z [B,nc,H,W] -> sigmoid patch map [B,1,H/2-2,W/2-2]."""
    with ops.pack_scope(m, "fwd", z.device):
        return _discriminator_forward(m, z, training, need_ctx)


def _discriminator_forward(m, z: torch.Tensor, training: bool, need_ctx: bool):
    if z.dim() != 4 or z.shape[1] != m.nc:
        raise ValueError("D expects a [B,%d,H,W] input, got %s" % (m.nc, tuple(z.shape)))
    if not z.is_cuda or z.dtype != torch.float32:
        raise RuntimeError("fdgan_b200 runs on CUDA fp32 tensors only (no CPU fallback)")
    dev = z.device
    nf = m.nf
    B, _, H, W = z.shape
    zin = View.from_nchw(z)
    l1, l2, l3, l4, l5 = m.layer_params()
    fpool = _Pool(4 * 6 * nf + 64, torch.float32, dev)
    spool = _Pool(2 * 6 * nf + 16, torch.float64, dev)
    nbt = []
    H1, W1 = (H + 2 - 4) // 2 + 1, (W + 2 - 4) // 2 + 1
    if H1 < 3 or W1 < 3:
        raise RuntimeError("D: input %dx%d too small" % (H, W))
    Y1 = View.alloc(B, H1, W1, nf, dev)
    w, ld = ops.pack_weight(l1.weight, 0)
    ops.conv2d(zin, w, ld, 4, 4, 2, 1, nf, Y1)
    Y2 = View.alloc(B, H1, W1, 2 * nf, dev)
    S2 = spool.take(4 * nf)
    w, ld = ops.pack_weight(l2.conv.weight, 0)
    ops.conv2d(Y1, w, ld, 3, 3, 1, 1, 2 * nf, Y2, slope=0.2, stats=S2 if training else None, stats_ld=2 * nf)
    bn2 = _bn_run(l2.bn, S2, 2 * nf, 2 * nf, B * H1 * W1, training, fpool, nbt)
    Y3 = View.alloc(B, H1, W1, 4 * nf, dev)
    S3 = spool.take(8 * nf)
    w, ld = ops.pack_weight(l3.conv.weight, 0)
    ops.conv2d(Y2, w, ld, 3, 3, 1, 1, 4 * nf, Y3, scale=bn2.scale, shift=bn2.shift, slope=0.2,
               stats=S3 if training else None, stats_ld=4 * nf)
    bn3 = _bn_run(l3.bn, S3, 4 * nf, 4 * nf, B * H1 * W1, training, fpool, nbt)
    Y4 = View.alloc(B, H1 - 1, W1 - 1, 8 * nf, dev)
    w, ld = ops.pack_weight(l4.weight, 0)
    ops.conv2d(Y3, w, ld, 4, 4, 1, 1, 8 * nf, Y4, scale=bn3.scale, shift=bn3.shift, slope=0.2)
    out = torch.empty((B, 1, H1 - 2, W1 - 2), dtype=torch.float32, device=dev)
    if ops.USE_UMMA and TAP_DECOMPOSE_L5 and (8 * nf) % 8 == 0:
        # one output channel: 1x1 convolution 8nf -> 16 (a column per filter tap; the OIHW weight [1][8nf][4][4] is that operand as it
        # stands) on the tensor cores, then the 16 taps are summed over shifted pixels -- the input is read once instead of 16 weight tiles
        # being streamed against an N = 32 tile of which one column is used
        S5 = View.alloc(B, Y4.H, Y4.W, 16, dev)
        ops.conv2d(Y4, l5.weight, 16, 1, 1, 1, 0, 16, S5, slope=0.2)
        ops.tap_sum(S5, View.from_nchw(out), 4, 4, 1, act=ACT_SIGMOID)
    else:
        w, ld = ops.pack_weight(l5.weight, 0)
        ops.conv2d(Y4, w, ld, 4, 4, 1, 1, 1, View.from_nchw(out), slope=0.2, act=ACT_SIGMOID)
    if nbt:
        torch._foreach_add_(nbt, 1)
    if not need_ctx:
        return out, None
    ctx = DCtx()
    ctx.z, ctx.out, ctx.Y1, ctx.Y2, ctx.Y3, ctx.Y4, ctx.bn2, ctx.bn3 = z, out, Y1, Y2, Y3, Y4, bn2, bn3
    ctx.keep = (fpool, spool)
    ctx.training = training
    return out, ctx


class _NoGrads(dict):
    """grads mapping used when a network's parameters are frozen: every lookup yields None."""

    def __getitem__(self, k):
        return None

    def get(self, k, default=None):
        return None


def discriminator_backward(m, ctx: DCtx, dout: torch.Tensor, grads, need_dx: bool):
    q = _SideQueue(dout.device, ctx.z.shape[0] * ctx.z.shape[2] * ctx.z.shape[3], ctx.z.shape[0], grads)
    _TLS_Q.q = q if q.on else None
    try:
        with ops.pack_scope(m, "bwd", dout.device):
            return _discriminator_backward(m, ctx, dout, grads, need_dx)
    finally:
        _TLS_Q.q = None
        q.join()


def _discriminator_backward(m, ctx: DCtx, dout: torch.Tensor, grads, need_dx: bool):
    if not ctx.training:
        raise RuntimeError("fdgan_b200: backward through D in eval() mode is not implemented")
    need_w = grads is not None
    if grads is None:
        grads = _NoGrads()
    dev = dout.device
    nf = m.nf
    l1, l2, l3, l4, l5 = m.layer_params()
    dpool = _Pool(2 * 6 * nf + 16, torch.float64, dev)
    Y1, Y2, Y3, Y4, bn2, bn3 = ctx.Y1, ctx.Y2, ctx.Y3, ctx.Y4, ctx.bn2, ctx.bn3
    dpre = torch.empty_like(ctx.out)
    ops.act_bwd(dout.contiguous(), ctx.out, dpre, ACT_SIGMOID)
    g5 = View.from_nchw(dpre)
    if need_w and ops.USE_UMMA and TAP_DECOMPOSE_L5:
        G5 = View.alloc(Y4.N, Y4.H, Y4.W, 16, dev)           # the patch-map gradient spread over the 16 taps
        ops.tap_spread(g5, G5, 4, 4, 1)
        _wgrad(Y4, G5, 1, 1, 1, 0, grads["main.layer5.conv.weight"], slope=0.2, transposed=True)      # [8nf][16] = OIHW [1][8nf][4][4]
    elif need_w:
        _wgrad(Y4, g5, 4, 4, 1, 1, grads["main.layer5.conv.weight"], slope=0.2)
    dY4 = View.alloc(Y4.N, Y4.H, Y4.W, Y4.C, dev)
    wd, ldd = ops.pack_weight(l5.weight, 1)
    ops.conv2d(g5, wd, ldd, 4, 4, 1, 2, Y4.C, dY4, e=Y4, eslope=0.2)
    split_g = need_w and ops.USE_UMMA and SPLIT_GRADS
    if split_g:
        # layer 4 / layer 3 weight gradients run 18 / 6 CTAs per pixel range, each converting the SAME gradient rows to bf16 hi / lo;
        # one pass writes them as split planes and the CTAs take them through bulk tensor loads (FdgWgrad.g_split)
        p4 = torch.empty(dY4.N * dY4.H * dY4.W * dY4.C, dtype=torch.float32, device=dev)
        ops.ew_bwd(dY4, dY4, slope=1.0, out_split=p4)
        _wgrad(Y3, View.nhwc(p4, dY4.N, dY4.H, dY4.W, dY4.C), 4, 4, 1, 1, grads["main.layer4.conv.weight"], scale=bn3.scale, shift=bn3.shift,
                  slope=0.2, g_split=p4)
    elif need_w:
        _wgrad(Y3, dY4, 4, 4, 1, 1, grads["main.layer4.conv.weight"], scale=bn3.scale, shift=bn3.shift, slope=0.2)
    dY3 = View.alloc(Y3.N, Y3.H, Y3.W, Y3.C, dev)
    wd, ldd = ops.pack_weight(l4.weight, 1)
    ops.conv2d(dY4, wd, ldd, 4, 4, 1, 2, Y3.C, dY3)
    p3 = torch.empty(dY3.N * dY3.H * dY3.W * dY3.C, dtype=torch.float32, device=dev) if split_g else None
    _bn_bwd(dY3, Y3, bn3, dY3, dpool, grads if need_w else None, "main.layer3.layer3.bn", slope=0.2, out_split=p3, both=True)
    if need_w:
        _wgrad(Y2, dY3, 3, 3, 1, 1, grads["main.layer3.layer3.conv.weight"], scale=bn2.scale, shift=bn2.shift, slope=0.2, g_split=p3)
    dY2 = View.alloc(Y2.N, Y2.H, Y2.W, Y2.C, dev)
    wd, ldd = ops.pack_weight(l3.conv.weight, 1)
    ops.conv2d(dY3, wd, ldd, 3, 3, 1, 1, Y2.C, dY2)
    _bn_bwd(dY2, Y2, bn2, dY2, dpool, grads if need_w else None, "main.layer2.layer2.bn", slope=0.2)
    if need_w:
        _wgrad(Y1, dY2, 3, 3, 1, 1, grads["main.layer2.layer2.conv.weight"], slope=0.2)
    dY1 = View.alloc(Y1.N, Y1.H, Y1.W, Y1.C, dev)
    wd, ldd = ops.pack_weight(l2.conv.weight, 1)
    ops.conv2d(dY2, wd, ldd, 3, 3, 1, 1, Y1.C, dY1, e=Y1, eslope=0.2)
    zin = View.from_nchw(ctx.z)
    if need_w:
        _wgrad(zin, dY1, 4, 4, 2, 1, grads["main.layer1.conv.weight"])
    if not need_dx:
        return None
    # channels-last memory, returned as a logical NCHW tensor
    dzv = View.alloc(zin.N, zin.H, zin.W, zin.C, dev)
    ops.dgrad_strided(dY1, l1.weight, 2, 1, dzv)
    return dzv.as_nchw()


# ======================================================================================================
# VGG16 feature extractor
# ======================================================================================================

VGG_STAGES = (("conv1_1", "conv1_2"), ("conv2_1", "conv2_2"), ("conv3_1", "conv3_2", "conv3_3"),
              ("conv4_1", "conv4_2", "conv4_3"))


class VCtx:
    pass


def vgg_forward(m, x: torch.Tensor, need_ctx: bool):
    """Vgg16.forward (myutils/vgg16.py:27-49) -> [relu1_2, relu2_2, relu3_3, relu4_3] as NCHW-shaped views."""
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError("Vgg16 expects a [B,3,H,W] input, got %s" % (tuple(x.shape),))
    if not x.is_cuda or x.dtype != torch.float32:
        raise RuntimeError("fdgan_b200 runs on CUDA fp32 tensors only (no CPU fallback)")
    dev = x.device
    B, _, H, W = x.shape
    cur = View.from_nchw(x)
    feats, acts = [], []
    for si, stage in enumerate(VGG_STAGES):
        if si > 0:
            pooled = View.alloc(B, cur.H // 2, cur.W // 2, cur.C, dev)
            if pooled.H < 1 or pooled.W < 1:
                raise RuntimeError("Vgg16: input %dx%d too small" % (H, W))
            ops.maxpool2_fwd(cur, pooled)
            cur = pooled
        stage_acts = [cur]
        for name in stage:
            conv = getattr(m, name)
            cout = conv.weight.shape[0]
            y = View.alloc(B, cur.H, cur.W, cout, dev)
            w, ld = ops.pack_weight(conv.weight, 0)
            ops.conv2d(cur, w, ld, 3, 3, 1, 1, cout, y, bias=conv.bias, act=ACT_RELU)
            cur = y
            stage_acts.append(y)
        acts.append(stage_acts)
        feats.append(cur)
    outs = [f.as_nchw() for f in feats]
    ctx = VCtx()
    ctx.feats = feats                      # the four NHWC feature buffers (always returned)
    ctx.x, ctx.acts = (x, acts) if need_ctx else (None, None)
    return outs, ctx


def vgg_backward(m, ctx: VCtx, gouts, grads, need_dx: bool):
    """gouts: 4 tensors or None (NCHW-shaped, any strides).  grads: dict for weight/bias gradients or None
    (frozen extractor: data gradient only)."""
    dev = ctx.x.device
    need_w = grads is not None
    g_next = None  # gradient w.r.t. the pooled input of the following stage
    for si in reversed(range(len(VGG_STAGES))):
        stage = VGG_STAGES[si]
        acts = ctx.acts[si]
        top = acts[-1]
        if g_next is None and gouts[si] is None:
            continue
        # dF = (gouts + unpool(g_next)) * [top > 0], every element written exactly once by its first contributor (no zero fill, no
        # separate mask pass): the un-pooling routes to the block maximum and applies the mask itself (top is post-ReLU)
        if g_next is None:
            dF = View.alloc(top.N, top.H, top.W, top.C, dev)
            ops.ew_bwd(View.from_nchw(gouts[si]), top, out=dF, slope=0.0)
        elif top.H % 2 == 0 and top.W % 2 == 0:
            dF = View.alloc(top.N, top.H, top.W, top.C, dev)
            ops.maxpool2_bwd(top, g_next, dF, accumulate=False, relu_mask=True)
            if gouts[si] is not None:
                ops.ew_bwd(View.from_nchw(gouts[si]), top, out=dF, slope=0.0, accumulate=True)
        else:       # odd sizes: the last row / column is not covered by the 2x2 blocks
            dF = View.alloc(top.N, top.H, top.W, top.C, dev, zero=True)
            if gouts[si] is not None:
                ops.copy4d(View.from_nchw(gouts[si]), dF, accumulate=True)
            ops.maxpool2_bwd(top, g_next, dF, accumulate=True)
            ops.ew_bwd(dF, top, out=dF, slope=0.0)       # ReLU mask of the stage output
        g = dF
        for li in reversed(range(len(stage))):
            name = stage[li]
            conv = getattr(m, name)
            xin = acts[li]
            if need_w:
                ops.wgrad(xin, g, 3, 3, 1, 1, grads[name + ".weight"], dbias=grads[name + ".bias"])
            first = (si == 0 and li == 0)
            if first and not need_dx:
                g = None
                break
            wd, ldd = ops.pack_weight(conv.weight, 1)
            if first:
                dx = torch.empty_like(ctx.x, memory_format=torch.contiguous_format)
                ops.conv2d(g, wd, ldd, 3, 3, 1, 1, 3, View.from_nchw(dx))
                return dx
            gi = View.alloc(xin.N, xin.H, xin.W, xin.C, dev)
            if li > 0:
                ops.conv2d(g, wd, ldd, 3, 3, 1, 1, xin.C, gi, e=xin, eslope=0.0)   # fused ReLU mask of the producer
            else:
                ops.conv2d(g, wd, ldd, 3, 3, 1, 1, xin.C, gi)                       # pooled input: mask applied after un-pooling
            g = gi
        g_next = g
    return None
