"""GPU parity of every C-ABI entry point against plain torch fp32/fp64 CPU arithmetic (unit level).
All calls go through the C ABI (fdgan_b200.ops -> ctypes -> libfdgan_b200.so)."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests.util import maxabs, seeded

pytestmark = pytest.mark.gpu


def _ops():
    from fdgan_b200 import ops
    return ops


def cl(t):
    """NCHW logical tensor in channels-last memory on the GPU."""
    return t.cuda().contiguous(memory_format=torch.channels_last)


def ref_prologue(x, scale, shift, slope):
    v = x if scale is None else x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    return torch.where(v > 0, v, slope * v)


CONV_CASES = [
    # Cin, Cout, R, stride, pad, H, W, gather, affine, slope, bias, act, store, stats, mask
    (128, 32, 3, 1, 1, 12, 20, 0, True, 0.0, False, 0, 0, True, False),    # K1 dense-layer conv2
    (96, 128, 1, 1, 0, 9, 7, 0, True, 0.0, False, 0, 0, True, False),      # K2 dense-layer conv1
    (64, 32, 1, 1, 0, 10, 12, 1, False, 1.0, True, 0, 0, False, False),    # conv_refin2 (avg-pool gather)
    (256, 128, 1, 1, 0, 8, 8, 1, True, 0.0, False, 0, 0, True, False),     # transition
    (3, 64, 3, 1, 1, 11, 13, 0, False, 1.0, True, 1, 0, True, False),      # conv_refin1 (K=27)
    (16, 3, 3, 1, 1, 10, 10, 0, False, 1.0, True, 2, 0, False, False),     # conv_refin3 + tanh
    (160, 128, 3, 1, 1, 6, 6, 0, False, 1.0, True, 0, 0, True, False),     # conv_refine4
    (96, 16, 1, 1, 0, 5, 6, 0, False, 0.0, False, 0, 1, False, False),     # TransitionBlockdy (up2 store)
    (64, 64, 1, 1, 0, 5, 6, 0, False, 0.0, False, 1, 1, False, False),     # up2 store + relu epilogue
    (9, 36, 4, 2, 1, 16, 18, 0, False, 1.0, False, 0, 0, False, False),    # D layer 1
    (36, 72, 3, 1, 1, 8, 9, 0, False, 0.2, False, 0, 0, True, False),      # D layer 2
    (144, 288, 4, 1, 1, 8, 9, 0, True, 0.2, False, 0, 0, False, False),    # D layer 4
    (288, 1, 4, 1, 1, 7, 8, 0, False, 0.2, False, 3, 0, False, False),     # D layer 5 + sigmoid
    (32, 128, 3, 1, 1, 7, 9, 0, False, 1.0, False, 0, 0, False, True),     # dgrad-like with ReLU mask epilogue
    (128, 72, 1, 1, 0, 6, 5, 0, False, 1.0, False, 0, 2, False, True),     # accumulate store + mask
    (16, 24, 1, 1, 0, 4, 6, 2, False, 1.0, False, 0, 0, False, False),     # up2 gather
    (64, 600, 3, 1, 1, 4, 4, 0, False, 1.0, True, 1, 0, False, False),     # wide N (several N tiles), VGG-like
    (1024, 256, 3, 1, 1, 9, 11, 0, False, 0.0, False, 0, 0, False, False), # dense_block4.conv2 (K = 9216)
    (224, 128, 1, 1, 0, 33, 31, 0, True, 0.0, False, 0, 0, True, False),   # ragged M (tail tile), padded K chunk
    (128, 32, 3, 1, 1, 64, 64, 0, True, 0.0, False, 0, 0, True, False),    # many tiles + statistics
    (512, 128, 3, 1, 1, 8, 8, 0, False, 0.0, False, 0, 2, False, True),    # dense_block5.conv2 dgrad-style accumulate
    (128, 32, 3, 1, 1, 37, 29, 0, True, 0.0, False, 0, 0, True, False),    # halo kernel: ragged 16x8 tiles, statistics
    (144, 288, 4, 1, 2, 21, 19, 0, True, 0.2, False, 0, 0, False, False),  # D layer 4 data-gradient geometry (4x4, pad 2)
    (32, 128, 3, 1, 1, 40, 40, 0, False, 1.0, False, 0, 0, False, True),   # dense-layer conv2 data gradient + ReLU mask
    (64, 64, 3, 1, 1, 5, 3, 0, False, 1.0, True, 1, 0, False, False),      # image smaller than one halo tile
    (192, 128, 1, 1, 0, 40, 44, 0, True, 0.0, False, 0, 0, True, False),   # bulk-tensor-fed 1x1 loader: many tiles, 3 chunks, statistics
    (72, 96, 1, 1, 0, 23, 17, 0, True, 0.2, True, 0, 0, False, False),     # ... Cin % 16 == 8 (half K slice), LeakyReLU prologue, bias
    (512, 300, 1, 1, 0, 13, 10, 0, False, 0.0, False, 1, 0, False, False), # ... several N tiles (BottleneckBlockdy conv1 shape family)
    (384, 128, 1, 1, 0, 12, 12, 0, False, 0.0, False, 1, 1, False, False), # ... TransitionBlockdy: ReLU prologue, up2 store
    (72, 144, 3, 1, 1, 19, 21, 0, True, 0.2, False, 0, 0, True, False),    # D layer 3: two 80-wide N tiles (16-channel tail groups), statistics
    (288, 144, 4, 1, 2, 13, 11, 0, False, 1.0, False, 0, 0, False, True),  # D layer 4 data gradient: 80-wide tiles + LeakyReLU mask
    (144, 72, 3, 1, 1, 17, 9, 0, False, 1.0, False, 0, 0, False, True),    # D layer 3 data gradient: one 80-wide tile, Cout tail inside it
    (128, 160, 3, 1, 1, 9, 10, 0, False, 1.0, True, 1, 2, False, False),   # conv_refine4 data-gradient geometry: 160 = 2 x 80, accumulate
    (144, 288, 4, 1, 1, 20, 18, 0, True, 0.2, False, 0, 0, True, False),   # D layer 4: three 96-wide N tiles, bulk tensor stores, statistics
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("variant", ["simt_nhwc", "simt_nchw", "tcgen05", "tcgen05_halo", "tcgen05_pertap"])
def test_conv2d(case, variant):
    """fdg_conv2d against fp64 torch on every shape family of the path, through both kernels: the fp32 SIMT path
    (2e-5) and the tcgen05 bf16x3 path (5e-5; ~16 operand mantissa bits, fp32 accumulation in TMEM)."""
    ops = _ops()
    Cin, Cout, R, stride, pad, H, W, gather, affine, slope, bias, act, store, stats, mask = case
    nchw_in = variant == "simt_nchw"
    impl = ops.IMPL_UMMA if variant.startswith("tcgen05") else ops.IMPL_SIMT
    if variant.startswith("tcgen05") and not (Cin % 8 == 0 and Cin >= 16):
        pytest.skip("shape not covered by the tcgen05 path (runs on the SIMT kernel)")
    halo_shape = gather == 0 and stride == 1 and 2 <= R <= 4
    if variant == "tcgen05_pertap" and not halo_shape:
        pytest.skip("same kernel as the tcgen05 variant for this shape")
    k1_shape = halo_shape and R == 3 and pad == 1 and Cout <= 32
    if variant == "tcgen05_halo" and not k1_shape:
        pytest.skip("same kernel as the tcgen05 variant for this shape")
    ops.USE_K1 = variant == "tcgen05"        # filter-row-concatenated kernel vs the halo-tile kernel for the growth convolutions
    from fdgan_b200 import _lib
    _lib.set_option("halo", 0 if variant == "tcgen05_pertap" else 1)   # halo-tile kernel vs generic per-tap kernel
    N = 2
    ph, pw = (2 * H, 2 * W) if gather == 1 else ((H + 1) // 2, (W + 1) // 2) if gather == 2 else (H, W)
    if gather == 2:
        H, W = 2 * ph, 2 * pw
    x = seeded((N, Cin, ph, pw), 1, -1.0, 1.0)
    w = seeded((Cout, Cin, R, R), 2, -1.0, 1.0) / math.sqrt(Cin * R * R)
    b = seeded((Cout,), 3, -0.5, 0.5) if bias else None
    sc = seeded((Cin,), 4, 0.5, 1.5) if affine else None
    sh = seeded((Cin,), 5, -0.3, 0.3) if affine else None
    # ---- reference (CPU, fp64)
    a = ref_prologue(x.double(), sc.double() if affine else None, sh.double() if affine else None, slope)
    if gather == 1:
        a = F.avg_pool2d(a, 2)
    elif gather == 2:
        a = F.interpolate(a, scale_factor=2, mode="nearest")
    y = F.conv2d(a, w.double(), b.double() if bias else None, stride=stride, padding=pad)
    y = {0: lambda t: t, 1: torch.relu, 2: torch.tanh, 3: torch.sigmoid}[act](y)
    OH, OW = y.shape[-2:]
    e = seeded((N, Cout, OH, OW), 6, -1.0, 1.0) if mask else None
    if mask:
        y = y * torch.where(e.double() > 0, 1.0, 0.3)
    ysum, ysq = y.sum((0, 2, 3)), (y * y).sum((0, 2, 3))
    if store == 1:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    y0 = seeded(tuple(y.shape), 7, -1.0, 1.0)
    if store == 2:
        y = y + y0.double()
    # ---- device
    xd = x.cuda() if nchw_in else cl(x)
    wp, ld = ops.pack_weight(w.cuda(), 0)
    yd = cl(y0.clone())
    st = torch.zeros(2 * Cout + 6, dtype=torch.float64, device="cuda") if stats else None
    ops.conv2d(ops.View.from_nchw(xd), wp, ld, R, R, stride, pad, Cout, ops.View.from_nchw(yd), gather=gather,
               scale=sc.cuda() if affine else None, shift=sh.cuda() if affine else None, slope=slope,
               bias=b.cuda() if bias else None, act=act, e=ops.View.from_nchw(cl(e)) if mask else None, eslope=0.3,
               store=store, stats=st, stats_ld=Cout + 3, impl=impl)
    torch.cuda.synchronize()
    _lib.set_option("halo", 1)
    ops.USE_K1 = True
    assert maxabs(yd, y) <= (5e-5 if variant.startswith("tcgen05") else 2e-5)
    if stats:
        tol = 2e-6 * float(N * OH * OW) + 1e-3   # fp32 partial sums over the tile, fp64 across tiles
        assert maxabs(st[:Cout], ysum) <= tol and maxabs(st[Cout + 3:2 * Cout + 3], ysq) <= tol


@pytest.mark.parametrize("cin,ctot,c0", [(64, 256, 0), (96, 256, 32), (224, 256, 0), (128, 512, 384)])
def test_conv2d_1x1_on_channel_slices_of_a_concat_buffer(cin, ctot, c0):
    """Dense-layer conv1 as the generator runs it: the input is channels [c0, c0+cin) of a wider NHWC concat buffer (the bulk-tensor
    loader's map clips at cin: the neighbouring channels must not leak in), the output goes into a channel slice of another one."""
    ops = _ops()
    N, H, W, Cout = 2, 19, 27, 128
    buf = seeded((N, H, W, ctot), 1, -1.0, 1.0)
    w = seeded((Cout, cin, 1, 1), 2, -1.0, 1.0) / math.sqrt(cin)
    sc, sh = seeded((cin,), 4, 0.5, 1.5), seeded((cin,), 5, -0.3, 0.3)
    x = buf[..., c0:c0 + cin].permute(0, 3, 1, 2).double()
    y = F.conv2d(ref_prologue(x, sc.double(), sh.double(), 0.0), w.double())
    bd = buf.cuda()
    xv = ops.View.nhwc(bd, N, H, W, ctot).ch(c0, c0 + cin)
    out = torch.full((N * H * W * 160,), 7.0, dtype=torch.float32, device="cuda")
    yv = ops.View.nhwc(out, N, H, W, 160).ch(16, 16 + Cout)
    wp, ld = ops.pack_weight(w.cuda(), 0)
    st = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    ops.conv2d(xv, wp, ld, 1, 1, 1, 0, Cout, yv, scale=sc.cuda(), shift=sh.cuda(), slope=0.0, stats=st, stats_ld=Cout, impl=ops.IMPL_UMMA)
    torch.cuda.synchronize()
    o = out.view(N, H, W, 160)
    assert maxabs(o[..., 16:16 + Cout].permute(0, 3, 1, 2), y) <= 5e-5
    assert bool((o[..., :16] == 7.0).all()) and bool((o[..., 16 + Cout:] == 7.0).all())      # neighbours of the output slice untouched
    assert maxabs(st[:Cout], y.sum((0, 2, 3))) <= 2e-6 * N * H * W + 1e-3


@pytest.mark.parametrize("Cout,R,pad,mask", [(288, 4, 2, True), (288, 4, 2, False), (8, 4, 1, True), (288, 3, 1, True)])
def test_conv2d_single_input_channel(Cout, R, pad, mask):
    """Cin == 1 direct kernels (data gradient of Fusion-D layer 5: 1 -> 288 channels, 4x4, LeakyReLU mask of the layer
    input): the register-weight 4x4 kernel and the generic one (3x3), automatic dispatch, fp32 arithmetic."""
    ops = _ops()
    N, H, W = 2, 9, 10
    x = seeded((N, 1, H, W), 1, -1.0, 1.0)
    w = seeded((Cout, 1, R, R), 2, -1.0, 1.0) / R
    y = F.conv2d(x.double(), w.double(), padding=pad)
    OH, OW = y.shape[-2:]
    e = seeded((N, Cout, OH, OW), 6, -1.0, 1.0) if mask else None
    if mask:
        y = y * torch.where(e.double() > 0, 1.0, 0.2)
    wp, ld = ops.pack_weight(w.cuda(), 0)
    yd = cl(torch.full((N, Cout, OH, OW), 7.0))
    ops.conv2d(ops.View.from_nchw(cl(x)), wp, ld, R, R, 1, pad, Cout, ops.View.from_nchw(yd), alpha=0.5,
               e=ops.View.from_nchw(cl(e)) if mask else None, eslope=0.2)
    assert maxabs(yd, 0.5 * y) <= 2e-6


def test_conv2d_nchw_output_and_errors():
    ops = _ops()
    x = seeded((1, 16, 9, 9), 1, -1, 1)
    w = seeded((3, 16, 3, 3), 2, -1, 1) / 12
    y = torch.empty(1, 3, 9, 9, device="cuda")
    wp, ld = ops.pack_weight(w.cuda(), 0)
    ops.conv2d(ops.View.from_nchw(cl(x)), wp, ld, 3, 3, 1, 1, 3, ops.View.from_nchw(y), impl=ops.IMPL_SIMT)
    assert maxabs(y, F.conv2d(x, w, padding=1)) <= 1e-5
    with pytest.raises(ValueError):
        ops.conv2d(ops.View.from_nchw(cl(x)), wp, ld, 3, 3, 1, 1, 4, ops.View.from_nchw(y))
    with pytest.raises(RuntimeError):  # C ABI rejects an inconsistent descriptor (w_ld < Cout)
        ops.conv2d(ops.View.from_nchw(cl(x)), wp, 2, 3, 3, 1, 1, 3, ops.View.from_nchw(y))


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_pack_weight(mode):
    ops = _ops()
    if mode == 2:
        w = seeded((24, 10, 1, 1), 1)
        want = w[:, :, 0, 0].t()                      # [co][ci]
    else:
        w = seeded((10, 24, 3, 3), 1)
        if mode == 0:
            want = w.permute(2, 3, 1, 0).reshape(9 * 24, 10)
        else:
            want = w.flip(2, 3).permute(2, 3, 0, 1).reshape(9 * 10, 24)
    out, ld = ops.pack_weight(w.cuda(), mode)
    got = out.view(-1, ld)[:, :want.shape[1]]
    assert maxabs(got, want) == 0.0
    assert float(out.view(-1, ld)[:, want.shape[1]:].abs().sum()) == 0.0


WGRAD_CASES = [
    # Cin, Cout, R, stride, pad, H, W, gather, affine, slope, transposed, dbias
    (128, 32, 3, 1, 1, 12, 10, 0, True, 0.0, False, False),
    (96, 128, 1, 1, 0, 9, 7, 0, True, 0.0, False, False),
    (256, 128, 1, 1, 0, 6, 6, 1, True, 0.0, False, False),
    (3, 64, 3, 1, 1, 9, 11, 0, False, 1.0, False, True),
    (96, 16, 1, 1, 0, 7, 5, 0, False, 0.0, True, False),
    (9, 36, 4, 2, 1, 16, 14, 0, False, 1.0, False, False),
    (144, 288, 4, 1, 1, 7, 7, 0, True, 0.2, False, False),
    (160, 130, 3, 1, 1, 5, 6, 0, False, 1.0, False, True),
    (160, 136, 3, 1, 1, 9, 11, 0, False, 1.0, False, True),      # two channel blocks (padded), two co tiles at NT=128? (136 -> NT 256)
    (64, 32, 1, 1, 0, 40, 24, 1, False, 1.0, False, True),       # conv_refin2 (avg-pool gather), many pixel chunks
    (1024, 256, 3, 1, 1, 6, 5, 0, False, 0.0, False, False),     # dense_block4.conv2
    (512, 64, 1, 1, 0, 12, 12, 0, False, 0.0, True, False),      # trans_block5 (ConvTranspose layout)
    (128, 32, 3, 1, 1, 33, 29, 0, True, 0.0, False, False),      # K1 with ragged pixel count
    (288, 1, 4, 1, 1, 9, 10, 0, False, 0.2, False, False),       # D layer 5 (single output channel, scalar gradient loads)
    (16, 3, 3, 1, 1, 12, 12, 0, False, 1.0, False, True),        # conv_refin3 (NCHW gradient)
    (128, 32, 3, 1, 1, 64, 48, 0, True, 0.0, False, False),      # K1, many 8x8 pixel blocks (halo weight-gradient kernel)
    (160, 40, 3, 1, 1, 19, 21, 0, True, 0.2, False, False),      # two channel blocks, two output-channel tiles, ragged blocks
    (64, 24, 4, 1, 1, 12, 12, 0, False, 1.0, False, False),      # 4x4 filter: 16 taps = all 512 TMEM columns
    (160, 32, 3, 1, 1, 19, 37, 0, True, 0.2, False, False),      # growth-convolution kernel: two channel blocks (second one padded), ragged 16x8 tiles
    (64, 8, 3, 1, 1, 17, 35, 0, False, 1.0, False, True),        # ... Cout < 32 (zero-padded gradient channels), half-empty channel block, bias gradient
    (128, 32, 3, 1, 1, 8, 16, 0, True, 0.0, False, False),       # ... exactly one tile per image
]


@pytest.mark.parametrize("case", WGRAD_CASES)
@pytest.mark.parametrize("variant", ["simt", "tcgen05", "tcgen05_pertap"])
def test_wgrad(case, variant):
    ops = _ops()
    Cin, Cout, R, stride, pad, H, W, gather, affine, slope, transposed, dbias = case
    if variant.startswith("tcgen05") and not (Cin % 8 == 0 and Cin >= 16):
        pytest.skip("shape not covered by the tcgen05 weight-gradient path (runs on the SIMT kernel)")
    halo_shape = gather == 0 and stride == 1 and 2 <= R <= 4 and Cout <= 64 and not transposed
    if variant == "tcgen05_pertap" and not halo_shape:
        pytest.skip("same kernel as the tcgen05 variant for this shape")
    from fdgan_b200 import _lib
    _lib.set_option("halo", 0 if variant == "tcgen05_pertap" else 1)
    impl = ops.IMPL_UMMA if variant.startswith("tcgen05") else ops.IMPL_SIMT
    N = 3
    ph, pw = (2 * H, 2 * W) if gather == 1 else (H, W)
    x = seeded((N, Cin, ph, pw), 1, -1, 1).double()
    sc = seeded((Cin,), 4, 0.5, 1.5).double() if affine else None
    sh = seeded((Cin,), 5, -0.3, 0.3).double() if affine else None
    a = ref_prologue(x, sc, sh, slope)
    if gather == 1:
        a = F.avg_pool2d(a, 2)
    w = torch.zeros(Cout, Cin, R, R, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(a, w, stride=stride, padding=pad)
    g = seeded(tuple(y.shape), 8, -1, 1).double()
    (y * g).sum().backward()
    want = w.grad
    if transposed:
        want = want[:, :, 0, 0].t().reshape(Cin, Cout, 1, 1)
    dw = torch.zeros(tuple(want.shape), device="cuda")
    db = torch.zeros(Cout, device="cuda") if dbias else None
    nchw_small = Cin < 16
    xd = x.float().cuda() if nchw_small else cl(x.float())
    ops.wgrad(ops.View.from_nchw(xd), ops.View.from_nchw(cl(g.float())), R, R, stride, pad, dw, gather=gather,
              scale=sc.float().cuda() if affine else None, shift=sh.float().cuda() if affine else None, slope=slope,
              transposed=transposed, dbias=db, impl=impl)
    _lib.set_option("halo", 1)
    assert maxabs(dw, want) <= 2e-4 * max(1.0, float(want.abs().max()))
    if dbias:
        assert maxabs(db, g.sum((0, 2, 3))) <= 1e-3


def test_bn_finalize_and_backward_pieces():
    ops = _ops()
    N, C, H, W = 3, 20, 6, 5
    x = seeded((N, C, H, W), 1, -2, 3).double().requires_grad_(True)
    gamma = seeded((C,), 2, 0.5, 1.5).double().requires_grad_(True)
    beta = seeded((C,), 3, -0.5, 0.5).double().requires_grad_(True)
    rm, rv = seeded((C,), 4, -0.1, 0.1).double(), seeded((C,), 5, 0.5, 1.5).double()
    rm0, rv0 = rm.clone(), rv.clone()
    y = F.leaky_relu(F.batch_norm(x, rm, rv, gamma, beta, True, 0.1, 1e-5), 0.2)
    g = seeded((N, C, H, W), 6, -1, 1).double()
    (y * g).sum().backward()
    # device: stats -> finalize -> bwd stats -> bwd finalize -> apply
    xd = cl(x.detach().float())
    st = torch.stack([x.detach().sum((0, 2, 3)), (x.detach() ** 2).sum((0, 2, 3))]).reshape(-1).cuda()
    buf = torch.zeros(4 * C, device="cuda")
    rmd, rvd = rm0.float().cuda(), rv0.float().cuda()
    ops.bn_finalize(st, C, C, N * H * W, gamma.detach().float().cuda(), beta.detach().float().cuda(), 1e-5, 0.1, rmd, rvd,
                    True, buf[:C], buf[C:2 * C], buf[2 * C:3 * C], buf[3 * C:])
    assert maxabs(rmd, rm) <= 1e-6 and maxabs(rvd, rv) <= 1e-6
    xv, gv = ops.View.from_nchw(xd), ops.View.from_nchw(cl(g.float()))
    st2 = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    ops.ew_bwd(gv, xv, stats=st2, scale=buf[:C], shift=buf[C:2 * C], slope=0.2)
    coef = torch.empty(3 * C, device="cuda")
    dgm, dbt = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.bn_bwd_finalize(st2, C, N * H * W, gamma.detach().float().cuda(), buf[2 * C:3 * C], buf[3 * C:], coef, dgm, dbt)
    base = seeded((N, C, H, W), 9, -1, 1)
    outd = cl(base.clone())
    ops.ew_bwd(gv, xv, out=ops.View.from_nchw(outd), scale=buf[:C], shift=buf[C:2 * C], slope=0.2, coef=coef, accumulate=True)
    assert maxabs(outd, x.grad + base.double()) <= 2e-5
    assert maxabs(dgm, gamma.grad) <= 1e-4 and maxabs(dbt, beta.grad) <= 1e-4
    # eval mode uses the running statistics
    ops.bn_finalize(None, 0, C, 1, gamma.detach().float().cuda(), beta.detach().float().cuda(), 1e-5, 0.1, rmd, rvd, False,
                    buf[:C], buf[C:2 * C])
    want = gamma.detach() / torch.sqrt(rv + 1e-5)
    assert maxabs(buf[:C], want) <= 1e-5


@pytest.mark.parametrize("shape", [(2, 128, 96, 9, 7, 0.0), (1, 128, 224, 33, 31, 0.0), (2, 64, 160, 12, 10, 0.2), (2, 128, 64, 10, 6, 0.0)])
def test_conv2d_bn_backward_epilogue(shape):
    """1x1 data-gradient conv with the BatchNorm-backward epilogue (FdgConv.e_scale) + fdg_bn_bwd_finalize + fdg_affine_accum
    == autograd through conv1x1(leaky_relu(batch_norm(x))) w.r.t. x (dense-layer norm1/conv1 backward, torchvision
    densenet.py:_DenseLayer), accumulated into an existing gradient buffer."""
    ops = _ops()
    N, Cmid, Cx, H, W, slope = shape
    x = seeded((N, Cx, H, W), 1, -2, 3).double().requires_grad_(True)
    gamma = seeded((Cx,), 2, 0.5, 1.5).double().requires_grad_(True)
    beta = seeded((Cx,), 3, -0.5, 0.5).double().requires_grad_(True)
    w = (seeded((Cmid, Cx, 1, 1), 4, -1, 1) / math.sqrt(Cx)).double()
    gout = seeded((N, Cmid, H, W), 5, -1, 1).double()            # dL/d(conv output)
    y = F.conv2d(F.leaky_relu(F.batch_norm(x, None, None, gamma, beta, True, 0.1, 1e-5), slope), w)
    (y * gout).sum().backward()
    base = seeded((N, Cx, H, W), 9, -1, 1)
    # ---- device: forward statistics -> scale/shift, then the fused backward
    cnt = N * H * W
    xd = cl(x.detach().float())
    st = torch.stack([x.detach().sum((0, 2, 3)), (x.detach() ** 2).sum((0, 2, 3))]).reshape(-1).cuda()
    buf = torch.zeros(4 * Cx, device="cuda")
    ops.bn_finalize(st, Cx, Cx, cnt, gamma.detach().float().cuda(), beta.detach().float().cuda(), 1e-5, 0.1, None, None, True,
                    buf[:Cx], buf[Cx:2 * Cx], buf[2 * Cx:3 * Cx], buf[3 * Cx:])
    dxd = cl(base.clone())
    st2 = torch.zeros(2 * Cx, dtype=torch.float64, device="cuda")
    wd = w.float().cuda().reshape(Cmid, Cx).contiguous()          # [K = Cmid][Cx]: the 1x1 data-gradient operand
    xv = ops.View.from_nchw(xd)
    ops.conv2d(ops.View.from_nchw(cl(gout.float())), wd, Cx, 1, 1, 1, 0, Cx, ops.View.from_nchw(dxd), store=ops.STORE_ACCUM,
               e=xv, eslope=slope, e_scale=buf[:Cx], e_shift=buf[Cx:2 * Cx], stats=st2, stats_ld=Cx)
    coef = torch.empty(3 * Cx, device="cuda")
    dgm, dbt = torch.zeros(Cx, device="cuda"), torch.zeros(Cx, device="cuda")
    ops.bn_bwd_finalize(st2, Cx, cnt, gamma.detach().float().cuda(), buf[2 * Cx:3 * Cx], buf[3 * Cx:], coef, dgm, dbt)
    assert maxabs(coef[:Cx], buf[:Cx]) <= 1e-6                     # alpha == the forward scale (what the epilogue multiplied by)
    ops.affine_accum(xv, ops.View.from_nchw(dxd), coef[Cx:2 * Cx], coef[2 * Cx:])
    torch.cuda.synchronize()
    assert maxabs(dxd, x.grad + base.double()) <= 1e-4
    scale_g = float(gamma.grad.abs().max())
    assert maxabs(dgm, gamma.grad) <= 2e-4 * max(1.0, scale_g) and maxabs(dbt, beta.grad) <= 2e-4 * max(1.0, float(beta.grad.abs().max()))


@pytest.mark.parametrize("shape", [(2, 32, 128, 19, 13, 0.0), (1, 32, 128, 40, 24, 0.0), (2, 48, 96, 9, 11, 0.2),
                                   # two output-channel tiles / LeakyReLU / a 16-channel gradient on the two-epilogue-set instantiation (conv_halo<128,3,BN2>)
                                   (1, 32, 256, 17, 9, 0.2), (3, 16, 128, 8, 8, 0.0), (16, 32, 128, 32, 32, 0.0)])
def test_conv2d_3x3_bn_backward_epilogue(shape):
    """3x3 data-gradient conv (halo-tile kernel) with the BatchNorm-backward epilogue and a NORMAL store + fdg_bn_bwd_finalize(unit_alpha) +
    one mask-free fdg_ew_bwd pass == autograd through conv3x3(leaky_relu(batch_norm(t))) w.r.t. t (dense-layer norm2 / conv2 backward)."""
    ops = _ops()
    N, Cout, Ct, H, W, slope = shape
    t = seeded((N, Ct, H, W), 1, -2, 3).double().requires_grad_(True)
    gamma = seeded((Ct,), 2, 0.5, 1.5).double().requires_grad_(True)
    beta = seeded((Ct,), 3, -0.5, 0.5).double().requires_grad_(True)
    w = (seeded((Cout, Ct, 3, 3), 4, -1, 1) / math.sqrt(9 * Ct)).double()
    gout = seeded((N, Cout, H, W), 5, -1, 1).double()
    y = F.conv2d(F.leaky_relu(F.batch_norm(t, None, None, gamma, beta, True, 0.1, 1e-5), slope), w, padding=1)
    (y * gout).sum().backward()
    cnt = N * H * W
    td = cl(t.detach().float())
    st = torch.stack([t.detach().sum((0, 2, 3)), (t.detach() ** 2).sum((0, 2, 3))]).reshape(-1).cuda()
    buf = torch.zeros(4 * Ct, device="cuda")
    ops.bn_finalize(st, Ct, Ct, cnt, gamma.detach().float().cuda(), beta.detach().float().cuda(), 1e-5, 0.1, None, None, True,
                    buf[:Ct], buf[Ct:2 * Ct], buf[2 * Ct:3 * Ct], buf[3 * Ct:])
    wd, ld = ops.pack_weight(w.float().cuda(), 1)                  # flipped [(r,s,co)][ci] operand of the data gradient
    dz = cl(torch.zeros(N, Ct, H, W))
    st2 = torch.zeros(2 * Ct, dtype=torch.float64, device="cuda")
    tv = ops.View.from_nchw(td)
    ops.conv2d(ops.View.from_nchw(cl(gout.float())), wd, ld, 3, 3, 1, 1, Ct, ops.View.from_nchw(dz), e=tv, eslope=slope,
               e_scale=buf[:Ct], e_shift=buf[Ct:2 * Ct], stats=st2, stats_ld=Ct)
    coef = torch.empty(3 * Ct, device="cuda")
    dgm, dbt = torch.zeros(Ct, device="cuda"), torch.zeros(Ct, device="cuda")
    ops.bn_bwd_finalize(st2, Ct, cnt, gamma.detach().float().cuda(), buf[2 * Ct:3 * Ct], buf[3 * Ct:], coef, dgm, dbt, unit_alpha=True)
    assert maxabs(coef[:Ct], torch.ones(Ct)) == 0.0
    dx = cl(torch.zeros(N, Ct, H, W))
    ops.ew_bwd(ops.View.from_nchw(dz), tv, out=ops.View.from_nchw(dx), coef=coef, slope=1.0)
    torch.cuda.synchronize()
    assert maxabs(dx, t.grad) <= 1e-4
    assert maxabs(dgm, gamma.grad) <= 2e-4 * max(1.0, float(gamma.grad.abs().max())) and maxabs(dbt, beta.grad) <= 2e-4 * max(1.0, float(beta.grad.abs().max()))


def test_split_bf16_operands():
    """Split-bf16 planes (fdg_ew_bwd out_split) feeding fdg_conv2d (x_split, BatchNorm-backward epilogue) and
    fdg_conv2d_wgrad (g_split) give the same results as the fp32 tensors they replace: the kernels apply exactly this
    split to fp32 operands themselves."""
    ops = _ops()
    N, H, W, Cm, Cx = 2, 20, 13, 128, 160
    P = N * H * W
    g = seeded((N, Cm, H, W), 1, -1, 1)
    t = seeded((N, Cm, H, W), 2, -1, 1)
    coef = torch.cat([seeded((Cm,), 3, 0.5, 1.5), seeded((Cm,), 4, -0.2, 0.2), seeded((Cm,), 5, -0.1, 0.1)]).cuda()
    gv, tv = ops.View.from_nchw(cl(g)), ops.View.from_nchw(cl(t))
    # fp32 apply pass vs split apply pass
    d32 = ops.View.alloc(N, H, W, Cm, "cuda")
    ops.ew_bwd(gv, tv, out=d32, slope=0.0, coef=coef)
    planes = torch.empty(P * Cm, dtype=torch.float32, device="cuda")
    ops.ew_bwd(gv, tv, slope=0.0, coef=coef, out_split=planes)
    bf = planes.view(torch.bfloat16)
    rec = (bf[:P * Cm].float() + bf[P * Cm:].float()).view(N, H, W, Cm)
    assert maxabs(rec, d32.base.view(N, H, W, Cm)) <= 1e-5
    # conv (BatchNorm-backward epilogue) from the planes vs from the fp32 tensor
    x = seeded((N, Cx, H, W), 6, -2, 2)
    xv = ops.View.from_nchw(cl(x))
    sc, sh = seeded((Cx,), 7, 0.5, 1.5).cuda(), seeded((Cx,), 8, -0.3, 0.3).cuda()
    wd = (seeded((Cm, Cx), 9, -1, 1) / math.sqrt(Cm)).cuda().contiguous()
    outs = []
    for use_split in (False, True):
        y = cl(seeded((N, Cx, H, W), 10, -1, 1))
        st = torch.zeros(2 * Cx, dtype=torch.float64, device="cuda")
        src = ops.View.nhwc(planes, N, H, W, Cm) if use_split else d32
        ops.conv2d(src, wd, Cx, 1, 1, 1, 0, Cx, ops.View.from_nchw(y), store=ops.STORE_ACCUM, e=xv, eslope=0.0, e_scale=sc, e_shift=sh,
                   stats=st, stats_ld=Cx, x_split=planes if use_split else None)
        outs.append((y, st))
    torch.cuda.synchronize()
    assert maxabs(outs[1][0], outs[0][0]) <= 2e-5 and maxabs(outs[1][1], outs[0][1]) <= 1e-2
    # weight gradient with the gradient operand from the planes
    w1 = torch.zeros(Cm, Cx, 1, 1, device="cuda")
    w2 = torch.zeros_like(w1)
    ops.wgrad(xv, d32, 1, 1, 1, 0, w1, scale=sc, shift=sh, slope=0.0)
    ops.wgrad(xv, ops.View.nhwc(planes, N, H, W, Cm), 1, 1, 1, 0, w2, scale=sc, shift=sh, slope=0.0, g_split=planes)
    torch.cuda.synchronize()
    assert maxabs(w2, w1) <= 1e-4 * max(1.0, float(w1.abs().max()))


@pytest.mark.parametrize("cin,cout,H,W", [(36, 72, 21, 19), (36, 40, 16, 32), (20, 96, 9, 17), (132, 24, 12, 12)])
def test_wgrad_3x3_with_cin_multiple_of_4(cin, cout, H, W):
    """Fusion-D layer 2 (36 -> 72, 3x3) and relatives: Cin % 8 != 0 shapes run on the growth-convolution weight-gradient kernel with up to
    three 32-channel output tiles (automatic dispatch), LeakyReLU prologue."""
    ops = _ops()
    N = 2
    x = seeded((N, cin, H, W), 1, -1, 1)
    g = seeded((N, cout, H, W), 2, -1, 1)
    ref = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    (F.conv2d(F.leaky_relu(x.double(), 0.2), ref, padding=1) * g.double()).sum().backward()
    dw = torch.zeros(cout, cin, 3, 3, device="cuda")
    ops.wgrad(ops.View.from_nchw(cl(x)), ops.View.from_nchw(cl(g)), 3, 3, 1, 1, dw, slope=0.2)
    torch.cuda.synchronize()
    assert maxabs(dw, ref.grad) <= 5e-5 * max(1.0, float(ref.grad.abs().max()))


@pytest.mark.parametrize("cin,cout,R,pad", [(144, 288, 4, 1), (72, 144, 3, 1), (64, 72, 3, 1)])
def test_wgrad_wide_tiles_take_split_planes(cin, cout, R, pad):
    """Fusion-D layer 4 / layer 3 weight gradients (wide output-channel tiles, Cout % 64 != 0): gradient operand from split-bf16 planes
    written by fdg_ew_bwd (planes AND the fp32 tensor in one pass) == the fp32 gradient operand."""
    ops = _ops()
    N, H, W = 2, 13, 11
    x = seeded((N, cin, H, W), 1, -1, 1)
    OH = H + 2 * pad - R + 1
    OW = W + 2 * pad - R + 1
    g = seeded((N, cout, OH, OW), 2, -1, 1)
    t = seeded((N, cout, OH, OW), 3, -1, 1)
    xv, gv, tv = ops.View.from_nchw(cl(x)), ops.View.from_nchw(cl(g)), ops.View.from_nchw(cl(t))
    P = N * OH * OW
    planes = torch.empty(P * cout, dtype=torch.float32, device="cuda")
    d32 = ops.View.alloc(N, OH, OW, cout, "cuda")
    ops.ew_bwd(gv, tv, out=d32, slope=0.2, out_split=planes)       # LeakyReLU mask; both outputs
    want = g * torch.where(t > 0, 1.0, 0.2)
    assert maxabs(d32.as_nchw(), want) <= 1e-6
    bf = planes.view(torch.bfloat16)
    assert maxabs((bf[:P * cout].float() + bf[P * cout:].float()).view(N, OH, OW, cout).permute(0, 3, 1, 2), want) <= 1e-5
    w1 = torch.zeros(cout, cin, R, R, device="cuda")
    w2 = torch.zeros_like(w1)
    ops.wgrad(xv, d32, R, R, 1, pad, w1, slope=0.2)
    ops.wgrad(xv, ops.View.nhwc(planes, N, OH, OW, cout), R, R, 1, pad, w2, slope=0.2, g_split=planes)
    torch.cuda.synchronize()
    ref = torch.zeros(cout, cin, R, R, dtype=torch.float64, requires_grad=True)
    (F.conv2d(F.leaky_relu(x.double(), 0.2), ref, padding=pad) * want.double()).sum().backward()
    assert maxabs(w1, ref.grad) <= 5e-5 * max(1.0, float(ref.grad.abs().max()))
    assert maxabs(w2, ref.grad) <= 5e-5 * max(1.0, float(ref.grad.abs().max()))


def test_ew_bwd_pooled_gradient_scalar_path():
    ops = _ops()
    N, C, H, W = 2, 9, 6, 8   # C=9: scalar path
    x = seeded((N, C, H, W), 1, -1, 1)
    g = seeded((N, C, H // 2, W // 2), 2, -1, 1)
    out = torch.zeros(N, C, H, W, device="cuda")
    ops.ew_bwd(ops.View.from_nchw(g.cuda()), ops.View.from_nchw(x.cuda()), out=ops.View.from_nchw(out), slope=0.0,
               g_gather=ops.GATHER_UP2, gscale=0.25)
    want = 0.25 * F.interpolate(g, scale_factor=2, mode="nearest") * (x > 0)
    assert maxabs(out, want) <= 1e-6


def test_ew_bwd_pooled_gradient_vector_path():
    """Transition backward: gradient at half resolution (adjoint of the 2x2 average pool folded into the gather), BatchNorm
    scale/shift + ReLU mask, statistics pass and apply pass with (alpha, beta, delta); 128-bit generic kernel."""
    ops = _ops()
    N, C, H, W = 3, 24, 10, 14
    x = seeded((N, C, H, W), 1, -1, 1)
    g = seeded((N, C, H // 2, W // 2), 2, -1, 1)
    sc, sh = seeded((C,), 3, 0.5, 1.5), seeded((C,), 4, -0.3, 0.3)
    coef = torch.cat([seeded((C,), 5, 0.5, 1.5), seeded((C,), 6, -0.2, 0.2), seeded((C,), 7, -0.1, 0.1)])
    gv, xv = ops.View.from_nchw(cl(g)), ops.View.from_nchw(cl(x))
    v = x.double() * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)
    dz = 0.25 * F.interpolate(g.double(), scale_factor=2, mode="nearest") * (v > 0)
    st = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    ops.ew_bwd(gv, xv, stats=st, scale=sc.cuda(), shift=sh.cuda(), slope=0.0, g_gather=ops.GATHER_UP2, gscale=0.25)
    assert maxabs(st[:C], dz.sum((0, 2, 3))) <= 1e-4 and maxabs(st[C:], (dz * x.double()).sum((0, 2, 3))) <= 1e-4
    out0 = seeded((N, C, H, W), 8, -1, 1)
    out = cl(out0.clone())
    ops.ew_bwd(gv, xv, out=ops.View.from_nchw(out), scale=sc.cuda(), shift=sh.cuda(), slope=0.0, g_gather=ops.GATHER_UP2, gscale=0.25,
               coef=coef.cuda(), accumulate=True)
    a_, b_, d_ = (coef[i * C:(i + 1) * C].double().view(1, -1, 1, 1) for i in range(3))
    assert maxabs(out, out0.double() + a_ * dz + b_ * x.double() + d_) <= 1e-5


def test_ew_bwd_pooled_gradient_on_channel_slices():
    """Same pass on channel slices of wider NHWC buffers (the transition backward accumulates into a dense block's gradient buffer),
    several CTAs per channel group, LeakyReLU slope, plain (non-accumulating) store."""
    ops = _ops()
    N, C, H, W, CT = 2, 256, 36, 28, 320
    xb = seeded((N, H, W, CT), 1, -1, 1)
    gb = seeded((N, H // 2, W // 2, CT), 2, -1, 1)
    sc, sh = seeded((C,), 3, 0.5, 1.5), seeded((C,), 4, -0.3, 0.3)
    coef = torch.cat([seeded((C,), 5, 0.5, 1.5), seeded((C,), 6, -0.2, 0.2), seeded((C,), 7, -0.1, 0.1)])
    xd, gd = xb.cuda(), gb.cuda()
    xv = ops.View.nhwc(xd, N, H, W, CT).ch(32, 32 + C)
    gv = ops.View.nhwc(gd, N, H // 2, W // 2, CT).ch(64, 64 + C)
    x = xb[..., 32:32 + C].permute(0, 3, 1, 2).double()
    g = gb[..., 64:64 + C].permute(0, 3, 1, 2).double()
    v = x * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)
    dz = 0.25 * F.interpolate(g, scale_factor=2, mode="nearest") * torch.where(v > 0, 1.0, 0.1)
    st = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    ops.ew_bwd(gv, xv, stats=st, scale=sc.cuda(), shift=sh.cuda(), slope=0.1, g_gather=ops.GATHER_UP2, gscale=0.25)
    assert maxabs(st[:C], dz.sum((0, 2, 3))) <= 2e-4 and maxabs(st[C:], (dz * x).sum((0, 2, 3))) <= 2e-4
    ob = torch.full((N, H, W, CT), 3.0, device="cuda")
    ov = ops.View.nhwc(ob, N, H, W, CT).ch(16, 16 + C)
    ops.ew_bwd(gv, xv, out=ov, scale=sc.cuda(), shift=sh.cuda(), slope=0.1, g_gather=ops.GATHER_UP2, gscale=0.25, coef=coef.cuda())
    a_, b_, d_ = (coef[i * C:(i + 1) * C].double().view(1, -1, 1, 1) for i in range(3))
    assert maxabs(ob[..., 16:16 + C].permute(0, 3, 1, 2), a_ * dz + b_ * x + d_) <= 1e-5
    assert bool((ob[..., :16] == 3.0).all()) and bool((ob[..., 16 + C:] == 3.0).all())


def test_pool2_bn_act():
    """fdg_pool2_bn_act == avg_pool2d(leaky_relu(x * scale + shift)) on a channel slice of a wider buffer, with and without the affine."""
    ops = _ops()
    N, H, W, C, CT = 2, 12, 20, 64, 96
    xb = seeded((N, H, W, CT), 1, -1, 1)
    sc, sh = seeded((C,), 2, 0.5, 1.5), seeded((C,), 3, -0.3, 0.3)
    x = xb[..., 16:16 + C].permute(0, 3, 1, 2).double()
    xv = ops.View.nhwc(xb.cuda(), N, H, W, CT).ch(16, 16 + C)
    for affine, slope in ((True, 0.0), (False, 1.0), (True, 0.2)):
        v = x * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1) if affine else x
        want = F.avg_pool2d(torch.where(v > 0, v, slope * v), 2)
        y = ops.View.alloc(N, H // 2, W // 2, C, "cuda")
        ops.pool2_bn_act(xv, y, sc.cuda() if affine else None, sh.cuda() if affine else None, slope)
        assert maxabs(y.as_nchw(), want) <= 1e-6


@pytest.mark.parametrize("cin,R,pad,H,W", [(288, 4, 1, 15, 13), (64, 3, 1, 9, 20), (72, 4, 2, 8, 8)])
def test_single_output_channel_conv_by_taps(cin, R, pad, H, W):
    """Fusion-D layer 5 re-associated: 1x1 convolution Cin -> R*S on the tensor cores + fdg_tap_sum == conv2d with one output channel;
    fdg_tap_spread + transposed 1x1 weight gradient == its weight gradient (the OIHW parameter is used as the [Cin][R*S] operand as is)."""
    ops = _ops()
    N, T = 2, R * R
    x = seeded((N, cin, H, W), 1, -1, 1)
    w = (seeded((1, cin, R, R), 2, -1, 1) / math.sqrt(cin * T)).requires_grad_(True)
    y = torch.sigmoid(F.conv2d(F.leaky_relu(x.double(), 0.2), w.double(), padding=pad))
    OH, OW = y.shape[-2:]
    g = seeded((N, 1, OH, OW), 3, -1, 1)
    (F.conv2d(F.leaky_relu(x.double(), 0.2), w.double(), padding=pad) * g.double()).sum().backward()
    xv = ops.View.from_nchw(cl(x))
    wd = w.detach().cuda().contiguous()
    s = ops.View.alloc(N, H, W, T, "cuda")
    ops.conv2d(xv, wd, T, 1, 1, 1, 0, T, s, slope=0.2, impl=ops.IMPL_UMMA)
    out = torch.empty(N, 1, OH, OW, device="cuda")
    ops.tap_sum(s, ops.View.from_nchw(out), R, R, pad, act=ops.ACT_SIGMOID)
    assert maxabs(out, y) <= 5e-5
    gs = ops.View.alloc(N, H, W, T, "cuda")
    ops.tap_spread(ops.View.from_nchw(g.cuda()), gs, R, R, pad)
    dw = torch.zeros(1, cin, R, R, device="cuda")
    ops.wgrad(xv, gs, 1, 1, 1, 0, dw, slope=0.2, transposed=True)
    torch.cuda.synchronize()
    assert maxabs(dw, w.grad) <= 5e-5 * max(1.0, float(w.grad.abs().max()))


def test_maxpool_copy_colsum_actbwd():
    ops = _ops()
    x = seeded((2, 12, 9, 10), 1, -1, 1).requires_grad_(True)
    y = F.max_pool2d(x, 2, 2)
    g = seeded(tuple(y.shape), 2, -1, 1)
    (y * g).sum().backward()
    xd = cl(x.detach())
    yd = cl(torch.zeros_like(y))
    ops.maxpool2_fwd(ops.View.from_nchw(xd), ops.View.from_nchw(yd))
    assert maxabs(yd, y) == 0.0
    gx = cl(torch.zeros_like(x))
    ops.maxpool2_bwd(ops.View.from_nchw(xd), ops.View.from_nchw(cl(g)), ops.View.from_nchw(gx), accumulate=True)
    assert maxabs(gx, x.grad) == 0.0
    gx2 = cl(torch.full_like(x.detach(), 7.0))       # first-writer mode with the ReLU mask of a post-ReLU input
    ops.maxpool2_bwd(ops.View.from_nchw(xd), ops.View.from_nchw(cl(g)), ops.View.from_nchw(gx2), accumulate=False, relu_mask=True)
    assert maxabs(gx2[:, :, :8, :], (x.grad * (x.detach() > 0))[:, :, :8, :]) == 0.0
    x3 = seeded((2, 3, 8, 6), 9, -1, 1)              # scalar kernels (C = 3, NCHW)
    g3 = seeded((2, 3, 4, 3), 10, -1, 1)
    gx3 = torch.full((2, 3, 8, 6), 7.0, device="cuda")
    ops.maxpool2_bwd(ops.View.from_nchw(x3.cuda()), ops.View.from_nchw(g3.cuda()), ops.View.from_nchw(gx3), accumulate=False, relu_mask=True)
    x3r = x3.clone().requires_grad_(True)
    (F.max_pool2d(x3r, 2, 2) * g3).sum().backward()
    assert maxabs(gx3, x3r.grad * (x3 > 0)) == 0.0
    # copy4d: adjoint of nearest x2 and of avg-pool
    a = seeded((2, 8, 6, 6), 3, -1, 1)
    o = cl(torch.zeros(2, 8, 3, 3))
    ops.copy4d(ops.View.from_nchw(cl(a)), ops.View.from_nchw(o), gather=ops.GATHER_AVGPOOL2, scale=4.0)
    assert maxabs(o, 4 * F.avg_pool2d(a, 2)) <= 1e-6
    o2 = cl(torch.ones(2, 8, 12, 12))
    ops.copy4d(ops.View.from_nchw(cl(a)), ops.View.from_nchw(o2), gather=ops.GATHER_UP2, scale=0.25, slope=0.0, accumulate=True)
    assert maxabs(o2, 1 + 0.25 * F.interpolate(torch.relu(a), scale_factor=2, mode="nearest")) <= 1e-6
    # direct copy between channel slices of NHWC buffers (four channels per thread) and between NCHW views (scalar path)
    src = seeded((2, 16, 5, 7), 6, -1, 1)
    dst = cl(torch.zeros(2, 24, 5, 7))
    ops.copy4d(ops.View.from_nchw(cl(src)).ch(4, 12), ops.View.from_nchw(dst).ch(8, 16), slope=0.0)
    exp = torch.zeros(2, 24, 5, 7)
    exp[:, 8:16] = torch.relu(src[:, 4:12])
    assert maxabs(dst, exp) == 0.0
    ops.copy4d(ops.View.from_nchw(cl(src)).ch(2, 8), ops.View.from_nchw(dst).ch(16, 22), scale=0.5, accumulate=True)   # slices off 16 bytes
    exp[:, 16:22] += 0.5 * src[:, 2:8]
    assert maxabs(dst, exp) == 0.0
    d2 = torch.ones(2, 3, 5, 7, device="cuda")
    ops.copy4d(ops.View.from_nchw(src[:, :3].contiguous().cuda()), ops.View.from_nchw(d2), accumulate=True)
    assert maxabs(d2, 1 + src[:, :3]) == 0.0
    cs = torch.zeros(8, device="cuda")
    ops.colsum(ops.View.from_nchw(cl(a)), cs)
    assert maxabs(cs, a.sum((0, 2, 3))) <= 1e-4
    # 128-bit column sums: channel-group counts that do not divide the CTA, several channel tiles, accumulate, a channel slice;
    # scalar kernel for C = 3
    for C_, c0, c1 in ((144, 0, 144), (520, 0, 520), (64, 16, 48), (3, 0, 3)):
        b_ = seeded((2, C_, 9, 11), 7, -1, 1)
        cs = torch.ones(c1 - c0, device="cuda")
        ops.colsum(ops.View.from_nchw(cl(b_)).ch(c0, c1), cs, accumulate=True)
        assert maxabs(cs, 1 + b_[:, c0:c1].double().sum((0, 2, 3))) <= 2e-4
    yv, gv = torch.tanh(seeded((1000,), 4, -2, 2)), seeded((1000,), 5, -1, 1)
    od = torch.empty(1000, device="cuda")
    ops.act_bwd(gv.cuda(), yv.cuda(), od, ops.ACT_TANH)
    assert maxabs(od, gv * (1 - yv * yv)) <= 1e-6
    ops.act_bwd(gv.cuda(), yv.abs().cuda(), od, ops.ACT_SIGMOID)
    assert maxabs(od, gv * yv.abs() * (1 - yv.abs())) <= 1e-6


def test_dgrad_strided():
    ops = _ops()
    x = seeded((2, 9, 16, 18), 1, -1, 1).requires_grad_(True)
    w = seeded((36, 9, 4, 4), 2, -1, 1) / 12
    y = F.conv2d(x, w, stride=2, padding=1)
    g = seeded(tuple(y.shape), 3, -1, 1)
    (y * g).sum().backward()
    dx = torch.zeros(2, 9, 16, 18, device="cuda")
    ops.dgrad_strided(ops.View.from_nchw(cl(g)), w.cuda(), 2, 1, ops.View.from_nchw(dx))
    assert maxabs(dx, x.grad) <= 1e-5
    # four same-parity pixels per thread: ragged column groups, odd sizes, channels-last output
    for shape in ((1, 9, 15, 21), (3, 9, 34, 40), (2, 5, 9, 7)):
        x = seeded(shape, 4, -1, 1).requires_grad_(True)
        w = seeded((36, shape[1], 4, 4), 5, -1, 1) / 12
        y = F.conv2d(x, w, stride=2, padding=1)
        g = seeded(tuple(y.shape), 6, -1, 1)
        (y * g).sum().backward()
        dx = cl(torch.full(shape, 3.0))
        ops.dgrad_strided(ops.View.from_nchw(cl(g)), w.cuda(), 2, 1, ops.View.from_nchw(dx))
        assert maxabs(dx, x.grad) <= 1e-5, shape


@pytest.mark.parametrize("shape", [(2, 3, 40, 56), (1, 3, 8, 9), (1, 3, 33, 65)])
def test_freq_concat_fwd_bwd(shape):
    from fdgan_b200.loss import freq_concat
    from oracle import fdgan_oracle as O
    x = seeded(shape, 1).requires_grad_(True)
    z = O.freq_concat(x)
    g = seeded(tuple(z.shape), 2, -1, 1)
    (z * g).sum().backward()
    xd = x.detach().cuda().requires_grad_(True)
    zd = freq_concat(xd)
    assert tuple(zd.shape) == tuple(z.shape)
    assert maxabs(zd, z) <= 1e-5
    (zd * g.cuda()).sum().backward()
    assert maxabs(xd.grad, x.grad) <= 1e-4


@pytest.mark.parametrize("shape,layout", [((2, 3, 40, 56), "nchw"), ((1, 3, 33, 65), "cl"), ((1, 1, 7, 9), "nchw")])
def test_ssim_loss_grad(shape, layout):
    """fdg_ssim_loss_grad against autograd through the oracle's restatement of pytorch_ssim._ssim (which
    tests/test_oracle_golden.py pins to the reference module's own output): value of sum(ssim_map) and d/dx."""
    from oracle import fdgan_oracle as O
    ops = _ops()
    x = seeded(shape, 1, 0.0, 1.0).double().requires_grad_(True)
    y = seeded(shape, 2, 0.0, 1.0).double()
    n = x.numel()
    val = O.ssim(x, y)
    (0.7 * (1 - val)).backward()
    mk = (lambda t: cl(t)) if layout == "cl" else (lambda t: t.cuda().contiguous())
    xd, yd = mk(x.detach().float()), mk(y.float())
    base = seeded(shape, 3, -1, 1)
    gd = mk(base.clone())
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    ops.ssim_loss_grad(ops.View.from_nchw(xd), ops.View.from_nchw(yd), -0.7 / n, -0.7 / n, loss, ops.View.from_nchw(gd), accumulate=True)
    torch.cuda.synchronize()
    assert abs((0.7 + float(loss)) - float(0.7 * (1 - val))) <= 2e-6
    assert maxabs(gd, x.grad + base.double()) <= 1e-6 + 2e-3 * float(x.grad.abs().max())


def test_adam_flat():
    ops = _ops()
    p = seeded((1000,), 1, -1, 1)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=2e-4, betas=(0.5, 0.999))
    pd, m, v = p.cuda(), torch.zeros(1000, device="cuda"), torch.zeros(1000, device="cuda")
    for step in range(1, 4):
        g = seeded((1000,), 10 + step, -1, 1)
        ref.grad = g.clone()
        opt.step()
        ops.adam_flat(pd, (2 * g).cuda(), m, v, 2e-4, 0.5, 0.999, 1e-8, step, grad_scale=0.5)
    assert maxabs(pd, ref.detach()) <= 1e-6


THIN_CASES = [
    # Cin, Cout, R, stride, pad, H, W, x_nchw, g_nchw, affine, slope, dbias
    (3, 64, 3, 1, 1, 37, 70, True, False, False, 1.0, True),      # stem conv_refin1 / Vgg16 conv1_1: NCHW image, ragged 4x32 tiles
    (9, 36, 4, 2, 1, 22, 74, False, False, False, 1.0, False),    # Fusion-D layer 1: 4x4 stride 2, NHWC 9-channel input
    (16, 3, 3, 1, 1, 21, 45, False, True, False, 1.0, True),      # head conv_refin3: NCHW 3-channel gradient, bias gradient by planes
    (12, 16, 3, 1, 1, 9, 33, False, False, True, 0.2, False),     # BatchNorm + LeakyReLU prologue, zero padding after it
    (3, 64, 3, 1, 1, 128, 96, True, False, False, 1.0, True),     # many tiles per CTA (every slice group accumulates)
]


@pytest.mark.parametrize("case", THIN_CASES)
def test_thin_layers_auto_dispatch(case):
    """The direct kernels of the thin layers (conv_direct.cu: compile-time-filter forward, tiled weight gradient, planar bias
    gradient) are reached only with impl = auto; forward and weight gradient against fp64 torch."""
    ops = _ops()
    Cin, Cout, R, stride, pad, H, W, x_nchw, g_nchw, affine, slope, dbias = case
    N = 3
    x = seeded((N, Cin, H, W), 1, -1, 1).double()
    sc = seeded((Cin,), 4, 0.5, 1.5).double() if affine else None
    sh = seeded((Cin,), 5, -0.3, 0.3).double() if affine else None
    a = ref_prologue(x, sc, sh, slope)
    w = (seeded((Cout, Cin, R, R), 2, -1, 1) / math.sqrt(Cin * R * R)).double().requires_grad_(True)
    b = seeded((Cout,), 3, -0.5, 0.5).double()
    y = F.conv2d(a, w, b, stride=stride, padding=pad)
    g = seeded(tuple(y.shape), 8, -1, 1).double()
    (y * g).sum().backward()
    xd = x.float().cuda() if x_nchw else cl(x.float())
    xv = ops.View.from_nchw(xd)
    # forward (the prologue of the thin forward kernel is an activation only)
    if not affine and Cout in (16, 36, 64):
        wp, ld = ops.pack_weight(w.detach().float().cuda(), 0)
        yd = cl(torch.zeros(tuple(y.shape)))
        ops.conv2d(xv, wp, ld, R, R, stride, pad, Cout, ops.View.from_nchw(yd), slope=slope, bias=b.float().cuda())
        assert maxabs(yd, y.detach()) <= 2e-5
    gd = g.float().cuda() if g_nchw else cl(g.float())
    dw = torch.zeros(Cout, Cin, R, R, device="cuda")
    db = torch.zeros(Cout, device="cuda") if dbias else None
    ops.wgrad(xv, ops.View.from_nchw(gd), R, R, stride, pad, dw, scale=sc.float().cuda() if affine else None,
              shift=sh.float().cuda() if affine else None, slope=slope, dbias=db)
    scale = float(w.grad.abs().max())
    assert maxabs(dw, w.grad) <= 2e-5 * max(1.0, scale)
    if dbias:
        assert maxabs(db, g.sum((0, 2, 3))) <= 2e-5 * max(1.0, float(g.sum((0, 2, 3)).abs().max()))


def test_colsum_nchw_planes():
    ops = _ops()
    for shape in ((4, 3, 64, 64), (2, 5, 70, 91)):
        b_ = seeded(shape, 7, -1, 1)
        cs = torch.ones(shape[1], device="cuda")
        ops.colsum(ops.View.from_nchw(b_.cuda()), cs, accumulate=True)
        assert maxabs(cs, 1 + b_.double().sum((0, 2, 3))) <= 5e-4


@pytest.mark.parametrize("cin,cout,R,pad,H,W,slope", [
    (160, 128, 3, 1, 7, 32, 1.0),      # conv_refine4: 2.5 channel units per tap (padded rows), one output tile
    (512, 128, 3, 1, 5, 64, 0.0),      # dense_block5.conv2: ReLU prologue folded into the planes, two chunks per row
    (192, 256, 3, 1, 4, 32, 0.0),      # 256-wide output tile
    (640, 512, 3, 1, 3, 32, 1.0),      # conv_refin6: two 256-wide output tiles
    (72, 136, 4, 1, 6, 33, 0.2),       # 4x4 filter, OW = 32, channel tail inside a unit, Cout tail inside the 256-wide tile
])
def test_wgrad_both_operands_from_split_planes(cin, cout, R, pad, H, W, slope):
    """FdgWgrad.x_split: wide stride-1 RxS weight gradients fed by bulk tensor loads alone (filter taps = box coordinates, borders
    zero-filled by the tensor map) == fp64 torch; also through a channel slice of a wider gradient buffer."""
    ops = _ops()
    N = 2
    x = seeded((N, cin, H, W), 1, -1, 1)
    OH, OW = H + 2 * pad - R + 1, W + 2 * pad - R + 1
    assert OW % 32 == 0
    g = seeded((N, cout + 8, OH, OW), 2, -1, 1)
    xv = ops.View.from_nchw(cl(x))
    gv = ops.View.from_nchw(cl(g)).ch(8, cout + 8)
    assert ops.wgrad_planes_ok(xv, gv, R, R, 1, pad)
    xs, gs = ops.split_planes(xv, slope), ops.split_planes(gv, 1.0)
    dw = torch.zeros(cout, cin, R, R, device="cuda")
    db = torch.zeros(cout, device="cuda")
    ops.wgrad(xv, gv, R, R, 1, pad, dw, slope=slope, dbias=db, x_split=xs, g_split=gs)
    torch.cuda.synchronize()
    ref = torch.zeros(cout, cin, R, R, dtype=torch.float64, requires_grad=True)
    gd = g[:, 8:].double()
    (F.conv2d(F.leaky_relu(x.double(), slope), ref, padding=pad) * gd).sum().backward()
    assert maxabs(dw, ref.grad) <= 5e-5 * max(1.0, float(ref.grad.abs().max()))
    assert maxabs(db, gd.sum((0, 2, 3))) <= 1e-4 * max(1.0, float(gd.sum((0, 2, 3)).abs().max()))
