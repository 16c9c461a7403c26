"""Micro-benchmark (not a pytest file): times fdg_conv2d / fdg_conv2d_wgrad on the dominant shapes of the step at
B=16, 256x256.  python tests/bench_conv.py [simt|umma|both]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fdgan_b200 import ops
from fdgan_b200.ops import View

B = int(os.environ.get("BENCH_B", "16"))
SHAPES = [
    # name, Cin, Cout, R, pad, H, W, affine, stats
    ("K1 3x3 128->32 @256", 128, 32, 3, 1, 256, 256, True, True),
    ("K1 3x3 128->32 @128", 128, 32, 3, 1, 128, 128, True, True),
    ("K1 3x3 128->32 @64", 128, 32, 3, 1, 64, 64, True, True),
    ("K2 1x1 64->128 @256", 64, 128, 1, 0, 256, 256, True, True),
    ("K2 1x1 224->128 @256", 224, 128, 1, 0, 256, 256, True, True),
    ("K2 1x1 480->128 @128", 480, 128, 1, 0, 128, 128, True, True),
    ("K2 1x1 992->128 @64", 992, 128, 1, 0, 64, 64, True, True),
    ("dgrad 3x3 32->128 @256", 32, 128, 3, 1, 256, 256, False, False),
    ("dgrad 1x1 128->224 @256", 128, 224, 1, 0, 256, 256, False, False),
    ("refine4 3x3 160->128 @128", 160, 128, 3, 1, 128, 128, False, True),
    ("D L4 4x4 144->288 @128", 144, 288, 4, 1, 128, 128, True, False),
    ("vgg 3x3 64->64 @256", 64, 64, 3, 1, 256, 256, False, False),
    ("vgg 3x3 256->256 @64", 256, 256, 3, 1, 64, 64, False, False),
    ("vgg 3x3 512->512 @32", 512, 512, 3, 1, 32, 32, False, False),
]


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "both"
    only = sys.argv[2] if len(sys.argv) > 2 else ""
    dev = "cuda"
    for name, Cin, Cout, R, pad, H, W, affine, stats in SHAPES:
        if only and only not in name:
            continue
        x = View.alloc(B, H, W, Cin, dev); x.base.normal_()
        OH = H + 2 * pad - R + 1
        y = View.alloc(B, OH, OH, Cout, dev)
        w = torch.randn(Cout, Cin, R, R, device=dev) / (Cin * R * R) ** 0.5
        wp, ld = ops.pack_weight(w, 0)
        sc = torch.rand(Cin, device=dev) + 0.5 if affine else None
        sh = torch.rand(Cin, device=dev) - 0.5 if affine else None
        st = torch.zeros(2 * Cout, dtype=torch.float64, device=dev) if stats else None
        flops = 2.0 * B * OH * OH * Cout * Cin * R * R
        out = "%-28s" % name
        for impl, tag in ((ops.IMPL_SIMT, "simt"), (ops.IMPL_UMMA, "umma")):
            if which not in (tag, "both"):
                continue
            wu = wk = None
            if impl == ops.IMPL_UMMA:
                if ops.k1_eligible(x, Cout, R, R, 1, pad, 0): wk = ops.pack_weight_k1(wp, ld, Cin, Cout, dev)
                else: wu = ops.pack_weight_umma(wp, ld, R * R, Cin, Cout, dev)
            ms = timeit(lambda: ops.conv2d(x, wp, ld, R, R, 1, pad, Cout, y, scale=sc, shift=sh, slope=0.0 if affine else 1.0,
                                           stats=st, stats_ld=Cout, impl=impl, w_umma=wu, w_k1=wk))
            out += "  %s %8.3f ms %7.1f TF/s" % (tag, ms, flops / ms / 1e9)
        # weight gradient (SIMT today)
        if which in ("both", "wgrad") or os.environ.get("BENCH_WGRAD"):
            g = View.alloc(B, OH, OH, Cout, dev); g.base.normal_()
            dw = torch.zeros_like(w)
            gs = ops.split_planes(g, 1.0) if (os.environ.get("BENCH_GSPLIT") and Cout % 8 == 0) else None    # FAST 1x1 loader needs planes
            ms = timeit(lambda: ops.wgrad(x, g, R, R, 1, pad, dw, scale=sc, shift=sh, slope=0.0 if affine else 1.0, g_split=gs), iters=5)
            out += "  wgrad %8.3f ms %7.1f TF/s" % (ms, flops / ms / 1e9)
        print(out, flush=True)


if __name__ == "__main__":
    main()
