"""Ablation timing of the tcgen05 kernels (not a pytest file): which side bounds each kernel?

    python tests/ablate.py [filter]

dbg bits (fdg_set_option("dbg", v)): 1 = loaders skip the global loads, 2 = loaders skip split + shared stores,
4 = the MMA thread skips tcgen05.mma (commits only), 8 = the epilogue skips TMEM loads + global stores.
Results are garbage in every mode but 0; only the times matter.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fdgan_b200 import ops, _lib
from fdgan_b200.ops import View

B = int(os.environ.get("BENCH_B", "16"))
SHAPES = [
    # name, kind, Cin, Cout, R, pad, H, affine, stats
    ("K1 3x3 128->32 @256", "conv", 128, 32, 3, 1, 256, True, True),
    ("K1 3x3 128->32 @64", "conv", 128, 32, 3, 1, 64, True, True),
    ("K2 1x1 64->128 @256", "conv", 64, 128, 1, 0, 256, True, True),
    ("K2 1x1 224->128 @256", "conv", 224, 128, 1, 0, 256, True, True),
    ("K2 1x1 992->128 @64", "conv", 992, 128, 1, 0, 64, True, True),
    ("dgrad 3x3 32->128 @256", "conv", 32, 128, 3, 1, 256, False, False),
    ("dgrad 1x1 128->224 @256", "conv", 128, 224, 1, 0, 256, False, False),
    ("vgg 3x3 64->64 @256", "conv", 64, 64, 3, 1, 256, False, False),
    ("D L4 4x4 144->288 @128", "conv", 144, 288, 4, 1, 128, True, False),
    ("D L4 dgrad 4x4 288->144 @127", "conv", 288, 144, 4, 2, 127, False, False),
    ("D L3 3x3 72->144 @128", "conv", 72, 144, 3, 1, 128, True, False),
    ("vgg 3x3 128->128 @128", "conv", 128, 128, 3, 1, 128, False, False),
    ("vgg 3x3 256->256 @64", "conv", 256, 256, 3, 1, 64, False, False),
    ("vgg 3x3 512->512 @32", "conv", 512, 512, 3, 1, 32, False, False),
    ("wgrad K1 3x3 128->32 @256", "wgrad", 128, 32, 3, 1, 256, True, False),
    ("wgrad K2 1x1 224->128 @256", "wgrad", 224, 128, 1, 0, 256, True, False),
    ("wgrad K2 1x1 992->128 @64", "wgrad", 992, 128, 1, 0, 64, True, False),
    ("wgrad vgg 3x3 64->64 @256", "wgrad", 64, 64, 3, 1, 256, False, False),
    ("wgrad W D L4 4x4 144->288 @128", "wgrad", 144, 288, 4, 1, 128, True, False),
    ("wgrad W D L3 3x3 72->144 @128", "wgrad", 72, 144, 3, 1, 128, True, False),
    ("wgrad W refine4 3x3 160->128 @128", "wgrad", 160, 128, 3, 1, 128, False, False),
    ("wgrad W bdy5 3x3 512->128 @64", "wgrad", 512, 128, 3, 1, 64, False, False),
    ("wgrad W bdy4 3x3 1024->256 @32", "wgrad", 1024, 256, 3, 1, 32, False, False),
    ("wgrad W refin6 3x3 640->512 @32", "wgrad", 640, 512, 3, 1, 32, False, False),
]
MODES = [int(m) for m in os.environ.get("ABL_MODES", "0,1,3,4,8,12,7,15").split(",")]


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    dev = "cuda"
    print("%-34s" % "shape" + "".join("  dbg=%-2d  " % m for m in MODES) + "   (ms; TF/s at dbg=0)")
    for name, kind, Cin, Cout, R, pad, H, affine, stats in SHAPES:
        if only and not any(f in name for f in only.split(",")):
            continue
        x = View.alloc(B, H, H, Cin, dev); x.base.normal_()
        OH = H + 2 * pad - R + 1
        y = View.alloc(B, OH, OH, Cout, dev)
        w = torch.randn(Cout, Cin, R, R, device=dev) / (Cin * R * R) ** 0.5
        wp, ld = ops.pack_weight(w, 0)
        sc = torch.rand(Cin, device=dev) + 0.5 if affine else None
        sh = torch.rand(Cin, device=dev) - 0.5 if affine else None
        st = torch.zeros(2 * Cout, dtype=torch.float64, device=dev) if stats else None
        flops = 2.0 * B * OH * OH * Cout * Cin * R * R
        out = "%-34s" % name
        t0 = None
        for m in MODES:
            _lib.set_option("dbg", m)
            if kind == "conv":
                wu = wk = None
                if ops.k1_eligible(x, Cout, R, R, 1, pad, 0): wk = ops.pack_weight_k1(wp, ld, Cin, Cout, dev)
                else: wu = ops.pack_weight_umma(wp, ld, R * R, Cin, Cout, dev)
                ms = timeit(lambda: ops.conv2d(x, wp, ld, R, R, 1, pad, Cout, y, scale=sc, shift=sh, slope=0.0 if affine else 1.0,
                                               stats=st, stats_ld=Cout, impl=ops.IMPL_UMMA, w_umma=wu, w_k1=wk))
            else:
                y.base.normal_()
                dw = torch.zeros_like(w)
                if os.environ.get("ABL_PLANES") and ops.wgrad_planes_ok(x, y, R, R, 1, pad) and not affine:
                    def run():
                        xs, gs = ops.split_planes(x, 1.0), ops.split_planes(y, 1.0)
                        ops.wgrad(x, y, R, R, 1, pad, dw, x_split=xs, g_split=gs)
                    ms = timeit(run, iters=3)
                else:
                    ms = timeit(lambda: ops.wgrad(x, y, R, R, 1, pad, dw, scale=sc, shift=sh, slope=0.0 if affine else 1.0), iters=3)
            if m == 0:
                t0 = ms
            out += " %8.3f " % ms
        _lib.set_option("dbg", 0)
        print(out + "  %7.1f TF/s" % (flops / t0 / 1e9), flush=True)


if __name__ == "__main__":
    main()
