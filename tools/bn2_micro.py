"""Micro-benchmark of the dense-layer conv2 data gradient (3x3, 32 -> 128): plain, and with the fused norm2 backward epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fdgan_b200 import ops
from fdgan_b200.ops import View

def timeit(fn, iters=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def main():
    dev = "cuda"
    B = 16
    for H in (256, 128, 64):
        g = View.alloc(B, H, H, 32, dev); g.base.normal_()
        T = View.alloc(B, H, H, 128, dev); T.base.normal_()
        y = View.alloc(B, H, H, 128, dev)
        w = torch.randn(32, 128, 3, 3, device=dev) / 34.0
        wd, ld = ops.pack_weight(w, 1)
        wu = ops.pack_weight_umma(wd, ld, 9, 32, 128, dev)
        sc = torch.rand(128, device=dev) + 0.5; sh = torch.rand(128, device=dev) - 0.5
        st = torch.zeros(256, dtype=torch.float64, device=dev)
        plain = timeit(lambda: ops.conv2d(g, wd, ld, 3, 3, 1, 1, 128, y, w_umma=wu))
        fused = timeit(lambda: ops.conv2d(g, wd, ld, 3, 3, 1, 1, 128, y, w_umma=wu, e=T, eslope=0.0, e_scale=sc, e_shift=sh, stats=st, stats_ld=128))
        red = timeit(lambda: ops.ew_bwd(y, T, stats=st, scale=sc, shift=sh, slope=0.0))
        print("H=%d  plain dgrad %.3f ms   fused BN2 %.3f ms   reduce pass %.3f ms   (FDG_HALO_CLUSTER=%s)" % (H, plain, fused, red, os.environ.get("FDG_HALO_CLUSTER", "1")), flush=True)

main()
