"""Micro-timing of the thin-layer kernels (not a pytest file).  FDG_THIN=0 selects the generic SIMT kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fdgan_b200 import ops
from fdgan_b200.ops import View

B = 16
dev = "cuda"


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


x = torch.rand(B, 3, 256, 256, device=dev)
xin = View.from_nchw(x)
w = torch.randn(64, 3, 3, 3, device=dev)
wp, ld = ops.pack_weight(w, 0)
y = View.alloc(B, 256, 256, 256, dev)
bias = torch.randn(64, device=dev)
st = torch.zeros(512, dtype=torch.float64, device=dev)
print("stem fwd 3->64      %.3f ms" % timeit(lambda: ops.conv2d(xin, wp, ld, 3, 3, 1, 1, 64, y.ch(0, 64), bias=bias, act=ops.ACT_RELU, stats=st, stats_ld=256)))
g = View.alloc(B, 256, 256, 64, dev); g.base.normal_()
dw = torch.zeros_like(w); db = torch.zeros(64, device=dev)
print("stem wgrad          %.3f ms" % timeit(lambda: ops.wgrad(xin, g, 3, 3, 1, 1, dw, dbias=db)))
z = View.alloc(B, 256, 256, 9, dev); z.base.normal_()
w1 = torch.randn(36, 9, 4, 4, device=dev)
w1p, ld1 = ops.pack_weight(w1, 0)
y1 = View.alloc(B, 128, 128, 36, dev)
print("D L1 fwd 9->36 s2   %.3f ms" % timeit(lambda: ops.conv2d(z, w1p, ld1, 4, 4, 2, 1, 36, y1)))
g1 = View.alloc(B, 128, 128, 36, dev); g1.base.normal_()
dw1 = torch.zeros_like(w1)
print("D L1 wgrad          %.3f ms" % timeit(lambda: ops.wgrad(z, g1, 4, 4, 2, 1, dw1)))
dz = View.alloc(B, 256, 256, 9, dev)
print("D L1 dgrad strided  %.3f ms" % timeit(lambda: ops.dgrad_strided(g1, w1, 2, 1, dz)))
g5 = torch.randn(B, 1, 126, 126, device=dev)
w5 = torch.randn(1, 288, 4, 4, device=dev)
w5p, ld5 = ops.pack_weight(w5, 1)
y4 = View.alloc(B, 127, 127, 288, dev); y4.base.normal_()
d4 = View.alloc(B, 127, 127, 288, dev)
print("D L5 dgrad 1->288   %.3f ms" % timeit(lambda: ops.conv2d(View.from_nchw(g5), w5p, ld5, 4, 4, 1, 2, 288, d4, e=y4, eslope=0.2)))
dw5 = torch.zeros_like(w5)
print("D L5 wgrad 288->1   %.3f ms" % timeit(lambda: ops.wgrad(y4, View.from_nchw(g5), 4, 4, 1, 1, dw5, slope=0.2)))
y5 = torch.empty(B, 1, 126, 126, device=dev)
w5f, ld5f = ops.pack_weight(w5, 0)
print("D L5 fwd 288->1     %.3f ms" % timeit(lambda: ops.conv2d(y4, w5f, ld5f, 4, 4, 1, 1, 1, View.from_nchw(y5), slope=0.2, act=ops.ACT_SIGMOID)))

x6 = View.alloc(B, 256, 256, 16, dev); x6.base.normal_()
gh = torch.randn(B, 3, 256, 256, device=dev)
dwh = torch.zeros(3, 16, 3, 3, device=dev); dbh = torch.zeros(3, device=dev)
print("head wgrad 16->3    %.3f ms" % timeit(lambda: ops.wgrad(x6, View.from_nchw(gh), 3, 3, 1, 1, dwh, dbias=dbh)))
print("head colsum only    %.3f ms" % timeit(lambda: ops.colsum(View.from_nchw(gh), dbh, accumulate=True)))
