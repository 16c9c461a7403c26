"""Diagnostic (not a pytest file): forward / gradient parity of FDGAN, D, Vgg16 vs the CPU oracle for the fp32 SIMT
path and the tcgen05 bf16x3 path.  Run on the GPU box: python tests/diag_parity.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fdgan_b200
from fdgan_b200 import ops
from oracle import fdgan_oracle as O
from tests.util import seeded


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    d = a - b
    return float(d.norm() / max(float(b.norm()), 1e-30)), float(d.abs().max() / max(float(b.abs().max()), 1e-30)), float(d.abs().max())


def fdgan_case(shape, dt=torch.float32):
    sd = O.make_fdgan_state(0, dt)
    for k in O.fdgan_used_param_names():
        sd[k].requires_grad_(True)
    x = seeded(shape, 5).to(dt).requires_grad_(True)
    r = seeded(shape, 6, -1, 1).to(dt)
    y = O.fdgan_forward(sd, x, True, False)
    (y * r).sum().backward()
    return sd, x, r, y


for shape in ((1, 3, 32, 32), (2, 3, 32, 32), (2, 3, 64, 48), (1, 3, 128, 128)):
    sd, xo, r, yo = fdgan_case(shape)
    sd64, xo64, _, yo64 = fdgan_case(shape, torch.float64)
    print("FDGAN %s: oracle fp32 vs fp64: y %.2e dx relL2 %.2e" % (shape, float((yo.double() - yo64).abs().max()), rel(xo.grad, xo64.grad)[0]), flush=True)
    for use in (False, True):
        ops.USE_UMMA = use
        net = fdgan_b200.FDGAN()
        net.load_state_dict(O.make_fdgan_state(0))
        net = net.cuda().train()
        xd = xo.detach().float().cuda().requires_grad_(True)
        y = net(xd)
        (y * r.cuda()).sum().backward()
        ey = float((y.detach().cpu().double() - yo64).abs().max())
        l2s = []
        for k, p in net.named_parameters():
            if sd64[k].grad is None:
                continue
            if float(sd64[k].grad.abs().max()) < 1e-6:
                continue
            l2s.append((rel(p.grad, sd64[k].grad)[0], k))
        l2s.sort()
        print("   %-8s y max-abs vs fp64 %.2e | dx relL2 %.2e relmax %.2e | param-grad relL2 median %.2e p90 %.2e max %.2e (%s) n>3e-2: %d" % (
            "tcgen05" if use else "simt", ey, *rel(xd.grad, xo64.grad)[:2], l2s[len(l2s) // 2][0], l2s[int(len(l2s) * .9)][0], l2s[-1][0], l2s[-1][1],
            sum(1 for v, _ in l2s if v > 3e-2)), flush=True)

# VGG + D forward/backward
for use in (False, True):
    ops.USE_UMMA = use
    v = fdgan_b200.Vgg16(); v.load_state_dict(O.make_vgg_state(2)); v = v.cuda()
    for p in v.parameters(): p.requires_grad_(False)
    x = seeded((2, 3, 64, 64), 9)
    vsd = {k: t.double() for k, t in O.make_vgg_state(2).items()}
    xo = x.double().requires_grad_(True)
    fo = O.vgg16_forward(vsd, xo)
    sum((f ** 2).mean() for f in fo).backward()
    xd = x.cuda().requires_grad_(True)
    fd = v(xd)
    sum((f ** 2).mean() for f in fd).backward()
    print("VGG %-8s feats rel-max %s | dx relL2 %.2e relmax %.2e" % ("tcgen05" if use else "simt", ["%.1e" % rel(a, b)[1] for a, b in zip(fd, fo)], *rel(xd.grad, xo.grad)[:2]), flush=True)
    d = fdgan_b200.D(9, 36); d.load_state_dict(O.make_d_state(9, 36, 1)); d = d.cuda().train()
    dsd = {k: (t.double() if t.is_floating_point() else t) for k, t in O.make_d_state(9, 36, 1).items()}
    for k in O.d_param_names(dsd): dsd[k].requires_grad_(True)
    z = seeded((2, 9, 64, 64), 7, -1, 1)
    zo = z.double().requires_grad_(True)
    po = O.d_forward(dsd, zo, True, False)
    (po * po).sum().backward()
    zd = z.cuda().requires_grad_(True)
    pd = d(zd)
    (pd * pd).sum().backward()
    gl = [(rel(p.grad, dsd[k].grad)[0], k) for k, p in d.named_parameters()]
    print("D   %-8s out max-abs %.2e | dz relL2 %.2e | param-grad relL2 max %.2e (%s)" % ("tcgen05" if use else "simt", rel(pd, po)[2], rel(zd.grad, zo.grad)[0], max(gl)[0], max(gl)[1]), flush=True)
