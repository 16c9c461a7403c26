"""Diagnostic (not a pytest file): exercises the tcgen05 conv path on structured inputs and prints what it finds.
Run on the GPU box:  timeout 120 python tests/diag_umma.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from fdgan_b200 import ops
from fdgan_b200.ops import View


def run(name, Cin, Cout, R, pad, H, W, N=1, ident=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((N, Cin, H, W), generator=g) * 2 - 1
    if ident:
        w = torch.zeros(Cout, Cin, R, R)
        for co in range(Cout):
            w[co, co % Cin, R // 2, R // 2] = 1.0
    else:
        w = (torch.rand((Cout, Cin, R, R), generator=g) * 2 - 1) / (Cin * R * R) ** 0.5
    want = F.conv2d(x.double(), w.double(), padding=pad)
    xd = x.cuda().contiguous(memory_format=torch.channels_last)
    wp, ld = ops.pack_weight(w.cuda(), 0)
    OH, OW = want.shape[-2:]
    y = torch.zeros(N, Cout, OH, OW, device="cuda").contiguous(memory_format=torch.channels_last)
    ops.conv2d(View.from_nchw(xd), wp, ld, R, R, 1, pad, Cout, View.from_nchw(y), impl=ops.IMPL_UMMA)
    torch.cuda.synchronize()
    err = (y.cpu().double() - want).abs()
    print("%-28s max err %.3e (max |want| %.3e)  mean err %.3e" % (name, err.max().item(), want.abs().max().item(), err.mean().item()), flush=True)
    if err.max() > 1e-3 and ident:
        yy = y.cpu()[0].permute(1, 2, 0).reshape(-1, Cout)      # [pixel][co]
        xx = x[0].permute(1, 2, 0).reshape(-1, Cin)
        for pix in (0, 1, 8, 9, 33):
            row = yy[pix]
            hits = []
            for co in (0, 1, 2, 8, 9, 17):
                d = (xx - row[co]).abs()
                idx = int(d.argmin())
                hits.append((co, idx // Cin, idx % Cin, float(d.min())))
            print("   pixel", pix, "-> (co, src pixel, src ci, |diff|):", hits, flush=True)
    return err.max().item()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    run("1x1 ident 64->32 128px", 64, 32, 1, 0, 8, 16, ident=True)
    run("1x1 rand 64->32 128px", 64, 32, 1, 0, 8, 16)
    run("1x1 rand 128->32 (2 chunks)", 128, 32, 1, 0, 8, 16)
    run("1x1 rand 256->128 (4 chunks)", 256, 128, 1, 0, 16, 16)
    run("1x1 rand 512->64 (8 chunks)", 512, 64, 1, 0, 10, 13, N=2)
    run("1x1 rand 96->256 (pad chunk)", 96, 256, 1, 0, 9, 9)
    run("3x3 ident 64->64", 64, 64, 3, 1, 8, 16, ident=True)
    run("3x3 rand 128->32", 128, 32, 3, 1, 12, 20, N=2)
    run("3x3 rand 160->128", 160, 128, 3, 1, 6, 6)
    run("4x4 rand 144->288 p1", 144, 288, 4, 1, 9, 8, N=2)
    print("done", flush=True)
