"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules here.

TEST INFRASTRUCTURE ONLY; runs in the authoring container (needs /root/reference).
    python -m oracle.make_golden
The reference has no golden vectors of its own, so these outputs of the
reference itself -- FDGAN (models/dehaze1113.py:702-801), D (:188-230), Vgg16
(myutils/vgg16.py) and pytorch_ssim -- are what pins oracle/fdgan_oracle.py.
Weights come from oracle.fdgan_oracle.make_state (seeded per entry name), inputs from
seeded torch generators; both are regenerated identically inside the tests.
"""
from __future__ import annotations

import os
import warnings

import numpy as np
import torch

from . import fdgan_oracle as O
from . import ref_import as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

G_GRAD_KEYS = [
    "conv_refin1.weight", "conv_refin1.bias", "conv_refin2.weight", "conv_refine4.bias",
    "dense_block1.denselayer1.norm1.weight", "dense_block1.denselayer1.conv1.weight",
    "dense_block1.denselayer6.conv2.weight", "trans_block1.norm.bias", "trans_block1.conv.weight",
    "dense_block2.denselayer12.norm2.weight", "trans_block2.conv.weight",
    "dense_block3.denselayer24.norm1.bias", "dense_block3.denselayer1.conv2.weight", "trans_block3.norm.weight",
    "conv_refin5.weight", "conv_refin6.bias", "dense_block4.conv1.weight", "dense_block4.conv2.weight",
    "trans_block4.conv1.weight", "dense_block5.conv2.weight", "trans_block5.conv1.weight",
    "dense_block6.conv1.weight", "trans_block6.conv1.weight", "conv_refin3.weight", "conv_refin3.bias",
]
G_STAT_KEYS = [
    "dense_block1.denselayer1.norm1.running_mean", "dense_block1.denselayer1.norm1.running_var",
    "dense_block2.denselayer7.norm2.running_var", "trans_block3.norm.running_mean",
    "dense_block3.denselayer24.norm1.running_var", "dense_block3.denselayer24.norm1.num_batches_tracked",
]


MAX_SAMPLES = 2048


def sample(t):
    """Small fixture of a big tensor: [sum, l2 norm] followed by a strided subsample.
    tests/ recompute the same three things from the tensor under test."""
    a = np.asarray(t.detach().numpy() if hasattr(t, "detach") else t, dtype=np.float64).reshape(-1)
    stride = max(1, a.size // MAX_SAMPLES)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum())], a[::stride]]).astype(np.float64)


def seeded(shape, seed, lo=0.0, hi=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g) * (hi - lo) + lo


def gen_fdgan(batch, hw, tag):
    net = R.load_state(R.ref_fdgan(), O.make_fdgan_state(0))
    net.train()  # README.md:38
    x = seeded((batch, 3, hw, hw), 5).requires_grad_(True)
    r = seeded((batch, 3, hw, hw), 6, -1.0, 1.0)
    y = net(x)
    (y * r).sum().backward()
    params = dict(net.named_parameters())
    out = {"y": y.detach().numpy(), "dx": x.grad.numpy()}
    for k in G_GRAD_KEYS:
        out["grad:" + k] = sample(params[k].grad)
    sd = net.state_dict()
    for k in G_STAT_KEYS:
        out["stat:" + k] = sd[k].numpy()
    unused = [k for k, p in params.items() if p.grad is None]
    out["n_unused"] = np.array(len(unused))
    np.savez_compressed(os.path.join(OUT, "fdgan_%s.npz" % tag), **out)
    print("fdgan", tag, "y range", float(y.min()), float(y.max()), "unused", len(unused))


def gen_d():
    for nf in (36, 64):
        net = R.load_state(R.ref_d(9, nf), O.make_d_state(9, nf, 1))
        net.train()
        x = seeded((2, 9, 32, 32), 7, -1.0, 1.0).requires_grad_(True)
        y = net(x)
        r = seeded(tuple(y.shape), 8, -1.0, 1.0)
        (y * r).sum().backward()
        out = {"y": y.detach().numpy(), "dx": x.grad.numpy()}
        for k, p in net.named_parameters():
            out["grad:" + k] = sample(p.grad)
        for k, v in net.state_dict().items():
            if "running" in k:
                out["stat:" + k] = v.numpy()
        np.savez_compressed(os.path.join(OUT, "d_nf%d.npz" % nf), **out)
        print("D nf", nf, tuple(y.shape))


def gen_vgg():
    net = R.load_state(R.ref_vgg16(), O.make_vgg_state(2))
    x = seeded((2, 3, 16, 16), 9).requires_grad_(True)
    feats = net(x)
    loss = sum((f * seeded(tuple(f.shape), 10 + i, -1.0, 1.0)).sum() for i, f in enumerate(feats))
    loss.backward()
    out = {"dx": x.grad.numpy()}
    for i, f in enumerate(feats):
        out["f%d" % i] = sample(f)
    np.savez_compressed(os.path.join(OUT, "vgg16.npz"), **out)
    print("vgg", [tuple(f.shape) for f in feats])


def gen_ssim():
    ssim = R.ref_ssim()
    a = seeded((2, 3, 24, 24), 20).requires_grad_(True)
    b = seeded((2, 3, 24, 24), 21)
    v = ssim(a, b)
    v.backward()
    np.savez_compressed(os.path.join(OUT, "ssim.npz"), v=v.detach().numpy(), da=a.grad.numpy())
    print("ssim", float(v))


TRAIN_G_KEYS = ["conv_refin1.weight", "dense_block1.denselayer1.norm1.weight", "dense_block2.denselayer12.conv2.weight",
                "trans_block3.conv.weight", "dense_block4.conv2.weight", "trans_block6.conv1.weight", "conv_refin3.bias"]
TRAIN_D_KEYS = ["main.layer1.conv.weight", "main.layer2.layer2.bn.weight", "main.layer3.layer3.conv.weight", "main.layer5.conv.weight"]


def train_inputs(batch=2, hw=32):
    return seeded((batch, 3, hw, hw), 30), seeded((batch, 3, hw, hw), 31)


def gen_train_step():
    """The reconstructed FD-GAN iteration composed from the REFERENCE's own modules (FDGAN, D, Vgg16, pytorch_ssim),
    torch.optim.Adam (demo.py:43-46 flags) and torch's loss functions; only the frequency decomposition comes from the oracle
    (loss.py survives as bytecode).  Two consecutive steps, with and without the SSIM term: pins oracle.train_step's composition
    (D update first, detach, updated-D-as-constant in the G update, three BatchNorm running-stat updates of D per step) and its
    Adam arithmetic."""
    import torch.nn.functional as F
    for tag, w_ssim in (("plain", 0.0), ("ssim", 0.1)):
        netG = R.load_state(R.ref_fdgan(), O.make_fdgan_state(0)).train()
        netD = R.load_state(R.ref_d(9, 36), O.make_d_state(9, 36, 1)).train()
        vgg = R.load_state(R.ref_vgg16(), O.make_vgg_state(2))
        ssim = R.ref_ssim()
        for p in vgg.parameters():
            p.requires_grad_(False)
        optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.5, 0.999))
        optD = torch.optim.Adam(netD.parameters(), lr=2e-4, betas=(0.5, 0.999))
        hazy, clean = train_inputs()
        out = {}
        for it in range(2):
            fake = netG(hazy)
            optD.zero_grad()
            pr, pf = netD(O.freq_concat(clean)), netD(O.freq_concat(fake.detach()))
            loss_d = F.binary_cross_entropy(pr, torch.ones_like(pr)) + F.binary_cross_entropy(pf, torch.zeros_like(pf))
            loss_d.backward()
            gd = {k: p.grad.clone() for k, p in netD.named_parameters()}
            optD.step()
            optG.zero_grad()
            for p in netD.parameters():
                p.requires_grad_(False)
            l1 = F.l1_loss(fake, clean)
            fv, cv = vgg(fake), vgg(clean)
            perc = sum(F.mse_loss(fv[k], cv[k].detach()) for k in (1, 3))
            pg = netD(O.freq_concat(fake))
            adv = F.binary_cross_entropy(pg, torch.ones_like(pg))
            loss_g = 1.0 * l1 + 0.5 * perc + 0.01 * adv
            if w_ssim:
                loss_g = loss_g + w_ssim * (1 - ssim(fake, clean))
            loss_g.backward()
            for p in netD.parameters():
                p.requires_grad_(True)
            gg = {k: p.grad.clone() for k, p in netG.named_parameters() if p.grad is not None}
            optG.step()
            out["it%d:losses" % it] = np.array([float(loss_d), float(l1), float(0.5 * perc), float(0.01 * adv), float(loss_g)])
            for k in TRAIN_G_KEYS:
                out["it%d:gradG:%s" % (it, k)] = sample(gg[k])
                out["it%d:G:%s" % (it, k)] = sample(dict(netG.named_parameters())[k])
            for k in TRAIN_D_KEYS:
                out["it%d:gradD:%s" % (it, k)] = sample(gd[k])
                out["it%d:D:%s" % (it, k)] = sample(dict(netD.named_parameters())[k])
            out["it%d:fake" % it] = sample(fake)
        sdD = netD.state_dict()
        out["D:running_var"] = sdD["main.layer2.layer2.bn.running_var"].numpy()
        out["D:nbt"] = sdD["main.layer2.layer2.bn.num_batches_tracked"].numpy()
        out["G:running_mean"] = netG.state_dict()["dense_block1.denselayer1.norm1.running_mean"].numpy()
        np.savez_compressed(os.path.join(OUT, "train_step_%s.npz" % tag), **out)
        print("train step", tag, out["it0:losses"], out["it1:losses"])


def main():
    assert R.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    warnings.simplefilter("ignore")
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    gen_fdgan(1, 32, "b1_32")
    gen_fdgan(2, 32, "b2_32")
    gen_d()
    gen_vgg()
    gen_ssim()
    gen_train_step()


if __name__ == "__main__":
    main()
