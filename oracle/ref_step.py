"""The reconstructed FD-GAN iteration composed from the REFERENCE's own modules -- FDGAN / D
(models/dehaze1113.py:702-801, :188-230), Vgg16 (myutils/vgg16.py:27-49) -- with torch.optim.Adam (demo.py:43-46 flags)
and torch's loss functions; only the frequency decomposition comes from the oracle (loss.py survives as bytecode).

TEST / BASELINE INFRASTRUCTURE ONLY.  Same composition as oracle/make_golden.py:gen_train_step (which pins
oracle.fdgan_oracle.train_step); here it is a reusable object so that bench.py can time the reference itself:
  * `bench.py --impl reference`   -> RefStep(device="cpu")  : the reference's CPU path on the host cores
  * `gpu_baseline` of the bench line -> RefStep(device="cuda"): the reference's torch / cuDNN path on the same B200
    (cudnn.benchmark = True as demo.py:11-12 sets it; TF32 off = the reference's fp32 arithmetic, or on).
The modules come from /root/reference or its byte-for-byte copy under baseline/_ref (oracle/vendor_ref.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import fdgan_oracle as O
from . import ref_import as R


def available() -> bool:
    return R.available()


class RefStep:
    def __init__(self, device="cpu", lr=2e-4, betas=(0.5, 0.999), weights=None, perc_layers=(1, 3)):
        self.dev = torch.device(device)
        self.netG = R.load_state(R.ref_fdgan(), O.make_fdgan_state(0)).train().to(self.dev)      # README.md:38: always train mode
        self.netD = R.load_state(R.ref_d(9, 36), O.make_d_state(9, 36, 1)).train().to(self.dev)
        self.vgg = R.load_state(R.ref_vgg16(), O.make_vgg_state(2)).to(self.dev)
        for p in self.vgg.parameters():
            p.requires_grad_(False)
        self.optG = torch.optim.Adam(self.netG.parameters(), lr=lr, betas=betas)
        self.optD = torch.optim.Adam(self.netD.parameters(), lr=lr, betas=betas)
        self.w = dict(O.DEFAULT_LOSS_WEIGHTS)
        if weights:
            self.w.update(weights)
        self.perc_layers = tuple(perc_layers)

    def step(self, hazy, clean):
        netG, netD, vgg, w = self.netG, self.netD, self.vgg, self.w
        fake = netG(hazy)
        self.optD.zero_grad()
        pr, pf = netD(O.freq_concat(clean)), netD(O.freq_concat(fake.detach()))
        loss_d = F.binary_cross_entropy(pr, torch.ones_like(pr)) + F.binary_cross_entropy(pf, torch.zeros_like(pf))
        loss_d.backward()
        self.optD.step()
        self.optG.zero_grad()
        for p in netD.parameters():
            p.requires_grad_(False)
        loss_g = w["l1"] * F.l1_loss(fake, clean)
        if w["perc"] != 0.0 and self.perc_layers:
            fv, cv = vgg(fake), vgg(clean)
            loss_g = loss_g + w["perc"] * sum(F.mse_loss(fv[k], cv[k].detach()) for k in self.perc_layers)
        pg = netD(O.freq_concat(fake))
        loss_g = loss_g + w["adv"] * F.binary_cross_entropy(pg, torch.ones_like(pg))
        loss_g.backward()
        for p in netD.parameters():
            p.requires_grad_(True)
        self.optG.step()
        return loss_d.detach(), loss_g.detach(), fake.detach()
