"""The FD-GAN training step on fdgan_b200 kernels (the reference's train.py is not in its tree; this follows
the reconstruction in SURVEY 3.3 from demo.py:43-46 flags, misc.py helpers, loss.pyc and facades/network.png).

    fake = G(hazy)
    D step:  lossD = BCE(D([clean,LF,HF]), 1) + BCE(D([fake.detach(),LF,HF]), 0);  Adam(D)
    G step:  lossG = w_l1 L1(fake, clean) + w_perc sum_k MSE(vgg_k(fake), vgg_k(clean)) + w_adv BCE(D([fake,LF,HF]), 1);  Adam(G)

The step drives the network executors directly (no autograd graph): loss values and their gradients come from
one fused kernel each, parameter gradients land in ONE flat fp32 buffer per network, which is all-reduced once
(NCCL over NVLink when world_size > 1) and consumed by a fused flat Adam.
"""
from __future__ import annotations

import torch

from . import dist as fdist
from . import engine, ops
from .ops import View

DEFAULT_WEIGHTS = dict(l1=1.0, perc=0.5, adv=0.01)


def dataparallel_state_dict(module) -> dict:
    """State dict with the 'module.' prefix nn.DataParallel checkpoints carry: what the reference's demo.py:78-86 expects (it strips
    the first 7 characters of EVERY key unconditionally, so an un-prefixed checkpoint would be mangled there)."""
    return {"module." + k: v.detach().clone() for k, v in module.state_dict().items()}


class FlatState:
    """Used parameters of one network re-homed into a single flat fp32 buffer (16-byte aligned slices), with a
    flat gradient buffer of the same layout (views by parameter name) and flat Adam moments."""

    def __init__(self, module):
        named = module._used_named_parameters()
        dev = named[0][1].device
        offs, off = [], 0
        for _n, p in named:
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.n = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad_views = {}
        self.slices = {}                    # name -> (offset, numel, shape) in the flat buffers
        for (n, p), o in zip(named, offs):
            v = self.flat[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            self.grad_views[n] = self.grad[o:o + p.numel()].view(p.shape)
            self.slices[n] = (o, p.numel(), tuple(p.shape))
        self.step = 0
        self.dev_state = None
        self.module = module

    def zero_grad(self):
        self.grad.zero_()

    def adam(self, lr, beta1, beta2, eps, grad_scale):
        self.step += 1
        if self.dev_state is not None:      # CUDA-graph mode: the step counter lives on the device
            ops.adam_flat_dev(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, lr, beta1, beta2, eps, self.dev_state, grad_scale)
        else:
            ops.adam_flat(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, lr, beta1, beta2, eps, self.step, grad_scale)
        if ops._FROZEN:     # the kernel writes through raw pointers (no _version bump): drop cached images of these parameters
            ops.drop_frozen_in_range(self.flat.data_ptr(), self.flat.data_ptr() + 4 * self.n)

    def state_dict(self):
        """Adam state by parameter name (the layout of the flat buffers is an implementation detail): the moments and the
        step count, as torch.optim.Adam keeps them per parameter."""
        def by_name(buf):
            return {n: buf[o:o + k].view(shape).clone() for n, (o, k, shape) in self.slices.items()}
        return {"step": int(self.step), "exp_avg": by_name(self.exp_avg), "exp_avg_sq": by_name(self.exp_avg_sq)}

    def load_state_dict(self, sd):
        for key in ("exp_avg", "exp_avg_sq"):
            if set(sd[key]) != set(self.slices):
                missing, extra = set(self.slices) - set(sd[key]), set(sd[key]) - set(self.slices)
                raise KeyError("optimizer state %s: missing %s, unexpected %s" % (key, sorted(missing)[:3], sorted(extra)[:3]))
        with torch.no_grad():
            for key, buf in (("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq)):
                for n, (o, k, shape) in self.slices.items():
                    t = sd[key][n]
                    if tuple(t.shape) != shape:
                        raise ValueError("optimizer state %s[%s]: shape %s, expected %s" % (key, n, tuple(t.shape), shape))
                    buf[o:o + k].view(shape).copy_(t)      # in place: a captured CUDA graph keeps pointing at these buffers
            self.step = int(sd["step"])
            if self.dev_state is not None:
                self.dev_state[0] = float(self.step)

    def use_device_step(self):
        """Move the step counter / bias corrections to the device (idempotent); continues from the current step."""
        if self.dev_state is None:
            self.dev_state = torch.zeros(3, dtype=torch.float32, device=self.flat.device)
            self.dev_state[0] = float(self.step)


OVERLAP_CLEAN_BRANCH = bool(int(__import__("os").environ.get("FDG_OVERLAP", "1")))
OVERLAP_PERC_BRANCH = bool(int(__import__("os").environ.get("FDG_OVERLAP_PERC", "1")))     # Vgg16(fake) forward / backward beside the D step
_AUX_STREAMS = {}


def _aux_stream(dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    s = _AUX_STREAMS.get(key)
    if s is None:
        s = _AUX_STREAMS[key] = torch.cuda.Stream(device=dev)
    return s


class GANTrainer:
    def __init__(self, netG, netD, vgg, lr=2e-4, betas=(0.5, 0.999), eps=1e-8, weights=None, perc_layers=(1, 3),
                 process_group=None):
        self.G, self.D, self.V = netG, netD, vgg
        self.lr, self.betas, self.eps = lr, betas, eps
        self.w = dict(DEFAULT_WEIGHTS)
        if weights:
            self.w.update(weights)
        self.perc_layers = tuple(perc_layers)
        self.group = process_group
        self.world = fdist.world_size(process_group)
        self.sG, self.sD = FlatState(netG), FlatState(netD)
        fdist.broadcast_flat_(self.sG.flat, 0, process_group)
        fdist.broadcast_flat_(self.sD.flat, 0, process_group)
        dev = self.sG.flat.device
        self.loss_buf = torch.zeros(5, dtype=torch.float64, device=dev)   # lossD, weighted l1 / perceptual / adversarial / -w*mean(ssim) terms of lossG
        self.last = {}

    # ------------------------------------------------------------------
    def step_graphed(self, hazy: torch.Tensor, clean: torch.Tensor, sync_losses: bool = True):
        """``step`` replayed from a captured CUDA graph (single-GPU; fixed input shape).  At small batch the step is
        launch-bound (~1150 kernel launches); the graph removes the Python / launch overhead.  The first call warms up
        eagerly and captures; the step counter of Adam lives on the device so that replays stay correct."""
        if self.world != 1:
            raise RuntimeError("step_graphed: single-process only (the NCCL all-reduce is not captured)")
        key = (tuple(hazy.shape), tuple(clean.shape))
        if getattr(self, "_graph_key", None) != key:
            self.sG.use_device_step()
            self.sD.use_device_step()
            self._static_hazy, self._static_clean = hazy.clone(), clean.clone()
            # the warm-up below is a real step: snapshot everything it changes and put it back afterwards
            state = [self.sG.flat, self.sG.exp_avg, self.sG.exp_avg_sq, self.sG.dev_state, self.sD.flat, self.sD.exp_avg,
                     self.sD.exp_avg_sq, self.sD.dev_state] + list(self.G.buffers()) + list(self.D.buffers())
            snap = [t.clone() for t in state]
            step0 = (self.sG.step, self.sD.step)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.step(self._static_hazy, self._static_clean, sync_losses=False)      # warm-up: attributes, caches, pools
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            # the graph bakes in raw device pointers of the cached operand images of frozen parameters (Vgg16): hold strong
            # references for the life of the graph, so that evictions from ops._FROZEN cannot free memory a replay still reads
            self._graph_keepalive = []
            ops._CAPTURE_KEEPALIVE = self._graph_keepalive
            try:
                with torch.cuda.graph(self._graph):
                    self._static_fake = self.step(self._static_hazy, self._static_clean, sync_losses=False)
            finally:
                ops._CAPTURE_KEEPALIVE = None
            self._graph_key = key
            for t_, s_ in zip(state, snap):
                t_.copy_(s_)
            self.sG.step, self.sD.step = step0
        self._static_hazy.copy_(hazy)
        self._static_clean.copy_(clean)
        self._graph.replay()
        self.sG.step += 1
        self.sD.step += 1
        if sync_losses:
            self._read_losses()
        return self._static_fake

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        """Everything a resumed run needs: both networks (reference-keyed state dicts, loadable by the reference's modules'
        load_state_dict), both Adam states by parameter name, and the hyper-parameters for a consistency check.  The keys carry no
        'module.' prefix; for a checkpoint the reference's demo.py:78-86 can consume (it strips 7 characters from every key) use
        ``torch.save(dataparallel_state_dict(trainer.G), path)``."""
        return {"version": 1, "netG": {k: v.detach().clone() for k, v in self.G.state_dict().items()},
                "netD": {k: v.detach().clone() for k, v in self.D.state_dict().items()},
                "optG": self.sG.state_dict(), "optD": self.sD.state_dict(),
                "hyper": {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weights": dict(self.w),
                          "perc_layers": tuple(self.perc_layers)}}

    def load_state_dict(self, sd, strict_hyper: bool = False):
        """In place: parameters stay views of the flat buffers (module.load_state_dict copies), so flat gradients, the fused
        Adam and a captured CUDA graph remain valid."""
        if sd.get("version") != 1:
            raise ValueError("unknown trainer checkpoint version %r" % (sd.get("version"),))
        if strict_hyper:
            mine = self.state_dict_hyper()
            if sd["hyper"] != mine:
                raise ValueError("hyper-parameters differ: checkpoint %s, trainer %s" % (sd["hyper"], mine))
        self.G.load_state_dict(sd["netG"])
        self.D.load_state_dict(sd["netD"])
        self.sG.load_state_dict(sd["optG"])
        self.sD.load_state_dict(sd["optD"])

    def state_dict_hyper(self):
        return {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weights": dict(self.w), "perc_layers": tuple(self.perc_layers)}

    def _read_losses(self):
        v = self.loss_buf.tolist()    # the step's device->host read; slots 1..3 hold the WEIGHTED generator terms
        w_ssim = float(self.w.get("ssim", 0.0))
        ssim_w = (w_ssim + v[4]) if w_ssim != 0.0 else 0.0
        self.last = dict(loss_d=v[0], l1_weighted=v[1], perc_weighted=v[2], adv_weighted=v[3], ssim_weighted=ssim_w,
                         loss_g=v[1] + v[2] + v[3] + ssim_w)

    def step(self, hazy: torch.Tensor, clean: torch.Tensor, sync_losses: bool = True, apply: bool = True, allreduce: bool = True):
        """One D update and one G update on this rank's shard.  hazy, clean: [b,3,H,W] fp32 CUDA.
        ``apply=False`` stops after the gradients (no Adam: ``sD.grad`` / ``sG.grad`` hold the summed-over-ranks gradients of
        the CURRENT parameters); ``allreduce=False`` keeps them local.  Both exist for bench.py's data-parallel self-check."""
        G, D, V = self.G, self.D, self.V
        dev = hazy.device
        lb = self.loss_buf
        lb.zero_()
        b1, b2 = self.betas
        gscale = 1.0 / self.world

        if tuple(clean.shape) != tuple(hazy.shape):
            raise ValueError("clean %s does not match hazy %s" % (tuple(clean.shape), tuple(hazy.shape)))
        clean = clean.contiguous()
        B, _, H, W = clean.shape
        want_perc = self.w["perc"] != 0.0 and bool(self.perc_layers)
        # Everything that depends on the clean image only -- its frequency decomposition, D(real) forward, Vgg16(clean) -- runs on an
        # auxiliary stream beside the generator forward: the generator's small-map kernels (64x64 / 32x32 maps, one-CTA BatchNorm
        # finalisations) leave SMs idle that these tensor-bound kernels fill; at batch 1 the two chains simply run side by side.
        # Ordering: the auxiliary stream starts after everything enqueued so far (previous optimiser steps) and is joined before the
        # first consumer; tensors it allocates are reused only by its own later allocations, i.e. after the next fork.
        overlap = OVERLAP_CLEAN_BRANCH and dev.type == "cuda"
        main = torch.cuda.current_stream(dev) if overlap else None
        aux = _aux_stream(dev) if overlap else None
        cctx = None

        def clean_branch():
            z_real = View.alloc(B, H, W, 9, dev)
            ops.freq_concat_fwd(View.from_nchw(clean), z_real)
            pr_, ctx_r_ = engine.discriminator_forward(D, z_real.as_nchw(), True, True)
            cc = engine.vgg_forward(V, clean, False)[1] if want_perc else None
            return z_real, pr_, ctx_r_, cc

        if overlap:
            aux.wait_stream(main)
            with torch.cuda.stream(aux):
                z_real, pr, ctx_r, cctx = clean_branch()
        fake, gctx = engine.generator_forward(G, hazy, True, True)
        if tuple(clean.shape) != tuple(fake.shape):
            raise ValueError("clean %s does not match G(hazy) %s" % (tuple(clean.shape), tuple(fake.shape)))
        z_fake = View.alloc(B, H, W, 9, dev)
        ops.freq_concat_fwd(View.from_nchw(fake), z_fake)

        # ---------------- D step
        self.sD.zero_grad()
        if overlap:
            main.wait_stream(aux)
        else:
            z_real, pr, ctx_r, cctx = clean_branch()

        # The perceptual branch -- Vgg16(fake), the feature-space MSE gradients, the Vgg16 data gradient -- needs `fake` only: it runs on
        # the auxiliary stream beside the whole discriminator step and the adversarial gradient (second fork; joined where dfake takes
        # the perceptual gradient).  At batch 1 this hides ~45 latency-bound launches; at batch 16 the kernels just interleave.
        def perc_branch():
            _fo, vctx = engine.vgg_forward(V, fake, True)
            gouts = [None, None, None, None]
            for k in self.perc_layers:
                fk, ck = vctx.feats[k], cctx.feats[k]
                n_k = fk.N * fk.H * fk.W * fk.C
                gk = View.alloc(fk.N, fk.H, fk.W, fk.C, dev)
                ops.loss_grad(ops.LOSS_MSE, fk.base, ck.base, n_k, self.w["perc"] / n_k, lb[2:3], gk.base)
                gouts[k] = gk.as_nchw()
            return engine.vgg_backward(V, vctx, gouts, None, True)

        overlap_p = overlap and want_perc and OVERLAP_PERC_BRANCH
        dxv = None
        if overlap_p:
            aux.wait_stream(main)
            with torch.cuda.stream(aux):
                dxv = perc_branch()
        pf, ctx_f = engine.discriminator_forward(D, z_fake.as_nchw(), True, True)
        n_p = pr.numel()
        dpr, dpf = torch.empty_like(pr), torch.empty_like(pf)
        ops.loss_grad(ops.LOSS_BCE, pr, None, n_p, 1.0 / n_p, lb[0:1], dpr, target=1.0)
        ops.loss_grad(ops.LOSS_BCE, pf, None, n_p, 1.0 / n_p, lb[0:1], dpf, target=0.0)
        engine.discriminator_backward(D, ctx_r, dpr, self.sD.grad_views, False)
        engine.discriminator_backward(D, ctx_f, dpf, self.sD.grad_views, False)
        del ctx_r, ctx_f
        if allreduce:
            fdist.allreduce_flat_(self.sD.grad, self.group)
        if apply:
            self.sD.adam(self.lr, b1, b2, self.eps, gscale)

        # ---------------- G step (D frozen: data gradient only)
        self.sG.zero_grad()
        pf2, ctx_f2 = engine.discriminator_forward(D, z_fake.as_nchw(), True, True)
        dpf2 = torch.empty_like(pf2)
        ops.loss_grad(ops.LOSS_BCE, pf2, None, n_p, self.w["adv"] / n_p, lb[3:4], dpf2, target=1.0)
        dz = engine.discriminator_backward(D, ctx_f2, dpf2, None, True)
        del ctx_f2
        dfake = torch.empty_like(fake)
        scratch = torch.empty(B * 3 * H * W, dtype=torch.float32, device=dev)
        ops.freq_concat_bwd(View.from_nchw(dz), View.from_nchw(dfake), scratch)
        n_img = fake.numel()
        ops.loss_grad(ops.LOSS_L1, fake, clean, n_img, self.w["l1"] / n_img, lb[1:2], dfake, accumulate=True)
        w_ssim = float(self.w.get("ssim", 0.0))
        if w_ssim != 0.0:      # w * (1 - mean ssim_map(fake, clean)), models/pytorch_ssim/__init__.py:17-37
            ops.ssim_loss_grad(View.from_nchw(fake), View.from_nchw(clean), -w_ssim / n_img, -w_ssim / n_img, lb[4:5],
                               View.from_nchw(dfake), accumulate=True)
        if want_perc:
            if overlap_p:
                main.wait_stream(aux)
            else:
                dxv = perc_branch()
            cctx = None
            ops.copy4d(View.from_nchw(dxv), View.from_nchw(dfake), accumulate=True)
        engine.generator_backward(G, gctx, dfake, self.sG.grad_views, False)
        del gctx
        if allreduce:
            fdist.allreduce_flat_(self.sG.grad, self.group)
        if apply:
            self.sG.adam(self.lr, b1, b2, self.eps, gscale)

        if sync_losses:
            self._read_losses()
        return fake
