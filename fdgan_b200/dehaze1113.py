"""``models/dehaze1113.py`` surface of the reference, backed by fdgan_b200 kernels.

Drop-in for ``import models.dehaze1113 as net; net.FDGAN()`` (demo.py:18,73) and ``net.D(nc, nf)``
(models/dehaze1113.py:188-230): same constructor signatures, same attribute paths, same state-dict keys
and shapes (786 entries for FDGAN), so ``load_state_dict`` of a reference checkpoint (after the
``module.`` strip of demo.py:82-84, which ``load_state_dict`` here also does itself) works unchanged.

The sub-modules below only HOLD parameters/buffers under the reference's names; none of them computes.
``forward`` runs the whole network through fdgan_b200.engine inside one autograd node.
"""
from __future__ import annotations

import math
import re
from collections import OrderedDict

import torch
import torch.nn as nn

from . import engine

# --------------------------------------------------------------------------------------------------
# parameter holders (names follow torch.nn / torchvision so that state-dict keys match the reference)
# --------------------------------------------------------------------------------------------------


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("%s only holds parameters; the enclosing fdgan_b200 network runs the kernels" % type(self).__name__)


class ConvParams(_Holder):
    """weight [Cout,Cin,k,k] (+ bias) of an nn.Conv2d; transposed=True: nn.ConvTranspose2d weight [Cin,Cout,k,k]."""

    def __init__(self, c_in, c_out, k, bias, transposed=False, init="default"):
        super().__init__()
        shape = (c_in, c_out, k, k) if transposed else (c_out, c_in, k, k)
        self.weight = nn.Parameter(torch.empty(shape))
        fan_in = shape[1] * k * k
        if init == "kaiming_normal":      # torchvision DenseNet convs
            nn.init.kaiming_normal_(self.weight)
        else:                             # torch.nn.Conv2d default
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            self.bias = nn.Parameter(torch.empty(c_out))
            bound = 1.0 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)
        else:
            self.register_parameter("bias", None)


class BatchNormParams(_Holder):
    """weight/bias/running_mean/running_var/num_batches_tracked of an nn.BatchNorm2d (eps 1e-5, momentum 0.1)."""

    def __init__(self, c):
        super().__init__()
        self.num_features = c
        self.eps = 1e-5
        self.momentum = 0.1
        self.track_running_stats = True
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        # PyTorch-0.3-era checkpoints (README.md:24) have no num_batches_tracked: same leniency as nn.BatchNorm2d
        key = prefix + "num_batches_tracked"
        if key not in state_dict:
            state_dict[key] = torch.tensor(0, dtype=torch.long)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


class _DenseLayerParams(_Holder):
    def __init__(self, c_in, growth=32, bn_size=4):
        super().__init__()
        self.norm1 = BatchNormParams(c_in)
        self.conv1 = ConvParams(c_in, bn_size * growth, 1, False, init="kaiming_normal")
        self.norm2 = BatchNormParams(bn_size * growth)
        self.conv2 = ConvParams(bn_size * growth, growth, 3, False, init="kaiming_normal")


class _DenseBlockParams(_Holder):
    def __init__(self, n_layers, c_in, growth=32):
        super().__init__()
        self.n_layers, self.c_in = n_layers, c_in
        for i in range(n_layers):
            self.add_module("denselayer%d" % (i + 1), _DenseLayerParams(c_in + i * growth, growth))


class _TransitionParams(_Holder):
    def __init__(self, c_in, c_out):
        super().__init__()
        self.norm = BatchNormParams(c_in)
        self.conv = ConvParams(c_in, c_out, 1, False, init="kaiming_normal")


class BottleneckBlockdy(_Holder):
    """models/dehaze1113.py:256-275 (bn1/bn2 are registered by the reference but never executed)."""

    def __init__(self, in_planes, out_planes, dropRate=0.0):
        super().__init__()
        if dropRate != 0.0:
            raise NotImplementedError("dropRate > 0 is not used by FDGAN")
        inter = out_planes * 4
        self.bn1 = BatchNormParams(in_planes)
        self.conv1 = ConvParams(in_planes, inter, 1, False)
        self.bn2 = BatchNormParams(inter)
        self.conv2 = ConvParams(inter, out_planes, 3, False)
        self.droprate = dropRate


class TransitionBlockdy(_Holder):
    """models/dehaze1113.py:358-370."""

    def __init__(self, in_planes, out_planes, dropRate=0.0):
        super().__init__()
        if dropRate != 0.0:
            raise NotImplementedError("dropRate > 0 is not used by FDGAN")
        self.bn1 = BatchNormParams(in_planes)
        self.conv1 = ConvParams(in_planes, out_planes, 1, False, transposed=True)
        self.droprate = dropRate


_LEGACY_DENSE = re.compile(r"^(.*denselayer\d+\.(?:norm|relu|conv))\.((?:[12])\.(?:weight|bias|running_mean|running_var))$")


def _remap_checkpoint_keys(state_dict):
    """`module.` prefix of DataParallel checkpoints (demo.py:82-84) and torchvision-0.2 dense-layer names
    `norm.1/conv.1/norm.2/conv.2` (the regex torchvision applies in its own loader)."""
    out = OrderedDict()
    meta = getattr(state_dict, "_metadata", None)
    for k, v in state_dict.items():
        if k.startswith("module."):
            k = k[7:]
        mt = _LEGACY_DENSE.match(k)
        if mt:
            k = mt.group(1) + mt.group(2)
        out[k] = v
    if meta is not None:
        out._metadata = meta
    return out


# --------------------------------------------------------------------------------------------------
# autograd nodes
# --------------------------------------------------------------------------------------------------


def _alloc_grads(named_params):
    """One zero-filled flat buffer holding the gradient of every used parameter, plus name -> view."""
    total = sum(p.numel() for _n, p in named_params)
    dev = named_params[0][1].device
    flat = torch.zeros(total, dtype=torch.float32, device=dev)
    views, off = {}, 0
    for n, p in named_params:
        views[n] = flat[off:off + p.numel()].view(p.shape)
        off += p.numel()
    return flat, views


class _NetFn(torch.autograd.Function):
    """One autograd node for a whole network: inputs (module, x, *used parameters)."""

    @staticmethod
    def forward(ctx, mod, x, *params):
        need = any(ctx.needs_input_grad[1:])
        out, ectx = mod._run_forward(x, need)
        ctx.mod, ctx.ectx = mod, ectx
        return out

    @staticmethod
    def backward(ctx, dout):
        mod = ctx.mod
        need_dx = ctx.needs_input_grad[1]
        need_w = any(ctx.needs_input_grad[2:])
        named = mod._used_named_parameters()
        grads = None
        sink = None if getattr(mod, "_is_replica", False) else mod._grad_sink   # a replica's gradients go back through autograd
        if need_w:
            if sink is not None:
                grads = sink                     # caller-owned flat gradient buffer (training step fast path)
            else:
                _flat, grads = _alloc_grads(named)
        if ctx.ectx is None:      # the saved activations are released by the first backward (they are tens of MB to GB)
            raise RuntimeError("fdgan_b200: trying to backward through %s a second time: the saved activations of this forward have already "
                               "been freed (retain_graph=True is not supported; run the forward again)" % type(mod).__name__)
        dx = mod._run_backward(ctx.ectx, dout, grads, need_dx)
        ctx.ectx = None
        if not need_w or sink is not None:
            return (None, dx) + (None,) * len(named)
        return (None, dx) + tuple(grads[n] if ctx.needs_input_grad[2 + i] else None for i, (n, _p) in enumerate(named))


class _KernelNet(nn.Module):
    """Common plumbing: used-parameter list, flat-gradient sink, checkpoint key remap."""

    _grad_sink = None

    def _used_named_parameters(self):
        if getattr(self, "_is_replica", False):
            # nn.DataParallel replica (demo.py:89 with several GPUs): torch.nn.parallel.replicate copies __dict__ (the cache
            # below would still name the ORIGINAL module's parameters), empties _parameters and stores this device's copies as
            # plain attributes, listed in _former_parameters.  Same names, same order as named_parameters() of the original.
            out = []
            for mname, m in self.named_modules():
                for k, v in getattr(m, "_former_parameters", {}).items():
                    n = mname + "." + k if mname else k
                    if self._is_used(n):
                        out.append((n, v))
            return out
        cache = getattr(self, "_used_cache", None)
        if cache is None:
            cache = [(n, p) for n, p in self.named_parameters() if self._is_used(n)]
            object.__setattr__(self, "_used_cache", cache)
        return cache

    def _apply(self, fn, *a, **k):
        object.__setattr__(self, "_used_cache", None)   # .cuda()/.to() replace parameter storage
        return super()._apply(fn, *a, **k)

    def _is_used(self, name):
        return True

    def set_grad_sink(self, views):
        """Route parameter gradients into caller-owned, pre-zeroed tensors (name -> view of a flat buffer) instead
        of returning them through autograd; used by fdgan_b200.train for the single NCCL all-reduce."""
        object.__setattr__(self, "_grad_sink", views)

    def load_state_dict(self, state_dict, strict=True, **kw):
        return super().load_state_dict(_remap_checkpoint_keys(state_dict), strict=strict, **kw)

    def _align_replica_tensors(self):
        """nn.DataParallel replicas on devices other than the first receive their parameters as views into ONE coalesced broadcast
        buffer (torch.nn.parallel.replicate -> comm.broadcast_coalesced), i.e. at arbitrary 4-byte offsets; the kernels read weights,
        biases and BatchNorm vectors with 128-bit loads and bulk copies.  The executors read parameters through the attribute path
        (``lyr.conv1.weight``), so every misaligned tensor of a replica is shadowed there by a 16-byte aligned copy (one multi-tensor
        copy); the autograd node keeps the broadcast outputs as its inputs, so gradients still flow back to the original."""
        if getattr(self, "_replica_aligned", False):
            return
        mods, keys, src = [], [], []
        for m in self.modules():
            for k, v in list(getattr(m, "_former_parameters", {}).items()) + list(m._buffers.items()):
                if v is not None and v.is_floating_point() and v.data_ptr() % 16 != 0:
                    mods.append(m); keys.append(k); src.append(v.detach())
        if src:
            dst = [torch.empty_like(v, memory_format=torch.contiguous_format) for v in src]
            torch._foreach_copy_(dst, src)
            for m, k, d in zip(mods, keys, dst):
                if k in m._buffers:
                    m._buffers[k] = d           # running statistics of a replica are scratch (DataParallel keeps the first device's)
                else:
                    m.__dict__[k] = d
        object.__setattr__(self, "_replica_aligned", True)

    def _call(self, x):
        if getattr(self, "_is_replica", False):
            self._align_replica_tensors()
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for _n, p in self._used_named_parameters())):
            return _NetFn.apply(self, x, *[p for _n, p in self._used_named_parameters()])
        out, _ = self._run_forward(x, False)
        return out


# --------------------------------------------------------------------------------------------------
# FDGAN generator
# --------------------------------------------------------------------------------------------------


class FDGAN(_KernelNet):
    """models/dehaze1113.py:702-801.  ``FDGAN()`` takes no arguments, like the reference; the DenseNet-121
    pieces are initialised like torchvision's (the reference additionally downloads ImageNet weights at
    dehaze1113.py:707 -- load them through ``load_state_dict`` if you have them)."""

    _UNUSED = ("conv0.", "dense_block31.", "dense_norm31.", "dense_block4.bn", "dense_block5.bn", "dense_block6.bn",
               "trans_block4.bn", "trans_block5.bn", "trans_block6.bn")

    def __init__(self):
        super().__init__()
        self.conv0 = ConvParams(3, 64, 7, False, init="kaiming_normal")          # registered, never executed
        self.dense_block1 = _DenseBlockParams(6, 64)
        self.trans_block1 = _TransitionParams(256, 128)
        self.dense_block2 = _DenseBlockParams(12, 128)
        self.trans_block2 = _TransitionParams(512, 256)
        self.dense_block3 = _DenseBlockParams(24, 256)
        self.trans_block3 = _TransitionParams(1024, 512)
        self.dense_block31 = _DenseBlockParams(16, 512)                            # registered, never executed
        self.dense_norm31 = BatchNormParams(1024)                                  # registered, never executed
        self.dense_block4 = BottleneckBlockdy(512, 256)
        self.trans_block4 = TransitionBlockdy(768, 128)
        self.dense_block5 = BottleneckBlockdy(384, 128)
        self.trans_block5 = TransitionBlockdy(512, 64)
        self.dense_block6 = BottleneckBlockdy(64, 32)
        self.trans_block6 = TransitionBlockdy(96, 16)
        self.conv_refin1 = ConvParams(3, 64, 3, True)
        self.conv_refin6 = ConvParams(640, 512, 3, True)
        self.conv_refin5 = ConvParams(256, 128, 1, True)
        self.conv_refin3 = ConvParams(16, 3, 3, True)
        self.conv_refin2 = ConvParams(64, 32, 1, True)
        self.conv_refine4 = ConvParams(160, 128, 3, True)

    def _is_used(self, name):
        return not name.startswith(self._UNUSED)

    def _run_forward(self, x, need_ctx):
        return engine.generator_forward(self, x, self.training, need_ctx)

    def _run_backward(self, ectx, dout, grads, need_dx):
        return engine.generator_backward(self, ectx, dout, grads, need_dx)

    def forward(self, x):
        return self._call(x)


# --------------------------------------------------------------------------------------------------
# Fusion-discriminator
# --------------------------------------------------------------------------------------------------


class _Named(_Holder):
    """A container whose single child carries the reference's second-level name (main.layer2.layer2.conv ...)."""


class D(_KernelNet):
    """models/dehaze1113.py:188-230: conv4x4 s2 -> [LeakyReLU -> conv3x3 -> BN] x2 -> LeakyReLU -> conv4x4 s1 ->
    LeakyReLU -> conv4x4 s1 -> Sigmoid.  State-dict keys: main.layer1.conv.weight, main.layer2.layer2.{conv,bn}.*,
    main.layer3.layer3.{conv,bn}.*, main.layer4.conv.weight, main.layer5.conv.weight."""

    def __init__(self, nc, nf):
        super().__init__()
        self.nc, self.nf = nc, nf
        main = _Named()
        main.layer1 = _Named()
        main.layer1.conv = ConvParams(nc, nf, 4, False)
        for idx, c in ((2, nf), (3, 2 * nf)):
            outer, inner = _Named(), _Named()
            inner.conv = ConvParams(c, 2 * c, 3, False)
            inner.bn = BatchNormParams(2 * c)
            outer.add_module("layer%d" % idx, inner)
            main.add_module("layer%d" % idx, outer)
        main.layer4 = _Named()
        main.layer4.conv = ConvParams(4 * nf, 8 * nf, 4, False)
        main.layer5 = _Named()
        main.layer5.conv = ConvParams(8 * nf, 1, 4, False)
        self.main = main

    def layer_params(self):
        mn = self.main
        return mn.layer1.conv, mn.layer2.layer2, mn.layer3.layer3, mn.layer4.conv, mn.layer5.conv

    def _run_forward(self, x, need_ctx):
        return engine.discriminator_forward(self, x, self.training, need_ctx)

    def _run_backward(self, ectx, dout, grads, need_dx):
        return engine.discriminator_backward(self, ectx, dout, grads, need_dx)

    def forward(self, x):
        return self._call(x)
