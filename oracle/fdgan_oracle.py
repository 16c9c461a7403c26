"""Functional CPU restatement of the FD-GAN hot path (TEST INFRASTRUCTURE ONLY).

Plain ``torch.nn.functional`` on fp32 (or fp64) CPU tensors, written as pure
functions over a reference-keyed state dict so the same dict loads into the
reference modules, into this oracle and into ``fdgan_b200``.

What each function follows (paths relative to /root/reference):
  * ``fdgan_forward``      models/dehaze1113.py:758-801 (FDGAN.forward) with the
                           torchvision DenseNet-121 pieces it borrows at :707-728
                           (dense layer / transition arithmetic as restated in-tree
                           at models/densenet.py:179-242)
  * ``bottleneck_dy``      models/dehaze1113.py:268-275 (in-place ReLU => the skip
                           half of the concat is relu(x))
  * ``transition_dy``      models/dehaze1113.py:366-370
  * ``d_forward``          models/dehaze1113.py:188-230 with blockUNet1 :29-43
  * ``vgg16_forward``      myutils/vgg16.py:27-49
  * ``blur`` / ``laplacian``  __pycache__/loss.cpython-36.pyc (source lines
                           L122-162 / L205-304 recovered in SURVEY Appendix B) --
                           bytecode only, "parity unpinned"
  * ``ssim``               models/pytorch_ssim/__init__.py:7-37
  * ``train_step``         RECONSTRUCTED; the reference has no train.py, so the choice of loss terms
                           and weights is unpinned; the arithmetic is pinned by tests/golden/train_step_*.npz (two
                           iterations of the reference's own modules + torch.optim.Adam, make_golden.py:gen_train_step)

BatchNorm always uses batch statistics when ``train=True`` (README.md:38, demo.py
never calls .eval()).
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
CONV_GAIN = 2.0  # widens the BN-free decoder output to about +-0.9 (a shrunken output makes the 1e-3 bar vacuous)

# --------------------------------------------------------------------------------------
# parameter tables (names/shapes of the reference state dicts)
# --------------------------------------------------------------------------------------


def _bn_specs(prefix, c):
    return [
        (prefix + ".weight", (c,), "bn_w"),
        (prefix + ".bias", (c,), "bn_b"),
        (prefix + ".running_mean", (c,), "bn_rm"),
        (prefix + ".running_var", (c,), "bn_rv"),
        (prefix + ".num_batches_tracked", (), "bn_nbt"),
    ]


def _dense_block_specs(prefix, n_layers, c_in, growth=32, bn_size=4):
    out = []
    for i in range(n_layers):
        c = c_in + growth * i
        p = "%s.denselayer%d" % (prefix, i + 1)
        out += _bn_specs(p + ".norm1", c)
        out.append((p + ".conv1.weight", (bn_size * growth, c, 1, 1), "conv_tv"))
        out += _bn_specs(p + ".norm2", bn_size * growth)
        out.append((p + ".conv2.weight", (growth, bn_size * growth, 3, 3), "conv_tv"))
    return out


def _transition_specs(prefix, c_in, c_out):
    return _bn_specs(prefix + ".norm", c_in) + [(prefix + ".conv.weight", (c_out, c_in, 1, 1), "conv_tv")]


def _bdy_specs(prefix, c_in, c_out):
    inter = 4 * c_out
    return (
        _bn_specs(prefix + ".bn1", c_in)
        + [(prefix + ".conv1.weight", (inter, c_in, 1, 1), "conv")]
        + _bn_specs(prefix + ".bn2", inter)
        + [(prefix + ".conv2.weight", (c_out, inter, 3, 3), "conv")]
    )


def _tdy_specs(prefix, c_in, c_out):
    # ConvTranspose2d weight is [Cin, Cout, 1, 1]
    return _bn_specs(prefix + ".bn1", c_in) + [(prefix + ".conv1.weight", (c_in, c_out, 1, 1), "convT")]


def _conv_b_specs(prefix, c_out, c_in, k):
    return [(prefix + ".weight", (c_out, c_in, k, k), "conv"), (prefix + ".bias", (c_out,), "bias:%d" % (c_in * k * k))]


def fdgan_specs():
    """(name, shape, kind) for every FDGAN state-dict entry (models/dehaze1113.py:703-755)."""
    s = [("conv0.weight", (64, 3, 7, 7), "conv_tv")]
    s += _dense_block_specs("dense_block1", 6, 64)
    s += _transition_specs("trans_block1", 256, 128)
    s += _dense_block_specs("dense_block2", 12, 128)
    s += _transition_specs("trans_block2", 512, 256)
    s += _dense_block_specs("dense_block3", 24, 256)
    s += _transition_specs("trans_block3", 1024, 512)
    s += _dense_block_specs("dense_block31", 16, 512)
    s += _bn_specs("dense_norm31", 1024)
    s += _bdy_specs("dense_block4", 512, 256)
    s += _tdy_specs("trans_block4", 768, 128)
    s += _bdy_specs("dense_block5", 384, 128)
    s += _tdy_specs("trans_block5", 512, 64)
    s += _bdy_specs("dense_block6", 64, 32)
    s += _tdy_specs("trans_block6", 96, 16)
    s += _conv_b_specs("conv_refin1", 64, 3, 3)
    s += _conv_b_specs("conv_refin6", 512, 640, 3)
    s += _conv_b_specs("conv_refin5", 128, 256, 1)
    s += _conv_b_specs("conv_refin3", 3, 16, 3)
    s += _conv_b_specs("conv_refin2", 32, 64, 1)
    s += _conv_b_specs("conv_refine4", 128, 160, 3)
    return s


def d_specs(nc, nf):
    """Fusion-discriminator D(nc, nf) state dict (models/dehaze1113.py:188-226)."""
    s = [("main.layer1.conv.weight", (nf, nc, 4, 4), "conv")]
    s.append(("main.layer2.layer2.conv.weight", (2 * nf, nf, 3, 3), "conv"))
    s += _bn_specs("main.layer2.layer2.bn", 2 * nf)
    s.append(("main.layer3.layer3.conv.weight", (4 * nf, 2 * nf, 3, 3), "conv"))
    s += _bn_specs("main.layer3.layer3.bn", 4 * nf)
    s.append(("main.layer4.conv.weight", (8 * nf, 4 * nf, 4, 4), "conv"))
    s.append(("main.layer5.conv.weight", (1, 8 * nf, 4, 4), "conv"))
    return s


VGG_CFG = [
    ("conv1_1", 3, 64), ("conv1_2", 64, 64),
    ("conv2_1", 64, 128), ("conv2_2", 128, 128),
    ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256),
    ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512),
    ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512),
]


def vgg_specs():
    """Vgg16 state dict (myutils/vgg16.py:9-25)."""
    s = []
    for name, ci, co in VGG_CFG:
        s += [(name + ".weight", (co, ci, 3, 3), "conv_tv"), (name + ".bias", (co,), "bias:%d" % (ci * 9))]
    return s


def make_state(specs, seed=0, dtype=torch.float32):
    """Deterministic weights, independent of construction order: every entry is drawn
    from its own generator seeded by crc32(name) ^ seed.  Magnitudes follow the default
    initialisers of the layers involved (torchvision kaiming-normal for its convs,
    PyTorch kaiming-uniform(a=sqrt 5) elsewhere) with non-trivial BN affine/running
    values so that every term of the arithmetic is exercised."""
    sd = OrderedDict()
    for name, shape, kind in specs:
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
        if kind == "bn_nbt":
            sd[name] = torch.zeros((), dtype=torch.long)
            continue
        if kind == "conv_tv":
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g, dtype=torch.float64) * math.sqrt(2.0 / fan_in)
        elif kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            b = CONV_GAIN / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * b
        elif kind == "convT":
            fan_in = shape[0]
            b = CONV_GAIN / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * b
        elif kind.startswith("bias:"):
            b = 1.0 / math.sqrt(int(kind.split(":")[1]))
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * b
        elif kind == "bn_w":
            t = 0.8 + 0.4 * torch.rand(shape, generator=g, dtype=torch.float64)
        elif kind == "bn_b":
            t = 0.2 * torch.rand(shape, generator=g, dtype=torch.float64) - 0.1
        elif kind == "bn_rm":
            t = 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif kind == "bn_rv":
            t = 0.8 + 0.4 * torch.rand(shape, generator=g, dtype=torch.float64)
        else:
            raise KeyError(kind)
        sd[name] = t.to(dtype)
    return sd


def make_fdgan_state(seed=0, dtype=torch.float32):
    return make_state(fdgan_specs(), seed, dtype)


def make_d_state(nc=9, nf=36, seed=1, dtype=torch.float32):
    return make_state(d_specs(nc, nf), seed, dtype)


def make_vgg_state(seed=2, dtype=torch.float32):
    return make_state(vgg_specs(), seed, dtype)


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------


def _bn(sd, prefix, x, train, update):
    """nn.BatchNorm2d forward.  In train mode uses batch statistics and (optionally)
    mutates running_mean/var in ``sd`` the way torch does (momentum 0.1, unbiased var)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if train:
        if update:
            y = F.batch_norm(x, rm, rv, w, b, True, BN_MOMENTUM, BN_EPS)
            sd[prefix + ".num_batches_tracked"] += 1
        else:
            y = F.batch_norm(x, None, None, w, b, True, BN_MOMENTUM, BN_EPS)
        return y
    return F.batch_norm(x, rm, rv, w, b, False, BN_MOMENTUM, BN_EPS)


def dense_layer(sd, p, feats, train, update):
    """torchvision _DenseLayer (spec copy: models/densenet.py:188-211)."""
    x = torch.cat(feats, 1)
    y = F.conv2d(F.relu(_bn(sd, p + ".norm1", x, train, update)), sd[p + ".conv1.weight"])
    y = F.conv2d(F.relu(_bn(sd, p + ".norm2", y, train, update)), sd[p + ".conv2.weight"], padding=1)
    return y


def dense_block(sd, p, x, n_layers, train, update):
    """torchvision _DenseBlock (spec copy: models/densenet.py:224-242)."""
    feats = [x]
    for i in range(n_layers):
        feats.append(dense_layer(sd, "%s.denselayer%d" % (p, i + 1), feats, train, update))
    return torch.cat(feats, 1)


def transition(sd, p, x, train, update):
    """torchvision _Transition: BN -> ReLU -> 1x1 conv -> AvgPool2 (models/densenet.py:214-221)."""
    y = F.conv2d(F.relu(_bn(sd, p + ".norm", x, train, update)), sd[p + ".conv.weight"])
    return F.avg_pool2d(y, 2, 2)


def bottleneck_dy(sd, p, x):
    """BottleneckBlockdy.forward (models/dehaze1113.py:268-275).  nn.ReLU(inplace=True)
    mutates x, so the concat carries relu(x); bn1/bn2 exist but are not executed."""
    xr = F.relu(x)
    out = F.conv2d(xr, sd[p + ".conv1.weight"])
    out = F.conv2d(F.relu(out), sd[p + ".conv2.weight"], padding=1)
    return torch.cat([xr, out], 1)


def transition_dy(sd, p, x):
    """TransitionBlockdy.forward (models/dehaze1113.py:366-370): ReLU -> ConvTranspose 1x1 -> nearest x2."""
    out = F.conv_transpose2d(F.relu(x), sd[p + ".conv1.weight"])
    return F.interpolate(out, scale_factor=2, mode="nearest")


def fdgan_forward(sd, x, train=True, update_running=True, taps=None):
    """FDGAN.forward (models/dehaze1113.py:758-801).  ``taps``: optional dict that
    receives intermediate activations (NCHW) for block-level parity tests."""
    t = taps if taps is not None else {}
    x0 = F.relu(F.conv2d(x, sd["conv_refin1.weight"], sd["conv_refin1.bias"], padding=1))
    x01 = F.conv2d(F.avg_pool2d(x0, 2), sd["conv_refin2.weight"], sd["conv_refin2.bias"])
    t["x0"], t["x01"] = x0, x01
    b1 = dense_block(sd, "dense_block1", x0, 6, train, update_running)
    x1 = transition(sd, "trans_block1", b1, train, update_running)
    t["b1"], t["x1"] = b1, x1
    x10 = F.conv2d(torch.cat([x01, x1], 1), sd["conv_refine4.weight"], sd["conv_refine4.bias"], padding=1)
    x2 = transition(sd, "trans_block2", dense_block(sd, "dense_block2", x10, 12, train, update_running), train, update_running)
    t["x10"], t["x2"] = x10, x2
    x3 = transition(sd, "trans_block3", dense_block(sd, "dense_block3", x2, 24, train, update_running), train, update_running)
    x22 = F.conv2d(F.avg_pool2d(x2, 2), sd["conv_refin5.weight"], sd["conv_refin5.bias"])
    t["x3"], t["x22"] = x3, x22
    x6in = F.conv2d(torch.cat([x3, x22], 1), sd["conv_refin6.weight"], sd["conv_refin6.bias"], padding=1)
    x4 = transition_dy(sd, "trans_block4", bottleneck_dy(sd, "dense_block4", x6in))
    t["x4"] = x4
    x42 = torch.cat([x4, x2], 1)
    x5 = transition_dy(sd, "trans_block5", bottleneck_dy(sd, "dense_block5", x42))
    t["x5"] = x5
    x6 = transition_dy(sd, "trans_block6", bottleneck_dy(sd, "dense_block6", x5))
    t["x6"] = x6
    return torch.tanh(F.conv2d(x6, sd["conv_refin3.weight"], sd["conv_refin3.bias"], padding=1))


def fdgan_used_param_names():
    """Parameters FDGAN.forward touches (the rest never get a gradient)."""
    names = []
    for name, _shape, kind in fdgan_specs():
        if kind.startswith("bn_r") or kind == "bn_nbt":
            continue
        if name.startswith(("conv0.", "dense_block31.", "dense_norm31.")):
            continue
        if name.startswith(("dense_block4.bn", "dense_block5.bn", "dense_block6.bn",
                            "trans_block4.bn", "trans_block5.bn", "trans_block6.bn")):
            continue
        names.append(name)
    return names


def d_forward(sd, x, train=True, update_running=True):
    """D.forward (models/dehaze1113.py:188-230): conv4x4 s2 -> [LReLU -> conv3x3 -> BN] x2
    -> LReLU -> conv4x4 s1 -> LReLU -> conv4x4 s1 -> sigmoid."""
    y = F.conv2d(x, sd["main.layer1.conv.weight"], stride=2, padding=1)
    y = F.conv2d(F.leaky_relu(y, 0.2), sd["main.layer2.layer2.conv.weight"], padding=1)
    y = _bn(sd, "main.layer2.layer2.bn", y, train, update_running)
    y = F.conv2d(F.leaky_relu(y, 0.2), sd["main.layer3.layer3.conv.weight"], padding=1)
    y = _bn(sd, "main.layer3.layer3.bn", y, train, update_running)
    y = F.conv2d(F.leaky_relu(y, 0.2), sd["main.layer4.conv.weight"], stride=1, padding=1)
    y = F.conv2d(F.leaky_relu(y, 0.2), sd["main.layer5.conv.weight"], stride=1, padding=1)
    return torch.sigmoid(y)


def vgg16_forward(sd, x):
    """Vgg16.forward (myutils/vgg16.py:27-49) -> [relu1_2, relu2_2, relu3_3, relu4_3]."""
    def c(name, h):
        return F.relu(F.conv2d(h, sd[name + ".weight"], sd[name + ".bias"], padding=1))
    h = c("conv1_2", c("conv1_1", x)); r1 = h
    h = F.max_pool2d(h, 2, 2)
    h = c("conv2_2", c("conv2_1", h)); r2 = h
    h = F.max_pool2d(h, 2, 2)
    h = c("conv3_3", c("conv3_2", c("conv3_1", h))); r3 = h
    h = F.max_pool2d(h, 2, 2)
    h = c("conv4_3", c("conv4_2", c("conv4_1", h))); r4 = h
    return [r1, r2, r3, r4]


# --------------------------------------------------------------------------------------
# frequency decomposition (loss.pyc) -- parity unpinned
# --------------------------------------------------------------------------------------

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def isotropic_gaussian_kernel(l=15, sigma=3.0, dtype=torch.float32):
    """loss.pyc@L153-159: exp(-(xx^2+yy^2)/(2 sigma^2)) / sum, float64 math then cast."""
    ax = torch.arange(-l // 2 + 1.0, l // 2 + 1.0, dtype=torch.float64)
    xx, yy = torch.meshgrid(ax, ax, indexing="xy")
    k = torch.exp(-(xx ** 2 + yy ** 2) / (2.0 * sigma ** 2))
    return (k / k.sum()).to(dtype)


def blur(x, l=15, sigma=3.0, use_input_norm=True):
    """Blur.forward (loss.pyc@L142-151): ImageNet-normalise, ReflectionPad2d(l//2),
    one l x l Gaussian applied to every (batch, channel) plane."""
    b, c, h, w = x.shape
    if use_input_norm:
        mean = torch.tensor(IMAGENET_MEAN, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
        x = (x - mean) / std
    pad = F.pad(x, (l // 2,) * 4, mode="reflect")
    k = isotropic_gaussian_kernel(l, sigma, x.dtype).view(1, 1, l, l).to(x.device)
    hp, wp = pad.shape[-2:]
    return F.conv2d(pad.reshape(c * b, 1, hp, wp), k).view(b, c, h, w)


def laplacian(x, kernel_size=3):
    """Laplacian.forward (loss.pyc@L286-301): depth-wise ones(k,k) with centre 1-k^2, zero pad."""
    if x.dim() != 4:
        raise ValueError("Invalid input shape, we expect BxCxHxW. Got: {}".format(tuple(x.shape)))
    c = x.shape[1]
    k = torch.ones(kernel_size, kernel_size, dtype=x.dtype, device=x.device)
    k[kernel_size // 2, kernel_size // 2] = 1 - kernel_size ** 2
    return F.conv2d(x, k.view(1, 1, kernel_size, kernel_size).repeat(c, 1, 1, 1), padding=(kernel_size - 1) // 2, groups=c)


def freq_concat(x):
    """[x, LF, HF] along channels -- the Fusion-discriminator input (facades/network.png)."""
    return torch.cat([x, blur(x), laplacian(x)], 1)


# --------------------------------------------------------------------------------------
# SSIM loss (models/pytorch_ssim/__init__.py)
# --------------------------------------------------------------------------------------


def ssim_window(window_size=11, sigma=1.5, dtype=torch.float32):
    g = torch.tensor([math.exp(-(i - window_size // 2) ** 2 / float(2 * sigma ** 2)) for i in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).to(dtype)


def ssim(img1, img2, window_size=11):
    """pytorch_ssim._ssim (models/pytorch_ssim/__init__.py:17-37), size_average=True."""
    c = img1.shape[1]
    w = ssim_window(window_size, 1.5, img1.dtype).to(img1.device).expand(c, 1, window_size, window_size).contiguous()
    p = window_size // 2
    mu1 = F.conv2d(img1, w, padding=p, groups=c)
    mu2 = F.conv2d(img2, w, padding=p, groups=c)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = F.conv2d(img1 * img1, w, padding=p, groups=c) - mu1_sq
    s2 = F.conv2d(img2 * img2, w, padding=p, groups=c) - mu2_sq
    s12 = F.conv2d(img1 * img2, w, padding=p, groups=c) - mu12
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu12 + c1) * (2 * s12 + c2)) / ((mu1_sq + mu2_sq + c1) * (s1 + s2 + c2))
    return m.mean()


# --------------------------------------------------------------------------------------
# reconstructed training step
# --------------------------------------------------------------------------------------

DEFAULT_LOSS_WEIGHTS = dict(l1=1.0, ssim=0.0, perc=0.5, adv=0.01)


def _bce(p, target):
    return F.binary_cross_entropy(p, torch.full_like(p, target))


def d_param_names(d_sd):
    return [k for k, v in d_sd.items() if v.is_floating_point() and "running" not in k]


def train_step(g_sd, d_sd, v_sd, hazy, clean, state_g, state_d, weights=None, perc_layers=(1, 3), lr=2e-4,
               betas=(0.5, 0.999)):
    """One reconstructed FD-GAN iteration: D update on (clean, fake.detach()), then G update
    through blur/laplace, the UPDATED D (parameters constant) and the frozen VGG.  Mutates g_sd / d_sd.
    Returns (losses dict, D gradients, G gradients)."""
    wts = dict(DEFAULT_LOSS_WEIGHTS)
    if weights:
        wts.update(weights)
    g_names, d_names = fdgan_used_param_names(), d_param_names(d_sd)
    for k in g_names:
        g_sd[k].requires_grad_(True)
    for k in d_names:
        d_sd[k].requires_grad_(True)
    fake = fdgan_forward(g_sd, hazy, True, True)
    real_in = freq_concat(clean)
    fake_in_d = freq_concat(fake.detach())
    loss_d = _bce(d_forward(d_sd, real_in, True, True), 1.0) + _bce(d_forward(d_sd, fake_in_d, True, True), 0.0)
    grads_d = torch.autograd.grad(loss_d, [d_sd[k] for k in d_names])
    adam_step([d_sd[k] for k in d_names], grads_d, state_d, lr, betas)
    d_const = OrderedDict((k, v.detach()) for k, v in d_sd.items())
    l1 = F.l1_loss(fake, clean)
    loss_g = wts["l1"] * l1
    parts = {"l1_weighted": float(wts["l1"] * l1.detach())}
    if wts["ssim"] != 0.0:
        loss_g = loss_g + wts["ssim"] * (1 - ssim(fake, clean))
    perc = 0.0
    if wts["perc"] != 0.0 and perc_layers:
        v_const = OrderedDict((k, v.detach()) for k, v in v_sd.items())
        fv, cv = vgg16_forward(v_const, fake), vgg16_forward(v_const, clean)
        for k in perc_layers:
            perc = perc + F.mse_loss(fv[k], cv[k].detach())
        loss_g = loss_g + wts["perc"] * perc
    adv = _bce(d_forward(d_const, freq_concat(fake), True, True), 1.0)
    loss_g = loss_g + wts["adv"] * adv
    grads_g = torch.autograd.grad(loss_g, [g_sd[k] for k in g_names])
    adam_step([g_sd[k] for k in g_names], grads_g, state_g, lr, betas)
    parts.update(loss_d=float(loss_d.detach()), loss_g=float(loss_g.detach()),
                 perc_weighted=float(wts["perc"] * (perc.detach() if torch.is_tensor(perc) else perc)),
                 adv_weighted=float(wts["adv"] * adv.detach()))
    return parts, dict(zip(d_names, grads_d)), dict(zip(g_names, grads_g)), fake.detach()


def adam_step(params, grads, state, lr=2e-4, betas=(0.5, 0.999), eps=1e-8):
    """torch.optim.Adam arithmetic (flags --lrG/--lrD 2e-4, --beta1 0.5: demo.py:43-46)."""
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    b1, b2 = betas
    for i, (p, g) in enumerate(zip(params, grads)):
        m = state.setdefault(("m", i), torch.zeros_like(p))
        v = state.setdefault(("v", i), torch.zeros_like(p))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
        p.data.addcdiv_(m, denom, value=-lr / (1 - b1 ** t))
