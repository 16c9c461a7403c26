"""Engine-level gradient parity with an fp64 arbiter.

Per-parameter gradients of FDGAN at B >= 2 are chaotic in the REFERENCE's own arithmetic (ReLU masks flip on near-zero
pre-activations), which is why tests/util.py:grad_close holds single parameters to a few per cent of relative L2.  These tests
justify that bar by measurement and add checks a backward WIRING error (a dropped skip branch, a missing 0.25 pool factor, a
missing deferred BatchNorm affine term) cannot hide under:

  * the fp64 oracle is the arbiter: the GPU's error against it must stay within a small multiple of the error the reference's own
    fp32 CPU arithmetic (the fp32 oracle) has against it, per parameter group and over the whole gradient vector;
  * on a larger input (B=4, 64x64: ~16k pixels per BatchNorm at full resolution, so single mask flips average out) EVERY
    parameter gradient of the exact-fp32 SIMT path is held to 3e-2 and of the tcgen05 path to 5e-2 relative L2;
  * a directional derivative: <grad, v> from the backward pass against a central finite difference of the fp64 oracle's loss along
    the same random direction v (a scalar that integrates over all masks).
"""
import pytest
import torch

from oracle import fdgan_oracle as O
from tests.util import seeded

pytestmark = pytest.mark.gpu


def _gpu_grads(x, r, seed, umma):
    import fdgan_b200
    from fdgan_b200 import ops
    old = ops.USE_UMMA
    ops.USE_UMMA = umma
    try:
        net = fdgan_b200.FDGAN()
        net.load_state_dict(O.make_fdgan_state(seed))
        net = net.cuda().train()
        xd = x.detach().clone().cuda().requires_grad_(True)
        y = net(xd)
        (y * r.detach().cuda()).sum().backward()
        torch.cuda.synchronize()
        return {k: p.grad.detach().double().cpu() for k, p in net.named_parameters() if p.grad is not None}, xd.grad.double().cpu()
    finally:
        ops.USE_UMMA = old


def _oracle_grads(x, r, seed, dtype):
    sd = type(O.make_fdgan_state(seed))((k, (v.to(dtype) if v.is_floating_point() else v.clone())) for k, v in O.make_fdgan_state(seed).items())
    names = O.fdgan_used_param_names()
    for k in names:
        sd[k].requires_grad_(True)
    xo = x.detach().clone().to(dtype).requires_grad_(True)
    y = O.fdgan_forward(sd, xo, True, False)
    (y * r.to(dtype)).sum().backward()
    return {k: sd[k].grad.double() for k in names}, xo.grad.double()


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _vec_rel(ga, gb, keys):
    num = sum(float((ga[k] - gb[k]).square().sum()) for k in keys)
    den = sum(float(gb[k].square().sum()) for k in keys)
    return (num / den) ** 0.5


ZERO = ("conv_refine4.bias",)      # analytically zero gradient (a bias feeding only BatchNorms)


@pytest.mark.parametrize("umma", [False, True], ids=["simt_fp32", "tcgen05_bf16x3"])
def test_gradient_error_within_fp32_reference_spread(umma):
    """fp64 arbiter at B=2, 64x48 (the shape of test_fdgan_forward_backward_vs_oracle)."""
    shape = (2, 3, 64, 48)
    x, r = seeded(shape, 11), seeded(shape, 12, -1.0, 1.0)
    g64, dx64 = _oracle_grads(x, r, 3, torch.float64)
    g32, dx32 = _oracle_grads(x, r, 3, torch.float32)
    gg, dxg = _gpu_grads(x, r, 3, umma)
    keys = [k for k in g64 if k not in ZERO]
    assert set(gg) == set(g64)
    e_ref = {k: _rel(g32[k], g64[k]) for k in keys}
    e_gpu = {k: _rel(gg[k], g64[k]) for k in keys}
    ref_all, gpu_all = _vec_rel(g32, g64, keys), _vec_rel(gg, g64, keys)
    ref_rms = (sum(v * v for v in e_ref.values()) / len(keys)) ** 0.5
    gpu_rms = (sum(v * v for v in e_gpu.values()) / len(keys)) ** 0.5
    print("fp64 arbiter (%s): whole-vector rel-L2 gpu %.3e / fp32-reference %.3e; per-parameter rms gpu %.3e / ref %.3e; max gpu %.3e / ref %.3e; dx gpu %.3e / ref %.3e"
          % ("tcgen05" if umma else "simt", gpu_all, ref_all, gpu_rms, ref_rms, max(e_gpu.values()), max(e_ref.values()), _rel(dxg, dx64), _rel(dx32, dx64)))
    # The GPU's distance from the exact gradient is a small multiple K of the distance the reference's own fp32 arithmetic has.
    # Measured on a B200 (round 2, four runs): fp32 SIMT path 3.5e-3 / 6.3e-3 / 6.8e-3 / 1.04e-2 against the reference's 2.3e-3 (1.6x .. 4.6x:
    # the order of the fp64 statistics atomics changes last bits of BatchNorm scale / shift from run to run, and at this size a last-bit
    # change of the forward flips enough ReLU masks to move the gradient by up to 1 % -- the fp32 reference's 2.3e-3 is one draw from
    # that same distribution); tcgen05 bf16x3 path 1.56e-2 (6.9x: its 5e-5 forward error, fp32: 4e-6, flips ~10x more masks).
    K = 10.0
    assert gpu_all <= K * ref_all + 2e-3, (gpu_all, ref_all)
    assert gpu_rms <= K * ref_rms + 2e-3, (gpu_rms, ref_rms)
    assert max(e_gpu.values()) <= K * max(e_ref.values()) + 5e-3
    assert _rel(dxg, dx64) <= K * _rel(dx32, dx64) + 2e-3


@pytest.mark.parametrize("umma,bar", [(False, 3e-2), (True, 5e-2)], ids=["simt_fp32", "tcgen05_bf16x3"])      # measured 7.3e-3 / 2.1e-2 (a wiring error is O(1))
def test_every_parameter_gradient_tight_on_larger_input(umma, bar):
    """B=4, 64x64: every one of the 361 used parameters individually, against the fp64 oracle."""
    shape = (4, 3, 64, 64)
    x, r = seeded(shape, 21), seeded(shape, 22, -1.0, 1.0)
    g64, dx64 = _oracle_grads(x, r, 0, torch.float64)
    gg, dxg = _gpu_grads(x, r, 0, umma)
    worst, worst_k = 0.0, None
    for k in g64:
        if k in ZERO:
            assert float(gg[k].abs().max()) <= 1e-3
            continue
        e = _rel(gg[k], g64[k])
        if e > worst:
            worst, worst_k = e, k
    print("B=4 64x64 (%s): worst per-parameter rel-L2 vs fp64 %.3e (%s); dx %.3e" % ("tcgen05" if umma else "simt", worst, worst_k, _rel(dxg, dx64)))
    assert worst <= bar, (worst, worst_k)
    assert _rel(dxg, dx64) <= bar


def test_directional_derivative_matches_finite_difference():
    """<dL/dtheta, v> from the GPU backward vs (L(theta + h v) - L(theta - h v)) / 2h of the fp64 oracle, v a random direction over ALL
    used parameters (scaled per parameter to its own magnitude).  A dropped or mis-scaled branch changes this scalar by O(1)."""
    shape = (2, 3, 32, 32)
    x, r = seeded(shape, 31), seeded(shape, 32, -1.0, 1.0)
    names = O.fdgan_used_param_names()
    base = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in O.make_fdgan_state(0).items()}
    g = torch.Generator().manual_seed(77)
    v = {k: torch.randn(base[k].shape, generator=g, dtype=torch.float64) * base[k].abs().mean().clamp_min(1e-3) for k in names}

    def loss(h):
        sd = type(O.make_fdgan_state(0))((k, (base[k] + h * v[k]) if k in v else base[k].clone()) for k in base)
        with torch.no_grad():
            return float((O.fdgan_forward(sd, x.double(), True, False) * r.double()).sum())

    h = 1e-6      # fp64: the loss is piecewise smooth (ReLU kinks), so the step must be small; measured 0.2 % from the analytic value
    fd = (loss(h) - loss(-h)) / (2 * h)
    fd2 = (loss(2 * h) - loss(-2 * h)) / (4 * h)
    for umma in (False, True):
        gg, _ = _gpu_grads(x, r, 0, umma)
        dd = sum(float((gg[k] * v[k]).sum()) for k in names)
        print("directional derivative (%s): backward %.6e, finite difference %.6e (h) / %.6e (2h)" % ("tcgen05" if umma else "simt", dd, fd, fd2))
        # measured: fp64 analytic -101.65, finite difference -101.86, SIMT backward -102.49 (0.6 %), tcgen05 backward -105.18 (3.3 %: the
        # whole-vector gradient error of that path, 1.5 %, projected on one random direction)
        assert abs(dd - fd) <= (6e-2 if umma else 2e-2) * abs(fd) + 10 * abs(fd - fd2), (dd, fd, fd2)
