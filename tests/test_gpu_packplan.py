"""Pack plans (ops.PackPlan / fdg_pack_batch): all weight-operand repacks of a network pass in one launch per level must give
bit-identical results to the per-layer repack launches, follow parameter updates, and survive re-homed parameters."""
import pytest
import torch

from oracle import fdgan_oracle as O
from tests.util import seeded

pytestmark = pytest.mark.gpu


def _run(net, x, r):
    xd = x.clone().requires_grad_(True)
    y = net(xd)
    (y * r).sum().backward()
    g = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    for p in net.parameters():
        p.grad = None
    return y.detach().clone(), xd.grad.clone(), g


@pytest.mark.parametrize("which", ["FDGAN", "D"])
def test_pack_plan_is_bit_identical_and_tracks_updates(which):
    import fdgan_b200
    from fdgan_b200 import _lib as L, ops
    if which == "FDGAN":
        mk = lambda: fdgan_b200.FDGAN()
        sd = O.make_fdgan_state(0)
        x, r = seeded((2, 3, 32, 32), 5).cuda(), seeded((2, 3, 32, 32), 6, -1, 1).cuda()
    else:
        mk = lambda: fdgan_b200.D(9, 36)
        sd = O.make_d_state(9, 36, 1)
        x, r = seeded((2, 9, 32, 32), 5).cuda(), seeded((2, 1, 14, 14), 6, -1, 1).cuda()
    a, b = mk(), mk()
    a.load_state_dict(sd); b.load_state_dict(sd)
    a, b = a.cuda().train(), b.cuda().train()
    launches = {}
    for it in range(3):
        outs = {}
        for name, net, plan in (("plan", a, True), ("single", b, False)):
            ops.USE_PACK_PLAN = plan
            try:
                n0 = L.launch_count()
                outs[name] = _run(net, x, r)
                launches[(name, it)] = L.launch_count() - n0
            finally:
                ops.USE_PACK_PLAN = True
        ya, dxa, ga = outs["plan"]
        yb, dxb, gb = outs["single"]
        assert torch.equal(ya, yb) and torch.equal(dxa, dxb), it      # forward + data gradient: no atomics on these paths
        for k in gb:
            # weight gradients accumulate split-K partials with atomics: identical up to summation order
            assert float((ga[k] - gb[k]).abs().max()) <= 1e-5 * max(1.0, float(gb[k].abs().max())), (it, k)
        with torch.no_grad():      # an "optimiser step": the plan must pick up the new values on the next pass
            for net in (a, b):
                for p in net.parameters():
                    p.mul_(1.0 + 0.01 * (it + 1))
    assert launches[("plan", 0)] == launches[("single", 0)]                      # first pass records (per-layer launches)
    assert launches[("plan", 1)] < launches[("single", 1)] - (100 if which == "FDGAN" else 8)
    print(which, "launches per fwd+bwd: plan", launches[("plan", 1)], "single", launches[("single", 1)])


def test_pack_plan_survives_rehomed_parameters_and_path_switch():
    import fdgan_b200
    from fdgan_b200 import ops
    from fdgan_b200.train import FlatState
    net = fdgan_b200.D(9, 36)
    net.load_state_dict(O.make_d_state(9, 36, 1))
    net = net.cuda().train()
    x, r = seeded((2, 9, 32, 32), 5).cuda(), seeded((2, 1, 14, 14), 6, -1, 1).cuda()
    y0, dx0, _ = _run(net, x, r)
    _run(net, x, r)                                   # plan ready
    FlatState(net)                                    # parameters move into one flat buffer: every plan key misses
    y1, dx1, _ = _run(net, x, r)                      # served by single launches, plan re-records afterwards
    y2, dx2, _ = _run(net, x, r)
    y3, dx3, _ = _run(net, x, r)
    assert torch.equal(y0, y1) and torch.equal(y0, y2) and torch.equal(y0, y3) and torch.equal(dx0, dx3)
    old = ops.USE_UMMA
    try:
        ops.USE_UMMA = False                          # other conv path -> other plan
        ys, _, _ = _run(net, x, r)
    finally:
        ops.USE_UMMA = old
    assert float((ys - y0).abs().max()) <= 2e-4
    y4, _, _ = _run(net, x, r)
    assert torch.equal(y0, y4)
