"""Checkpoint / resume of the training step's state (host logic, CPU): parameters stay views of the flat buffers after a load,
Adam state travels by parameter name, a mismatching checkpoint is rejected.  The kernels are not involved (GANTrainer.step needs
a GPU); tests/test_gpu_train.py covers the step itself."""
import io

import pytest
import torch


@pytest.fixture(scope="module")
def nets():
    import __graft_entry__ as ge
    ge.build()
    import fdgan_b200
    return fdgan_b200


def _trainer(pkg, seed):
    from fdgan_b200.train import GANTrainer
    torch.manual_seed(seed)
    G, D, V = pkg.FDGAN(), pkg.D(9, 36), pkg.Vgg16()
    return GANTrainer(G, D, V, weights=dict(ssim=0.1))


def test_trainer_state_round_trip(nets):
    a, b = _trainer(nets, 1), _trainer(nets, 2)
    g = torch.Generator().manual_seed(3)
    for st, n in ((a.sG, 7), (a.sD, 9)):
        st.exp_avg.copy_(torch.randn(st.n, generator=g))
        st.exp_avg_sq.copy_(torch.rand(st.n, generator=g))
        st.step = n
    buf = io.BytesIO()
    torch.save(a.state_dict(), buf)                      # the checkpoint is plain tensors / dicts: torch.save round trip
    buf.seek(0)
    sd = torch.load(buf, weights_only=False)
    assert not torch.equal(a.sG.flat, b.sG.flat)
    flat_ptr, first = b.sG.flat.data_ptr(), next(iter(b.G._used_named_parameters()))[1]
    b.load_state_dict(sd, strict_hyper=True)
    for x, y in ((a.sG, b.sG), (a.sD, b.sD)):
        assert torch.equal(x.flat, y.flat) and x.step == y.step
        sx, sy = x.state_dict(), y.state_dict()          # by name: the alignment padding between slices does not travel
        for key in ("exp_avg", "exp_avg_sq"):
            assert list(sx[key]) == list(sy[key]) and all(torch.equal(sx[key][n], sy[key][n]) for n in sx[key])
    # parameters are still views of the same flat buffer (flat gradients / fused Adam / a captured graph stay valid)
    assert b.sG.flat.data_ptr() == flat_ptr and first.data_ptr() == flat_ptr
    for (k1, v1), (k2, v2) in zip(a.G.state_dict().items(), b.G.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    # the network part is a reference-keyed state dict: loads into a fresh module, with or without the DataParallel prefix
    fresh = nets.FDGAN()
    fresh.load_state_dict({"module." + k: v for k, v in sd["netG"].items()})
    assert torch.equal(fresh.conv_refin1.weight, a.G.conv_refin1.weight)


def test_trainer_state_rejects_mismatch(nets):
    a, b = _trainer(nets, 1), _trainer(nets, 2)
    sd = a.state_dict()
    bad = dict(sd)
    bad["version"] = 7
    with pytest.raises(ValueError):
        b.load_state_dict(bad)
    bad = dict(sd)
    bad["optG"] = dict(sd["optG"])
    bad["optG"]["exp_avg"] = dict(list(sd["optG"]["exp_avg"].items())[1:])
    with pytest.raises(KeyError):
        b.load_state_dict(bad)
    bad = dict(sd)
    bad["hyper"] = dict(sd["hyper"], lr=1e-3)
    with pytest.raises(ValueError):
        b.load_state_dict(bad, strict_hyper=True)
    b.load_state_dict(bad)                               # hyper-parameters are advisory unless strict_hyper
    name, t = next(iter(sd["optD"]["exp_avg_sq"].items()))
    bad = dict(sd)
    bad["optD"] = dict(sd["optD"], exp_avg_sq=dict(sd["optD"]["exp_avg_sq"]))
    bad["optD"]["exp_avg_sq"][name] = t.reshape(-1)[:-1]
    with pytest.raises(ValueError):
        b.load_state_dict(bad)


def test_fused_adam_drops_cached_images_of_its_parameters(nets):
    """The flat Adam kernel writes parameters through raw pointers (no ``_version`` bump): operand images cached for a phase in
    which such a parameter was frozen must not survive the update.  Host bookkeeping only."""
    from fdgan_b200 import ops
    ops._FROZEN.clear()
    ops._FROZEN[(1000, (4,), 0)] = ("inside",)
    ops._FROZEN[(1012, (4,), 1)] = ("inside",)
    ops._FROZEN[(2000, (4,), 0)] = ("other network",)
    assert ops.drop_frozen_in_range(1000, 1016) == 2
    assert list(ops._FROZEN) == [(2000, (4,), 0)]
    ops._FROZEN.clear()
