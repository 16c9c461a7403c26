"""nn.DataParallel with several GPUs (demo.py:89) runs the module through torch.nn.parallel.replicate: replicas get a COPY of
the original's __dict__, an empty _parameters and this device's parameter copies as plain attributes listed in
_former_parameters.  The host glue must hand the autograd node the replica's own tensors, in the original's order, and must
not reuse the original's cached list or its flat gradient sink.  CPU: the replica is built the way replicate() builds it."""
from collections import OrderedDict

import pytest
import torch


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as ge
    ge.build()
    import fdgan_b200
    return fdgan_b200


def _replicate_like_torch(net):
    """torch/nn/parallel/replicate.py, single replica, parameter copies = 2 * p (non-leaf, like Broadcast's outputs)."""
    modules = list(net.modules())
    index = {m: i for i, m in enumerate(modules)}
    copies = []
    for m in modules:
        r = m._replicate_for_data_parallel()
        r._former_parameters = OrderedDict()
        copies.append(r)
    for i, m in enumerate(modules):
        for key, child in m._modules.items():
            setattr(copies[i], key, None if child is None else copies[index[child]])
        for key, p in m._parameters.items():
            if p is not None:
                c = p * 2.0
                setattr(copies[i], key, c)
                copies[i]._former_parameters[key] = c
        for key, b in m._buffers.items():
            if b is not None:
                setattr(copies[i], key, b.clone())
    return copies[0]


@pytest.mark.parametrize("which", ["FDGAN", "D", "Vgg16"])
def test_replica_uses_its_own_parameters(pkg, which):
    net = {"FDGAN": pkg.FDGAN, "D": lambda: pkg.D(9, 36), "Vgg16": pkg.Vgg16}[which]()
    named = net._used_named_parameters()                     # fills the cache that replicate() will copy into the replica
    net.set_grad_sink({n: torch.zeros_like(p) for n, p in named})
    rep = _replicate_like_torch(net)
    assert rep._is_replica and list(rep.parameters()) == []
    rnamed = rep._used_named_parameters()
    assert [n for n, _t in rnamed] == [n for n, _p in named]
    for (n, t), (_n, p) in zip(rnamed, named):
        assert t is not p and not isinstance(t, torch.nn.Parameter) and torch.equal(t, 2.0 * p.detach()), n
        mod, attr = n.rsplit(".", 1)
        assert rep.get_submodule(mod).__dict__[attr] is t     # the tensor the executors read through the attribute path
    assert net._used_named_parameters() is named             # the original keeps its cached list


def test_replica_shadows_misaligned_tensors_with_aligned_copies(pkg):
    """Replicas on the other devices get views into one coalesced broadcast buffer (arbitrary 4-byte offsets); the executors must see
    16-byte aligned tensors of the same values while the autograd inputs stay the broadcast outputs."""
    net = pkg.D(9, 36)
    rep = _replicate_like_torch(net)
    flat = torch.zeros(sum(t.numel() for _n, t in rep._used_named_parameters()) + 64)
    off = 1                                              # start misaligned, as after a 3-float bias in the real buffer
    originals = {}
    for m in rep.modules():
        for k, v in list(m._former_parameters.items()):
            view = flat[off:off + v.numel()].view(v.shape)
            view.copy_(v)
            off += v.numel()
            m._former_parameters[k] = view
            m.__dict__[k] = view
            originals[(id(m), k)] = view
    assert any(v.data_ptr() % 16 for v in originals.values())
    rep._align_replica_tensors()
    for m in rep.modules():
        for k, v in m._former_parameters.items():
            seen = m.__dict__[k]
            assert seen.data_ptr() % 16 == 0 and torch.equal(seen, v)
            assert v is originals[(id(m), k)]            # the autograd inputs are untouched
    named = rep._used_named_parameters()
    assert all(t is originals[(id(rep.get_submodule(n.rsplit(".", 1)[0])), n.rsplit(".", 1)[1])] for n, t in named)
