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
