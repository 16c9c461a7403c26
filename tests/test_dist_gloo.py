"""world_size-2 gloo test (CPU) of the multi-GPU host logic: batch sharding, flat gradient layout and the single
all-reduce per network.  The rank-averaged gradients of per-rank-BatchNorm replicas must equal the gradient of the
mean of the per-rank losses computed in one process."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from oracle import fdgan_oracle as O
from tests.util import seeded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_loss(d_sd, z):
    p = O.d_forward(d_sd, z, True, False)
    return torch.nn.functional.binary_cross_entropy(p, torch.ones_like(p))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import fdgan_b200
    from fdgan_b200 import dist as fdist
    from fdgan_b200.train import FlatState
    r, _lr, w = fdist.init_process_group("gloo")
    assert (r, w) == (rank, world) and fdist.world_size() == world
    torch.manual_seed(100 + rank)            # replicas start different; rank 0's parameters must win
    net = fdgan_b200.D(9, 8)
    st = FlatState(net)
    fdist.broadcast_flat_(st.flat, 0)
    names = [n for n, _p in net._used_named_parameters()]
    d_sd = {k: v.clone() for k, v in net.state_dict().items()}
    for n in names:
        d_sd[n].requires_grad_(True)
    zg = seeded((4, 9, 24, 24), 77, -1, 1)
    lo, hi = fdist.shard_range(4, rank, world)
    loss = _rank_loss(d_sd, zg[lo:hi])
    grads = torch.autograd.grad(loss, [d_sd[n] for n in names])
    for n, g in zip(names, grads):
        st.grad_views[n].copy_(g)
    fdist.allreduce_flat_(st.grad)
    st.grad.mul_(1.0 / world)
    torch.save({"flat": st.flat.clone(), "grad": {n: st.grad_views[n].clone() for n in names},
                "state": {k: v.clone() for k, v in net.state_dict().items()}}, os.path.join(out_dir, "r%d.pt" % rank))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_allreduce_equals_mean_of_rank_losses(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(tmp_path, "r0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "r1.pt"))
    assert torch.equal(r0["flat"], r1["flat"])                       # broadcast from rank 0
    names = list(r0["grad"].keys())
    d_sd = {k: v.clone() for k, v in r0["state"].items()}
    for n in names:
        d_sd[n].requires_grad_(True)
    zg = seeded((4, 9, 24, 24), 77, -1, 1)
    loss = 0.5 * (_rank_loss(d_sd, zg[0:2]) + _rank_loss(d_sd, zg[2:4]))   # per-rank BatchNorm statistics
    want = torch.autograd.grad(loss, [d_sd[n] for n in names])
    for n, g in zip(names, want):
        assert torch.allclose(r0["grad"][n], g, rtol=1e-4, atol=1e-6), n
        assert torch.equal(r0["grad"][n], r1["grad"][n]), n


def test_shard_range():
    from fdgan_b200.dist import shard_range
    assert [shard_range(64, r, 8) for r in (0, 7)] == [(0, 8), (56, 64)]
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)
