"""nn.DataParallel over TWO GPUs in one process (demo.py:89 as the reference runs on a multi-GPU box): replicas get their own
parameter copies, worker threads launch on their own device.  Skipped on a single-GPU box (there DataParallel calls the module
directly); the replica plumbing itself is covered on CPU by tests/test_dataparallel_replica.py.
Run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dataparallel.py` (round 2: passes on 2 x B200; profiles/r02_multigpu.md)."""
import pytest
import torch

from oracle import fdgan_oracle as O
from tests.util import maxabs, seeded

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_fdgan_under_dataparallel_two_gpus():
    import fdgan_b200
    net = fdgan_b200.FDGAN()
    net.load_state_dict(O.make_fdgan_state(0))
    x = seeded((4, 3, 32, 32), 5)
    dp = torch.nn.DataParallel(net, device_ids=[0, 1]).cuda()
    y = dp(x.cuda())                                           # scatter 2 + 2, replicate, gather on cuda:0
    # BatchNorm statistics are per replica: the expected output is the oracle on each half
    want = torch.cat([O.fdgan_forward(O.make_fdgan_state(0), x[:2], True, False), O.fdgan_forward(O.make_fdgan_state(0), x[2:], True, False)])
    assert y.device.index == 0 and maxabs(y, want) <= 2e-4
    (y * seeded((4, 3, 32, 32), 6, -1, 1).cuda()).sum().backward()   # gradients flow back through Broadcast to the original
    g = net.conv_refin3.weight.grad
    assert g is not None and g.device.index == 0 and bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0
