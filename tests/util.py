"""Shared helpers for the parity tests (golden fixtures, seeded inputs, sampling)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAX_SAMPLES = 2048


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def seeded(shape, seed, lo=0.0, hi=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g) * (hi - lo) + lo


def sample(t):
    """Same reduction as oracle/make_golden.py:sample."""
    a = np.asarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float64).reshape(-1)
    stride = max(1, a.size // MAX_SAMPLES)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum())], a[::stride]]).astype(np.float64)


def assert_sample_close(got, want, rtol, atol, what=""):
    got = sample(got)
    scale = max(1.0, float(np.abs(want[2:]).max()))
    # element samples
    err = np.abs(got[2:] - want[2:]).max()
    assert err <= atol + rtol * scale, "%s: sampled max-abs %.3e (scale %.3e)" % (what, err, scale)
    # l2 norm (robust global check)
    assert abs(got[1] - want[1]) <= atol + rtol * max(1.0, want[1]), "%s: l2 %.6e vs %.6e" % (what, got[1], want[1])


def maxabs(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max())
