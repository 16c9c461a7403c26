"""Numerics of the tensor-core operand scheme, emulated on the CPU (every fp32 operand is split hi = bf16(x),
lo = bf16(x - hi) and a product is hi*hi + hi*lo + lo*hi with fp32 accumulation).  The test states WHY the kernels pay three
tensor-core passes: on the contraction lengths of the path the scheme stays inside the 5e-5 unit-test tolerance (and the 1e-3
end-to-end bar of BASELINE.json with margin), while plain bf16 operands and a two-term split do not."""
import math

import pytest
import torch


def _split(x):
    hi = x.to(torch.bfloat16).float()
    lo = (x - hi).to(torch.bfloat16).float()
    return hi, lo


def _gemm_terms(a, b, terms):
    ah, al = _split(a)
    bh, bl = _split(b)
    ops = {"hh": (ah, bh), "hl": (ah, bl), "lh": (al, bh), "ll": (al, bl)}
    acc = torch.zeros(a.shape[0], b.shape[1], dtype=torch.float32)
    for t in terms:
        x, y = ops[t]
        acc = acc + x @ y            # fp32 accumulate, like the TMEM accumulator
    return acc


@pytest.mark.parametrize("K", [64, 1152, 9216])      # 1x1 64->128, dense-layer 3x3 (9 x 128), dense_block4.conv2 (9 x 1024)
def test_three_term_split_meets_the_tolerance_and_cheaper_schemes_do_not(K):
    g = torch.Generator().manual_seed(K)
    a = torch.rand(256, K, generator=g) * 2 - 1                       # activations of unit scale
    b = (torch.rand(K, 128, generator=g) * 2 - 1) / math.sqrt(K)      # weights scaled like the layers' initialisers
    want = a.double() @ b.double()
    err3 = float((_gemm_terms(a, b, ("hh", "hl", "lh")).double() - want).abs().max())
    err2 = float((_gemm_terms(a, b, ("hh", "hl")).double() - want).abs().max())
    err1 = float((_gemm_terms(a, b, ("hh",)).double() - want).abs().max())
    fp32 = float(((a @ b).double() - want).abs().max())
    assert err3 <= 1e-5, (K, err3)             # ~6e-6 per unit-scale layer: inside the 5e-5 tests/test_gpu_ops.py holds the kernels to
    assert fp32 < err3 < 32 * fp32, (K, err3, fp32)   # ~16 of fp32's 24 mantissa bits: an order of magnitude above fp32 rounding
    assert err1 > 1e-3 and err2 > 1e-3, (K, err1, err2)   # bf16 operands, or one dropped cross term: a single layer already spends
    assert err3 < err2 / 100                              #   the whole 1e-3 budget of BASELINE.json (the path stacks ~100 layers)


def test_split_is_exact_to_sixteen_mantissa_bits():
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(1 << 16, generator=g) * 2 - 1) * torch.logspace(-6, 3, 1 << 16)
    hi, lo = _split(x)
    rel = ((hi + lo).double() - x.double()).abs() / x.double().abs().clamp_min(1e-30)
    assert float(rel.max()) <= 2.0 ** -16                              # what the dropped lo*lo term and the second rounding cost
