"""-m gpu: the callers that add / multiply and re-round in loops: Tensor.__add__/__sub__/__mul__ block /
Kronecker cores on the device, the fused sum+round node, tn.reduce (tools.py:460-512)."""
import operator

import numpy as np
import pytest
import torch

from oracle import cases

pytestmark = pytest.mark.gpu


def _tt(shape, rank, seed, dtype=torch.float64):
    return [torch.as_tensor(c).to(dtype).cuda() for c in cases.random_tt(shape, rank, seed)]


def test_add_sub_mul_match_dense_arithmetic():
    import tntorch_b200 as tnb

    shape = (6, 5, 7, 4)
    a, b = tnb.Tensor(_tt(shape, 3, 1)), tnb.Tensor(_tt(shape, [2, 4, 2], 2))
    A, Bd = a.torch(), b.torch()
    s = a + b
    assert list(s.ranks_tt) == [1, 5, 7, 5, 1]  # block cores: ranks add (tensor.py:445-520)
    assert float((s.torch() - (A + Bd)).abs().max()) < 1e-12 * float((A + Bd).abs().max())
    assert float(((a - b).torch() - (A - Bd)).abs().max()) < 1e-12 * float(A.abs().max() + Bd.abs().max())
    assert float(((a * 2.5).torch() - 2.5 * A).abs().max()) < 1e-12 * float(A.abs().max())
    assert float(((a + 1.5).torch() - (A + 1.5)).abs().max()) < 1e-12 * float(A.abs().max() + 1.5)
    assert float(((-a).torch() + A).abs().max()) == 0.0
    h = a * b
    assert list(h.ranks_tt) == [1, 6, 12, 6, 1]  # Kronecker cores: ranks multiply
    assert float((h.torch() - A * Bd).abs().max()) < 1e-12 * float((A * Bd).abs().max())
    # one mode
    v, w = tnb.Tensor(_tt((9,), 1, 3)), tnb.Tensor(_tt((9,), 1, 4))
    assert float(((v + w).torch() - (v.torch() + w.torch())).abs().max()) < 1e-14
    # rounding the sum of a tensor with itself gives the ranks back (reference tests/test_round.py:41-59)
    d = a + a
    d.round_tt(eps=1e-10)
    assert max(d.ranks_tt) <= 3
    assert float((d.torch() - 2 * A).abs().max()) < 1e-9 * float(A.abs().max())


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_reduce_add_is_the_fused_sum_round(dtype):
    """tn.reduce(ts, operator.add, eps, rmax): same result as summing densely; every tree node is one tnb_tt_sum_round."""
    import tntorch_b200 as tnb

    shape = (8, 7, 6, 5, 6)
    base = [tnb.Tensor(_tt(shape, 2, 10 + k, dtype)) for k in range(3)]
    coef = np.random.default_rng(0).standard_normal((11, 3))
    ts = [base[0] * float(c[0]) + base[1] * float(c[1]) + base[2] * float(c[2]) for c in coef]  # rank 6 each, true rank <= 6
    dense = sum(t.torch().double() for t in ts)
    tol = 1e-9 if dtype == torch.float64 else 2e-4
    r = tnb.reduce(ts, operator.add, eps=1e-10 if dtype == torch.float64 else 1e-5)
    assert max(r.ranks_tt) <= 6
    assert float(torch.linalg.vector_norm(r.torch().double() - dense) / torch.linalg.vector_norm(dense)) < tol
    r2 = tnb.reduce(ts, operator.add, rmax=4)  # bounded ranks
    assert max(r2.ranks_tt) <= 4
    # a generic function goes through function(a, b) + round
    r3 = tnb.reduce(ts[:4], lambda x, y: x + y, eps=1e-10 if dtype == torch.float64 else 1e-5)
    d4 = sum(t.torch().double() for t in ts[:4])
    assert float(torch.linalg.vector_norm(r3.torch().double() - d4) / torch.linalg.vector_norm(d4)) < tol


def test_fused_sum_round_equals_add_then_round():
    from tntorch_b200 import ops

    shape = (10, 9, 8, 7)
    a, b, c = _tt(shape, 4, 20), _tt(shape, 3, 21), _tt(shape, 2, 22)
    fused = ops.tt_sum_round([a, b, c], alpha=[1.0, -0.5, 2.0], eps=1e-12)
    plain = ops.tt_round(ops.tt_sum([a, b, c], alpha=[1.0, -0.5, 2.0]), eps=1e-12)
    assert [x.shape for x in fused] == [x.shape for x in plain]
    dense = cases.tt_full([x.cpu().numpy() for x in a]) - 0.5 * cases.tt_full([x.cpu().numpy() for x in b]) \
        + 2.0 * cases.tt_full([x.cpu().numpy() for x in c])
    got = cases.tt_full([x.cpu().numpy() for x in fused])
    assert np.abs(got - dense).max() < 1e-10 * np.abs(dense).max()
