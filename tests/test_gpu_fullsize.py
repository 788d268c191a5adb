"""-m gpu: BASELINE.json's full-size workload through size-independent properties (the oracle cannot run 2^30
elements in test time), plus mirrors of the reference's own self-consistency tests for the path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_full_size_64x5_properties():
    """randn(64^5) fp32 -> TT-rank 32 (the bench workload): ranks, gauge, norm bookkeeping, error identity,
    idempotence of a second rounding."""
    import tntorch_b200 as tnb
    from tntorch_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(7)
    X = torch.randn((64,) * 5, generator=g, device="cuda")
    cores, info = ops.ttsvd(X, rmax=32, return_info=True)
    assert [c.shape[0] for c in cores] + [1] == [1, 32, 32, 32, 32, 1]
    assert info["tc_grams"] >= 2 and info["fused_filters"] >= 1  # the tensor-core path ran, not a fallback
    # gauge: cores 1.. have orthonormal right unfoldings
    for c in cores[1:]:
        M = c.reshape(c.shape[0], -1).double()
        assert (M @ M.T - torch.eye(M.shape[0], device="cuda", dtype=torch.float64)).abs().max().item() < 2e-5
    # the first Gram's trace is ||X||^2 — of the TF32-truncated operands: (1 - c) ||X||^2 with c ~ 7e-4 .. 1.1e-3
    # (tests/test_model.py::test_tf32_gram_is_a_uniform_scaling); the rank rule only uses ratios of it
    xn = float(torch.linalg.vector_norm(X.double()))
    assert 0.0 < 1.0 - info["norm"] / xn < 1e-3
    # orthogonal projection: ||X - T||^2 = ||X||^2 - ||T||^2, and ||T|| = ||core 0|| in this gauge
    err = ops.tt_relative_error(X, cores)
    nt = float(torch.linalg.vector_norm(cores[0].double()))
    assert abs(err - np.sqrt(max(0.0, 1.0 - (nt / xn) ** 2))) < 1e-5
    assert 0.999 < err < 1.0  # a random Gaussian tensor is incompressible
    # rounding the result again at the same rank changes nothing
    t = tnb.Tensor(cores)
    t2 = tnb.round_tt(t, rmax=32)
    assert list(t2.ranks_tt) == list(t.ranks_tt)
    assert tnb.relative_error(t.torch(), t2) < 1e-5
    del X


def test_reference_test_round_tt_doubling():
    """tests/test_round.py:41-59 of the reference: t = gt + gt (block-diagonal cores), round_tt(1e-8); the ranks come
    back and relative_error(gt, t/2) <= 1e-7 ('eig' bar), on random 8-mode fp64 TTs."""
    import tntorch_b200 as tnb

    rng = np.random.default_rng(0)
    for trial in range(6):
        N = 8
        shape = rng.integers(2, 6, N)
        ranks = [1] + list(rng.integers(1, 5, N - 1)) + [1]
        gt = [torch.as_tensor(rng.standard_normal((ranks[k], shape[k], ranks[k + 1]))).cuda() for k in range(N)]
        doubled = []
        for k, c in enumerate(gt):  # the TT of gt + gt: block-diagonal interior cores, stacked end cores
            if k == 0:
                doubled.append(torch.cat([c, c], dim=2))
            elif k == N - 1:
                doubled.append(torch.cat([c, c], dim=0))
            else:
                z = torch.zeros_like(c)
                doubled.append(torch.cat([torch.cat([c, z], dim=2), torch.cat([z, c], dim=2)], dim=0))
        t = tnb.Tensor(doubled)
        t.round_tt(eps=1e-8)
        assert max(t.ranks_tt) <= max(ranks)
        full_gt = tnb.Tensor(gt).torch()
        assert float(torch.linalg.vector_norm(full_gt - t.torch() / 2) / torch.linalg.vector_norm(full_gt)) <= 1e-7


def test_reference_test_init_reproduces_dense():
    """tests/test_init.py:7-13: Tensor(gt) with no rank argument is the exact TT: .torch() reproduces gt to 1e-7."""
    import tntorch_b200 as tnb

    rng = np.random.default_rng(1)
    for trial in range(8):
        shape = tuple(rng.integers(1, 6, rng.integers(1, 6)))
        gt = torch.as_tensor(rng.standard_normal(shape)).cuda()
        t = tnb.Tensor(gt)
        assert float((t.torch() - gt).abs().max()) <= 1e-7 * max(1.0, float(gt.abs().max()))


def test_reference_test_truncated_svd_batch():
    """tests/test_round.py:21-38: batched truncated_svd equals the per-sample calls."""
    import tntorch_b200 as tnb

    g = torch.Generator().manual_seed(5)
    M = torch.rand(2, 32, 32, generator=g, dtype=torch.float64).cuda()
    for alg in ("svd", "eig"):
        Lb, Rb = tnb.truncated_svd(M, rmax=8, batch=True, algorithm=alg)
        for b in range(2):
            L1, R1 = tnb.truncated_svd(M[b], rmax=8, algorithm=alg)
            assert torch.allclose(Lb[b] @ Rb[b], L1 @ R1, atol=1e-9)
