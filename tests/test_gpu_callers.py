"""-m gpu: the remaining SURVEY.md §8(f)-3 callers on the device path, against the reference's own outputs
(tests/golden/callers.npz, written by oracle/gen_callers.py): metrics.hadamard_sum, tools.shift_mode, TTMatrix."""
import os

import numpy as np
import pytest
import torch

from oracle import cases

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "callers.npz"))


def _tensor(cores, dtype=torch.float64):
    import tntorch_b200 as tnb

    return tnb.Tensor([torch.as_tensor(c).to(dtype).cuda() for c in cores])


@pytest.mark.parametrize("name", list(cases.HADAMARD_SUM_CASES))
def test_hadamard_sum_matches_reference(name):
    import tntorch_b200 as tnb

    ts = [_tensor(c) for c in cases.hadamard_operands(cases.HADAMARD_SUM_CASES[name])]
    ref = float(GOLD[f"{name}/exact"])
    assert abs(tnb.hadamard_sum(ts) - ref) <= 1e-11 * abs(ref)                      # default algorithm = 'exact'
    for alg in ("svd", "eig"):
        got = tnb.hadamard_sum(ts, algorithm=alg, eps=1e-8)
        assert abs(got - ref) <= 1e-5 * abs(ref)                                   # north-star tolerance
        assert abs(got - float(GOLD[f"{name}/svd"])) <= 1e-5 * abs(ref)             # the reference's approximate value
    # a loose eps still lands within eps * ||product|| * sqrt(numel) (Cauchy-Schwarz on the rounding error)
    if len(ts) == 2:
        scale = float(np.sqrt(np.prod(ts[0].shape))) * float((ts[0] * ts[1]).torch().norm())
        got = tnb.hadamard_sum(ts, algorithm="svd", eps=1e-3)
        assert abs(got - ref) <= 1e-3 * scale


def test_hadamard_sum_rejects_batches_and_shape_mismatch():
    import tntorch_b200 as tnb

    a = _tensor(cases.random_tt((4, 4, 4), 2, 1))
    b = _tensor(cases.random_tt((4, 5, 4), 2, 2))
    with pytest.raises(AssertionError):
        tnb.hadamard_sum([a, b])
    assert abs(tnb.hadamard_sum([a]) - float(a.torch().sum())) < 1e-10
    assert abs(float(tnb.dot(a, a)) - float((a.torch() ** 2).sum())) < 1e-9 * float((a.torch() ** 2).sum())


@pytest.mark.parametrize("name", list(cases.SHIFT_MODE_CASES))
def test_shift_mode_matches_reference(name):
    import tntorch_b200 as tnb

    spec, n, shift, eps = cases.SHIFT_MODE_CASES[name]
    cores = cases.shift_mode_input(spec)
    t = _tensor(cores)
    out = tnb.shift_mode(t, n, shift, eps=eps)
    assert out is t                                                                # in place, like the reference
    ref = GOLD[f"{name}/full"]
    got = t.torch().cpu().numpy()
    assert got.shape == ref.shape
    assert np.linalg.norm(got - ref) <= 1e-5 * np.linalg.norm(ref)
    assert t.ranks_tt.tolist() == GOLD[f"{name}/ranks"].tolist()
    if not isinstance(eps, str):  # lossless up to eps: the dense tensor is the input with the mode moved
        dense = np.moveaxis(cases.tt_full(cores), n, n + shift)
        assert np.linalg.norm(got - dense) <= 10 * eps * np.linalg.norm(dense)


def test_shift_mode_zero_and_tucker_factors():
    import tntorch_b200 as tnb

    t = _tensor(cases.random_tt((5, 6, 7), 3, 5))
    before = t.torch().clone()
    assert tnb.shift_mode(t, 1, 0) is t
    assert torch.equal(t.torch(), before)
    t.round_tucker(rmax=4)
    dense = t.torch().clone()
    tnb.shift_mode(t, 2, -2, eps=1e-10)
    assert all(U is None for U in t.Us)
    assert float(torch.dist(t.torch(), dense.permute(2, 0, 1))) <= 1e-8 * float(dense.norm())
    with pytest.raises(ValueError):
        tnb.shift_mode(t, 0, 1, eps=-1.0)


@pytest.mark.parametrize("name", list(cases.TTMATRIX_CASES))
def test_ttmatrix_matches_reference(name):
    import tntorch_b200 as tnb

    spec = cases.TTMATRIX_CASES[name]
    M = torch.as_tensor(cases.ttmatrix_input(spec)).cuda()
    ttm = tnb.TTMatrix(M, ranks=spec["ranks"], input_dims=spec["input_dims"], output_dims=spec["output_dims"])
    assert ttm.batch == bool(spec["batch"])
    assert [list(c.shape)[-4:] for c in ttm.cores] == GOLD[f"{name}/core_shapes"].tolist()
    assert [int(r) for r in ttm.ranks] == GOLD[f"{name}/ranks"].tolist()
    ref = GOLD[f"{name}/full"]
    got = ttm.torch().cpu().numpy()
    assert got.shape == ref.shape
    assert np.linalg.norm(got - ref) <= 1e-5 * np.linalg.norm(ref)
    # explicit cores round-trip (matrix.py:47-57)
    again = tnb.TTMatrix(ttm.cores, ranks=spec["ranks"], input_dims=spec["input_dims"], output_dims=spec["output_dims"])
    assert torch.equal(again.torch(), ttm.torch())


def test_ttmatrix_full_rank_is_exact():
    import tntorch_b200 as tnb

    M = torch.randn(24, 30, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    ttm = tnb.TTMatrix(M, ranks=[64], input_dims=[4, 6], output_dims=[5, 6])
    assert float(torch.dist(ttm.torch(), M)) <= 1e-8 * float(M.norm())
