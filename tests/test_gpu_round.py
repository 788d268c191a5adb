"""-m gpu: TT rounding of TT input and truncated_svd through the drop-in API, against golden vectors."""
import os

import numpy as np
import pytest
import torch

from gpu_util import ranks_of, relerr64
from oracle import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", list(cases.ROUND_CASES))
def test_round_tt_matches_reference(name):
    import tntorch_b200 as tnb

    g = np.load(os.path.join(GOLD, "round_tt.npz"))
    spec = cases.ROUND_CASES[name]
    cores = cases.make_tt(spec)
    dense = cases.tt_full(cores)
    t = tnb.Tensor([torch.as_tensor(c).cuda() for c in cores])
    kw = {k: spec[k] for k in ("eps", "rmax") if k in spec}
    t2 = tnb.round_tt(t, **kw)
    assert ranks_of(t2.cores) == list(g[f"{name}/svd/ranks"])
    tol = 1e-5
    assert abs(relerr64(dense, t2.cores) - float(g[f"{name}/svd/relerr"])) <= tol
    # round_tt(t) must not touch t (round.py:7-19 clones first)
    assert ranks_of(t.cores) == [1] + [c.shape[2] for c in cores]


@pytest.mark.parametrize("name", list(cases.TSVD_CASES))
@pytest.mark.parametrize("lo", [True, False])
def test_truncated_svd_matches_reference(name, lo):
    import tntorch_b200 as tnb

    g = np.load(os.path.join(GOLD, "truncated_svd.npz"))
    spec = cases.TSVD_CASES[name]
    M = cases.make_matrix(spec)
    kw = {k: spec[k] for k in ("eps", "delta", "rmax") if k in spec}
    left, right = tnb.truncated_svd(torch.as_tensor(M).cuda(), left_ortho=lo, **kw)
    key = f"{name}/svd/{'L' if lo else 'R'}"
    assert left.shape[1] == int(g[key + "/rank"])
    prod = (left.double() @ right.double()).cpu().numpy()
    tol = 1e-4 if M.dtype == np.float32 else 1e-7
    np.testing.assert_allclose(prod, g[key + "/prod"], atol=tol * max(1.0, np.abs(M).max()))
    if not spec.get("zero"):
        r = left.shape[1]
        eye = torch.eye(r, device="cuda", dtype=left.dtype)
        orth = (left.T @ left - eye) if lo else (right @ right.T - eye)
        assert orth.abs().max().item() < (1e-4 if M.dtype == np.float32 else 1e-8)


def test_orthogonalization_preserves_tensor():
    """tests/test_round.py:7-18: left_/right_/orthogonalize leave the tensor unchanged (<= 1e-7) and produce
    orthonormal unfoldings."""
    import tntorch_b200 as tnb

    cores = cases.random_tt((6, 5, 7, 4, 6), 5, seed=77)
    gt = cases.tt_full(cores)
    t = tnb.Tensor([torch.as_tensor(c).cuda() for c in cores])
    R = t.left_orthogonalize(1)
    assert R.shape[0] == t.cores[1].shape[-1]
    assert relerr64(gt, t.cores) <= 1e-7
    L = t.right_orthogonalize(3)
    assert L.shape[1] == t.cores[3].shape[0]
    assert relerr64(gt, t.cores) <= 1e-7
    t.orthogonalize(2)
    assert relerr64(gt, t.cores) <= 1e-7
    for k in (0, 1):
        M = t.cores[k].reshape(-1, t.cores[k].shape[-1])
        assert (M.T @ M - torch.eye(M.shape[1], device="cuda", dtype=M.dtype)).abs().max().item() < 1e-10
    for k in (3, 4):
        M = t.cores[k].reshape(t.cores[k].shape[0], -1)
        assert (M @ M.T - torch.eye(M.shape[0], device="cuda", dtype=M.dtype)).abs().max().item() < 1e-10
    with pytest.raises(AssertionError):
        t.left_orthogonalize(4)
