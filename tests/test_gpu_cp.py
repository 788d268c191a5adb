"""-m gpu: CP-ALS through the drop-in API against the fp64 reference (golden vectors, fixed sweep count;
SURVEY §0.5: the reference's own fp32 ALS collapses, so parity is taken against fp64)."""
import os

import numpy as np
import pytest
import torch

from oracle import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _cp_full(factors):
    letters = "abcdefgh"[: len(factors)]
    return np.einsum(",".join(f"{l}r" for l in letters) + "->" + letters, *[f.double().cpu().numpy() for f in factors])


@pytest.mark.parametrize("name", list(cases.CP_CASES))
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_cp_als_matches_fp64_reference(name, dtype):
    import tntorch_b200 as tnb

    g = np.load(os.path.join(GOLD, "cp_als.npz"))
    spec = cases.CP_CASES[name]
    X = cases.make_cp_dense(spec)
    t = tnb.Tensor(torch.as_tensor(X).to(dtype).cuda(), ranks_cp=spec["R"], max_iter=spec["sweeps"], tol=float("-inf"))
    assert [tuple(c.shape) for c in t.cores] == [(s, spec["R"]) for s in spec["shape"]]
    assert t.cores[0].dtype == dtype
    err = np.linalg.norm(X - _cp_full(t.cores)) / np.linalg.norm(X)
    ref = float(g[f"{name}/relerr"])
    assert abs(err - ref) <= (1e-5 if dtype == torch.float64 else 1e-4), (err, ref)
    # Tensor.torch() on CP cores reproduces the same reconstruction (tensor.py:1666-1680)
    assert np.abs(t.torch().double().cpu().numpy() - _cp_full(t.cores)).max() <= 1e-4 * np.abs(X).max()


def test_cp_als_error_trace_and_stopping():
    from tntorch_b200 import ops

    spec = cases.CP_CASES["cp_16x4_R5"]
    X = torch.as_tensor(cases.make_cp_dense(spec)).cuda()
    fac, info = ops.cp_als(X, 5, max_iter=25, tol=1e-4, return_info=True)
    errs = info["errors"]
    assert 2 <= info["iters"] <= 25 and len(errs) == info["iters"]
    assert errs[-2] - errs[-1] < 1e-4 or info["iters"] == 25  # tensor.py:380-381
    true = np.linalg.norm(X.cpu().numpy() - _cp_full(fac)) / np.linalg.norm(X.cpu().numpy())
    assert abs(errs[-1] - true) < 1e-6  # the Gram-based error formula equals the reconstruction error


def test_cp_als_rank_larger_than_mode():
    """I_n < R: missing HOSVD columns are filled pseudo-randomly (tensor.py:258-272)."""
    from tntorch_b200 import ops

    g = torch.Generator().manual_seed(0)
    fs = [torch.randn(s, 6, generator=g, dtype=torch.float64) for s in (4, 9, 8)]
    X = torch.einsum("ar,br,cr->abc", *fs).cuda()
    fac, info = ops.cp_als(X, 6, max_iter=60, tol=float("-inf"), return_info=True)
    assert info["errors"][-1] < 0.2


def test_cp_on_tucker_core():
    """tensor.py:278-302: Tensor(X, ranks_cp=R, ranks_tucker=S) = Tucker compression, then ALS on the dense Tucker core
    from random factors; the result is a CP-Tucker tensor (cores [S_n, R], factors [I_n, S_n])."""
    import tntorch_b200 as tnb

    rng = np.random.default_rng(9)
    fac = [rng.standard_normal((s, 3)) for s in (14, 12, 10)]
    X = torch.as_tensor(np.einsum("ar,br,cr->abc", *fac)).cuda()
    torch.manual_seed(0)
    t = tnb.Tensor(X, ranks_cp=3, ranks_tucker=4, max_iter=200, tol=1e-12)
    assert all(c.dim() == 2 and c.shape[1] == 3 and c.shape[0] <= 4 for c in t.cores)  # CP factors of the Tucker core
    assert [U.shape[0] for U in t.Us] == [14, 12, 10] and all(U.shape[1] == c.shape[0] for U, c in zip(t.Us, t.cores))
    assert list(t.shape) == [14, 12, 10]
    err = float(torch.linalg.vector_norm(X - t.torch()) / torch.linalg.vector_norm(X))
    assert err < 1e-3, err
    # ops.cp_als with a given start: zero sweeps return the start itself
    from tntorch_b200 import ops

    init = [torch.as_tensor(f).cuda() for f in fac]
    out = ops.cp_als(X, 3, max_iter=0, init=init)
    assert all(torch.equal(a, b) for a, b in zip(out, init))
    out, info = ops.cp_als(X, 3, max_iter=3, tol=float("-inf"), init=init, return_info=True)
    assert info["errors"][-1] < 1e-6  # started at the exact factors: stays there (the Gram-form error estimate resolves ~1e-8)
