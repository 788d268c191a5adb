"""-m gpu: parity of the CUDA TT-SVD (through the drop-in API -> C-ABI) with the reference's golden
vectors and with the oracle on the same seeded inputs.  Bar: ranks equal, relative reconstruction
error within 1e-5 of the reference (BASELINE.json north_star)."""
import os

import numpy as np
import pytest
import torch

from gpu_util import ranks_of, relerr64
from oracle import cases
from oracle import tt_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5


def _run(name, use_tc=True):
    import tntorch_b200 as tnb
    from tntorch_b200 import ops

    spec = cases.TTSVD_CASES[name]
    X = cases.make_dense(spec)
    Xd = torch.as_tensor(X).cuda()
    if spec.get("eps") is not None:
        t = tnb.Tensor(Xd, eps=spec["eps"])  # = round_tt + round_tucker in the reference (tensor.py:436-439)
        ranks = list(t.ranks_tt)
        cores = t.decompress_tucker_factors().cores
        assert ranks_of(cores) == ranks
    else:
        cores = ops.ttsvd(Xd, rmax=spec["ranks_tt"], use_tensorcore=use_tc)
    return X, cores


@pytest.mark.parametrize("name", [k for k, v in cases.TTSVD_CASES.items() if v["kind"] != "zeros"])
def test_ttsvd_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, "ttsvd.npz"))
    X, cores = _run(name)
    alg = "svd" if f"{name}/svd/relerr" in g.files else "eig"
    assert ranks_of(cores) == list(g[f"{name}/{alg}/ranks"])
    assert cores[0].dtype == torch.as_tensor(X).dtype
    assert abs(relerr64(X, cores) - float(g[f"{name}/{alg}/relerr"])) <= TOL


@pytest.mark.parametrize("name", ["randn32x5_r32_f32", "twin32x5_r32_f32", "randn64x4_r32_f32", "twin_16x5_f32"])
def test_ttsvd_cuda_core_path_matches_too(name):
    """Same cases with the tensor-core Gram disabled (generic fp64-accumulating kernels)."""
    g = np.load(os.path.join(GOLD, "ttsvd.npz"))
    X, cores = _run(name, use_tc=False)
    alg = "svd" if f"{name}/svd/relerr" in g.files else "eig"
    assert ranks_of(cores) == list(g[f"{name}/{alg}/ranks"])
    assert abs(relerr64(X, cores) - float(g[f"{name}/{alg}/relerr"])) <= TOL


@pytest.mark.parametrize("name", ["cfg1_randn16x4_f32", "ragged_f32", "twin_small_f64", "two_modes"])
def test_ttsvd_matches_oracle_live(name):
    spec = cases.TTSVD_CASES[name]
    X, cores = _run(name)
    oc = orc.tt_svd(X, ranks_tt=spec["ranks_tt"])
    assert ranks_of(cores) == [1] + [c.shape[2] for c in oc]
    assert abs(relerr64(X, cores) - orc.relative_error(X, oc)) <= TOL


def test_gauge_matches_reference():
    """cores[1:] have orthonormal right unfoldings, cores[0] carries the norm."""
    X, cores = _run("twin_small_f64")
    for c in cores[1:]:
        M = c.reshape(c.shape[0], -1)
        assert (M @ M.T - torch.eye(M.shape[0], device=M.device, dtype=M.dtype)).abs().max().item() < 1e-9
    assert abs(torch.linalg.norm(cores[0]).item() - np.linalg.norm(X)) / np.linalg.norm(X) < 1e-4


def test_zero_tensor():
    from tntorch_b200 import ops

    cores = ops.ttsvd(torch.zeros(6, 5, 4, dtype=torch.float64, device="cuda"), rmax=3)
    assert [tuple(c.shape) for c in cores] == [(1, 6, 1), (1, 5, 1), (1, 4, 1)]
    assert all(float(c.abs().max()) == 0.0 for c in cores)


def test_device_relative_error_kernel():
    from tntorch_b200 import ops

    X, cores = _run("twin_16x5_f32")
    e_dev = ops.tt_relative_error(torch.as_tensor(X).cuda(), cores)
    assert abs(e_dev - relerr64(X, cores)) < 2e-6


def test_batch_mode_matches_per_sample():
    """tests/test_tensor.py:28-49: batched TT-SVD equals the per-sample decomposition."""
    import tntorch_b200 as tnb

    g = torch.Generator().manual_seed(3)
    X = torch.randn(4, 5, 5, 5, 5, generator=g, dtype=torch.float64).cuda()
    tb = tnb.Tensor(X, ranks_tt=3, batch=True)
    for b in range(4):
        ts = tnb.Tensor(X[b], ranks_tt=3)
        assert torch.allclose(tb.torch()[b], ts.torch(), atol=1e-9)


def test_unsupported_is_loud():
    from tntorch_b200 import ops

    with pytest.raises(NotImplementedError):  # eps-only on a >256 Gram whose rank exceeds 240: raised, never faked
        ops.ttsvd(torch.randn(600, 600, device="cuda"), rmax=None, eps=1e-3)
