"""The DEVICE algorithm (right-to-left Gram sweep + Chebyshev subspace iteration), stated in
NumPy in tests/sweep_model.py, reproduces the reference's ranks and relative error on the
golden vectors.  This pins the algorithm the CUDA kernels implement, without a GPU."""
import os

import numpy as np
import pytest

from oracle import cases
from oracle import tt_oracle as orc
from sweep_model import chfsi_topk, gram_sweep

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5  # BASELINE.json north_star: relative error within 1e-5 of the reference


@pytest.mark.parametrize("name", [k for k, v in cases.TTSVD_CASES.items() if v["kind"] != "zeros" and not (v.get("big") and "64" in k)])
def test_gram_sweep_matches_reference(name):
    g = np.load(os.path.join(GOLD, "ttsvd.npz"))
    spec = cases.TTSVD_CASES[name]
    X = cases.make_dense(spec)
    kw = {"eps": spec["eps"]} if spec.get("eps") is not None else {"ranks_tt": spec["ranks_tt"]}
    cores = gram_sweep(X, jacobi_max=128 if spec.get("big") else 256, **kw)
    alg = "svd" if f"{name}/svd/relerr" in g.files else "eig"
    assert [1] + [c.shape[2] for c in cores] == list(g[f"{name}/{alg}/ranks"])
    assert abs(orc.relative_error(X, cores) - float(g[f"{name}/{alg}/relerr"])) <= TOL


def test_gram_sweep_zero_tensor():
    cores = gram_sweep(np.zeros((6, 5, 4)), ranks_tt=3)
    assert [c.shape for c in cores] == [(1, 6, 1), (1, 5, 1), (1, 4, 1)]


@pytest.mark.parametrize("kind", ["flat", "decay", "lowrank"])
def test_chfsi_captures_optimal_energy(kind):
    rng = np.random.default_rng(0)
    n, r = 512, 16
    if kind == "flat":
        A = rng.standard_normal((8 * n, n))
    elif kind == "decay":
        A = rng.standard_normal((2 * n, n)) * np.logspace(0, -6, n)[None, :]
    else:
        A = rng.standard_normal((4 * n, r)) @ rng.standard_normal((r, n)) + 1e-3 * rng.standard_normal((4 * n, n))
    G = A.T @ A
    w = np.linalg.eigvalsh(G)[::-1]
    th, V, nprod = chfsi_topk(G, r)
    V = V[:, :r]
    assert np.abs(V.T @ V - np.eye(r)).max() < 1e-10
    deficit = (w[:r].sum() - np.trace(V.T @ G @ V)) / np.trace(G)
    assert deficit < 1e-7
    assert nprod < 400


def test_tf32_gram_is_a_uniform_scaling():
    """Why the tensor-core Gram (operands truncated to TF32 by the hardware) is safe for the rank rule, and where it
    stops being so (csrc/sweep.cuh, `tc_gram`): truncation shrinks every product by nearly the same factor, so
    G_tf32 = (1 - c) G + E with c ~ 7e-4 and ||E|| ~ 2e-6 ||G||: eigenvectors and eigenvalue RATIOS survive, absolute
    tails below ~1e-6 of the trace do not."""
    rng = np.random.default_rng(0)
    K, n, r = 100000, 64, 6
    A = (rng.standard_normal((K, r)) @ rng.standard_normal((r, n)) + 1e-3 * rng.standard_normal((K, n))).astype(np.float32)
    At = (A.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)
    A64 = A.astype(np.float64)
    G, Gt = A64.T @ A64, At.T @ At
    c = 1.0 - np.trace(Gt) / np.trace(G)
    assert 3e-4 < c < 1.5e-3
    assert np.linalg.norm(Gt - (1 - c) * G, 2) / np.linalg.norm(G, 2) < 5e-6
    w, wt = np.linalg.eigvalsh(G)[::-1], np.linalg.eigvalsh(Gt)[::-1]
    assert np.allclose(wt[:r] / w[:r], 1 - c, atol=5e-6)          # the signal spectrum is scaled, not distorted
    assert abs(wt[r:].sum() / wt.sum() - w[r:].sum() / w.sum()) < 2e-6  # tail fractions are good to ~1e-6 only
