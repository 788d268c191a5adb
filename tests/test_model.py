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
