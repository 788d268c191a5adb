"""CPU: the dimension-tree MTTKRP of csrc/cp_als.cuh (tests/cp_tree_model.py) equals the direct MTTKRP for every mode while
the factors change inside the sweep, and the sweep built on it follows the reference's error (tests/golden/cp_als.npz)
when started from the reference's kind of initialisation (leading eigenvectors of the mode Grams, tensor.py:217-277)."""
import os

import numpy as np
import pytest

from cp_tree_model import Tree, als_sweeps, direct_mttkrp
from oracle import cases

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("shape", [(3, 4, 5), (3, 4, 5, 6), (2, 3, 4, 5, 3), (4, 1, 3, 2), (5, 2, 1)])
def test_tree_mttkrp_equals_direct(shape):
    rng = np.random.default_rng(0)
    R = 3
    X = rng.standard_normal(shape)
    A = [rng.standard_normal((s, R)) for s in shape]
    tree = Tree(X)
    for n in range(len(shape)):  # modes > n still hold the old factors when mode n is computed, modes < n the new ones
        M = tree.mttkrp(n, A)
        D = direct_mttkrp(X, A, n)
        assert np.abs(M - D).max() <= 1e-12 * np.abs(D).max()
        A[n] = np.random.default_rng(10 + n).standard_normal(A[n].shape)  # "update"


@pytest.mark.parametrize("name", ["cp_16x4_R5", "cp_20x3_R8", "cp_5mode_R4"])
def test_tree_sweeps_follow_the_reference(name):
    g = np.load(os.path.join(GOLD, "cp_als.npz"))
    spec = cases.CP_CASES[name]
    X = cases.make_cp_dense(spec)
    N, R = X.ndim, spec["R"]
    A = []
    for n in range(N):  # HOSVD start: leading R eigenvectors of X_(n) X_(n)^T
        Xn = np.moveaxis(X, n, 0).reshape(X.shape[n], -1)
        lam, V = np.linalg.eigh(Xn @ Xn.T)
        A.append(V[:, ::-1][:, :R].copy())
    if any(a.shape[1] < R for a in A):
        pytest.skip("a mode shorter than R starts from random columns in the reference")
    fac, errs = als_sweeps(X, A, spec["sweeps"])
    letters = "abcdefgh"[:N]
    full = np.einsum(",".join(f"{l}r" for l in letters) + "->" + letters, *fac)
    true = np.linalg.norm(X - full) / np.linalg.norm(X)
    assert abs(errs[-1] - true) <= 1e-9          # the Gram-form error is the reconstruction error
    assert abs(true - float(g[f"{name}/relerr"])) <= 1e-5
