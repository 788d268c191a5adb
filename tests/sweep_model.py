"""NumPy model of the DEVICE algorithm (not of the reference) — test infrastructure.

The CUDA path does not run the reference's three-phase recipe (identity-flanked full-rank
TT -> QR sweep -> SVD sweep).  It runs the algebraically equivalent right-to-left
Gram sweep described in DESIGN.md §2.  This file states that sweep in NumPy so the
CPU-only test-suite can check, against the golden vectors taken from the real
reference, that the *algorithm* the kernels implement reproduces the reference's
results (ranks and relative error), independent of any GPU.  ``tests/test_model.py``
runs it; the ``-m gpu`` tests then check the kernels against the oracle directly.
"""
from __future__ import annotations

import numpy as np


def _eigh_desc(G):
    w, v = np.linalg.eigh(G)
    idx = np.argsort(w)[::-1]
    return w[idx], v[:, idx]


def svqb(X):
    """Orthonormalise the columns of X through the eigendecomposition of its (scaled) Gram
    matrix; two passes.  Mirrors tnb::orthonormalize_block (csrc/eig.cuh)."""
    dt = X.dtype
    for _ in range(2):
        X64 = X.astype(np.float64)
        S = X64.T @ X64
        d = 1.0 / np.sqrt(np.maximum(np.diag(S), 1e-300))
        S = S * d[:, None] * d[None, :]
        lam, Q = np.linalg.eigh(S)
        lam = np.maximum(lam, lam.max() * 1e-13)
        X = (X64 @ ((d[:, None] * Q) / np.sqrt(lam)[None, :])).astype(dt)
    return X


def chfsi_topk(G, r, b=None, tol=1e-6, max_outer=40, spread=1e4, mmax=40, dtype=np.float64, seed=1):
    """Chebyshev-filtered subspace iteration for the r largest eigenpairs of the PSD matrix G.
    Mirrors tnb::eig_topk_chfsi (csrc/eig.cuh): scaled Chebyshev filter damping [0, θ_b],
    degree chosen so the filter's dynamic range stays below `spread`, SVQB orthonormalisation,
    Rayleigh-Ritz, stop when the captured energy Σ_{i<r} θ_i grows by less than tol·trace."""
    n = G.shape[0]
    if b is None:
        b = min(n, max(2 * r, r + 16))
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, b)).astype(dtype)
    G = G.astype(dtype)
    tr = float(np.trace(G))
    X = svqb(X)
    H = (X.T @ (G @ X)).astype(np.float64)
    th, Q = _eigh_desc(H)
    X = (X @ Q).astype(dtype)
    prev = th[:r].sum()
    nprod = 1
    for it in range(max_outer):
        cut = max(th[-1], 0.0)
        top = th[0]
        hi = max(cut, 1e-30 * top + 1e-300)
        e = hi / 2
        c = hi / 2
        x1 = (top - c) / e
        m = int(np.clip(np.floor(np.log(2 * spread) / max(np.arccosh(max(x1, 1.0)), 1e-12)), 1, mmax))
        sigma1 = e / (top - c)
        sigma = sigma1
        Y = (G @ X - c * X) * (sigma1 / e)
        for _ in range(2, m + 1):
            sigma2 = 1.0 / (2.0 / sigma1 - sigma)
            Y, X = (2 * sigma2 / e) * (G @ Y - c * Y) - (sigma * sigma2) * X, Y
            sigma = sigma2
        nprod += m + 1
        X = svqb(Y.astype(dtype))
        H = (X.T @ (G @ X)).astype(np.float64)
        th, Q = _eigh_desc(H)
        X = (X @ Q).astype(dtype)
        cap = th[:r].sum()
        if it >= 1 and cap - prev <= tol * tr:
            break
        prev = cap
    return th, X, nprod


def rank_rule(S2, delta2, rmax):
    """round.py:147-158 on descending squared singular values S2."""
    L = len(S2)
    count_true = int(np.sum(np.cumsum(S2[::-1]) <= delta2))
    if rmax is None:
        rmax = np.iinfo(np.int32).max
    return max(1, int(min(rmax, L - count_true)))


def gram_sweep(T, ranks_tt=None, eps=1e-14, jacobi_max=256):
    """Right-to-left Gram sweep on a dense array; returns TT cores (same dtype as T)."""
    T = np.asarray(T)
    dt = T.dtype
    shape = T.shape
    N = T.ndim
    rmax = ranks_tt
    if not hasattr(rmax, "__len__"):
        rmax = [rmax] * (N - 1)
    normT = float(np.linalg.norm(T.astype(np.float64)))
    delta2 = (eps / max(1.0, np.sqrt(N - 1)) * normT) ** 2
    cores = [None] * N
    C = T.reshape(-1, shape[-1])
    r_next = 1
    for mu in range(N - 1, 0, -1):
        rows = int(np.prod(shape[:mu]))
        n = shape[mu] * r_next
        C = C.reshape(rows, n)
        C64 = C.astype(np.float64)
        L = min(rows, n)
        tall = rows >= n
        G = C64.T @ C64 if tall else C64 @ C64.T
        if L <= jacobi_max or rmax[mu - 1] is None:
            lam, vec = _eigh_desc(G)
            tails = None
        else:
            r = min(rmax[mu - 1], L)
            th, X, _ = chfsi_topk(G, r, dtype=np.float64)
            lam, vec = th, X
            tails = float(np.trace(G))
        if np.sqrt(max(lam[0], 0.0)) < 1e-13:  # round.py:137-145
            cores[mu] = np.zeros((1, shape[mu], r_next), dt)
            C = np.zeros((rows, 1), dt)
            r_next = 1
            continue
        lam = np.maximum(lam, 0.0)
        if tails is None:
            rank = rank_rule(lam[:L], delta2, rmax[mu - 1])
        else:
            # only the leading Ritz values are known: tail_k = trace - Σ_{i<=k} θ_i
            k = min(rmax[mu - 1], L)
            tail = tails - np.cumsum(lam[:k])
            count_ok = int(np.sum(tail[: k - 1] <= delta2)) if k > 1 else 0
            rank = max(1, k - count_ok)
        s = np.sqrt(lam[:rank])
        if tall:
            V = vec[:, :rank]
            cores[mu] = V.T.reshape(rank, shape[mu], r_next).astype(dt)
            C = (C64 @ V.astype(dt).astype(np.float64)).astype(dt)
        else:
            U = vec[:, :rank]
            right = (U.T @ C64) / s[:, None]
            cores[mu] = right.reshape(rank, shape[mu], r_next).astype(dt)
            C = (U * s[None, :]).astype(dt)
        r_next = rank
    cores[0] = C.reshape(1, shape[0], r_next).astype(dt)
    return cores
