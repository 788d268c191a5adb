"""NumPy statement of the CP-ALS sweep as libtnb200 runs it (csrc/cp_als.cuh: cp_tree_mttkrp + cp_als_impl): one projection
of X per sweep shared by modes 0..N-2 (right-to-left chain of Khatri-Rao reductions), mode N-1 through the transposed
copy, eig-based normal equations, Gram-form error.  Mirrors the kernels' index arithmetic (L, I, Q of every reduction)."""
import numpy as np


def khatri(Y, A, L, I, Q, R):
    """out[l, q, r] = sum_i Y[l, i, q, r] A[i, r]   (khatri_reduce*_kernel)"""
    return np.einsum("liqr,ir->lqr", Y.reshape(L, I, Q, R), A).reshape(-1)


class Tree:
    def __init__(self, X):
        self.X = X
        self.sh = X.shape
        self.N = X.ndim
        self.left = [int(np.prod(self.sh[:n])) for n in range(self.N)]
        self.XT = np.moveaxis(X, self.N - 1, 0).copy()  # cp_transpose_kernel, once per call

    def mttkrp(self, n, A):
        N, sh, left, R = self.N, self.sh, self.left, A[0].shape[1]
        if n == 0:
            self.Yk = (self.X.reshape(-1, sh[-1]) @ A[N - 1]).reshape(-1)  # project_any over X
            self.chain = [None] * (N - 2)
            src = self.Yk
            for k in range(N - 3, -1, -1):
                self.chain[k] = khatri(src, A[k + 1], left[k + 1], sh[k + 1], 1, R)
                src = self.chain[k]
            return self.chain[0].reshape(sh[0], R)
        if n <= N - 2:
            cur = self.Yk if n == N - 2 else self.chain[n]
            for lo in range(n):
                cur = khatri(cur, A[lo], 1, sh[lo], int(np.prod(sh[lo + 1:n + 1])), R)
            return cur.reshape(sh[n], R)
        cur = (self.XT.reshape(-1, sh[N - 2]) @ A[N - 2]).reshape(-1)  # project_any over XT
        for m in range(N - 3, -1, -1):
            cur = khatri(cur, A[m], sh[N - 1] * left[m], sh[m], 1, R)
        return cur.reshape(sh[N - 1], R)


def direct_mttkrp(X, A, n):
    letters = "abcdefgh"[: X.ndim]
    ops = [A[m] for m in range(X.ndim) if m != n]
    sub = letters + "," + ",".join(letters[m] + "r" for m in range(X.ndim) if m != n) + "->" + letters[n] + "r"
    return np.einsum(sub, X, *ops)


def als_sweeps(X, A, sweeps):
    """The sweep of cp_als_impl for N >= 3 from given factors; returns the factors and the per-sweep Gram-form error."""
    N = X.ndim
    A = [a.copy() for a in A]
    tree = Tree(X)
    grams = [a.T @ a for a in A]
    normX2 = float((X * X).sum())
    errs = []
    for _ in range(sweeps):
        for n in range(N):
            M = tree.mttkrp(n, A)
            P = np.ones_like(grams[0])
            for m in range(N):
                if m != n:
                    P = P * grams[m]
            lam, Q = np.linalg.eigh(P)
            inv = np.where(lam > 2.220446049250313e-16 * P.shape[0] * lam.max(), 1.0 / lam, 0.0)
            A[n] = M @ ((Q * inv) @ Q.T)
            grams[n] = A[n].T @ A[n]
        P = np.ones_like(grams[0])
        for m in range(N):
            P = P * grams[m]
        e2 = normX2 - 2.0 * float((M * A[N - 1]).sum()) + float(P.sum())
        errs.append(np.sqrt(max(e2, 0.0) / normX2))
    return A, errs
