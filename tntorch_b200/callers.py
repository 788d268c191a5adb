"""Callers of the rounding path that re-round in loops, on the device entry points:

  dot           metrics.py:28-116     <a, b> through the small interface matrices
  hadamard_sum  metrics.py:321-450    sum of the elementwise product of K tensors, re-rounding the running product
  shift_mode    tools.py:650-697      move one mode through a TT, one truncated_svd per position
  TTMatrix      matrix.py:12-140      matrix -> d-way (i_k x o_k) tensor -> TT-SVD (plain or batched)

Only the parts that sit on the hot path are mirrored (construction, decompression, the re-rounding loops); the
TT-matrix algebra (kron det/inv, tt_multiply) stays out of scope."""
from __future__ import annotations

import math
from typing import List, Sequence, Union

import torch

from . import ops
from .tensor import Tensor


def dot(t1: Tensor, t2: Tensor):
    """metrics.dot for two TT(-Tucker) tensors of the same shape; fp64 interface matrices."""
    assert list(t1.shape) == list(t2.shape)
    return t1._tt_dot(t2)


def _sum_all(cores: Sequence[torch.Tensor]):
    """Sum of every entry of a TT: each core is summed over its mode, the r x r' matrices are chained (fp64)."""
    f = torch.ones(1, cores[0].shape[0], dtype=torch.float64, device=cores[0].device)
    for c in cores:
        f = f @ c.sum(dim=1, dtype=torch.float64)
    return f.sum()


def _hadamard_sum_exact(cs: List[List[torch.Tensor]]):
    """The K-way interface tensor E[r_1..r_K] carried from the left: per mode, every operand's core is applied to its
    own axis for all mode values at once, then the mode is summed away (metrics.py:399-417)."""
    K, N = len(cs), len(cs[0])
    dev = cs[0][0].device
    E = torch.ones([1] * K, dtype=torch.float64, device=dev)
    for n in range(N):
        I = cs[0][n].shape[1]
        T = E[None].expand(I, *E.shape)
        for k in range(K):
            c = cs[k][n].double().permute(1, 0, 2)          # [I, r, r']
            T = T.movedim(k + 1, -1)                        # axis k last
            sh = T.shape
            T = torch.bmm(T.reshape(I, -1, sh[-1]), c).reshape(*sh[:-1], c.shape[-1]).movedim(-1, k + 1)
        E = T.sum(dim=0)
    return E.reshape(-1)[0]


def hadamard_sum(ts: Sequence[Tensor], algorithm: str = "exact", eps: float = None):
    """Sum over all entries of t_1 * t_2 * ... * t_K (metrics.hadamard_sum).

    'exact' contracts the K trains directly (cost grows with the product of the ranks).  'svd' / 'eig' keep a running
    product P <- round_tt(P * t_k, eps/sqrt(K-1)): Kronecker cores (tnb_tt_hadamard) followed by the rounding sweeps
    (tnb_tt_round), so the interface never exceeds the rounded rank times r_k; the result is within eps * ||prod|| *
    sqrt(numel) of the exact sum (the reference re-rounds the same intermediate quantities mode by mode)."""
    assert algorithm in ("svd", "eig", "exact")
    assert all(list(ts[0].shape) == list(t.shape) for t in ts[1:])
    if any(t.batch for t in ts):
        raise ValueError("Batched tensors are not supported.")
    cs = [t._tt_cores() for t in ts]
    if algorithm == "exact" or len(ts) == 1:
        return float(_hadamard_sum_exact(cs))
    step_eps = (1e-14 if eps is None else eps) / math.sqrt(len(ts) - 1)
    P = cs[0]
    for k in range(1, len(cs)):
        P = ops.tt_round(ops.tt_hadamard(P, cs[k]), eps=step_eps)
    return float(_sum_all(P))


def shift_mode(t: Tensor, n: int, shift: int, eps: Union[float, str] = 1e-3):
    """Move mode n by `shift` positions inside the train, in place (tools.shift_mode): orthogonalise at n, then per
    position merge the two neighbouring cores with swapped modes (library GEMM) and split them again with
    truncated_svd; eps='same' keeps every rank no larger than it was."""
    N = t.dim()
    assert 0 <= n + shift < N
    if shift == 0:
        return t
    if t.batch:
        raise ValueError("Batched tensors are not supported.")
    if any(U is not None for U in t.Us):
        d = t.decompress_tucker_factors()
        t.cores, t.Us = d.cores, [None] * N
    t.orthogonalize(n)
    cores = t.cores
    sign = 1 if shift > 0 else -1
    for i in range(n, n + shift, sign):
        c1, c2, left_ortho = (i, i + 1, True) if sign == 1 else (i - 1, i, False)
        R1, I1, R2 = cores[c1].shape
        _, I2, R3 = cores[c2].shape
        sc = ops.matmul(cores[c1].reshape(R1 * I1, R2), cores[c2].reshape(R2, I2 * R3))
        sc = sc.reshape(R1, I1, I2, R3).permute(0, 2, 1, 3).reshape(R1 * I2, I1 * R3).contiguous()
        if isinstance(eps, str):
            if eps != "same":
                raise ValueError("Relative error '{}' not recognized".format(eps))
            left, right = ops.truncated_svd(sc, eps=0, rmax=R2, left_ortho=left_ortho)
        elif eps >= 0:
            left, right = ops.truncated_svd(sc, eps=eps / math.sqrt(abs(shift)), left_ortho=left_ortho)
        else:
            raise ValueError("Relative error '{}' not recognized".format(eps))
        r = left.shape[1]
        cores[c1] = left.reshape(R1, I2, r)
        cores[c2] = right.reshape(r, I1, R3)
    return t


class TTMatrix:
    """A matrix of shape (prod i_k) x (prod o_k) stored as TT cores [r_k, i_k, o_k, r_{k+1}] (matrix.py:12-140):
    the matrix is viewed as the d-way tensor with modes (i_k o_k) and sent through the dense -> TT path
    (`Tensor(..., ranks_tt=ranks)`, batched through tnb_ttsvd_batch when a leading batch dimension is present)."""

    def __init__(self, t, ranks: List[int], input_dims: List[int], output_dims: List[int]):
        assert len(input_dims) == len(output_dims) and len(input_dims) > 0
        assert isinstance(ranks, list) and len(ranks) == len(input_dims) - 1
        self.input_dims = torch.tensor(input_dims)
        self.output_dims = torch.tensor(output_dims)
        self.d = d = len(input_dims)
        if isinstance(t, list):
            assert t[0].dim() in (4, 5)
            self.batch = t[0].dim() == 5
            self.cores = t
            self.ranks = torch.tensor([c.shape[-1] for c in t[:-1]])
            return
        M = t
        assert M.dim() in (2, 3)
        self.batch = M.dim() == 3
        assert math.prod(input_dims) == M.shape[-2] and math.prod(output_dims) == M.shape[-1]
        lead = [M.shape[0]] if self.batch else []
        off = len(lead)
        X = M.reshape(lead + list(input_dims) + list(output_dims))
        order = list(range(off)) + [off + k + s * d for k in range(d) for s in (0, 1)]     # i_0 o_0 i_1 o_1 ...
        X = X.permute(order).reshape(lead + [input_dims[k] * output_dims[k] for k in range(d)]).contiguous()
        tt = Tensor(X, ranks_tt=ranks, batch=self.batch)
        self.ranks = tt.ranks_tt[1:-1]
        self.cores = [c.reshape(*c.shape[: off + 1], input_dims[k], output_dims[k], c.shape[-1]) for k, c in enumerate(tt.cores)]

    def torch(self):
        """Back to the dense 2-D (or batched 3-D) matrix (matrix.py:120-160)."""
        d, off = self.d, 1 if self.batch else 0
        cores = [c.reshape(*c.shape[: off + 1], -1, c.shape[-1]) for c in self.cores]
        X = Tensor(cores, batch=self.batch).torch()
        idims, odims = self.input_dims.tolist(), self.output_dims.tolist()
        lead = [X.shape[0]] if self.batch else []
        X = X.reshape(lead + [s for k in range(d) for s in (idims[k], odims[k])])
        order = list(range(off)) + [off + 2 * k for k in range(d)] + [off + 2 * k + 1 for k in range(d)]
        return X.permute(order).reshape(lead + [math.prod(idims), math.prod(odims)])

    def numpy(self):
        return self.torch().detach().cpu().numpy()
