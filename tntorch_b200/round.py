"""Module-level mirrors of tntorch/round.py:7-187 (clone-then-round wrappers, truncated_svd)."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def round_tt(t, **kwargs):
    """round.py:7-19"""
    t2 = t.clone()
    t2.round_tt(**kwargs)
    return t2


def round_tucker(t, **kwargs):
    """round.py:22-34"""
    t2 = t.clone()
    t2.round_tucker(**kwargs)
    return t2


def round(t, **kwargs):
    """round.py:37-49"""
    t2 = t.clone()
    t2.round(**kwargs)
    return t2


def truncated_svd(
    M: torch.Tensor,
    delta: Optional[float] = None,
    eps: Optional[float] = None,
    rmax: Optional[int] = None,
    left_ortho: Optional[bool] = True,
    algorithm: Optional[str] = "svd",
    verbose: Optional[bool] = False,
    batch: Optional[bool] = False,
):
    """round.py:52-187.  `algorithm` is accepted for signature parity; both values run the same
    Gram + eigen kernels (the reference's own 'eig' formulation) with fp64 accumulation."""
    if delta is not None and eps is not None:
        raise ValueError("Provide either `delta` or `eps`")
    assert rmax is None or rmax >= 1
    assert algorithm in ("svd", "eig")
    if batch:
        # batch mode ignores eps/delta (round.py:149-150): rank = min(rmax, len(S))
        outs = [ops.truncated_svd(M[b], rmax=rmax, left_ortho=left_ortho, batch_mode=True, return_zero_flag=True)
                for b in range(M.shape[0])]
        if all(o[2] for o in outs):  # round.py:138-142: every sample is zero -> rank-1 zero factors
            return (torch.zeros(M.shape[0], M.shape[1], 1, dtype=M.dtype, device=M.device),
                    torch.zeros(M.shape[0], 1, M.shape[2], dtype=M.dtype, device=M.device))
        return torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs])
    return ops.truncated_svd(M, delta=delta, eps=eps, rmax=rmax, left_ortho=left_ortho)


def reduce(ts, function, eps=0, rmax=None, algorithm="svd", verbose=False, **kwargs):
    """tools.py:460-512: a function of all tensors of a sequence, climbing a binary tree and rounding every intermediate.
    For `operator.add` (and `operator.sub`'s accumulate form) the node `tn.round(a + b)` is ONE library call
    (tnb_tt_sum_round: block cores assembled in the workspace, then the rounding sweeps); any other function is applied
    as given and rounded with `round`."""
    import operator

    from .tensor import Tensor

    assert algorithm in ("svd", "eig")
    fused = function is operator.add and not kwargs

    def node(a, b):
        if fused and not a.batch and not b.batch:
            return Tensor(ops.tt_sum_round([a._tt_cores(), b._tt_cores()], eps=eps, rmax=rmax))
        out = function(a, b, **kwargs)
        out.round(eps=eps, rmax=rmax, algorithm=algorithm)
        return out

    d = dict()
    for elem in ts:
        climb = 0
        while climb in d:
            elem = node(d[climb], elem)
            d.pop(climb)
            climb += 1
        d[climb] = elem
    keys = list(d.keys())
    result = d[keys[0]]
    for key in keys[1:]:
        result = node(result, d[key])
    return result


def relative_error(gt, approx):
    """metrics.py:135-151 for (dense torch tensor, Tensor)."""
    from .tensor import Tensor

    if isinstance(gt, torch.Tensor) and isinstance(approx, Tensor) and not approx.batch and all(c.dim() == 3 for c in approx.cores):
        if any(U is not None for U in approx.Us):
            approx = approx.decompress_tucker_factors()
        return ops.tt_relative_error(gt.to(approx.cores[0].device), approx.cores)
    a = gt.torch() if isinstance(gt, Tensor) else gt
    b = approx.torch() if isinstance(approx, Tensor) else approx
    return float(torch.dist(a, b) / torch.norm(a))
