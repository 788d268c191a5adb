"""Host-side mirror of the reference's `tn.Tensor` for the decomposition / rounding path.

Same constructor signature, attribute names and error behaviour as rballester/tntorch
(tntorch/tensor.py:107-439), restricted to the hot path this package accelerates:
dense -> TT (``ranks_tt=`` / ``eps=``), TT cores in / out, ``round_tt``, ``round``, ``torch()``.
Everything numerical runs in libtnb200.so on the GPU; there is no CPU path.
"""
from __future__ import annotations

from typing import Any, Optional, Sequence, Union

import numpy as np
import torch

from . import ops


def _default_device(device):
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("tntorch_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"tntorch_b200 runs on CUDA devices only, got device={device}")
    return device


class Tensor(object):
    """TT tensor whose decomposition / rounding runs on B200 kernels (mirror of tntorch.Tensor)."""

    def __init__(
        self,
        data: Union[torch.Tensor, np.ndarray, Sequence[torch.Tensor]],
        Us: Optional[Union[torch.Tensor, Any]] = None,
        idxs: Optional[Any] = None,
        device: Optional[Any] = None,
        requires_grad: Optional[bool] = None,
        ranks_cp: int = None,
        ranks_tucker: Optional[Sequence[int]] = None,
        ranks_tt: Optional[Sequence[int]] = None,
        eps: Optional[float] = None,
        max_iter: Optional[int] = 25,
        tol: Optional[float] = 1e-4,
        verbose: Optional[bool] = False,
        batch: Optional[bool] = False,
        algorithm: Optional[str] = "svd",
    ):
        assert algorithm in ("svd", "eig")  # both map onto the same Gram/eigen kernels
        self.batch = batch
        if ranks_tucker is not None:
            raise NotImplementedError(
                "tntorch_b200 covers the TT / CP decomposition and TT rounding path (SURVEY.md §8); Tucker "
                "rounding is listed as a next row and is not built yet"
            )
        if isinstance(data, (list, tuple)):  # explicit cores (tensor.py:163-192)
            min_dim, max_dim = (3, 4) if batch else (2, 3)
            if not all(min_dim <= d.dim() <= max_dim for d in data):
                raise ValueError("All tensor cores must have 2 (for CP) or 3 (for TT) dimensions")
            d1, d2 = (1, 2) if batch else (0, 1)
            for n in range(len(data) - 1):  # tensor.py:177-191
                if (data[n + 1].dim() == max_dim and data[n].shape[-1] != data[n + 1].shape[d1]) or (
                    data[n + 1].dim() == min_dim and data[n].shape[-1] != data[n + 1].shape[d2]
                ):
                    raise ValueError("Core ranks do not match")
            dev = _default_device(device if device is not None else (data[0].device if data[0].is_cuda else None))
            self.cores = [c.to(dev) for c in data]
            N = len(data)
        else:
            if isinstance(data, np.ndarray):
                data = torch.as_tensor(data)
            elif not isinstance(data, torch.Tensor):
                raise ValueError(
                    "A tntorch.Tensor may be built either from a list of cores, one NumPy ndarray, or one PyTorch tensor"
                )
            dev = _default_device(device if device is not None else (data.device if data.is_cuda else None))
            data = data.to(dev)
            if data.dim() == 0:
                data = data * torch.ones(1, device=dev, dtype=data.dtype)
            if eps is not None and ranks_tt is not None:
                raise ValueError("Specify eps or ranks, but not both")  # tensor.py:436-438
            N = data.dim() - 1 if batch else data.dim()
            if ranks_cp is not None:  # CP-ALS (tensor.py:210-400)
                if ranks_tt is not None:
                    raise ValueError("ALS for CP-TT is not yet supported")
                assert not hasattr(ranks_cp, "__len__")
                if eps is not None:
                    raise ValueError("Specify eps or ranks, but not both")
                if batch:
                    self.cores = [torch.stack(f, dim=0) for f in zip(*[
                        ops.cp_als(data[b], ranks_cp, max_iter=max_iter, tol=tol) for b in range(data.shape[0])])]
                else:
                    self.cores = ops.cp_als(data, ranks_cp, max_iter=max_iter, tol=tol)
            elif batch:
                # the reference's batch mode: per-sample decomposition, rank = min(rmax, len(S)), no eps
                per = [ops.ttsvd(data[b], rmax=ranks_tt, batch_mode=True) for b in range(data.shape[0])]
                self.cores = [torch.stack([p[k] for p in per], dim=0) for k in range(N)]
            elif eps is not None:
                # Tensor(data, eps=...) -> round(eps) (tensor.py:436-439); the Tucker pass of round() is a no-op
                # for the ranks but not built here, so only the TT budget is spent.
                self.cores = ops.ttsvd(data, rmax=None, eps=eps)
            else:
                self.cores = ops.ttsvd(data, rmax=ranks_tt)
        if Us is None:
            Us = [None] * N
        self.Us = Us
        if requires_grad:
            for c in self.cores:
                c.requires_grad_()
        if idxs is None:
            idxs = [torch.arange(sh, device=self.cores[0].device) for sh in self.shape]
        self.idxs = idxs

    # ------------------------------------------------------------------ structure
    def dim(self):
        return len(self.cores)

    @property
    def shape(self):
        return torch.Size([c.shape[-2] for c in self.cores])

    @property
    def ranks_tt(self):
        d1 = 1 if self.batch else 0
        first = self.cores[0].shape[d1] if self.cores[0].dim() == (4 if self.batch else 3) else self.cores[0].shape[-1]
        return torch.tensor([first] + [c.shape[-1] for c in self.cores])

    def numcoef(self):
        return sum(c.numel() for c in self.cores)

    def clone(self):
        return Tensor([c.clone() for c in self.cores], batch=self.batch)

    def __repr__(self):
        return f"{self.dim()}D TT tensor (B200): shape {list(self.shape)}, TT ranks {self.ranks_tt.tolist()}"

    # ------------------------------------------------------------------ decompression (tensor.py:1639-1687)
    def torch(self):
        if self.batch:
            return torch.stack([Tensor([c[b] for c in self.cores]).torch() for b in range(self.cores[0].shape[0])])
        c0 = self.cores[0]
        r0 = c0.shape[0] if c0.dim() == 3 else c0.shape[1]
        f = torch.ones(1, r0, dtype=c0.dtype, device=c0.device)
        last = len(self.cores) - 1
        for n, c in enumerate(self.cores):  # tensor.py:1666-1680
            if c.dim() == 2:  # CP factor [I, R]
                if n < last:
                    f = torch.einsum("ai,bi->abi", f, c)
                else:
                    f = torch.einsum("ai,bi->ab", f, c)[..., None]
            else:
                f = torch.einsum("ai,ibj->abj", f, c)
            f = f.reshape(-1, f.shape[-1])
        f = f.sum(dim=-1) if f.shape[-1] > 1 else f[..., 0]
        return f.reshape(list(self.shape))

    def numpy(self):
        return self.torch().detach().cpu().numpy()

    # ------------------------------------------------------------------ orthogonalisation (tensor.py:1800-1909)
    def left_orthogonalize(self, mu: int):
        """Makes the mu-th core left-orthogonal and pushes the R factor to its right core; returns R
        (tensor.py:1800-1833).  Householder QR and the R push run in libtnb200 (tnb_qr_householder, tnb_matmul)."""
        assert 0 <= mu < self.dim() - 1
        if self.batch:
            raise NotImplementedError("batched orthogonalisation is not built")
        c = self.cores[mu]
        Q, R = ops.qr(c.reshape(-1, c.shape[-1]), return_r=True)
        Q, R = Q.to(c.dtype), R.to(c.dtype)
        self.cores[mu] = Q.reshape(c.shape[:-1] + (Q.shape[1],))
        nxt = self.cores[mu + 1]
        self.cores[mu + 1] = ops.matmul(R, nxt.reshape(nxt.shape[0], -1)).reshape((R.shape[0],) + nxt.shape[1:])
        return R

    def right_orthogonalize(self, mu: int):
        """Makes the mu-th core right-orthogonal and pushes the L factor to its left core; returns L
        (tensor.py:1835-1879)."""
        assert 1 <= mu < self.dim()
        if self.batch:
            raise NotImplementedError("batched orthogonalisation is not built")
        c = self.cores[mu]
        Q, L = ops.qr(c.reshape(c.shape[0], -1).t().contiguous(), return_r=True)
        L, Q = L.t().contiguous().to(c.dtype), Q.t().contiguous().to(c.dtype)
        self.cores[mu] = Q.reshape((Q.shape[0],) + c.shape[1:])
        prv = self.cores[mu - 1]
        self.cores[mu - 1] = ops.matmul(prv.reshape(-1, prv.shape[-1]), L).reshape(prv.shape[:-1] + (L.shape[1],))
        return L

    def orthogonalize(self, mu: int):
        """All left and right orthogonalisations needed to make the tensor mu-orthogonal; returns (R, L)
        (tensor.py:1881-1909)."""
        if mu < 0:
            mu += self.dim()
        dt, dev = self.cores[0].dtype, self.cores[0].device
        R = torch.ones(1, 1, dtype=dt, device=dev)
        L = torch.ones(1, 1, dtype=dt, device=dev)
        for i in range(mu):
            R = self.left_orthogonalize(i)
        for i in range(self.dim() - 1, mu, -1):
            L = self.right_orthogonalize(i)
        return R, L

    # ------------------------------------------------------------------ rounding (tensor.py:2008-2098)
    def round_tt(self, eps: float = 1e-14, rmax=None, algorithm: Optional[str] = "svd", verbose: Optional[bool] = False):
        """In place, by rebinding ``self.cores`` (same contract as the reference)."""
        assert algorithm in ("svd", "eig")
        N = self.dim()
        if not hasattr(rmax, "__len__"):
            rmax = [rmax] * (N - 1)
        assert len(rmax) == N - 1
        if self.batch:
            B = self.cores[0].shape[0]
            per = [ops.tt_round([c[b] for c in self.cores], eps=eps, rmax=rmax, batch_mode=True) for b in range(B)]
            self.cores = [torch.stack([p[k] for p in per], dim=0) for k in range(N)]
        else:
            self.cores = ops.tt_round(self.cores, eps=eps, rmax=rmax)

    def round(self, eps: float = 1e-14, **kwargs):
        """tensor.py:2085-2098: TT rounding, then Tucker rounding with the left-over budget.  Only the TT
        stage is built (Tucker factors are a 'next' row of SURVEY.md §8f)."""
        self.round_tt(eps, **kwargs)
