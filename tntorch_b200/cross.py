"""TT-cross approximation, host-side mirror of tntorch/cross.py:138-529 with the per-core numerics on the GPU.

Same signature, same sampling order of the global NumPy / torch RNGs, same `info` keys
(`nsamples, eval_time, val_epss, lsets, rsets, Rs, left_locals, total_time, val_eps`).  What changes is where
the work happens: the reference copies every QR factor to the host and runs a pure-NumPy maxvol
(cross.py:400-402, 432-434) followed by an `lstsq` (cross.py:403, 435); here the Householder QR
(`tnb_qr_householder`), maxvol and the interpolation core `Q inv(Q[local])` (`tnb_maxvol`) run on the device and
the index sets live in device tensors, so a sweep performs no device->host copy at all.
"""
from __future__ import annotations

import logging
import time

import numpy as np
import torch

from . import ops
from .tensor import Tensor


def meshgrid(domain, device=None):
    """tn.meshgrid for 1-D axes (tools.py): one rank-1 TT per axis."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    axes = [torch.as_tensor(a).to(dev) for a in domain]
    N = len(axes)
    out = []
    for k in range(N):
        cores = []
        for n in range(N):
            if n == k:
                cores.append(axes[n].reshape(1, -1, 1).clone())
            else:
                cores.append(torch.ones(1, len(axes[n]), 1, dtype=axes[k].dtype, device=dev))
        out.append(Tensor(cores))
    return out


def _evaluate_tt_at(cores, idxs):
    """Values of a TT at P multi-indices (idxs: list of N LongTensors of length P): Tensor[Xs].torch()."""
    P = idxs[0].shape[0]
    f = torch.ones(P, 1, cores[0].shape[0], dtype=cores[0].dtype, device=cores[0].device)
    for c, ix in zip(cores, idxs):
        f = torch.bmm(f, c[:, ix, :].permute(1, 0, 2))
    return f[:, 0, :].sum(dim=-1)


def _init_interfaces(tensors, rsets, N, device):
    """cross.py:113-135 with device index sets."""
    t_l, t_r = [], []
    for t in tensors:
        lin = [torch.ones(1, t.cores[0].shape[0], dtype=t.cores[0].dtype, device=device)] + [None] * (N - 1)
        rin = [None] * (N - 1) + [torch.ones(t.cores[-1].shape[-1], 1, dtype=t.cores[0].dtype, device=device)]
        for j in range(N - 1):
            M = torch.ones(t.cores[-1].shape[-1], rsets[j].shape[0], dtype=t.cores[0].dtype, device=device)
            for n in range(N - 1, j, -1):
                M = torch.einsum("iaj,ja->ia", t.cores[n][:, rsets[j][:, n - 1 - j], :], M)
            rin[j] = M
        t_l.append(lin)
        t_r.append(rin)
    return t_l, t_r


def cross(
    function=lambda x: x,
    domain=None,
    tensors=None,
    function_arg="vectors",
    ranks_tt=None,
    kickrank=3,
    rmax=100,
    eps=1e-6,
    max_iter=25,
    val_size=1000,
    verbose=True,
    return_info=False,
    record_samples=False,
    _minimize=False,
    device=None,
    suppress_warnings=False,
    detach_evaluations=False,
):
    """Cross-approximation of `function` over a tensor-product domain (see tntorch.cross)."""
    # _minimize (cross.py:342-359, 399-400): the samples are mapped through pi/2 - atan(f - min) before the maxvol
    # step, the running minimum and its multi-index are tracked, and the index search is the reference's
    # rect_maxvol(Q, maxK=r) (the device kernel of tnb_rect_maxvol; with maxK = r no row is added, i.e. maxvol stopped
    # after start_maxvol_iters = 10 swaps, maxvol.py:52-84).
    def pick_rows(Q):
        if _minimize:
            return ops.rect_maxvol(Q, maxK=Q.shape[1])
        return ops.maxvol(Q, max_iters=100)

    assert domain is not None or tensors is not None
    assert function_arg in ("vectors", "matrix")
    if device is None:
        if tensors is not None:
            t0 = tensors[0] if hasattr(tensors, "__len__") else tensors
            device = t0.cores[0].device
        else:
            device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("tntorch_b200.cross runs on CUDA devices only")
    if function_arg == "matrix":
        def f(*args):
            return function(torch.cat([arg[:, None] for arg in args], dim=1))
    else:
        f = function
    if detach_evaluations:
        inner = f

        def f(*args):  # noqa: F811
            res = inner(*args)
            return res.detach() if isinstance(res, torch.Tensor) else res

    if tensors is None:
        tensors = meshgrid(domain, device=device)
    if not hasattr(tensors, "__len__"):
        tensors = [tensors]
    for t in tensors:
        if t.batch:
            raise ValueError("Batched tensors are not supported.")  # cross.py:256-258
        if any(c.dim() != 3 for c in t.cores):
            raise NotImplementedError("cross over CP-format inputs is not built (TT cores only)")
    Is = list(tensors[0].shape)
    N = len(Is)

    if ranks_tt is None:
        ranks_tt = 1
    else:
        kickrank = None
    if not hasattr(ranks_tt, "__len__"):
        ranks_tt = [ranks_tt] * (N - 1)
    Rs = np.array([1] + list(ranks_tt) + [1])
    for n in list(range(1, N)) + list(range(N - 1, -1, -1)):
        Rs[n] = min(Rs[n - 1] * Is[n - 1], Rs[n], Is[n] * Rs[n + 1])

    dtype = tensors[0].cores[0].dtype
    cores = [torch.randn(int(Rs[n]), Is[n], int(Rs[n + 1])).to(device) for n in range(N)]  # same RNG draw as cross.py:277
    # index sets (cross.py:280-285): same NumPy RNG calls, kept on the device afterwards
    lsets = [torch.zeros(1, 1, dtype=torch.long, device=device)] + [None] * (N - 1)
    randint = np.hstack(
        [np.random.randint(0, Is[n + 1], [max(Rs), 1]) for n in range(N - 1)] + [np.zeros([max(Rs), 1], dtype=int)]
    )
    rsets = [torch.as_tensor(randint[: Rs[n + 1], n:], dtype=torch.long, device=device) for n in range(N - 1)] + [
        torch.zeros(1, 1, dtype=torch.long, device=device)
    ]
    t_lin, t_rin = _init_interfaces(tensors, rsets, N, device)

    Xs_val = [torch.as_tensor(np.random.choice(I, int(val_size))).to(device) for I in Is]
    ys_val = f(*[_evaluate_tt_at(t.cores, Xs_val) for t in tensors])
    if ys_val.dim() > 1:
        assert ys_val.dim() == 2 and ys_val.shape[1] == 1
        ys_val = ys_val[:, 0]
    assert len(ys_val) == val_size
    norm_ys_val = torch.norm(ys_val)

    if verbose:
        print("Cross-approximation over a {}D domain containing {:g} grid points:".format(N, float(np.prod(Is))))
    start = time.time()
    converged = False
    info = {"nsamples": 0, "eval_time": 0, "val_epss": [], "min": 0, "argmin": None}
    if record_samples:
        info["sample_positions"] = torch.zeros(0, N, device=device)
        info["sample_values"] = torch.zeros(0, device=device)

    def evaluate_function(j):  # cross.py:316-379
        Xs = []
        for k, t in enumerate(tensors):
            V = torch.einsum("ai,ibj,jc->abc", t_lin[k][j], t.cores[j], t_rin[k][j])
            Xs.append(V.flatten())
        t0 = time.time()
        evaluation = f(*Xs)
        if record_samples:
            info["sample_positions"] = torch.cat((info["sample_positions"], torch.stack(Xs, dim=1)), dim=0)
            info["sample_values"] = torch.cat((info["sample_values"], evaluation))
        info["eval_time"] += time.time() - t0
        if _minimize:
            shifted = np.pi / 2 - torch.atan(evaluation - info["min"])
            am = torch.argmax(shifted)
            eval_min = torch.tan(np.pi / 2 - shifted[am]) + info["min"]
            if (isinstance(info["min"], (int, float)) and info["min"] == 0) or bool(eval_min < info["min"]):
                c0, c1, c2 = np.unravel_index(int(am), [int(Rs[j]), Is[j], int(Rs[j + 1])])
                info["min"] = eval_min
                info["argmin"] = (tuple(int(v) for v in lsets[j][c0][1:].tolist()) + (int(c1),)
                                  + tuple(int(v) for v in rsets[j][c2][:-1].tolist()))
            evaluation = shifted
        if evaluation.dim() == 2:
            evaluation = evaluation[:, 0]
        bad = torch.isnan(evaluation) | torch.isinf(evaluation)
        if bool(bad.any()):  # cross.py:361-375 (this check is the sweep's only host synchronisation)
            k0 = int(torch.nonzero(bad)[0].item())
            raise ValueError(
                "Invalid return value for function {}: f({}) = {}".format(
                    function, ", ".join("{:g}".format(float(x[k0])) for x in Xs), float(evaluation[k0])
                )
            )
        V = evaluation.reshape(int(Rs[j]), Is[j], int(Rs[j + 1]))
        info["nsamples"] += V.numel()
        return V

    val_eps = torch.tensor(float("inf"))
    for it in range(max_iter):
        left_locals = []
        # ---- left-to-right (cross.py:391-420) ----
        for j in range(N - 1):
            V = evaluate_function(j).reshape(-1, int(Rs[j + 1]))
            Q = ops.qr(V)  # Householder QR on the device
            local, C = pick_rows(Q)  # local: int32 [R_{j+1}], C = Q inv(Q[local])  (= the lstsq of cross.py:403)
            cores[j] = C.to(V.dtype).reshape(int(Rs[j]), Is[j], int(Rs[j + 1]))
            local = local.long()
            left_locals.append(local)
            local_r = torch.div(local, Is[j], rounding_mode="floor")
            local_i = local - local_r * Is[j]
            lsets[j + 1] = torch.cat([lsets[j][local_r, :], local_i[:, None]], dim=1)
            for k, t in enumerate(tensors):
                t_lin[k][j + 1] = torch.einsum("ai,iaj->aj", t_lin[k][j][local_r, :], t.cores[j][:, local_i, :])
        # ---- right-to-left (cross.py:423-451) ----
        for j in range(N - 1, 0, -1):
            V = evaluate_function(j).reshape(int(Rs[j]), -1)
            Q = ops.qr(V.t().contiguous())
            local, C = pick_rows(Q)
            cores[j] = C.t().to(V.dtype).reshape(int(Rs[j]), Is[j], int(Rs[j + 1]))
            local = local.long()
            local_i = torch.div(local, int(Rs[j + 1]), rounding_mode="floor")
            local_r = local - local_i * int(Rs[j + 1])
            rsets[j - 1] = torch.cat([local_i[:, None], rsets[j][local_r, :]], dim=1)
            for k, t in enumerate(tensors):
                t_rin[k][j - 1] = torch.einsum("iaj,ja->ia", t.cores[j][:, local_i, :], t_rin[k][j][:, local_r])
        cores[0] = evaluate_function(0)  # cross.py:454-455

        val_eps = torch.norm(ys_val - _evaluate_tt_at(cores, Xs_val)) / norm_ys_val
        info["val_epss"].append(val_eps)
        if val_eps < eps:
            converged = True
        if verbose:
            print("iter: {: <3} | eps: {:.3e} | time: {:8.4f} | largest rank: {:3d}".format(
                it, float(val_eps), time.time() - start, int(max(Rs))), end="")
            print(" <- converged: eps < {}".format(eps) if converged else
                  (" <- max_iter was reached: {}".format(max_iter) if it == max_iter - 1 else ""))
        if converged:
            break
        elif it < max_iter - 1 and kickrank is not None:  # cross.py:481-498
            newRs = Rs.copy()
            newRs[1:-1] = np.minimum(rmax, newRs[1:-1] + kickrank)
            for n in list(range(1, N)) + list(range(N - 1, 0, -1)):
                newRs[n] = min(newRs[n - 1] * Is[n - 1], newRs[n], Is[n] * newRs[n + 1])
            extra = np.hstack(
                [np.random.randint(0, Is[n + 1], [max(newRs), 1]) for n in range(N - 1)]
                + [np.zeros([max(newRs), 1], dtype=int)]
            )
            for n in range(N - 1):
                if newRs[n + 1] > Rs[n + 1]:
                    add = torch.as_tensor(extra[: newRs[n + 1] - Rs[n + 1], n:], dtype=torch.long, device=device)
                    rsets[n] = torch.cat([rsets[n], add], dim=0)
            Rs = newRs
            t_lin, t_rin = _init_interfaces(tensors, rsets, N, device)

    if val_eps > eps and not suppress_warnings:
        logging.warning("eps={:g} (larger than {}) when cross-approximating {}".format(float(val_eps), eps, function))
    if verbose:
        print("Did {} function evaluations, which took {:.4g}s ({:.4g} evals/s)\n".format(
            info["nsamples"], info["eval_time"], info["nsamples"] / max(info["eval_time"], 1e-12)))

    ret = Tensor([c for c in cores])
    if return_info:
        info["lsets"] = [s.cpu().numpy() for s in lsets]
        info["rsets"] = [s.cpu().numpy() for s in rsets]
        info["Rs"] = Rs
        info["left_locals"] = [l.cpu().numpy() for l in left_locals]
        info["total_time"] = time.time() - start
        info["val_eps"] = val_eps
        return ret, info
    return ret


def cross_forward(info, function=lambda x: x, domain=None, tensors=None, function_arg="vectors", return_info=False):
    """cross.py:532-644: replay of a finished cross run.  Given the index sets `tn.cross(..., return_info=True)`
    recorded (`lsets`, `rsets`, `left_locals`, `Rs`), evaluate the function on those fibres again and assemble the TT
    by the cross-interpolation formula (core_j = V (V[left_local_j])^-1, last core = the fibres themselves) with torch
    operations only, so the result is differentiable w.r.t. whatever `function` closes over."""
    assert domain is not None or tensors is not None
    assert function_arg in ("vectors", "matrix")
    if function_arg == "matrix":
        def f(*args):
            return function(torch.cat([arg[:, None] for arg in args], dim=1))
    else:
        f = function
    device = None
    if tensors is None:
        device = domain[0].device
        if device.type != "cuda":
            device = torch.device("cuda")
        tensors = meshgrid(domain, device=device)
    if not hasattr(tensors, "__len__"):
        tensors = [tensors]
    if device is None:
        device = tensors[0].cores[0].device
    Is = list(tensors[0].shape)
    N = len(Is)
    lsets = [torch.as_tensor(np.asarray(l), dtype=torch.long, device=device) for l in info["lsets"]]
    rsets = [torch.as_tensor(np.asarray(r), dtype=torch.long, device=device) for r in info["rsets"]]
    left_locals = [torch.as_tensor(np.asarray(l), dtype=torch.long, device=device) for l in info["left_locals"]]
    Rs = [int(r) for r in info["Rs"]]
    if return_info:
        info["Xs"] = torch.zeros(0, N)
        info["shapes"] = []
    t_lin, t_rin = _init_interfaces(tensors, rsets, N, device)

    def evaluate_function(j):
        Xs = [torch.einsum("ai,ibj,jc->abc", t_lin[k][j], t.cores[j], t_rin[k][j]).flatten() for k, t in enumerate(tensors)]
        evaluation = f(*Xs)
        if return_info:
            info["Xs"] = torch.cat((info["Xs"], torch.stack(Xs, dim=1).detach().cpu()), dim=0)
            info["shapes"].append([Rs[j], Is[j], Rs[j + 1]])
        return evaluation.reshape(Rs[j], Is[j], Rs[j + 1])

    cores = []
    for j in range(N - 1):
        V = evaluate_function(j).reshape(-1, Rs[j + 1])
        A = V[left_locals[j], :]
        X = torch.linalg.solve(A.t(), V.t()).t()  # lstsq of a square, well-conditioned (maxvol) system: cross.py:620
        cores.append(X.reshape(Rs[j], Is[j], Rs[j + 1]))
        local_r = torch.div(left_locals[j], Is[j], rounding_mode="floor")
        local_i = left_locals[j] - local_r * Is[j]
        lsets[j + 1] = torch.cat([lsets[j][local_r, :], local_i[:, None]], dim=1)
        for k, t in enumerate(tensors):
            t_lin[k][j + 1] = torch.einsum("ai,iaj->aj", t_lin[k][j][local_r, :], t.cores[j][:, local_i, :])
    cores.append(evaluate_function(N - 1))
    out = Tensor(cores)
    return (out, info) if return_info else out
