"""Golden outputs of the REAL reference at the full sizes BASELINE.json names (configs 2-4).

Run once in the build container (needs ~30 GB RAM, a few minutes on 8 cores):

    cd /tmp && python /root/repo/oracle/gen_golden_full.py [ttsvd] [round] [cp]

Writes ``tests/golden/full.npz`` (merged with what is already there).  The inputs come from
``oracle/cases.py`` (NumPy PCG64 streams), so the GPU tests rebuild them bit-for-bit and only the
reference's ranks / relative errors are stored.  Test infrastructure, like everything under oracle/.
"""
import os
import sys
import time
import warnings

import numpy as np

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, REPO)

import torch  # noqa: E402
import tntorch as tn  # noqa: E402  (the real reference)

from oracle import cases  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden", "full.npz")
torch.set_num_threads(os.cpu_count())


def chunked_relerr(X, cores):
    """||X - TT(cores)|| / ||X|| in fp64, one slab of the first mode at a time."""
    c = [np.asarray(x.detach().numpy() if hasattr(x, "detach") else x, dtype=np.float64) for x in cores]
    rest = cases.tt_full_matrix(c[1:], left_open=True)  # r1 x prod(I[1:])
    num = den = 0.0
    Xf = X.reshape(X.shape[0], -1)
    for i in range(X.shape[0]):
        slab = c[0][0, i, :] @ rest
        x = Xf[i].astype(np.float64)
        num += float(np.sum((x - slab) ** 2))
        den += float(np.sum(x ** 2))
    return float(np.sqrt(num / den))


def tt_dot(a, b):
    m = np.ones((1, 1))
    for x, y in zip(a, b):
        m = np.einsum("ab,aic,bid->cd", m, np.asarray(x, np.float64), np.asarray(y, np.float64), optimize=True)
    return float(m[0, 0])


def tt_relerr(cores, out):
    aa, bb, ab = tt_dot(cores, cores), tt_dot(out, out), tt_dot(cores, out)
    return float(np.sqrt(max(aa + bb - 2 * ab, 0.0) / aa))


def gen_ttsvd(out):
    for name, spec in cases.FULL_TTSVD_CASES.items():
        X = cases.make_dense_big(spec)
        t0 = time.perf_counter()
        t = tn.Tensor(torch.from_numpy(X), ranks_tt=spec["ranks_tt"], algorithm="eig")
        dt = time.perf_counter() - t0
        out[f"{name}/eig/ranks"] = np.asarray(t.ranks_tt, dtype=np.int64)
        out[f"{name}/eig/relerr"] = np.float64(chunked_relerr(X, t.cores))
        out[f"{name}/eig/seconds_{os.cpu_count()}cores"] = np.float64(dt)
        print(name, list(t.ranks_tt), out[f"{name}/eig/relerr"], f"{dt:.1f}s", flush=True)
        del X, t


def gen_round(out):
    for name, spec in cases.FULL_ROUND_CASES.items():
        cores = cases.make_tt(spec)
        for alg in ("svd", "eig"):
            t = tn.Tensor([torch.as_tensor(c.copy()) for c in cores])
            kw = {k: spec[k] for k in ("eps", "rmax") if k in spec}
            t0 = time.perf_counter()
            t.round_tt(algorithm=alg, **kw)
            dt = time.perf_counter() - t0
            res = [c.numpy() for c in t.cores]
            out[f"{name}/{alg}/ranks"] = np.asarray(t.ranks_tt, dtype=np.int64)
            out[f"{name}/{alg}/relerr"] = np.float64(tt_relerr(cores, res))
            out[f"{name}/{alg}/seconds_{os.cpu_count()}cores"] = np.float64(dt)
            print(name, alg, list(t.ranks_tt), out[f"{name}/{alg}/relerr"], f"{dt:.2f}s", flush=True)


def gen_cp(out):
    for name, spec in cases.FULL_CP_CASES.items():
        X = cases.make_cp_dense(spec)
        torch.manual_seed(0)
        t0 = time.perf_counter()
        t = tn.Tensor(torch.as_tensor(X), ranks_cp=spec["R"], max_iter=spec["sweeps"], tol=float("-inf"))
        dt = time.perf_counter() - t0
        Xt = torch.as_tensor(X)
        out[f"{name}/relerr"] = np.float64(float(torch.norm(Xt - t.torch()) / torch.norm(Xt)))
        out[f"{name}/seconds_{os.cpu_count()}cores"] = np.float64(dt)
        print(name, out[f"{name}/relerr"], f"{dt:.1f}s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["round", "cp", "ttsvd"]
    out = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    for w in which:
        {"ttsvd": gen_ttsvd, "round": gen_round, "cp": gen_cp}[w](out)
        np.savez_compressed(OUT, **out)
