"""-m gpu: BASELINE.json configs 2-4 AT THE NAMED SIZES against outputs of the real reference
(tests/golden/full.npz, written by oracle/gen_golden_full.py in the build container).  Inputs are rebuilt here
from the same NumPy PCG64 streams; the bar is the north star's: relative error within 1e-5 of the reference."""
import os

import numpy as np
import pytest
import torch

from oracle import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "full.npz")


def _gold():
    return np.load(GOLD)


@pytest.mark.parametrize("name", list(cases.FULL_TTSVD_CASES))
def test_ttsvd_64x5_matches_reference(name):
    """config 2 stand-ins (64^5 fp32, TT-rank 32): the structured twin and the random tensor of the bench."""
    from tntorch_b200 import ops

    g = _gold()
    spec = cases.FULL_TTSVD_CASES[name]
    X = torch.from_numpy(cases.make_dense_big(spec)).cuda()
    cores = ops.ttsvd(X, rmax=spec["ranks_tt"])
    assert [1] + [int(c.shape[2]) for c in cores] == list(g[f"{name}/eig/ranks"])
    err = ops.tt_relative_error(X, cores)
    ref = float(g[f"{name}/eig/relerr"])
    assert abs(err - ref) <= 1e-5, (err, ref)
    del X


def _tt_dot(a, b):
    m = torch.ones(1, 1, dtype=torch.float64, device=a[0].device)
    for x, y in zip(a, b):
        m = torch.einsum("ab,aic,bid->cd", m, x.double(), y.double())
    return float(m[0, 0])


def _tt_relerr(cores, out):
    aa, bb, ab = _tt_dot(cores, cores), _tt_dot(out, out), _tt_dot(cores, out)
    return float(np.sqrt(max(aa + bb - 2 * ab, 0.0) / aa))


def test_round_tt_cfg3_as_named():
    """config 3: tn.randn([128]*10, ranks_tt=64) -> round_tt(rmax=16), fp64."""
    import tntorch_b200 as tnb

    g = _gold()
    name = "cfg3_128x10_r64to16_f64"
    spec = cases.FULL_ROUND_CASES[name]
    cores = [torch.as_tensor(c).cuda() for c in cases.make_tt(spec)]
    t = tnb.Tensor(cores)
    t2 = tnb.round_tt(t, rmax=spec["rmax"])
    assert list(t2.ranks_tt) == list(g[f"{name}/svd/ranks"])
    err = _tt_relerr(cores, t2.cores)
    for alg in ("svd", "eig"):
        assert abs(err - float(g[f"{name}/{alg}/relerr"])) <= 1e-7, (alg, err)
    # the input object is untouched (tn.round_tt clones, round.py:7-19)
    assert list(t.ranks_tt) == [1] + [64] * 9 + [1]


def test_round_tt_cfg3_doubled_twin():
    """reference tests/test_round.py:52-59 at config-3 size: (t + t).round_tt(eps=1e-8) returns t's ranks and 2t."""
    import tntorch_b200 as tnb

    g = _gold()
    name = "cfg3_doubled_128x10_r32_f64"
    spec = cases.FULL_ROUND_CASES[name]
    cores = [torch.as_tensor(c).cuda() for c in cases.make_tt(spec)]
    t = tnb.Tensor(cores)
    t.round_tt(eps=spec["eps"])
    assert list(t.ranks_tt) == list(g[f"{name}/svd/ranks"]) == [1] + [32] * 9 + [1]
    assert _tt_relerr(cores, t.cores) <= 1e-7


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_cp_als_R50_matches_fp64_reference(dtype):
    """config 4's rank (R=50) at the CPU-feasible 64^4: 10 sweeps, tol=-inf, fp64 reference trajectory."""
    import tntorch_b200 as tnb

    g = _gold()
    name = "cp_64x4_R50"
    spec = cases.FULL_CP_CASES[name]
    X = torch.as_tensor(cases.make_cp_dense(spec)).cuda()
    t = tnb.Tensor(X.to(dtype), ranks_cp=spec["R"], max_iter=spec["sweeps"], tol=float("-inf"))
    assert [tuple(c.shape) for c in t.cores] == [(s, spec["R"]) for s in spec["shape"]]
    f = [c.double() for c in t.cores]
    rec = torch.einsum("ar,br,cr,dr->abcd", *f)
    err = float(torch.linalg.vector_norm(X - rec) / torch.linalg.vector_norm(X))
    ref = float(g[f"{name}/relerr"])
    assert abs(err - ref) <= 1e-5, (err, ref)


def test_round_tt_cfg3_batch_of_64_throughput():
    """config 3 in batch form ("single instance is latency-bound; batch shows throughput"): 64 random
    128^10 rank-64 trains rounded to rank 16 by ONE tnb_tt_round_batch call (8 in flight, one synchronisation) against 64
    sequential calls; every result equals the single-call result and the reference's error."""
    import time

    from tntorch_b200 import ops

    g = _gold()
    name = "cfg3_128x10_r64to16_f64"
    spec = cases.FULL_ROUND_CASES[name]
    base = [torch.as_tensor(c).cuda() for c in cases.make_tt(spec)]
    B = 64
    gen = torch.Generator(device="cuda").manual_seed(5)
    batch = [base] + [[torch.randn(c.shape, generator=gen, device="cuda", dtype=torch.float64) for c in base] for _ in range(B - 1)]
    ops.tt_round_batch(batch[:8], rmax=16)  # warm-up
    one, info1 = None, None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, info = ops.tt_round_batch(batch, rmax=16, return_info=True)
    torch.cuda.synchronize()
    t_batch = time.perf_counter() - t0
    t0 = time.perf_counter()
    seq = [ops.tt_round(cores, rmax=16) for cores in batch[:16]]
    torch.cuda.synchronize()
    t_seq = (time.perf_counter() - t0) / 16
    ncoef = sum(c.numel() for c in base)
    print(f"round_tt 128^10 r64->16 fp64: batch of {B}: {t_batch * 1e3:.1f} ms = {t_batch / B * 1e3:.2f} ms per tensor "
          f"({B * ncoef / t_batch / 1e9:.2f} Gcoef/s); one call at a time: {t_seq * 1e3:.2f} ms per tensor")
    assert info["speculative"] == [1] * B
    assert [int(c.shape[2]) for c in out[0]] == [16] * 9 + [1]
    assert abs(_tt_relerr(base, out[0]) - float(g[f"{name}/svd/relerr"])) <= 1e-7
    for i in (1, 7, 15):
        assert abs(_tt_relerr(batch[i], out[i]) - _tt_relerr(batch[i], seq[i])) <= 1e-9
    assert t_batch / B < t_seq
