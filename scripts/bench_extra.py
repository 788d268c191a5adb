"""Timings of the other BASELINE.json configs (informational; the driver's headline is bench.py).
   python scripts/bench_extra.py [cfg3] [cfg4] [cfg5]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tntorch_b200 as tnb
from tntorch_b200 import ops

def ev_ms(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]
out = {}
if "cfg3" in which:  # tn.round_tt on random TT 128^10 rank 64 -> 16, fp64
    g = torch.Generator(device="cuda").manual_seed(0)
    rs = [1] + [64] * 9 + [1]
    cores = [torch.randn(rs[k], 128, rs[k + 1], generator=g, device="cuda", dtype=torch.float64) for k in range(10)]
    t = tnb.Tensor(cores)
    ms = ev_ms(lambda: tnb.round_tt(t, rmax=16), n=3)
    t2 = tnb.round_tt(t, rmax=16)
    coef = sum(c.numel() for c in cores)
    out["cfg3_round_tt_128^10_r64to16_f64"] = {"ms": ms, "Mcoef_per_s": coef / ms / 1e3, "ranks": t2.ranks_tt.tolist()}
    # batch of 64 such tensors through ONE tnb_tt_round_batch call (8 in flight, one synchronisation)
    B = 64
    batch = [[torch.randn(c.shape, generator=g, device="cuda", dtype=torch.float64) for c in cores] for _ in range(B)]
    msb = ev_ms(lambda: ops.tt_round_batch(batch, rmax=16), n=3)
    out["cfg3_batch64"] = {"ms_total": msb, "ms_per_tensor": msb / B, "Mcoef_per_s": B * coef / msb / 1e3}
    del batch
    print(json.dumps(out), flush=True)
if "cfg4" in which:  # CP-ALS R=50 on a synthetic rank-50 256^4 (16 GiB fp32); a few sweeps timed, per-sweep reported
    shape = (256,) * 4
    g = torch.Generator(device="cuda").manual_seed(1)
    fs = [torch.randn(s, 50, generator=g, device="cuda") for s in shape]
    X = torch.einsum("ar,br,cr,dr->abcd", *fs)
    X += 1e-2 * X.std() * torch.randn(shape, generator=g, device="cuda")
    del fs
    ops.cp_als(X, 50, max_iter=1, tol=float("-inf"))  # warm-up: allocates the workspace (the transposed copy is 16 GiB)
    def timed(sweeps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        fi = ops.cp_als(X, 50, max_iter=sweeps, tol=float("-inf"), return_info=True)
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / 1e3, fi[1]
    d1 = min(timed(1)[0] for _ in range(3))  # the HOSVD initialisation (1.1 s) dominates both: take the fastest of 3
    dt, info = min((timed(21) for _ in range(3)), key=lambda t: t[0])
    out["cfg4_cp_als_256^4_R50_f32"] = {"s_per_sweep": (dt - d1) / 20, "init_plus_1_sweep_s": d1, "21_sweeps_s": dt, "errors": info["errors"][:5]}
    print(json.dumps(out), flush=True)
    del X
if "cfg5" in which:  # TT-cross 32^6, ranks 10, 3 sweeps (the unit of BASELINE configs[4]); sequential problems
    dom = [torch.linspace(0, 1, 32, dtype=torch.float64) for _ in range(6)]
    np.random.seed(0); torch.manual_seed(0)
    nprob = 4
    torch.cuda.synchronize(); t0 = time.time()
    for b in range(nprob):
        sh = 1.0 + b / 512.0
        t, info = tnb.cross(lambda *xs: 1.0 / (sh + sum(xs)), domain=dom, ranks_tt=10, max_iter=3, verbose=False,
                            return_info=True, suppress_warnings=True)
    torch.cuda.synchronize(); dt = (time.time() - t0) / nprob
    out["cfg5_cross_32^6_r10_3sweeps_f64"] = {"s_per_problem": dt, "evals_per_s": info["nsamples"] / dt, "val_eps": float(info["val_eps"])}
    # B = 512 problems advanced together (tntorch_b200.cross_batch), three full sweeps (eps = 0)
    def family(pid, *xs):
        s = 1.0 + pid.double() / 512
        for x in xs:
            s = s + x
        return 1.0 / s
    tnb.cross_batch(family, dom, batch=8, batched_function=True, ranks_tt=10, max_iter=1, eps=0.0)
    torch.cuda.synchronize(); t0 = time.time()
    tb, infob = tnb.cross_batch(family, dom, batch=512, batched_function=True, ranks_tt=10, max_iter=3, eps=0.0, return_info=True)
    torch.cuda.synchronize(); dtb = time.time() - t0
    out["cfg5_cross_batch512"] = {"s_total": dtb, "problems_per_s": 512 / dtb, "evals_per_s": float(infob["nsamples"].sum()) / dtb,
                                  "max_val_eps": float(infob["val_eps"].max())}
    print(json.dumps(out), flush=True)
