"""Verbose GPU diagnostics (numbers, not just pass/fail) — run under gpurun, log to gpurun_out/."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from tntorch_b200 import ops

def t_ms(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
print(torch.cuda.get_device_name(0), "tc path:", ops.has_tensorcore_path(), flush=True)

if which in ("all", "blocks"):
    for shape in [(4096, 64), (70000, 96)]:
        A = torch.randn(*shape, device="cuda")
        G = ops.gram(A); ref = A.double().T @ A.double()
        print("gram generic", shape, "relerr", ((G-ref).abs().max()/ref.abs().max()).item(), flush=True)
    for n in (64, 128, 256):
        A = torch.randn(4*n, n, dtype=torch.float64, device="cuda"); G = A.T @ A
        w, V, sw = ops.eigh_jacobi(G, return_sweeps=True); wr = torch.linalg.eigvalsh(G).flip(0)
        print("jacobi", n, "sweeps", sw, "eigerr", ((w-wr).abs().max()/wr[0]).item(), "orth", (V.T@V-torch.eye(n,device='cuda',dtype=torch.float64)).abs().max().item(),
              "ms", t_ms(lambda: ops.eigh_jacobi(G)), flush=True)

if which in ("all", "tc"):
    for shape in [(8192, 128), (10000, 64), (4096, 256), (2048, 2048), (262144, 2048), (1 << 22, 64)]:
        A = torch.randn(*shape, device="cuda")
        G = ops.gram(A, tensorcore=True); torch.cuda.synchronize()
        ref = A.double().T @ A.double()
        err = ((G-ref).abs().max()/ref.diagonal().max()).item()
        ms = t_ms(lambda: ops.gram(A, tensorcore=True))
        flops = 2.0*shape[0]*shape[1]*shape[1]
        print("gram_tc", shape, "err", err, "ms", ms, "TFLOP/s(full)", flops/ms/1e9, "GB/s", shape[0]*shape[1]*4/ms/1e6, flush=True)
        del A, G, ref

if which in ("all", "ttsvd"):
    from oracle import cases
    from gpu_util import relerr64, ranks_of
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ttsvd.npz"))
    for name in ["cfg1_randn16x4_f32", "twin_16x5_f32", "randn32x5_r32_f32", "twin32x5_r32_f32", "randn64x4_r32_f32"]:
        spec = cases.TTSVD_CASES[name]; X = cases.make_dense(spec); Xd = torch.as_tensor(X).cuda()
        for tc in (False, True):
            t0 = time.time(); cores, info = ops.ttsvd(Xd, rmax=spec["ranks_tt"], use_tensorcore=tc, return_info=True); torch.cuda.synchronize(); dt = time.time()-t0
            alg = "svd" if f"{name}/svd/relerr" in g.files else "eig"
            e = relerr64(X, cores)
            print(name, "tc" if tc else "generic", ranks_of(cores), "err", e, "ref", float(g[f"{name}/{alg}/relerr"]), "d", e-float(g[f"{name}/{alg}/relerr"]), info, "wall_s", round(dt,3), flush=True)

if which in ("all", "big"):
    shape = (64,)*5
    X = torch.randn(*shape, device="cuda")
    for tc in (True, False):
        plan = ops.TTSVDPlan(shape, torch.float32, rmax=32, use_tensorcore=tc)
        t0 = time.time(); cores = plan.run(X); torch.cuda.synchronize(); print("first call s", time.time()-t0, flush=True)
        ms = t_ms(lambda: plan.run(X), n=3, warm=1)
        print("64^5 r=32", "tc" if tc else "generic", "ms", ms, "GElem/s", X.numel()/ms/1e6, "info", list(plan.info)[:4], "ranks", list(plan.ranks), flush=True)
        e = ops.tt_relative_error(X, cores)
        print("   device relerr", e, flush=True)
