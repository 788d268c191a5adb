import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tntorch_b200 import ops
X = torch.randn((64,) * 5, device="cuda")
for _ in range(2):
    cores, info = ops.ttsvd(X, rmax=32, return_info=True)
print("64^5 info:", info)
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(256, 64, dtype=torch.float64, device="cuda", generator=g); G = A.T @ A
w, V, sw = ops.eigh_jacobi(G, return_sweeps=True); print("random SPD 64 fp64: sweeps", sw, "us", round(t(lambda: ops.eigh_jacobi(G)), 1))
for off in (1e-1, 1e-2, 1e-3, 1e-4):
    D = torch.diag(torch.linspace(2.0, 1.0, 64, dtype=torch.float64, device="cuda"))
    E = torch.randn(64, 64, dtype=torch.float64, device="cuda", generator=g) * off; S = D + (E + E.T) / 2
    w, V, sw = ops.eigh_jacobi(S, return_sweeps=True); print(f"near-diagonal (off {off}) fp64: sweeps", sw, "us", round(t(lambda: ops.eigh_jacobi(S)), 1))
