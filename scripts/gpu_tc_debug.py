import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from tntorch_b200 import ops
    torch.manual_seed(0)
    for shape in [(4096, 128), (4096 + 37, 384), (8192, 64)]:
        rows, n = shape
        i = torch.arange(rows, dtype=torch.float64)[:, None]; j = torch.arange(n, dtype=torch.float64)[None, :]
        A = (((i * 7 + j * 13) % 17) - 8).float().cuda()
        G = ops.gram(A, tensorcore=True); torch.cuda.synchronize()
        ref = A.double().T @ A.double()
        eq = (G == ref).double().mean().item()
        print("variant", os.environ.get("TNB_TC_VARIANT", "0"), shape, "exact-match frac", eq, "maxabs G", G.abs().max().item(), "ref", ref.abs().max().item(),
              "G[0,:4]", G[0, :4].tolist(), "ref[0,:4]", ref[0, :4].tolist(), "G[1,:3]", G[1,:3].tolist(), "ref[1,:3]", ref[1,:3].tolist(), flush=True)
else:
    for v in ("0", "1", "2", "3"):
        env = dict(os.environ, TNB_TC_VARIANT=v)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=300)
        print(r.stdout[-3000:], r.stderr[-1500:], flush=True)
