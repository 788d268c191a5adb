"""Golden vectors for the §8(f)-3 callers (hadamard_sum, shift_mode, TTMatrix) from the REAL reference.

Build container only:   cd /tmp && python /root/repo/oracle/gen_callers.py   ->  tests/golden/callers.npz
Inputs come from oracle/cases.py; only the reference's outputs are stored."""
import os
import sys
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

out = {}
for name, spec in cases.HADAMARD_SUM_CASES.items():
    ops64 = cases.hadamard_operands(spec)
    ts = [tn.Tensor([torch.tensor(c) for c in cores]) for cores in ops64]
    # the reference allocates its interface / diagonal cores in torch's default dtype (metrics.py:365,402)
    torch.set_default_dtype(torch.float64)
    exact = tn.hadamard_sum(ts, algorithm="exact")
    approx = tn.hadamard_sum(ts, algorithm="svd", eps=1e-8)
    torch.set_default_dtype(torch.float32)
    dense = np.prod(np.stack([cases.tt_full(c) for c in ops64]), axis=0).sum()
    out[f"{name}/exact"] = np.float64(exact)
    out[f"{name}/dense"] = np.float64(dense)
    out[f"{name}/svd"] = np.float64(approx)
    print(name, exact, dense, approx, flush=True)

for name, (spec, n, shift, eps) in cases.SHIFT_MODE_CASES.items():
    cores = cases.shift_mode_input(spec)
    t = tn.Tensor([torch.tensor(c) for c in cores])
    tn.shift_mode(t, n, shift, eps=eps)
    out[f"{name}/full"] = t.torch().numpy()
    out[f"{name}/ranks"] = t.ranks_tt.numpy().astype(np.int64)
    print(name, list(t.shape), t.ranks_tt.tolist(), flush=True)

for name, spec in cases.TTMATRIX_CASES.items():
    M = torch.tensor(cases.ttmatrix_input(spec))
    kw = dict(ranks=spec["ranks"], input_dims=spec["input_dims"], output_dims=spec["output_dims"])
    if spec["batch"]:
        # the reference's batched constructor does not run on this torch (matrix.py:73 builds a tensor from a list that
        # mixes an int with tuples); a batch is per-sample TT-SVD with the same rank caps, so the golden is per sample
        per = [tn.TTMatrix(M[b], **kw) for b in range(spec["batch"])]
        out[f"{name}/full"] = np.stack([p.torch().numpy() for p in per])
        ttm = per[0]
    else:
        ttm = tn.TTMatrix(M, **kw)
        out[f"{name}/full"] = ttm.torch().numpy()
    out[f"{name}/ranks"] = np.asarray(ttm.ranks, dtype=np.int64)
    out[f"{name}/core_shapes"] = np.asarray([list(c.shape)[-4:] for c in ttm.cores], dtype=np.int64)
    print(name, [tuple(c.shape) for c in ttm.cores], float(torch.dist(ttm.torch(), M if not spec["batch"] else M[0]) / torch.norm(M if not spec["batch"] else M[0])), flush=True)

np.savez_compressed(os.path.join(REPO, "tests", "golden", "callers.npz"), **out)
