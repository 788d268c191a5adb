import sys; sys.path.insert(0,'.')
import numpy as np, torch
import tntorch_b200 as tnb
from tntorch_b200 import ops
rng = np.random.default_rng(9)
fac = [rng.standard_normal((s, 3)) for s in (14, 12, 10)]
X = torch.as_tensor(np.einsum("ar,br,cr->abc", *fac)).cuda()
for trial in range(3):
    torch.manual_seed(trial)
    t = tnb.Tensor(X, ranks_cp=3, ranks_tucker=4, max_iter=200, tol=1e-12)
    print('trial', trial, 'full err', float(torch.linalg.vector_norm(X - t.torch())/torch.linalg.vector_norm(X)), [tuple(c.shape) for c in t.cores])
    tt = tnb.Tensor(X, ranks_tucker=4)
    core = tt.tucker_core().contiguous()
    torch.manual_seed(trial)
    init = [torch.randn(sh, 3, dtype=core.dtype, device='cuda') for sh in core.shape]
    f, info = ops.cp_als(core, 3, max_iter=200, tol=1e-12, init=init, return_info=True)
    rec = torch.einsum("ar,br,cr->abc", *f)
    print('   direct: iters', info['iters'], 'errors', [round(e,6) for e in info['errors'][:4]], info['errors'][-2:], 'true', float(torch.linalg.vector_norm(core-rec)/torch.linalg.vector_norm(core)))
    # evaluate the CP-Tucker tensor by hand
    full = torch.einsum("ar,br,cr->abc", *[U @ c for U, c in zip(tt.Us, f)])
    print('   by hand', float(torch.linalg.vector_norm(X - full)/torch.linalg.vector_norm(X)))
