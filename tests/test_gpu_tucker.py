"""-m gpu: round_tucker / tn.round / Tensor(ranks_tucker=) / Tensor(eps=) against the reference."""
import os

import numpy as np
import pytest
import torch

from oracle import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", list(cases.TUCKER_CASES))
def test_tucker_rounding_matches_reference(name):
    import tntorch_b200 as tnb

    g = np.load(os.path.join(GOLD, "tucker.npz"))
    c = cases.TUCKER_CASES[name]
    if c["kind"] == "dense":
        X = cases.make_dense(c["spec"])
        kw = {k: c[k] for k in ("ranks_tucker", "eps") if k in c}
        t = tnb.Tensor(torch.as_tensor(X).cuda(), **kw)
        dense = X
    else:
        cores = cases.make_tt(c["spec"])
        dense = cases.tt_full(cores)
        t = tnb.Tensor([torch.as_tensor(x).cuda() for x in cores])
        if "round_tucker" in c:
            t = tnb.round_tucker(t, **c["round_tucker"])
        else:
            t = tnb.round(t, **c["round"])
    assert t.ranks_tt.tolist() == list(g[f"{name}/ranks_tt"])
    assert t.ranks_tucker.tolist() == list(g[f"{name}/ranks_tucker"])
    assert list(t.shape) == list(dense.shape)
    rec = t.torch().double().cpu().numpy()
    err = np.linalg.norm(dense - rec) / np.linalg.norm(dense)
    assert abs(err - float(g[f"{name}/relerr"])) <= 1e-5
    for U in t.Us:  # factors are orthonormal (left_ortho=True split)
        if U is not None:
            assert (U.T @ U - torch.eye(U.shape[1], device="cuda", dtype=U.dtype)).abs().max().item() < 1e-8


def test_eps_constructor_on_one_mode_data():
    """Tensor(data, eps=...) on 1-D / 0-D data used to raise (the error kernel needs two modes)."""
    import tntorch_b200 as tnb

    v = torch.arange(7, dtype=torch.float64).cuda()
    t = tnb.Tensor(v, eps=1e-3)
    assert list(t.shape) == [7] and float((t.torch() - v).abs().max()) < 1e-2 * float(v.abs().max())
    s = tnb.Tensor(torch.tensor(3.5).cuda(), eps=1e-3)
    assert abs(float(s.torch().reshape(-1)[0]) - 3.5) < 1e-2
