"""-m gpu: device maxvol / Householder QR against the reference's golden vectors, and tn.cross through the
drop-in API against the reference's cross on the same seeded RNG streams."""
import os

import numpy as np
import pytest
import torch

from oracle import cases
from oracle import tt_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", list(cases.MAXVOL_CASES))
def test_maxvol_index_sets_bit_exact(name):
    from tntorch_b200 import ops

    g = np.load(os.path.join(GOLD, "maxvol.npz"))
    A = cases.make_matrix(cases.MAXVOL_CASES[name])
    idx, C = ops.maxvol(torch.as_tensor(A).cuda())
    assert idx.cpu().tolist() == list(g[f"{name}/index"])  # integer / index work: bit exact
    assert abs(float(C.abs().max()) - float(g[f"{name}/absmax"])) < 1e-9
    # oracle on the same input
    oi, oC = orc.py_maxvol(A)
    assert idx.cpu().tolist() == list(oi)
    np.testing.assert_allclose(C.cpu().numpy(), oC, atol=1e-9)
    if A.shape[0] > A.shape[1]:  # C = A inv(A[idx])
        sub = A[idx.cpu().numpy()]
        np.testing.assert_allclose(C.cpu().numpy() @ sub, A, atol=1e-9 * np.abs(A).max())


def test_maxvol_batched_matches_single():
    from tntorch_b200 import ops

    rng = np.random.default_rng(5)
    A = torch.as_tensor(rng.standard_normal((16, 320, 10))).cuda()
    idx, C = ops.maxvol(A)
    for b in (0, 7, 15):
        i1, C1 = ops.maxvol(A[b])
        assert torch.equal(idx[b], i1) and torch.equal(C[b], C1)
        oi, _ = orc.py_maxvol(A[b].cpu().numpy())
        assert idx[b].cpu().tolist() == list(oi)


@pytest.mark.parametrize("shape", [(320, 10), (64, 64), (1000, 37), (5, 3), (40, 1)])
def test_householder_qr(shape):
    from tntorch_b200 import ops

    g = torch.Generator().manual_seed(3)
    A = torch.randn(*shape, generator=g, dtype=torch.float64).cuda()
    Q, R = ops.qr(A, return_r=True)
    k = min(shape)
    assert (Q.T @ Q - torch.eye(k, device="cuda", dtype=torch.float64)).abs().max().item() < 1e-13
    assert (Q @ R - A).abs().max().item() < 1e-12 * A.abs().max().item() * shape[0]
    Qr, Rr = torch.linalg.qr(A)  # same LAPACK sign convention
    assert (Q - Qr).abs().max().item() < 1e-10
    # rank-deficient input still yields a full orthonormal basis
    B = A.clone(); B[:, -1] = B[:, 0]
    Qb = ops.qr(B)
    assert (Qb.T @ Qb - torch.eye(k, device="cuda", dtype=torch.float64)).abs().max().item() < 1e-12


@pytest.mark.parametrize("name", list(cases.CROSS_CASES))
def test_cross_matches_reference(name):
    """Same NumPy/torch RNG streams as the reference run that produced the golden vectors."""
    import tntorch_b200 as tnb

    g = np.load(os.path.join(GOLD, "cross.npz"))
    spec = cases.CROSS_CASES[name]
    np.random.seed(spec["seed"])
    torch.manual_seed(spec["seed"])
    domain = [torch.linspace(spec["lo"], spec["hi"], spec["I"], dtype=torch.float64) for _ in range(spec["N"])]
    kw = {k: spec[k] for k in ("ranks_tt", "kickrank", "eps", "max_iter") if k in spec}
    fn = cases.cross_function(spec.get("shift", 0.0), spec.get("fn"))
    if spec.get("minimize"):
        kw["_minimize"] = True
    t, info = tnb.cross(fn, domain=domain, verbose=False, return_info=True, suppress_warnings=True, **kw)
    assert list(info["Rs"]) == list(g[f"{name}/Rs"])
    assert int(info["nsamples"]) == int(g[f"{name}/nsamples"])
    if spec.get("minimize"):  # cross.py:342-359: running minimum and its multi-index
        assert abs(float(info["min"]) - float(g[f"{name}/min"])) <= 1e-12
        assert tuple(int(v) for v in info["argmin"]) == tuple(int(v) for v in g[f"{name}/argmin"])
        return
    ref = float(g[f"{name}/val_eps"])
    got = float(info["val_eps"])
    assert got <= max(10 * ref, 1e-9) and (ref < 1e-6 or abs(got - ref) <= 0.5 * ref)
    if f"{name}/full_relerr" in g.files:
        grids = torch.meshgrid(*[d.cuda() for d in domain], indexing="ij")
        gt = fn(*grids)
        err = float(torch.norm(gt - t.torch()) / torch.norm(gt))
        fref = float(g[f"{name}/full_relerr"])
        assert err <= max(10 * fref, 1e-9)
    if spec.get("forward"):  # cross.py:532-644: the replay reproduces the cross result, and the reference's replay
        tf = tnb.cross_forward(info, fn, domain=domain)
        full = t.torch()
        assert float(torch.norm(tf.torch() - full) / torch.norm(full)) < 1e-9
        ref_fwd = torch.as_tensor(g[f"{name}/forward_full"]).cuda()
        assert float(torch.norm(tf.torch() - ref_fwd) / torch.norm(ref_fwd)) < 1e-6


def test_cross_forward_is_differentiable():
    """cross_forward exists so that the TT depends differentiably on parameters of the black-box function."""
    import tntorch_b200 as tnb

    np.random.seed(3); torch.manual_seed(3)
    dom = [torch.linspace(1, 2, 12, dtype=torch.float64) for _ in range(3)]
    a = torch.tensor(1.5, dtype=torch.float64, device="cuda", requires_grad=True)
    t, info = tnb.cross(lambda x, y, z: 1.0 / (x + y + 1.5 * z), domain=dom, ranks_tt=3, verbose=False, return_info=True,
                        suppress_warnings=True)
    tf = tnb.cross_forward(info, lambda x, y, z: 1.0 / (x + y + a * z), domain=dom)
    loss = sum((c ** 2).sum() for c in tf.cores)
    loss.backward()
    assert a.grad is not None and torch.isfinite(a.grad) and float(a.grad.abs()) > 0


def test_cross_reference_test_suite_cases():
    """tests/test_cross.py:7-17 (1/sum on 10^3, ranks 3) and :33-39 (1/t on a 32^4 grid)."""
    import tntorch_b200 as tnb

    np.random.seed(1); torch.manual_seed(1)
    dom = [torch.linspace(1, 10, 10, dtype=torch.float64) for _ in range(3)]
    t = tnb.cross(lambda x, y, z: 1.0 / (x + y + z), domain=dom, ranks_tt=3, verbose=False, suppress_warnings=True)
    X, Y, Z = torch.meshgrid(*[d.cuda() for d in dom], indexing="ij")
    gt = 1.0 / (X + Y + Z)
    assert float(torch.norm(gt - t.torch()) / torch.norm(gt)) < 5e-2
    dom = [torch.linspace(1, 2, 32, dtype=torch.float64) for _ in range(4)]
    t = tnb.cross(lambda *xs: torch.exp(-sum(xs)), domain=dom, verbose=False, suppress_warnings=True)
    G = torch.meshgrid(*[d.cuda() for d in dom], indexing="ij")
    gt = torch.exp(-sum(G))
    assert float(torch.norm(gt - t.torch()) / torch.norm(gt)) < 1e-4


def test_cross_rejects_invalid_function_values():
    import tntorch_b200 as tnb

    dom = [torch.linspace(0, 1, 8, dtype=torch.float64) for _ in range(3)]
    with pytest.raises(ValueError):  # cross.py:361-375
        tnb.cross(lambda x, y, z: torch.log(x - 0.5), domain=dom, ranks_tt=2, verbose=False, max_iter=1,
                  suppress_warnings=True)
