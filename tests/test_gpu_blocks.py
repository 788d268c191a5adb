"""-m gpu: building-block kernels through the C-ABI against fp64 references."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from tntorch_b200 import ops

    return ops


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("shape", [(1000, 7), (4096, 64), (333, 130), (70000, 96), (5, 300)])
def test_gram_generic(dtype, shape):
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    A = torch.randn(*shape, generator=g, dtype=dtype).cuda()
    G = ops.gram(A)
    ref = A.double().T @ A.double()
    tol = 1e-12 if dtype == torch.float64 else 1e-12  # fp64 accumulation of exact fp32 products
    assert (G - ref).abs().max().item() <= tol * ref.abs().max().item() * 50
    assert torch.equal(G, G.T)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_project(dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    A = torch.randn(5000, 96, generator=g, dtype=dtype).cuda()
    V = torch.randn(96, 17, generator=g, dtype=dtype).cuda()
    C = ops.project(A, V)
    ref = A.double() @ V.double()
    tol = 1e-5 if dtype == torch.float32 else 1e-13
    assert (C.double() - ref).abs().max().item() <= tol * ref.abs().max().item()


@pytest.mark.parametrize("n", [1, 2, 5, 64, 63, 104, 105, 128, 200, 256])
def test_eigh_jacobi(n):
    ops = _ops()
    g = torch.Generator().manual_seed(n)
    A = torch.randn(3 * n + 2, n, generator=g, dtype=torch.float64)
    if n >= 64:
        A = A * torch.logspace(0, -5, n, dtype=torch.float64)[None, :]
    G = (A.T @ A).cuda()
    w, V = ops.eigh_jacobi(G)
    wr = torch.linalg.eigvalsh(G).flip(0)
    assert (w - wr).abs().max().item() <= 1e-12 * wr[0].item()
    assert (V.T @ V - torch.eye(n, device="cuda", dtype=torch.float64)).abs().max().item() < 1e-12
    assert (G @ V - V * w[None, :]).abs().max().item() <= 1e-11 * wr[0].item()
    assert torch.all(w[:-1] >= w[1:])


@pytest.mark.parametrize("kind", ["flat", "decay", "lowrank"])
def test_eig_topk(kind):
    ops = _ops()
    rng = np.random.default_rng(0)
    n, r = 768, 24
    if kind == "flat":
        A = rng.standard_normal((8 * n, n))
    elif kind == "decay":
        A = rng.standard_normal((2 * n, n)) * np.logspace(0, -6, n)[None, :]
    else:
        A = rng.standard_normal((4 * n, r)) @ rng.standard_normal((r, n)) + 1e-3 * rng.standard_normal((4 * n, n))
    G = A.T @ A
    w_ref = np.linalg.eigvalsh(G)[::-1]
    w, V, info = ops.eig_topk(torch.as_tensor(G).cuda(), r)
    V = V[:, :r].cpu().numpy()
    assert np.abs(V.T @ V - np.eye(r)).max() < 1e-9
    deficit = (w_ref[:r].sum() - np.trace(V.T @ G @ V)) / np.trace(G)
    assert deficit < 1e-6, (deficit, info)
    assert info["converged"] == 1
