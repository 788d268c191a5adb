"""-m gpu: the batched decomposition behind the reference API (tn.Tensor(X[B, ...], ranks_tt=r, batch=True),
tensor.py:401-408 with the leading batch dimension) = one tnb_ttsvd_batch call with several samples in flight."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_batch_constructor_matches_per_sample_calls():
    import tntorch_b200 as tnb
    from tntorch_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(11)
    X = torch.randn(5, 24, 20, 16, 18, generator=g, device="cuda")
    t = tnb.Tensor(X, ranks_tt=6, batch=True)
    assert t.batch and [tuple(c.shape) for c in t.cores] == [(5, 1, 24, 6), (5, 6, 20, 6), (5, 6, 16, 6), (5, 6, 18, 1)]
    for b in range(5):
        ref = ops.ttsvd(X[b], rmax=6, batch_mode=True)
        e_b = ops.tt_relative_error(X[b], [c[b] for c in t.cores])
        e_r = ops.tt_relative_error(X[b], ref)
        assert abs(e_b - e_r) < 1e-6
    # fp64 + ragged ranks list
    Xd = torch.randn(3, 9, 8, 7, 6, generator=g, device="cuda", dtype=torch.float64)
    t = tnb.Tensor(Xd, ranks_tt=[3, 5, 4], batch=True)
    assert [tuple(c.shape[1:]) for c in t.cores] == [(1, 9, 3), (3, 8, 5), (5, 7, 4), (4, 6, 1)]
    full = t.torch()
    for b in range(3):
        ref = tnb.Tensor(Xd[b], ranks_tt=[3, 5, 4])
        e_b = ops.tt_relative_error(Xd[b], [c[b] for c in t.cores])
        assert abs(e_b - ops.tt_relative_error(Xd[b], ref.cores)) < 1e-10
        assert abs(float(torch.linalg.vector_norm(full[b] - Xd[b]) / torch.linalg.vector_norm(Xd[b])) - e_b) < 1e-10


def test_batch_with_a_sample_that_needs_the_host_driven_path():
    """one sample is zero outside a slice (its rank rule returns less than the cap): that sample is repeated on the
    host-driven path, the others keep their speculative result."""
    from tntorch_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(12)
    X = torch.randn(4, 12, 10, 8, generator=g, device="cuda", dtype=torch.float64)
    X[2] = 0
    X[2, :, :, 0] = torch.randn(12, 10, generator=g, device="cuda", dtype=torch.float64)
    out, info = ops.ttsvd_batch(X, rmax=4, return_info=True)
    assert info["speculative"] == [1, 1, 0, 1]
    assert [int(c.shape[2]) for c in out[2]] == [4, 1, 1]
    for b in range(4):
        ref = ops.ttsvd(X[b], rmax=4)
        assert [c.shape for c in out[b]] == [c.shape for c in ref]
        assert abs(ops.tt_relative_error(X[b], out[b]) - ops.tt_relative_error(X[b], ref)) < 1e-10


def test_batch_in_flight_is_faster_than_sequential_calls():
    """the in-flight throughput must be reachable through the API, without caller-side threads."""
    from tntorch_b200 import ops

    shape, B = (64, 64, 64, 64, 16), 6  # 1 GiB per tensor
    g = torch.Generator(device="cuda").manual_seed(13)
    X = torch.randn((B,) + shape, generator=g, device="cuda")
    plan = ops.TTSVDBatchPlan(shape, torch.float32, B, rmax=32, inflight=B)
    one = ops.TTSVDPlan(shape, torch.float32, rmax=32)
    for _ in range(2):
        plan.run(X)
        for b in range(B):
            one.run(X[b])
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(3):
        plan.run(X)
    e[1].record()
    for _ in range(3):
        for b in range(B):
            one.run(X[b])
    e[2].record()
    torch.cuda.synchronize()
    t_batch, t_seq = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    print(f"batch of {B} in flight: {t_batch / 3:.2f} ms, {B} sequential calls: {t_seq / 3:.2f} ms, ratio {t_seq / t_batch:.2f}")
    assert list(plan.spec) == [1] * B
    assert t_seq / t_batch > 1.1
