"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol declared in include/tnb200.h; compute entry points fail loudly without a GPU."""
import ctypes
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(REPO, "include", "tnb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tnb_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    from tntorch_b200.csrc import build

    path = build.build()
    assert os.path.exists(path)
    h = ctypes.CDLL(path)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/tnb200.h but not exported"


def test_binding_table_matches_header():
    from tntorch_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    assert _lib.lib().tnb_version() >= 100


def test_sizing_queries_run_without_gpu():
    from tntorch_b200 import _lib

    L = _lib.lib()
    sh = _lib.i64([64, 64, 64, 64, 64])
    rm = _lib.i32([32, 32, 32, 32])
    offs = (ctypes.c_int64 * 5)()
    cap = L.tnb_ttsvd_cores_capacity(5, sh, rm, offs)
    assert cap >= 64 * 32 + 3 * 32 * 64 * 32 + 32 * 64
    assert list(offs)[0] == 0 and all(offs[i] < offs[i + 1] for i in range(4))
    ws = L.tnb_ttsvd_workspace_bytes(_lib.TNB_F32, 5, sh, rm, 0)
    assert ws > 2 * 2**30  # at least the 2 GiB first carry
    assert L.tnb_ttsvd_cores_capacity(0, sh, rm, offs) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import tntorch_b200 as tnb
    from tntorch_b200 import _lib

    with pytest.raises(RuntimeError):
        tnb.Tensor(torch.randn(4, 4, 4), ranks_tt=2)
    with pytest.raises(RuntimeError):
        tnb.truncated_svd(torch.randn(4, 4))
    # straight through the C-ABI: must report an error, not compute
    L = _lib.lib()
    buf = (ctypes.c_double * 16)()
    rc = L.tnb_eigh_jacobi(ctypes.addressof(buf), 4, ctypes.addressof(buf), ctypes.addressof(buf),
                           ctypes.addressof(buf), 128, None)
    assert rc != 0
    assert b"no CUDA device" in L.tnb_last_error() or rc == _lib.ERR_CUDA


def test_reference_error_behaviour_at_the_boundary():
    import tntorch_b200 as tnb

    with pytest.raises(ValueError):  # round.py:77-78
        tnb.truncated_svd(torch.zeros(3, 3), delta=1.0, eps=1.0)
    with pytest.raises(ValueError):  # tensor.py:436-438 (checked before any device work)
        tnb.Tensor([torch.zeros(2, 3, 2), torch.zeros(3, 3, 1)])  # core ranks do not match (tensor.py:177-191)
