"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/txh.h declares;
compute entry points refuse to run without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "txh.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(txh_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(libtxh):
    from tx_fast_hydrology_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(libtxh, name), f"{name} declared in include/txh.h but not exported"
    assert set(_lib.SIGNATURES) == set(names), "ctypes binding and header disagree"
    assert libtxh.txh_version() >= 100
    assert libtxh.txh_row_stride(1) == 2 and libtxh.txh_row_stride(64) == 64 and libtxh.txh_row_stride(65) == 66


def test_no_cpu_fallback(libtxh):
    """Without a CUDA device the routing call fails loudly with TXH_E_NODEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from tx_fast_hydrology_b200._lib import TxhError
    from tx_fast_hydrology_b200.network import RiverNetwork
    assert libtxh.txh_device_count() == 0
    net = RiverNetwork(np.array([1, 2, 2], dtype=np.int64))
    net.compute_coeffs(np.full(3, 600.0), np.full(3, 0.2), 300.0)
    buf = (ctypes.c_double * 8)()
    rc = libtxh.txh_route_step(net.handle, buf, buf, 1, None, None)
    assert rc == -4 and b"no CPU fallback" in libtxh.txh_last_error()
    # ... and so do the fused entry points (dense filter chain, in-library assimilation loop)
    obs = (ctypes.c_int64 * 1)(1)
    z = (ctypes.c_double * 1)(0.0)
    rc = libtxh.txh_kf_filter(net.handle, buf, buf, None, buf, buf, obs, 1, z, buf, buf, buf, buf, buf, buf, None)
    assert rc == -4
    rc = libtxh.txh_run_assimilating(net.handle, buf, buf, 1, None, 0, 1, 2, 1, 1, obs, 1, buf, buf, buf, buf, 1,
                                     buf, buf, buf, buf, buf, buf, 0, None, None)
    assert rc == -4
    with pytest.raises((TxhError, RuntimeError)):
        from tx_fast_hydrology_b200.muskingum import Muskingum
        from tx_fast_hydrology_b200 import synthetic as S
        nd = S.make_network(10, 1)
        m = Muskingum(S.model_dict(nd, S.make_params(10, 1)))
        m.step(np.zeros(10))


def test_product_never_imports_oracle():
    """The product package must not reference the test oracle."""
    pkg = os.path.join(ROOT, "tx_fast_hydrology_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "txh_oracle" not in txt, f
