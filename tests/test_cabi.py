"""The C ABI: the built library exports every symbol include/mpcqp.h declares, the ctypes mirror of
mpcqp_params has the header's layout, and there is no CPU fallback (no compute calls here)."""
import ctypes
import os
import re

import pytest

import mpcqp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mpcqp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpcqp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = mpcqp.load()
    syms = _header_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(lib, s), "libmpcqp.so does not export %s" % s
    assert sorted(mpcqp.EXPORTS) == syms


def test_params_layout_matches_header():
    p = mpcqp.default_params()
    assert p.struct_size == ctypes.sizeof(mpcqp.Params)
    assert (p.n_steps, p.dt, p.T_gait, p.mu, p.fz_max) == (16, 0.02, 0.32, 0.9, 25.0)
    assert abs(p.mass - 2.50000279) < 1e-15 and abs(p.gI[4] - 5.106100e-2) < 1e-18
    assert abs(p.w_state[6] - 2 * 0.1 ** 0.5) < 1e-15 and p.w_force == 1e-5
    assert p.mode == mpcqp.MODE_ACTIVE_SET | mpcqp.MODE_STAGEWISE | mpcqp.MODE_IPM and p.max_sweeps == 16 and p.ipm_max_iter == 60
    assert b"sm_100a" in mpcqp.load().mpcqp_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the failure path is for GPU-less hosts")
    with pytest.raises(mpcqp.MpcqpError) as e:
        mpcqp.Engine(batch=2)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    import MPC
    with pytest.raises(mpcqp.MpcqpError):
        MPC.MPC(0.02, 16, 0.32, batch=1)


def test_bad_arguments_are_rejected_before_touching_the_gpu():
    lib = mpcqp.load()
    p = mpcqp.default_params()
    h = ctypes.c_void_p()
    p.struct_size = 8
    assert lib.mpcqp_create(ctypes.byref(p), ctypes.byref(h)) == -1
    assert b"struct_size" in lib.mpcqp_last_error()
    p = mpcqp.default_params(n_steps=65)            # horizons 1 .. 64
    assert lib.mpcqp_create(ctypes.byref(p), ctypes.byref(h)) == -1
    p = mpcqp.default_params(n_steps=24, mode=3)    # the dense stages exist for 16 and 32 steps only
    assert lib.mpcqp_create(ctypes.byref(p), ctypes.byref(h)) == -1
    assert b"dense" in lib.mpcqp_last_error()
    p = mpcqp.default_params(batch=0)
    assert lib.mpcqp_create(ctypes.byref(p), ctypes.byref(h)) == -1
    assert lib.mpcqp_run(None, 0.0, None, None, 0) == -1
