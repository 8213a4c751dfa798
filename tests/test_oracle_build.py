"""oracle/mpc_build.py (restated QP build) against what the UNMODIFIED reference MPC.py built
(tests/golden/solve_*.npz, made by tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import mpc_build

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "solve_*.npz")))


def test_golden_present():
    assert len(GOLD) >= 8


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[6:-4] for p in GOLD])
def test_build_matches_reference(path):
    g = np.load(path)
    N = g["x"].shape[1] // 24
    p = mpc_build.Params(n_steps=N)
    for t in range(len(g["ML_data"])):          # long-horizon fixtures keep the matrices of the first ticks only
        Pd, A, l, u, contact = mpc_build.build_qp(g["xref"][t], g["fsteps"][t], p, first_tick=(g["k"][t] == 0))
        assert np.array_equal(A.indices, g["ML_indices"]) and np.array_equal(A.indptr, g["ML_indptr"])
        assert A.nnz == 126 * N - 18
        np.testing.assert_allclose(A.data, g["ML_data"][t], rtol=0, atol=1e-15)
        np.testing.assert_allclose(u, g["NK"][t], rtol=0, atol=1e-15)
        fin = np.isfinite(g["NK_inf"][t])
        assert np.array_equal(np.isfinite(l), fin)
        np.testing.assert_allclose(l[fin], g["NK_inf"][t][fin], rtol=0, atol=1e-15)
        np.testing.assert_array_equal(Pd, g["P_data"])
        # the scatter maps the reference uses each tick (MPC.py:154-166)
        iB = g["i_update_B"][None, :] + 96 * np.arange(N)[:, None]
        assert iB.max() < A.nnz and g["i_update_S"].max() < A.nnz
        np.testing.assert_array_equal(A.data[g["i_update_S"]], 1.0 - np.repeat(contact.reshape(-1), 3))


def test_warm_start_shift_matches_reference():
    g = np.load(GOLD[-1] if "trot" in GOLD[-1] else [p for p in GOLD if p.endswith("solve_trot.npz")][0])
    N = 16
    for t in range(1, len(g["k"])):
        np.testing.assert_array_equal(mpc_build.shift_warm_start(g["x"][t - 1], N), g["warm_x"][t])


def test_extract_matches_reference():
    g = np.load([p for p in GOLD if p.endswith("solve_trot.npz")][0])
    for t in range(len(g["k"])):
        f, xr = mpc_build.extract(g["x"][t], g["xref"][t], 16)
        np.testing.assert_array_equal(f, g["f_applied"][t])
        np.testing.assert_allclose(xr, g["x_robot"][t], rtol=0, atol=1e-15)


def test_contact_table_edge_cases():
    fs = np.full((20, 13), np.nan)
    fs[:, 0] = 0
    c, rows = mpc_build.contact_table(fs, 16)                      # empty table: nobody in contact
    assert not c.any() and (rows == -1).all()
    fs[0, 0] = 20                                                   # phase longer than the horizon is clipped
    fs[0, 1:] = 0.1
    fs[0, 4] = 0.0                                                  # x == 0.0 counts as swing (MPC.py:650)
    c, rows = mpc_build.contact_table(fs, 16)
    assert c.shape == (16, 4) and (c[:, 0] == 1).all() and (c[:, 1] == 0).all() and (rows == 0).all()
