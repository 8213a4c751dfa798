"""General horizons and MPC time steps: n_steps = n_periods * T_gait / dt of the reference (main.py:20-23,
FootstepPlanner.py:52-63) is not limited to 16 / 32 / 64.  tests/golden/horizon_*.npz were made by running the unmodified
reference MPC.py at N = 8, 10, 12, 24, 32 (dt = 0.01) and 48 (make_golden.py `horizons`).

CPU (-m "not gpu"): the oracle's restated build (numpy and plain C) against what the reference built at these sizes.
GPU (-m gpu): the engine replays the fixtures through the C ABI -- default mode and with every tick forced through the
interior-point stage -- to the usual bars (1e-4 N, 1e-6 relative objective, identical rows holding with equality)."""
import glob
import os

import numpy as np
import pytest

from oracle import c_port, mpc_build

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "horizon_*.npz")))
IDS = [os.path.basename(p)[8:-4] for p in GOLD]


def _params(g):
    n = g["x"].shape[1] // 24
    return n, mpc_build.Params(dt=float(g["dt"]), n_steps=n, T_gait=float(g["T_gait"]))


def test_fixtures_cover_the_capacity_classes():
    ns = sorted(np.load(p)["x"].shape[1] // 24 for p in GOLD)
    assert ns == [8, 10, 12, 24, 32, 48]


@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_oracle_build_matches_reference_at_this_horizon(path):
    g = np.load(path)
    n, p = _params(g)
    m = c_port.MPC(n_steps=n, dt=p.dt)
    assert m.nnz == 126 * n - 18
    for t in range(len(g["ML_data"])):
        first = g["k"][t] == 0
        Pd, A, l, u, contact = mpc_build.build_qp(g["xref"][t], g["fsteps"][t], p, first_tick=first)
        assert np.array_equal(A.indices, g["ML_indices"]) and np.array_equal(A.indptr, g["ML_indptr"])
        np.testing.assert_allclose(A.data, g["ML_data"][t], rtol=0, atol=1e-15)
        np.testing.assert_allclose(u, g["NK"][t], rtol=0, atol=1e-15)
        np.testing.assert_array_equal(Pd, g["P_data"])
        Ap, Ai, Ax, lc, uc = m.build(g["xref"][t], g["fsteps"][t], first_tick=first)
        assert np.array_equal(Ai, g["ML_indices"]) and np.array_equal(Ap, g["ML_indptr"])
        np.testing.assert_allclose(Ax, g["ML_data"][t], rtol=0, atol=1e-15)
        np.testing.assert_allclose(uc, g["NK"][t], rtol=0, atol=1e-15)
    m.close()


@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_golden_optimum_is_certified_at_this_horizon(path):
    """The stored x is the reference's extraction of a KKT-certified optimum of the QP the reference built."""
    import scipy.sparse as sp
    from oracle import kkt
    g = np.load(path)
    n, p = _params(g)
    for t in range(len(g["k"])):
        assert g["cert_prim"][t] <= 1e-9 and g["cert_stat"][t] <= 1e-11 and g["cert_comp"][t] <= 1e-9
        Pd, A, l, u, _ = mpc_build.build_qp(g["xref"][t], g["fsteps"][t], p, first_tick=(g["k"][t] == 0))
        x = g["x"][t]
        Ax = A @ x
        assert (Ax <= u + 1e-9).all() and (Ax >= l - 1e-9).all()
        assert abs(0.5 * x @ (Pd * x) - g["obj"][t]) <= 1e-12 * max(1.0, abs(g["obj"][t]))


@pytest.mark.gpu
@pytest.mark.parametrize("ipm_only", [False, True], ids=["default", "ipm-only"])
@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_engine_replays_reference_at_this_horizon(path, ipm_only):
    import mpcqp
    from common import FORCE_TOL, OBJ_RTOL
    g = np.load(path)
    n, p = _params(g)
    kw = dict(max_sweeps=0) if ipm_only else {}
    eng = mpcqp.Engine(batch=1, n_steps=n, dt=p.dt, T_gait=p.T_gait, **kw)
    for t in range(len(g["k"])):
        eng.run(g["k"][t], g["xref"][t][None], g["fsteps"][t][None])
        f0, x, info = eng.forces()[0], eng.solution()[0], eng.info()
        assert info["status"][0] == 1, (t, info["status"], info["sweeps"], info["iters"])
        assert np.abs(x[12 * n:] - g["x"][t][12 * n:]).max() <= FORCE_TOL
        assert np.abs(f0 - g["f_applied"][t]).max() <= FORCE_TOL
        assert np.abs(x[:12 * n] - g["x"][t][:12 * n]).max() <= 1e-6
        assert abs(info["obj"][0] - g["obj"][t]) <= OBJ_RTOL * abs(g["obj"][t])
        _, A, l, u, contact = mpc_build.build_qp(g["xref"][t], g["fsteps"][t], p, first_tick=(g["k"][t] == 0))
        Ax = (A @ g["x"][t])[24 * n:]
        act = ((np.abs(Ax - u[24 * n:]) <= 1e-9) | (np.abs(Ax - l[24 * n:]) <= 1e-9)).reshape(n, 4, 5)
        np.testing.assert_array_equal(info["active"][0], act)
        np.testing.assert_array_equal(info["contact"][0], contact.astype(bool))
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_build_half_matches_reference_at_this_horizon(path):
    import mpcqp
    g = np.load(path)
    n, p = _params(g)
    T = len(g["ML_data"])
    eng = mpcqp.Engine(batch=T, n_steps=n, dt=p.dt, T_gait=p.T_gait)
    for first in (False, True):
        Bv, Sv, NK = eng.export_build(0.0 if first else 1.0, g["xref"][:T], g["fsteps"][:T])
        for t in range(T):
            if (g["k"][t] == 0) != first:
                continue
            ref_B = np.stack([g["ML_data"][t][g["i_update_B"] + 96 * k] for k in range(n)])
            np.testing.assert_allclose(Bv[t], ref_B, rtol=1e-13, atol=1e-16)
            np.testing.assert_array_equal(Sv[t], g["ML_data"][t][g["i_update_S"]])
            np.testing.assert_allclose(NK[t], g["NK"][t][:12 * n], rtol=0, atol=1e-15)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,dt,T_gait", [(20, 0.02, 0.4), (7, 0.04, 0.32), (40, 0.01, 0.4), (1, 0.02, 0.32)], ids=["N20", "N7", "N40", "N1"])
def test_closed_loop_certified_at_odd_horizons(n, dt, T_gait):
    """Horizons that fill no capacity class (run-time n inside the 16 / 32 / 64 instantiations), mixed gaits, odd batch:
    every robot solved, a sample certified by the oracle on the QP the reference would have built, masks identical."""
    import mpcqp
    from common import OBJ_RTOL, assert_certified, certify
    from scenario import Scenario
    B, T = 33, 6
    eng = mpcqp.Engine(batch=B, n_steps=n, dt=dt, T_gait=T_gait)
    sc = Scenario(B, n_steps=n, dt=dt, T_gait=T_gait, gaits=["trot", "pace", "bound", "walk"], seed=500 + n)
    par = mpc_build.Params(dt=dt, n_steps=n, T_gait=T_gait)
    for t in range(T):
        xref, fsteps = sc.inputs()
        eng.run(t, xref, fsteps)
        x, info = eng.solution(), eng.info()
        assert (info["status"] == 1).all(), (t, info["status"])
        for b in range(0, B, 4):
            cert = certify(xref[b], fsteps[b], x[b], info["y"][b], first_tick=(t == 0), params=par)
            assert_certified(cert, "N=%d tick %d robot %d" % (n, t, b))
            np.testing.assert_array_equal(cert["contact"].astype(bool), info["contact"][b])
            np.testing.assert_array_equal(cert["active"].reshape(n, 4, 5), info["active"][b])
            assert abs(cert["obj"] - info["obj"][b]) <= OBJ_RTOL * abs(cert["obj"])
        sc.advance(x[:, :12] + xref[:, :, 1])
    eng.close()
