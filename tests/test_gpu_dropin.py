"""The reference-facing Python API (MPC / MPC_Wrapper / MPC_Virtual drop-ins) on a B200, replaying the
golden run of the reference's own MPC.py."""
import os

import numpy as np
import pytest

from common import FORCE_TOL

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class _Planner:
    """what processing.process_mpc hands to mpc_wrapper.solve (processing.py:142)"""
    def __init__(self, xref, fsteps):
        self.xref, self.fsteps = xref, fsteps


def test_mpc_class_matches_reference_run():
    import MPC
    g = np.load(os.path.join(G, "solve_trot.npz"))
    mpc = MPC.MPC(0.02, 16, 0.32)
    assert mpc.n_steps == 16 and abs(mpc.h_ref - 0.2027682) < 1e-15 and mpc.x.shape == (384,)
    np.testing.assert_array_equal(mpc.P.data, g["P_data"])
    for t in range(len(g["k"])):
        fs = g["fsteps"][t].copy()
        assert mpc.run(g["k"][t], g["xref"][t], fs) == 0
        assert np.array_equal(np.isnan(fs), np.isnan(g["fsteps"][t]))        # caller's fsteps untouched
        assert mpc.f_applied.shape == (12,) and mpc.x.shape == (384,) and mpc.x_robot.shape == (12, 16)
        assert np.abs(mpc.f_applied - g["f_applied"][t]).max() <= FORCE_TOL
        assert np.abs(mpc.x - g["x"][t]).max() <= FORCE_TOL
        assert np.abs(mpc.x_robot - g["x_robot"][t]).max() <= 1e-6
        assert mpc.q_next.shape == (6, 1) and mpc.v_next.shape == (6, 1)
        np.testing.assert_allclose(mpc.q_next[:, 0], g["x_robot"][t][0:6, 0], atol=1e-6)
    # keyword form of the north-star paraphrase run(k, T_gait, fsteps, xref)
    assert mpc.run(k=1.0, T_gait=0.32, fsteps=g["fsteps"][3], xref=g["xref"][3]) == 0


def test_wrapper_and_virtual_batched():
    import MPC_Virtual
    names = ["trot_turn", "pace", "bound", "walk"]
    gs = [np.load(os.path.join(G, "solve_%s.npz" % n)) for n in names]
    mv = MPC_Virtual.MPC_Virtual(True, 0.02, 16, 20, 0.32)
    for t in range(8):
        planner = _Planner(np.stack([g["xref"][t] for g in gs]), np.stack([g["fsteps"][t] for g in gs]))
        assert mv.solve(20 * t, planner) == 0              # k counts TSID ticks, k / k_mpc reaches MPC.run
        f = mv.get_latest_result()
        if t == 0:
            np.testing.assert_array_equal(f, np.tile([0.0, 0.0, 8.0], (4, 4)))     # MPC_Wrapper.py:76-78
            f = mv.get_latest_result()
        assert f.shape == (4, 12)
        for b, g in enumerate(gs):
            assert np.abs(f[b] - g["f_applied"][t]).max() <= FORCE_TOL
    info = mv.solver.mpc.info
    assert (info["status"] == 1).all()


def test_asynchronous_wrapper_pipeline():
    """multiprocessing=True is a stream pipeline here (the reference's process-based version is dead code,
    MPC_Wrapper.py:48-51, 116-260).  solve() returns without waiting; get_latest_result() never blocks and hands out the
    newest forces that have arrived; after wait() they are those of the last tick, equal to the synchronous wrapper's."""
    import MPC_Wrapper
    g = np.load(os.path.join(G, "solve_trot.npz"))
    sync = MPC_Wrapper.MPC_Wrapper(0.02, 16, 20, 0.32, multiprocessing=False)
    asyn = MPC_Wrapper.MPC_Wrapper(0.02, 16, 20, 0.32, multiprocessing=True)
    seen = []
    for t in range(10):
        planner = _Planner(g["xref"][t], g["fsteps"][t])
        sync.solve(20 * t, planner)
        fs = sync.get_latest_result()
        if t == 0:
            fs = sync.get_latest_result()
        assert asyn.solve(20 * t, planner) == 0
        fa = asyn.get_latest_result()                       # tick 0: the reference's [0, 0, 8] x 4; later: whatever has landed
        if t == 0:
            np.testing.assert_array_equal(fa, np.array([0.0, 0.0, 8.0] * 4))
        assert fa.shape == (12,) and np.isfinite(fa).all()
        seen.append(fa.copy())
        fw = asyn.wait()                                    # everything issued so far has landed
        assert np.abs(fw - fs).max() <= 1e-9, t
        assert np.abs(fw - g["f_applied"][t]).max() <= FORCE_TOL
    # the non-blocking answers are always forces of SOME earlier-or-current tick
    for t, fa in enumerate(seen[1:], start=1):
        assert min(np.abs(fa - g["f_applied"][u]).max() for u in range(t + 1)) <= FORCE_TOL


def test_logger_cost_components():
    """Logger.log_cost_function (Logger.py:406-418) on the device, against its numpy formula on MPC.x."""
    import MPC
    g = np.load(os.path.join(G, "solve_walk.npz"))
    m = MPC.MPC(0.02, 16, 0.32)
    m.run(0, g["xref"][:3], g["fsteps"][:3])                # three robots at once
    cost = m._engine.cost_components()
    P = m.P.data
    for b in range(3):
        c = m.x[b] * P * m.x[b]                             # diag(x) diag(P) x
        ref = np.array([c[i:12 * 16:12].sum() for i in range(12)] + [c[12 * 16:].sum()])
        np.testing.assert_allclose(cost[b], ref, rtol=1e-12, atol=1e-15)


def test_world_pose_dead_reckoning_matches_reference():
    """MPC.q_w (MPC.py:58, 503-510) against the reference's own value, tick by tick, on a turning trot (N = 24, a horizon that
    fills no capacity class): single robot and the same robot twice in a batch.  Results are fetched lazily: q_w must come out
    the same whether or not anything was read between the ticks."""
    import MPC
    g = np.load(os.path.join(G, "horizon_trot_N24.npz"))
    dt, T_gait, n = float(g["dt"]), float(g["T_gait"]), g["x"].shape[1] // 24
    solo, lazy, pair = MPC.MPC(dt, n, T_gait), MPC.MPC(dt, n, T_gait), MPC.MPC(dt, n, T_gait)
    np.testing.assert_array_equal(solo.q_w[:, 0], [0.0, 0.0, 0.2027682, 0.0, 0.0, 0.0])
    for t in range(len(g["k"])):
        solo.run(g["k"][t], g["xref"][t], g["fsteps"][t])
        lazy.run(g["k"][t], g["xref"][t], g["fsteps"][t])          # nothing is read from `lazy` until the end
        pair.run(g["k"][t], np.stack([g["xref"][t]] * 2), np.stack([g["fsteps"][t]] * 2))
        assert solo.q_w.shape == (6, 1) and pair.q_w.shape == (2, 6, 1)
        np.testing.assert_allclose(solo.q_w[:, 0], g["q_w"][t], rtol=0, atol=1e-6)
        np.testing.assert_allclose(pair.q_w[1, :, 0], g["q_w"][t], rtol=0, atol=1e-6)
        np.testing.assert_allclose(solo.q_next[:, 0], g["x_robot"][t][0:6, 0], rtol=0, atol=1e-6)
        np.testing.assert_allclose(solo.v_next[:, 0], g["x_robot"][t][6:12, 0], rtol=0, atol=1e-6)
        assert solo.status == 1 and (pair.status == 1).all()
    np.testing.assert_allclose(lazy.q_w[:, 0], g["q_w"][-1], rtol=0, atol=1e-6)
    np.testing.assert_array_equal(lazy.q_w, solo.q_w)
    # a different batch size is a new set of robots: world poses restart from the initial pose instead of breaking
    pair.run(0, np.stack([g["xref"][0]] * 3), np.stack([g["fsteps"][0]] * 3))
    assert pair.q_w.shape == (3, 6, 1) and pair.f_applied.shape == (3, 12)
    np.testing.assert_allclose(pair.q_w[2, :, 0], g["q_w"][0], rtol=0, atol=1e-6)


def test_status_is_reported_and_unsolved_robots_warn():
    import MPC
    g = np.load(os.path.join(G, "solve_trot_N64.npz"))
    ok = MPC.MPC(0.02, 64, 0.32)
    ok.run(0, g["xref"][0], g["fsteps"][0])
    assert ok.status == 1
    starved = MPC.MPC(0.02, 64, 0.32, max_sweeps=1, ipm_max_iter=3)
    starved.run(0, g["xref"][0], g["fsteps"][0])
    with pytest.warns(RuntimeWarning):
        assert starved.status == 2
    assert np.isfinite(starved.f_applied).all() and np.isfinite(starved.x).all()
