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


def test_async_flag_matches_reference_behaviour():
    import MPC_Wrapper
    with pytest.raises(NotImplementedError):
        MPC_Wrapper.MPC_Wrapper(0.02, 16, 20, 0.32, multiprocessing=True)
