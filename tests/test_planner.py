"""mpc-tsid_b200/scenario.py (batched input generator) against the reference FootstepPlanner's
outputs stored in tests/golden/planner_trot.npz."""
import os

import numpy as np
import pytest

from scenario import GAIT_KINDS, Scenario, gait_sequence, reference_ramp

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "planner_trot.npz"))


@pytest.mark.parametrize("N", [16, 32])
def test_trot_planner_matches_reference(N):
    sc = Scenario(1, n_steps=N, gaits="trot", v_ref=G["v_ref"], phase=[0], random_commands=False)
    xr_g, fs_g = G["xref_N%d" % N], G["fsteps_N%d" % N]
    for k in range(xr_g.shape[0]):
        xr, fs = sc.inputs()
        np.testing.assert_allclose(sc.state[0], G["state_N%d" % N][k], rtol=0, atol=1e-15)
        assert np.array_equal(np.isnan(fs[0]), np.isnan(fs_g[k]))
        np.testing.assert_allclose(np.nan_to_num(fs[0]), np.nan_to_num(fs_g[k]), rtol=0, atol=1e-14)
        np.testing.assert_allclose(xr[0], xr_g[k], rtol=0, atol=1e-14)
        xn = xr_g[k][:, 1] + 0.01 * np.sin(np.arange(12) + k)
        sc.advance(xn[None])


def test_trot_planner_matches_reference_at_dt_001():
    """dt = 0.01 (main.py:20): the gait period is 32 steps.  Host twin against the reference planner's own xref / fsteps over more
    than two gait periods; and the period survives the packing the device planner consumes (two 64-bit words per robot)."""
    D = np.load(os.path.join(os.path.dirname(__file__), "golden", "planner_trot_dt01.npz"))
    sc = Scenario(1, n_steps=32, dt=0.01, T_gait=0.32, gaits="trot", v_ref=D["v_ref"], phase=[0], random_commands=False)
    assert sc.period == 32
    for k in range(D["xref"].shape[0]):
        xr, fs = sc.inputs()
        np.testing.assert_allclose(sc.state[0], D["state"][k], rtol=0, atol=1e-15)
        assert np.array_equal(np.isnan(fs[0]), np.isnan(D["fsteps"][k]))
        np.testing.assert_allclose(np.nan_to_num(fs[0]), np.nan_to_num(D["fsteps"][k]), rtol=0, atol=1e-14)
        np.testing.assert_allclose(xr[0], D["xref"][k], rtol=0, atol=1e-14)
        xn = D["xref"][k][:, 1] + 0.005 * np.sin(np.arange(12) + k)
        sc.advance(xn[None])
    bits = Scenario(3, n_steps=32, dt=0.01, T_gait=0.32, gaits=["trot", "walk", "pace"], random_commands=False)
    words = bits.seq_bits()
    assert words.shape == (3, 2) and words.dtype == np.uint64
    for b in range(3):
        for s_ in range(32):
            for j in range(4):
                assert ((int(words[b, s_ // 16]) >> (4 * (s_ % 16) + j)) & 1) == int(bits.seq[b, s_, j])
    assert Scenario(2, gaits="trot").seq_bits().shape == (2,)      # one word while the period fits it: the layout of rounds 1-2


def test_gait_tables():
    for kind in GAIT_KINDS:
        seq = gait_sequence(kind)
        assert seq.shape == (16, 4) and set(np.unique(seq)) <= {0.0, 1.0}
    assert gait_sequence("trot").sum() == 36            # 3 * 36 = 108 active force unknowns
    assert gait_sequence("walk").sum() == 48 and gait_sequence("static").sum() == 64
    sc = Scenario(3, gaits=["trot", "walk", "static"], phase=[0, 3, 7], random_commands=False)
    for _ in range(20):
        _, fs = sc.inputs()
        assert np.all(fs[:, :, 0].sum(axis=1) == 16)   # phase lengths always cover the horizon
        assert np.all(fs[2, 0, 1:] == fs[2, 0, 1:])    # static: no NaN in the only row
        sc.advance(np.zeros((3, 12)))


def test_batched_equals_single():
    a = Scenario(4, gaits=["trot", "pace"], seed=5)
    xr, fs = a.inputs()
    for b in range(4):
        s = Scenario(1, gaits=a.kinds[b], v_ref=a.v_ref[b], phase=[a.phase[b]], random_commands=False)
        x1, f1 = s.inputs()
        np.testing.assert_array_equal(x1[0], xr[b])
        np.testing.assert_array_equal(np.nan_to_num(f1[0]), np.nan_to_num(fs[b]))


def test_reference_ramp():
    assert reference_ramp(0) == 0.0 and reference_ramp(48) == 0.0
    assert abs(reference_ramp(100) - (2000 - 960) / 3500) < 1e-15 and reference_ramp(1000) == 1.0


def test_command_state_machine_and_reduced_polygon_match_reference():
    """Joystick commands that change in time, including vz / roll / pitch (getRefStates' state machine, FootstepPlanner.py:128-152:
    idle -> commanding -> released -> commanding again) and the `reduced` support polygon (FootstepPlanner.py:330-332), against
    the reference planner's own xref / fsteps, tick by tick over 60 ticks."""
    sc = Scenario(1, gaits="trot", v_ref=np.zeros(6), phase=[0], random_commands=False)
    flags = []
    for k in range(G["cmd_xref"].shape[0]):
        sc.set_v_ref(G["cmd_v_ref"][k])
        sc.reduced = bool(G["cmd_reduced"][k])
        xr, fs = sc.inputs()
        flags.append(int(sc.cmd_flag[0]))
        np.testing.assert_allclose(sc.state[0], G["cmd_state"][k], rtol=0, atol=1e-15)
        assert np.array_equal(np.isnan(fs[0]), np.isnan(G["cmd_fsteps"][k]))
        np.testing.assert_allclose(np.nan_to_num(fs[0]), np.nan_to_num(G["cmd_fsteps"][k]), rtol=0, atol=1e-14)
        np.testing.assert_allclose(xr[0], G["cmd_xref"][k], rtol=0, atol=1e-14)
        xn = G["cmd_xref"][k][:, 1] + 0.01 * np.sin(np.arange(12) + k)
        sc.advance(xn[None])
    assert set(flags) == {0, 1, 2} and flags[9] == 0 and flags[10] == 1 and flags[25] == 2 and flags[40] == 1
    assert np.abs(G["cmd_xref"][:, 2, 1] - 0.2027682).max() > 1e-3 and np.abs(G["cmd_xref"][:, 3, 5]).max() > 1e-3
