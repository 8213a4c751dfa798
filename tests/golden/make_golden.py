"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run once, in the build container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py

What runs unmodified from /root/reference:  MPC.py (MPC.run: construct_gait, create/update_ML,
create/update_NK, create_weight_matrices, call_solver, retrieve_result) and, for the trot planner
fixture, FootstepPlanner.py (update_fsteps, getRefStates).  What is substituted: the `osqp` module
(not installable here) by oracle/ref_shims/osqp.py = the restated OSQP algorithm at eps 1e-8 plus an
active-set polish, every returned point KKT-certified; plus import shims for matplotlib / pybullet
/ utils.getSkew (oracle/ref_shims/).  Inputs come from mpc-tsid_b200/scenario.py (itself checked
against the reference planner in the `planner_trot` fixture) so that gaits the reference's planner
cannot produce as shipped (its bound / pace / static constructors build a 6-row gait table that
does not fit its own 20-row fsteps, FootstepPlanner.py:186-282) are covered too.

Each solve_<case>.npz holds, per tick: the inputs (xref, fsteps, k), what MPC.py built (ML.data, NK,
NK_inf, P.data), what crossed the osqp boundary (warm-start x), and what came back after MPC.py's
own extraction (x, f_applied, x_robot) plus the certificate of the solution.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mpc-tsid_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(1, "/root/reference")
np.int = int          # MPC.py / FootstepPlanner.py predate numpy 1.24

import scipy.sparse          # noqa: E402
import scipy.sparse.csc      # noqa: E402,F401  (MPC.py uses scipy.sparse.csc.csc_matrix)
import MPC as RefMPC         # noqa: E402   -> /root/reference/MPC.py
import FootstepPlanner as RefPlanner   # noqa: E402
from scenario import Scenario, SHOULDERS, H_REF   # noqa: E402

assert RefMPC.__file__.startswith("/root/reference"), RefMPC.__file__


def run_case(name, scen, ticks, first_state=None, prefix="solve"):
    """Closed loop: scenario -> reference MPC.run -> scenario.advance(reference prediction)."""
    mpc = RefMPC.MPC(scen.dt, scen.N, scen.T_gait)
    if first_state is not None:
        scen.state[:] = first_state
    rec = {k: [] for k in ("xref", "fsteps", "k", "ML_data", "NK", "NK_inf", "x", "f_applied", "x_robot", "obj",
                           "warm_x", "x_admm", "cert_prim", "cert_stat", "cert_comp", "cert_sign", "osqp_iter", "q_w")}
    for t in range(ticks):
        xref, fsteps = scen.inputs()
        fs_in = fsteps[0].copy()
        mpc.run(t, xref[0].copy(), fs_in)          # MPC.py:327 overwrites NaNs in its argument
        prob = mpc.prob
        warm = [kw["x"] for nm, kw in prob.calls if nm == "warm_start"]
        rec["xref"].append(xref[0]); rec["fsteps"].append(fsteps[0]); rec["k"].append(float(t))
        rec["ML_data"].append(mpc.ML.data.copy()); rec["NK"].append(mpc.NK.ravel().copy()); rec["NK_inf"].append(mpc.NK_inf.copy())
        rec["x"].append(np.array(mpc.x)); rec["f_applied"].append(np.array(mpc.f_applied)); rec["x_robot"].append(mpc.x_robot.copy())
        rec["obj"].append(prob.last_cert["obj"]); rec["x_admm"].append(mpc.sol.x_admm.copy())
        rec["warm_x"].append(warm[-1] if (warm and t > 0) else np.zeros_like(mpc.x))
        for key, ck in (("cert_prim", "prim"), ("cert_stat", "stat"), ("cert_comp", "comp"), ("cert_sign", "bad_sign")):
            rec[key].append(prob.last_cert[ck])
        rec["osqp_iter"].append(mpc.sol.info.iter)
        rec["q_w"].append(mpc.q_w[:, 0].copy())                     # dead-reckoned world pose, MPC.py:503-510
        prob.calls.clear()
        scen.advance(mpc.x_robot[:, 0][None, :])
        print("  %s tick %2d  osqp-port iters %4d  |x_admm - x*| %.1e  stat %.1e  f0z %s" % (
            name, t, mpc.sol.info.iter, np.abs(mpc.sol.x_admm - mpc.x).max(), prob.last_cert["stat"],
            np.round(mpc.f_applied[2::3], 2)))
    out = {k: np.array(v) for k, v in rec.items()}
    out["P_data"] = mpc.P.data.copy()
    out["ML_indices"] = mpc.ML.indices.copy()
    out["ML_indptr"] = mpc.ML.indptr.copy()
    out["i_update_B"] = np.asarray(mpc.i_update_B)
    out["i_update_S"] = np.asarray(mpc.i_update_S)
    out["dt"], out["T_gait"] = np.float64(scen.dt), np.float64(scen.T_gait)
    if scen.N != 16 or prefix != "solve":
        for key in ("ML_data", "NK", "NK_inf", "x_admm", "warm_x"):     # keep the long-horizon fixture small
            out[key] = out[key][:2]
    np.savez_compressed(os.path.join(HERE, "%s_%s.npz" % (prefix, name)), **out)
    nact = int(((np.abs(out["x"][:, 12 * scen.N:].reshape(ticks, -1, 3)[:, :, 2] - 25.0) < 1e-9)).sum())
    print("wrote %s_%s.npz  (%d ticks, %d foot-steps at fz_max)" % (prefix, name, ticks, nact))


def planner_fixture():
    """The reference FootstepPlanner (trot is the only gait it can build as shipped) next to the same
    states fed to scenario.py; the fixture stores the reference's outputs."""
    rec = {}
    for nper in (1, 2):
        fp = RefPlanner.FootstepPlanner(0.02, nper)
        N = fp.n_steps
        sc = Scenario(1, n_steps=N, gaits="trot", v_ref=[0.4, -0.1, 0, 0, 0, 0.3], phase=[0], random_commands=False)
        xs, fs, states, feet = [], [], [], []
        for k in range(40):
            sc.inputs()
            st = sc.state[0].copy()
            l_feet = sc.local_feet()[0]
            fp.update_fsteps(k, l_feet, st[6:12].reshape(6, 1), sc.v_ref[0].reshape(6, 1), st[2], None, None, False)
            fp.getRefStates(k, 0.32, st[0:3].reshape(3, 1), st[3:6].reshape(3, 1), st[6:9].reshape(3, 1),
                            st[9:12].reshape(3, 1), sc.v_ref[0].reshape(6, 1))
            xs.append(fp.xref.copy()); fs.append(fp.fsteps.copy()); states.append(st); feet.append(l_feet.copy())
            xn = fp.xref[:, 1] + 0.01 * np.sin(np.arange(12) + k)
            sc.advance(xn[None])
        rec["xref_N%d" % N] = np.array(xs); rec["fsteps_N%d" % N] = np.array(fs)
        rec["state_N%d" % N] = np.array(states); rec["lfeet_N%d" % N] = np.array(feet)
    rec["v_ref"] = np.array([0.4, -0.1, 0, 0, 0, 0.3])
    # joystick commands that change in time, including vz / roll / pitch (the state machine of FootstepPlanner.py:128-152)
    # and the `reduced` support polygon (FootstepPlanner.py:330-332, toggled by Joystick.py:66-67)
    fp = RefPlanner.FootstepPlanner(0.02, 1)
    sc = Scenario(1, n_steps=16, gaits="trot", v_ref=np.zeros(6), phase=[0], random_commands=False)
    sched = [(0, [0.3, 0.0, 0.0, 0.0, 0.0, 0.2], False), (10, [0.3, 0.1, 0.1, 0.2, -0.15, 0.2], False),
             (25, [0.2, 0.0, 0.0, 0.1, 0.0, -0.3], False), (30, [0.2, 0.0, 0.03, 0.1, 0.0, -0.3], True),
             (40, [0.0, -0.2, -0.08, 0.0, 0.05, 0.0], True), (50, [0.0, 0.0, 0.0, 0.0, 0.0, 0.0], False)]
    xs, fs, states, vs, rs = [], [], [], [], []
    for k in range(60):
        v, red = [(vv, rr) for (t0, vv, rr) in sched if t0 <= k][-1]
        sc.set_v_ref(v); sc.reduced = red
        sc.inputs()
        st = sc.state[0].copy()
        vr = np.array(v, dtype=np.float64).reshape(6, 1)
        fp.update_fsteps(k, sc.local_feet()[0], st[6:12].reshape(6, 1), vr, st[2], None, None, red)
        fp.getRefStates(k, 0.32, st[0:3].reshape(3, 1), st[3:6].reshape(3, 1), st[6:9].reshape(3, 1), st[9:12].reshape(3, 1), vr)
        xs.append(fp.xref.copy()); fs.append(fp.fsteps.copy()); states.append(st); vs.append(np.array(v, dtype=np.float64)); rs.append(red)
        xn = fp.xref[:, 1] + 0.01 * np.sin(np.arange(12) + k)
        sc.advance(xn[None])
    rec["cmd_xref"], rec["cmd_fsteps"], rec["cmd_state"] = np.array(xs), np.array(fs), np.array(states)
    rec["cmd_v_ref"], rec["cmd_reduced"] = np.array(vs), np.array(rs)
    np.savez_compressed(os.path.join(HERE, "planner_trot.npz"), **rec)
    print("wrote planner_trot.npz")


def planner_fixture_dt01():
    """The reference FootstepPlanner at the MPC time step dt = 0.01 (main.py:20: gait period T_gait / dt = 32 steps, N = 32 with one
    period): same protocol as planner_fixture, its own file (planner_trot_dt01.npz)."""
    fp = RefPlanner.FootstepPlanner(0.01, 1)
    N = fp.n_steps
    assert N == 32
    sc = Scenario(1, n_steps=N, dt=0.01, T_gait=0.32, gaits="trot", v_ref=[0.35, 0.1, 0, 0, 0, -0.25], phase=[0], random_commands=False)
    xs, fs, states = [], [], []
    for k in range(70):
        sc.inputs()
        st = sc.state[0].copy()
        fp.update_fsteps(k, sc.local_feet()[0], st[6:12].reshape(6, 1), sc.v_ref[0].reshape(6, 1), st[2], None, None, False)
        fp.getRefStates(k, 0.32, st[0:3].reshape(3, 1), st[3:6].reshape(3, 1), st[6:9].reshape(3, 1),
                        st[9:12].reshape(3, 1), sc.v_ref[0].reshape(6, 1))
        xs.append(fp.xref.copy()); fs.append(fp.fsteps.copy()); states.append(st)
        xn = fp.xref[:, 1] + 0.005 * np.sin(np.arange(12) + k)
        sc.advance(xn[None])
    np.savez_compressed(os.path.join(HERE, "planner_trot_dt01.npz"), xref=np.array(xs), fsteps=np.array(fs), state=np.array(states),
                        v_ref=sc.v_ref[0].copy())
    print("wrote planner_trot_dt01.npz")


def main():
    planner_fixture()
    planner_fixture_dt01()
    # 1. nominal trot, the survey's known-answer start state, with noise
    s = Scenario(1, gaits="trot", v_ref=[0.3, 0, 0, 0, 0, 0.0], phase=[0], random_commands=False, seed=11)
    run_case("trot", s, 20, first_state=np.array([.01, -.005, .21, .02, -.03, 0, .1, .05, -.02, .1, -.1, .05]))
    # 2. trot with lateral + yaw command from a random gait phase
    s = Scenario(1, gaits="trot", v_ref=[0.6, -0.25, 0, 0, 0, 0.4], phase=[5], random_commands=False, seed=12)
    run_case("trot_turn", s, 12)
    # 3. other contact schedules
    for g, ph in (("pace", 3), ("bound", 9), ("walk", 2), ("static", 0)):
        s = Scenario(1, gaits=g, v_ref=[0.25, 0.1, 0, 0, 0, -0.2], phase=[ph], random_commands=False, seed=13)
        run_case(g, s, 8)
    # 4. aggressive: large tilt / velocity error so the friction pyramid and fz_max become active
    s = Scenario(1, gaits="trot", v_ref=[1.5, 0.6, 0, 0, 0, 1.0], phase=[1], random_commands=False, seed=14)
    run_case("aggressive", s, 10, first_state=np.array([0, 0, .14, .25, -.2, 0, -.6, .5, -.4, 1.0, -1.2, .8]))
    # 5. the reference's own manual scenario (test_motionless.py:28-47): all feet down, vz = 0.1
    s = Scenario(1, gaits="static", v_ref=[0, 0, 0, 0, 0, 0.0], phase=[0], random_commands=False, seed=15, noise=(0, 0, 0, 0))
    run_case("motionless", s, 4, first_state=np.array([0, 0, .2, 0, 0, 0, 0, 0, .1, 0, 0, 0]))


def long_horizon():
    # 6. long horizon (BASELINE configs[3]): n_periods = 2 -> N = 32, same T_gait
    s = Scenario(1, n_steps=32, gaits="trot", v_ref=[0.5, 0.1, 0, 0, 0, 0.3], phase=[3], random_commands=False, seed=16)
    run_case("trot_N32", s, 5)


def long_horizon_64():
    # 7. N = 64 (n_periods = 4, BASELINE configs[3])
    s = Scenario(1, n_steps=64, gaits="trot", v_ref=[0.4, -0.1, 0, 0, 0, -0.25], phase=[6], random_commands=False, seed=17)
    run_case("trot_N64", s, 3)


def general_horizons():
    # 8. horizons other than 16 / 32 / 64 and other MPC time steps: n_steps = n_periods * T_gait / dt (main.py:20-23,
    #    FootstepPlanner.py:52-63).  Files horizon_<case>.npz (they also carry dt and T_gait).
    cases = (("trot_N24", dict(n_steps=24, dt=0.02, T_gait=0.48, gaits="trot"), [0.4, 0.1, 0, 0, 0, 0.2], 5),
             ("trot_dt01", dict(n_steps=32, dt=0.01, T_gait=0.32, gaits="trot"), [0.5, -0.1, 0, 0, 0, 0.3], 7),
             ("walk_N8", dict(n_steps=8, dt=0.04, T_gait=0.32, gaits="walk"), [0.2, 0.05, 0, 0, 0, -0.2], 1),
             ("pace_N12", dict(n_steps=12, dt=0.02, T_gait=0.24, gaits="pace"), [0.3, 0.1, 0, 0, 0, 0.1], 4),
             ("bound_N10", dict(n_steps=10, dt=0.02, T_gait=0.2, gaits="bound"), [0.25, 0.0, 0, 0, 0, -0.3], 2),
             ("trot_N48", dict(n_steps=48, dt=0.02, T_gait=0.48, gaits="trot"), [0.6, 0.15, 0, 0, 0, -0.25], 9))
    for i, (name, kw, v, ph) in enumerate(cases):
        s = Scenario(1, v_ref=v, phase=[ph], random_commands=False, seed=30 + i, **kw)
        run_case(name, s, 4, prefix="horizon")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "planner":
        planner_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "planner_dt01":
        planner_fixture_dt01()
    elif len(sys.argv) > 1 and sys.argv[1] == "horizons":
        general_horizons()
    elif len(sys.argv) > 1 and sys.argv[1] == "long":
        long_horizon()
    elif len(sys.argv) > 1 and sys.argv[1] == "long64":
        long_horizon_64()
    else:
        main()
        long_horizon()
        long_horizon_64()
        general_horizons()
