"""The one-robot-per-lane active-set kernel (mpc-tsid_b200/csrc/mpcqp_lane.cuh; MPCQP_MODE_LANE, chosen automatically for batches
of tens of thousands of robots) -- needs a B200: pytest -m gpu.

Same bars as tests/test_gpu_parity.py: the golden fixtures made by the reference MPC.py at every horizon (|df| <= 1e-4 N,
objective 1e-6 relative, identical rows holding with equality), the vectorised KKT certificate on every robot of seeded
closed-loop batches, the edge cases, and agreement with the half-warp-per-robot kernel it replaces."""
import glob
import os

import numpy as np
import pytest

import mpcqp
import batch_kkt
from common import FORCE_TOL, OBJ_RTOL, assert_certified, certify
from scenario import Scenario
from oracle import mpc_build

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "solve_*.npz"))) + sorted(glob.glob(os.path.join(HERE, "golden", "horizon_*.npz")))
LANE = 13 | mpcqp.MODE_LANE


def _active(g, t, p):
    _, A, l, u, _ = mpc_build.build_qp(g["xref"][t], g["fsteps"][t], p, first_tick=(g["k"][t] == 0))
    n = p.n_steps
    Ax = (A @ g["x"][t])[24 * n:]
    l, u = l[24 * n:], u[24 * n:]
    return ((np.abs(Ax - u) <= 1e-9) | (np.abs(Ax - l) <= 1e-9)).reshape(n, 4, 5)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_golden_sequences_one_robot_per_lane(path):
    g = np.load(path)
    n = g["x"].shape[1] // 24
    dt = float(g["dt"]) if "dt" in g.files else 0.02
    T_gait = float(g["T_gait"]) if "T_gait" in g.files else 0.32
    p = mpc_build.Params(dt=dt, n_steps=n, T_gait=T_gait)
    eng = mpcqp.Engine(batch=1, n_steps=n, dt=dt, T_gait=T_gait, mode=LANE)
    for t in range(len(g["k"])):
        eng.run(g["k"][t], g["xref"][t][None], g["fsteps"][t][None])
        f0, x, info = eng.forces()[0], eng.solution()[0], eng.info()
        assert info["status"][0] == 1, (t, info["status"], info["sweeps"], info["iters"])
        assert np.abs(x[12 * n:] - g["x"][t][12 * n:]).max() <= FORCE_TOL
        assert np.abs(f0 - g["f_applied"][t]).max() <= FORCE_TOL
        assert np.abs(x[:12 * n] - g["x"][t][:12 * n]).max() <= 1e-6
        assert abs(info["obj"][0] - g["obj"][t]) <= OBJ_RTOL * abs(g["obj"][t])
        np.testing.assert_array_equal(info["active"][0], _active(g, t, p))
    eng.close()


@pytest.mark.parametrize("n,B", [(16, 1500), (7, 333), (32, 400), (64, 130)], ids=["N16", "N7", "N32", "N64"])
def test_lane_kernel_agrees_with_the_half_warp_kernel_and_certifies(n, B):
    """Seeded closed loop, all gaits, more robots than lanes in a warp and not a multiple of 32 (ragged last warp, refill from the
    work counter): every robot of every tick passes the vectorised KKT certificate; solutions agree with the half-warp kernel to
    1e-8, status / contact / rows holding with equality are identical, and so is the dead-reckoned world pose to 1e-9."""
    sc = Scenario(B, n_steps=n, gaits=["trot", "pace", "bound", "walk"], seed=57)
    a = mpcqp.Engine(batch=B, n_steps=n)
    b = mpcqp.Engine(batch=B, n_steps=n, mode=LANE)
    assert not (a.params.mode & mpcqp.MODE_LANE)
    p = mpc_build.Params(n_steps=n)
    for t in range(6):
        xref, fsteps = sc.inputs()
        a.run(t, xref, fsteps); b.run(t, xref, fsteps)
        xa, xb, ia, ib = a.solution(), b.solution(), a.info(), b.info()
        assert (ib["status"] == 1).all(), (t, np.bincount(ib["status"]))
        np.testing.assert_array_equal(ia["status"], ib["status"])
        np.testing.assert_array_equal(ia["contact"], ib["contact"])
        np.testing.assert_array_equal(ia["active"], ib["active"])
        np.testing.assert_allclose(xb, xa, rtol=0, atol=1e-8)
        np.testing.assert_allclose(ib["obj"], ia["obj"], rtol=1e-9)
        np.testing.assert_allclose(ib["y"], ia["y"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(b.forces(), a.forces(), rtol=0, atol=1e-8)
        np.testing.assert_allclose(b.world_pose(), a.world_pose(), rtol=0, atol=1e-9)
        if n == 16:
            cert = batch_kkt.certificate(xref, fsteps, xb, ib["y"], first_tick=(t == 0))
            batch_kkt.assert_batch_certified(cert, "tick %d" % t)
        else:
            for r in range(0, B, max(1, B // 12)):
                assert_certified(certify(xref[r], fsteps[r], xb[r], ib["y"][r], first_tick=(t == 0), params=p), "robot %d tick %d" % (r, t))
        sc.advance(xa[:, :12] + xref[:, :, 1])
    a.close(); b.close()


def test_lane_kernel_hands_cycling_robots_to_the_interior_point_stage():
    """max_sweeps = 0: no sweep at all, every robot goes through the queue to ipm_kernel; aggressive cold starts: some do."""
    B = 200
    sc = Scenario(B, gaits=["trot", "pace", "bound", "walk"], seed=58)
    a, b = mpcqp.Engine(batch=B, max_sweeps=0), mpcqp.Engine(batch=B, mode=LANE, max_sweeps=0)
    for t in range(3):
        xref, fsteps = sc.inputs()
        a.run(t, xref, fsteps); b.run(t, xref, fsteps)
        xa, xb = a.solution(), b.solution()
        assert b.fallback_count() == B and (b.info()["status"] == 1).all() and (b.info()["iters"] > 0).all()
        np.testing.assert_array_equal(xa, xb)            # same interior-point kernel, same inputs
        sc.advance(xa[:, :12] + xref[:, :, 1])
    a.close(); b.close()
    # cold starts at 1.5 m/s with velocity errors: the sweeps cycle on some robots; all come back solved and certified
    rng = np.random.default_rng(3)
    v = np.zeros((B, 6)); v[:, 0] = 1.5; v[:, 1] = rng.uniform(-0.3, 0.3, B); v[:, 5] = rng.uniform(-0.4, 0.4, B)
    sc = Scenario(B, gaits=["trot", "pace", "bound", "walk"], seed=59, v_ref=v)
    sc.state[:, 6:9] += rng.uniform(-0.3, 0.3, (B, 3))
    e = mpcqp.Engine(batch=B, mode=LANE)
    for t in range(3):
        xref, fsteps = sc.inputs()
        e.run(t, xref, fsteps)
        x, info = e.solution(), e.info()
        assert (info["status"] == 1).all()
        batch_kkt.assert_batch_certified(batch_kkt.certificate(xref, fsteps, x, info["y"], first_tick=(t == 0)), "cold tick %d" % t)
        sc.advance(x[:, :12] + xref[:, :, 1])
    e.close()


def test_edge_cases_one_robot_per_lane():
    eng = mpcqp.Engine(batch=6, mode=LANE)
    sc = Scenario(6, gaits="trot", seed=4)
    xref, fsteps = sc.inputs()
    fsteps[1, :, 0] = 0.0                # empty gait table -> no contact anywhere, forces 0, still "solved"
    xref[2, 3, 0] = np.nan               # NaN in the measured state -> BAD_INPUT, forces exactly 0 (never NaN)
    fsteps[3, 0, 1] = 0.0                # x == 0.0 marks swing (MPC.py:650)
    fsteps[4, 2:, 0] = 0.0               # phase table shorter than the horizon -> remaining steps without contact
    fsteps[5, 0, 0] = 1.5                # non-integer phase length -> BAD_INPUT
    inputs = (xref.copy(), fsteps.copy())
    eng.run(1, xref, fsteps)
    f0, x, info = eng.forces(), eng.solution(), eng.info()
    np.testing.assert_array_equal(inputs[0], xref)
    np.testing.assert_array_equal(inputs[1], fsteps)
    assert list(info["status"]) == [1, 1, 3, 1, 1, 3]
    assert np.all(np.isfinite(f0)) and np.all(f0[1] == 0) and np.all(f0[2] == 0) and np.all(f0[5] == 0)
    assert np.all(np.isfinite(x))
    assert not info["contact"][1].any() and not info["contact"][3][0, 0] and info["contact"][0][0, 0]
    assert not info["contact"][4][8:].any()
    for b in (0, 3, 4):
        assert_certified(certify(xref[b], fsteps[b], x[b], info["y"][b]), "edge %d" % b)
    ref = mpcqp.Engine(batch=6)
    ref.run(1, xref, fsteps)
    np.testing.assert_allclose(ref.solution(), x, rtol=0, atol=1e-8)
    np.testing.assert_array_equal(ref.info()["active"], info["active"])
    eng.close(); ref.close()


@pytest.mark.parametrize("noise", [(0.0, 0.0, 0.0, 0.0), (1e-3, 5e-3, 1e-2, 2e-2)], ids=["noiseless", "hash-noise"])
def test_device_closed_loop_one_robot_per_lane(noise):
    """Planner kernel + lane kernel + integration against the host loop driving the half-warp engine tick by tick."""
    B, T = 75, 14
    kw = dict(gaits=["trot", "pace", "bound", "walk", "static"], seed=31, noise=noise, noise_kind="hash")
    host_sc, dev_sc = Scenario(B, **kw), Scenario(B, **kw)
    host, dev = mpcqp.Engine(batch=B), mpcqp.Engine(batch=B, mode=LANE)
    dev.scenario_init(dev_sc)
    for t in range(T):
        xref, fsteps = host_sc.inputs()
        host.run(t, xref, fsteps)
        xh = host.solution()
        dev.scenario_run(1, emit_inputs=True)
        xd_ref, fd = dev.last_inputs()
        assert np.array_equal(np.isnan(fd), np.isnan(fsteps)), "tick %d: swing pattern" % t
        np.testing.assert_allclose(np.nan_to_num(fd), np.nan_to_num(fsteps), rtol=0, atol=1e-9)
        np.testing.assert_allclose(xd_ref, xref, rtol=0, atol=1e-9)
        assert (dev.info()["status"] == 1).all()
        np.testing.assert_allclose(dev.forces(), host.forces(), rtol=0, atol=1e-6)
        host_sc.advance(xh[:, :12] + xref[:, :, 1])
        st = dev.scenario_state()
        np.testing.assert_allclose(st["state"], host_sc.state, rtol=0, atol=1e-9)
        np.testing.assert_allclose(st["frame"], host_sc.frame, rtol=0, atol=1e-9)
    # several ticks in one call, no inputs emitted
    dev.scenario_run(5)
    for t in range(5):
        xref, fsteps = host_sc.inputs()
        host.run(T + t, xref, fsteps)
        xh = host.solution()
        host_sc.advance(xh[:, :12] + xref[:, :, 1])
    np.testing.assert_allclose(dev.scenario_state()["state"], host_sc.state, rtol=0, atol=1e-8)
    host.close(); dev.close()


def test_lane_kernel_host_chunks_and_refill():
    """More robots than resident lanes (18 944 on a B200: every lane refills from the work counter) and more than one host chunk
    (pageable inputs of a batch above two grids of lanes are staged and solved chunk by chunk on the side streams, each with its own
    workspace): same answers as the half-warp kernel."""
    B = 38000
    sc = Scenario(B, gaits=["trot", "pace", "bound", "walk"], seed=61)
    a, b = mpcqp.Engine(batch=B), mpcqp.Engine(batch=B, mode=LANE)
    for t in range(2):
        xref, fsteps = sc.inputs()
        a.run(t, xref, fsteps); b.run(t, xref, fsteps)
        xa, xb = a.solution(), b.solution()
        assert (b.info(with_y=False)["status"] == 1).all()
        np.testing.assert_allclose(xb, xa, rtol=0, atol=1e-8)
        sc.advance(xa[:, :12] + xref[:, :, 1])
    a.close(); b.close()
