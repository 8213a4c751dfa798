"""Parity of the CUDA engine (through the C ABI) with the reference -- needs a B200: pytest -m gpu.

Bars (BASELINE.json north_star): |df| <= 1e-4 N per force component, objective within 1e-6
relative, identical contact mask and identical active set, against
  (1) the golden fixtures produced by running the reference MPC.py itself (tests/golden/),
  (2) the oracle's KKT certificate on the reference-layout QP for seeded closed-loop batches,
  (3) size-independent properties at the full BASELINE batch (4096 instances).
"""
import glob
import os

import numpy as np
import pytest

import mpcqp
import batch_kkt
from common import FORCE_TOL, OBJ_RTOL, assert_certified, certify
from scenario import Scenario

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "solve_*.npz")))
N = 16


def _active_from_x(g, t, n):
    """Rows of the reference's inequality block that hold with equality at the golden optimum (matrices
    from oracle.mpc_build, which tests/test_oracle_build.py pins to the reference's own ML.data)."""
    from oracle import mpc_build
    _, A, l, u, _ = mpc_build.build_qp(g["xref"][t], g["fsteps"][t], mpc_build.Params(n_steps=n), first_tick=(g["k"][t] == 0))
    Ax = (A @ g["x"][t])[24 * n:]
    l, u = l[24 * n:], u[24 * n:]
    return ((np.abs(Ax - u) <= 1e-9) | (np.abs(Ax - l) <= 1e-9)).reshape(n, 4, 5)


@pytest.mark.parametrize("mode", [13, -13, 7, 3, 2], ids=["stagewise+ipm", "ipm-only", "stagewise+admm", "dense+admm", "admm-only"])
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[6:-4] for p in GOLD])
def test_golden_sequences(path, mode):
    """Replay each golden closed-loop sequence tick by tick (warm start carried like the reference).  mode 13 is the default
    (stage-wise sweeps, interior-point fallback); "ipm-only" switches the warm-started sweeps off so that EVERY tick is solved
    by the interior-point stage and the sweeps it seeds."""
    g = np.load(path)
    n = g["x"].shape[1] // 24                       # horizon of this fixture (16, or 32 / 64 for the long-horizon cases)
    if n == 64 and mode in (7, 3, 2):
        pytest.skip("the dense stages exist for N = 16 and 32")
    if mode == -13:
        eng = mpcqp.Engine(batch=1, n_steps=n, mode=13, max_sweeps=0)
    else:
        eng = mpcqp.Engine(batch=1, n_steps=n, mode=mode)
    for t in range(len(g["k"])):
        eng.run(g["k"][t], g["xref"][t][None], g["fsteps"][t][None])
        f0, x, info = eng.forces()[0], eng.solution()[0], eng.info()
        assert info["status"][0] == 1
        assert np.abs(x[12 * n:] - g["x"][t][12 * n:]).max() <= FORCE_TOL
        assert np.abs(f0 - g["f_applied"][t]).max() <= FORCE_TOL
        assert np.abs(x[:12 * n] - g["x"][t][:12 * n]).max() <= 1e-6
        assert abs(info["obj"][0] - g["obj"][t]) <= OBJ_RTOL * abs(g["obj"][t])
        np.testing.assert_array_equal(info["active"][0], _active_from_x(g, t, n))
    eng.close()


def test_build_half_matches_reference_coefficients():
    """K1 parity: the coefficients MPC.update_ML / update_NK write (ML.data[i_update_B], [i_update_S], NK)."""
    for path in GOLD:
        g = np.load(path)
        T, n = len(g["ML_data"]), g["x"].shape[1] // 24
        eng = mpcqp.Engine(batch=T, n_steps=n)
        for first in (False, True):
            ks = 0.0 if first else 1.0
            Bv, Sv, NK = eng.export_build(ks, g["xref"][:T], g["fsteps"][:T])
            for t in range(T):
                if (g["k"][t] == 0) != first:
                    continue
                ref_B = np.stack([g["ML_data"][t][g["i_update_B"] + 96 * k] for k in range(n)])
                np.testing.assert_allclose(Bv[t], ref_B, rtol=1e-13, atol=1e-16)
                np.testing.assert_array_equal(Sv[t], g["ML_data"][t][g["i_update_S"]])
                np.testing.assert_allclose(NK[t], g["NK"][t][:12 * n], rtol=0, atol=1e-15)
        eng.close()


@pytest.mark.parametrize("mode", [7, 3], ids=["stagewise", "dense"])
@pytest.mark.parametrize("gaits", [["trot"], ["pace", "bound", "walk", "static"]], ids=["trot", "mixed"])
def test_closed_loop_batch_certified(gaits, mode):
    """63 robots (odd: one half-warp of the stage-wise kernel idles) x 12 closed-loop ticks: every solution passes the
    oracle's KKT certificate on the QP the reference would have built; masks identical."""
    B, T = 63, 12
    eng = mpcqp.Engine(batch=B, mode=mode)
    sc = Scenario(B, gaits=gaits, seed=123)
    for t in range(T):
        xref, fsteps = sc.inputs()
        eng.run(t, xref, fsteps)
        x, info = eng.solution(), eng.info()
        assert (info["status"] == 1).all()
        for b in range(0, B, 3):
            cert = certify(xref[b], fsteps[b], x[b], info["y"][b], first_tick=(t == 0))
            assert_certified(cert, "tick %d robot %d" % (t, b))
            np.testing.assert_array_equal(cert["contact"].astype(bool), info["contact"][b])
            np.testing.assert_array_equal(cert["active"].reshape(N, 4, 5), info["active"][b])
            assert abs(cert["obj"] - info["obj"][b]) <= OBJ_RTOL * abs(cert["obj"])
        sc.advance(x[:, :12] + xref[:, :, 1])
    eng.close()


@pytest.mark.parametrize("mode", [7, 3], ids=["stagewise", "dense"])
def test_long_horizon_closed_loop_certified(mode):
    """BASELINE configs[3], N = 32 (n_periods = 2): 24 robots x 6 ticks, mixed gaits, oracle certificate."""
    from oracle import mpc_build
    B, T, n = 24, 6, 32
    eng = mpcqp.Engine(batch=B, n_steps=n, mode=mode)
    sc = Scenario(B, n_steps=n, gaits=["trot", "pace", "walk"], seed=77)
    par = mpc_build.Params(n_steps=n)
    for t in range(T):
        xref, fsteps = sc.inputs()
        eng.run(t, xref, fsteps)
        x, info = eng.solution(), eng.info()
        assert (info["status"] == 1).all()
        for b in range(0, B, 5):
            cert = certify(xref[b], fsteps[b], x[b], info["y"][b], first_tick=(t == 0), params=par)
            assert_certified(cert, "N=32 tick %d robot %d" % (t, b))
            np.testing.assert_array_equal(cert["contact"].astype(bool), info["contact"][b])
            np.testing.assert_array_equal(cert["active"].reshape(n, 4, 5), info["active"][b])
        sc.advance(x[:, :12] + xref[:, :, 1])
    eng.close()


def test_horizon_64_trot_closed_loop_certified():
    """BASELINE configs[3], N = 64 (n_periods = 4): both stages on the stage-wise factorisation (no dense ADMM factor fits),
    16 trot robots x 5 ticks released from rest with moderate commands; oracle certificate on a sample; every robot solved."""
    from oracle import mpc_build
    B, T, n = 16, 5, 64
    eng = mpcqp.Engine(batch=B, n_steps=n)
    assert eng.params.mode == 13
    rng = np.random.default_rng(64)
    v_ref = np.zeros((B, 6))
    v_ref[:, 0], v_ref[:, 1], v_ref[:, 5] = rng.uniform(-0.2, 0.45, B), rng.uniform(-0.15, 0.15, B), rng.uniform(-0.3, 0.3, B)
    sc = Scenario(B, n_steps=n, gaits=["trot"], seed=64, v_ref=v_ref)
    par = mpc_build.Params(n_steps=n)
    for t in range(T):
        xref, fsteps = sc.inputs()
        eng.run(t, xref, fsteps)
        x, info = eng.solution(), eng.info()
        assert (info["status"] == 1).all(), (t, info["status"], info["sweeps"], info["iters"])
        for b in (0, 7):
            cert = certify(xref[b], fsteps[b], x[b], info["y"][b], first_tick=(t == 0), params=par)
            assert_certified(cert, "N=64 tick %d robot %d" % (t, b))
            np.testing.assert_array_equal(cert["contact"].astype(bool), info["contact"][b])
        sc.advance(x[:, :12] + xref[:, :, 1])
    eng.close()


def test_interior_point_stage_alone_lands_on_the_same_optimum():
    """The fallback stage of the stage-wise path with the warm-started sweeps switched off (max_sweeps = 0 sends every robot to
    ipm_kernel): interior-point iterations, the signature they identify, the sweeps from it.  Same (unique) optimum as the dense
    ADMM stage, every robot certified by the same guard, cold and warm ticks, mixed gaits, odd batch."""
    B = 9
    sc = Scenario(B, gaits=["trot", "walk", "bound"], seed=5)
    dense, ipm = mpcqp.Engine(batch=B, mode=2), mpcqp.Engine(batch=B, mode=13, max_sweeps=0)
    for t in range(4):
        xref, fsteps = sc.inputs()
        dense.run(t, xref, fsteps)
        ipm.run(t, xref, fsteps)
        xd, xs, idn, iip = dense.solution(), ipm.solution(), dense.info(), ipm.info()
        assert (idn["status"] == 1).all() and (iip["status"] == 1).all()
        assert ipm.fallback_count() == B and (iip["iters"] > 0).all() and (iip["iters"] <= 60).all()
        assert np.abs(xd - xs).max() <= 1e-8
        np.testing.assert_array_equal(idn["active"], iip["active"])
        for b in (0, 4, 8):
            assert_certified(certify(xref[b], fsteps[b], xs[b], iip["y"][b], first_tick=(t == 0)), "tick %d robot %d" % (t, b))
        sc.advance(xd[:, :12] + xref[:, :, 1])
    dense.close(); ipm.close()


def _dynamics_residual(xref, fsteps, x, n, first_tick):
    """max |A x - b| over the 12 n dynamics rows of the QP the reference would have built (MPC.py:98-134, 362-378)."""
    from oracle import mpc_build
    _, A, l, u, _ = mpc_build.build_qp(xref, fsteps, mpc_build.Params(n_steps=n), first_tick=first_tick)
    return np.abs((A @ x)[:12 * n] - u[:12 * n]).max()


def test_hard_long_horizon_instances_are_solved():
    """N = 64 released from rest with a 0.7 .. 1 m/s command is close to bang-bang (two thirds of the stance foot-steps at the apex
    or at fz_max, near-degenerate multipliers): the warm-started sweeps cycle on some of them, the interior-point stage must
    then finish the job.  The reference's OSQP solves these (MPC.py:414-428), so must we: every robot SOLVED and certified
    by the oracle at the default settings."""
    from oracle import mpc_build
    B, n = 8, 64
    v_ref = np.zeros((B, 6))
    v_ref[:, 0] = np.linspace(0.7, 1.0, B)
    sc = Scenario(B, n_steps=n, gaits=["trot"], seed=3, v_ref=v_ref)
    eng = mpcqp.Engine(batch=B, n_steps=n)
    par = mpc_build.Params(n_steps=n)
    for t in range(3):
        xref, fsteps = sc.inputs()
        eng.run(t, xref, fsteps)
        x, info = eng.solution(), eng.info()
        assert (info["status"] == 1).all(), (t, info["status"], info["sweeps"], info["iters"])
        for b in range(B):
            assert_certified(certify(xref[b], fsteps[b], x[b], info["y"][b], first_tick=(t == 0), params=par), "tick %d robot %d" % (t, b))
        sc.advance(x[:, :12] + xref[:, :, 1])
    eng.close()


@pytest.mark.parametrize("n", [16, 32, 64])
def test_aggressive_cold_starts_are_solved(n):
    """Robots of every gait released with large velocity errors and commands up to 1.5 m/s (the regime of
    tests/golden/solve_aggressive.npz), cold start: every robot SOLVED, a sample certified by the oracle."""
    from oracle import mpc_build
    B = 256 if n < 64 else 64
    rng = np.random.default_rng(1000 + n)
    v_ref = np.zeros((B, 6))
    v_ref[:, 0], v_ref[:, 1], v_ref[:, 5] = rng.uniform(-0.8, 1.5, B), rng.uniform(-0.5, 0.5, B), rng.uniform(-0.8, 0.8, B)
    sc = Scenario(B, n_steps=n, gaits=["trot", "pace", "bound", "walk"], seed=100 + n, v_ref=v_ref)
    sc.state[:, 6:12] += rng.normal(0, 0.3, (B, 6))
    sc.state[:, 3:5] += rng.normal(0, 0.1, (B, 2))
    eng = mpcqp.Engine(batch=B, n_steps=n)
    par = mpc_build.Params(n_steps=n)
    for t in range(3):
        xref, fsteps = sc.inputs()
        eng.run(t, xref, fsteps)
        x, info = eng.solution(), eng.info()
        assert (info["status"] == 1).all(), (t, np.flatnonzero(info["status"] != 1), info["sweeps"].max(), info["iters"].max())
        batch_kkt.assert_batch_certified(batch_kkt.certificate(xref, fsteps, x, info["y"], par, first_tick=(t == 0)), "tick %d" % t)   # every robot
        for b in range(0, B, 37):
            assert_certified(certify(xref[b], fsteps[b], x[b], info["y"][b], first_tick=(t == 0), params=par), "tick %d robot %d" % (t, b))
        sc.advance(x[:, :12] + xref[:, :, 1])
    eng.close()


def test_uncertified_answers_are_feasible_and_consistent():
    """MPCQP_STATUS_MAX_ITER contract (forced here by starving both stages of iterations): finite forces inside the friction
    pyramid, and x[:12N] is the roll-out of exactly those forces -- the dynamics rows of the reference's QP hold."""
    B, n = 8, 64
    v_ref = np.zeros((B, 6))
    v_ref[:, 0] = np.linspace(0.7, 1.0, B)
    sc = Scenario(B, n_steps=n, gaits=["trot"], seed=3, v_ref=v_ref)
    eng = mpcqp.Engine(batch=B, n_steps=n, max_sweeps=1, ipm_max_iter=4)
    xref, fsteps = sc.inputs()
    eng.run(0, xref, fsteps)
    x, info = eng.solution(), eng.info()
    assert np.isfinite(x).all() and set(np.unique(info["status"])) <= {1, 2} and (info["status"] == 2).any()
    f = x[:, 12 * n:].reshape(B, n, 4, 3)
    mu = eng.params.mu
    assert (np.abs(f[..., 0]) <= mu * f[..., 2] + 1e-9).all() and (np.abs(f[..., 1]) <= mu * f[..., 2] + 1e-9).all()
    assert (f[..., 2] >= 0).all() and (f[..., 2] <= 25).all()
    for b in range(B):
        assert _dynamics_residual(xref[b], fsteps[b], x[b], n, True) <= 1e-9, b
    np.testing.assert_array_equal(eng.forces(), x[:, 12 * n:12 * n + 12])
    eng.close()


def test_stages_agree_and_warm_start_is_only_a_speedup():
    """Active-set stage, ADMM stage and a cold start must land on the same (unique) optimum."""
    B = 32
    sc = Scenario(B, gaits=["trot", "walk"], seed=9)
    a, b, c = mpcqp.Engine(batch=B, mode=3), mpcqp.Engine(batch=B, mode=2), mpcqp.Engine(batch=B, mode=3, warm_start=0)
    d = mpcqp.Engine(batch=B, mode=7)                      # stage-wise factorisation of the active-set stage
    for t in range(6):
        xref, fsteps = sc.inputs()
        xs = []
        for e in (a, b, c, d):
            e.run(t, xref, fsteps)
            xs.append(e.solution())
            assert (e.info()["status"] == 1).all()
        assert np.abs(xs[0] - xs[1]).max() <= 1e-7 and np.abs(xs[0] - xs[2]).max() <= 1e-7 and np.abs(xs[0] - xs[3]).max() <= 1e-7
        sc.advance(xs[0][:, :12] + xref[:, :, 1])
    assert b.info()["iters"].min() > 0 and a.info()["sweeps"].mean() <= c.info()["sweeps"].mean() + 1e-9
    for e in (a, b, c, d):
        e.close()


def test_edge_cases():
    eng = mpcqp.Engine(batch=6)
    sc = Scenario(6, gaits="trot", seed=4)
    xref, fsteps = sc.inputs()
    # 0: regular.  1: empty gait table -> no contact anywhere, forces 0, still "solved"
    fsteps[1, :, 0] = 0.0
    # 2: NaN in the measured state -> BAD_INPUT, forces exactly 0 (never NaN)
    xref[2, 3, 0] = np.nan
    # 3: x == 0.0 marks swing (MPC.py:650)
    fsteps[3, 0, 1] = 0.0
    # 4: phase table shorter than the horizon -> remaining steps without contact
    fsteps[4, 2:, 0] = 0.0
    # 5: non-integer phase length -> BAD_INPUT
    fsteps[5, 0, 0] = 1.5
    inputs = (xref.copy(), fsteps.copy())
    eng.run(1, xref, fsteps)
    f0, x, info = eng.forces(), eng.solution(), eng.info()
    np.testing.assert_array_equal(inputs[0], xref)            # inputs are never written (MPC.py:327 does)
    np.testing.assert_array_equal(inputs[1], fsteps)
    assert list(info["status"]) == [1, 1, 3, 1, 1, 3]
    assert np.all(np.isfinite(f0)) and np.all(f0[1] == 0) and np.all(f0[2] == 0) and np.all(f0[5] == 0)
    assert not info["contact"][1].any() and not info["contact"][3][0, 0] and info["contact"][0][0, 0]
    assert not info["contact"][4][8:].any()
    for b in (0, 3, 4):
        assert_certified(certify(xref[b], fsteps[b], x[b], info["y"][b]), "edge %d" % b)
    # the uncorrupted instance is unaffected by its neighbours
    solo = mpcqp.Engine(batch=1)
    solo.run(1, xref[:1], fsteps[:1])
    np.testing.assert_array_equal(solo.solution()[0], x[0])
    eng.close(); solo.close()


def test_long_gait_tables_take_the_full_copy_path():
    """Host inputs: when row 7 of every gait table is a terminator only rows 0..7 are copied to the device; a table that uses
    more rows (here 9 phases) must switch the whole batch to the full copy.  Same answers as the oracle either way, and the
    stale rows a long table leaves on the device must not leak into the next (short-table) tick."""
    B = 3
    sc = Scenario(B, gaits=["trot", "walk", "pace"], seed=8)
    eng = mpcqp.Engine(batch=B)
    for t in range(4):
        xref, fsteps = sc.inputs()
        if t in (1, 2):
            # robot 0: cut its table into 9 single- or double-step phases with the same contacts (equivalent QP, longer table)
            rows = []
            for r in range(20):
                c = int(fsteps[0, r, 0])
                if c == 0:
                    break
                for _ in range(c):
                    rows.append(fsteps[0, r, 1:].copy())
            phases = []                                    # [count, row]: at most 1 step per phase for the first four, then 2
            for row in rows:
                cap = 1 if len(phases) <= 4 else 2
                if phases and phases[-1][0] < cap and np.array_equal(np.isnan(row), np.isnan(phases[-1][1])):
                    phases[-1][0] += 1
                else:
                    phases.append([1, row])
            assert 9 <= len(phases) <= 19
            tab = np.full((20, 13), np.nan)
            tab[:, 0] = 0.0
            for r, (c, row) in enumerate(phases):
                tab[r, 0] = c
                tab[r, 1:] = row
            fsteps = fsteps.copy()
            fsteps[0] = tab
            assert fsteps[0, 7, 0] != 0.0
        eng.run(t, xref, fsteps)
        x, info = eng.solution(), eng.info()
        assert (info["status"] == 1).all()
        for b in range(B):
            assert_certified(certify(xref[b], fsteps[b], x[b], info["y"][b], first_tick=(t == 0)), "tick %d robot %d" % (t, b))
        sc.advance(x[:, :12] + xref[:, :, 1])
    eng.close()


@pytest.mark.parametrize("B,gaits,kw", [(5, ["trot", "walk", "pace", "bound"], {}), (2500, "trot", {}), (333, ["trot", "walk", "pace", "bound"], {"max_sweeps": 0})],
                         ids=["ragged-mixed", "two-waves", "ipm-only"])
def test_pinned_and_pageable_host_inputs_agree(B, gaits, kw):
    """Host inputs in page-locked buffers (what torch's pin_memory or mpcqp_host_alloc give) are fetched by the solve kernels
    themselves, robot by robot, straight from the caller's memory (rows 0..7 of a gait table first, the rest only for a table that
    is longer); pageable numpy arrays go through staged copies (2500 robots: the two-chunk path whose copies overlap the other
    chunk's solve).  Same robots, same ticks, both ways: every output must agree bit for bit, including a tick on which one gait
    table runs past row 7, and with every robot sent through the interior-point kernel (which fetches its inputs again)."""
    import torch
    staged, inplace = mpcqp.Engine(batch=B, **kw), mpcqp.Engine(batch=B, **kw)
    sc = Scenario(B, gaits=gaits, seed=31)
    px = torch.empty((B, 12, N + 1), dtype=torch.float64, pin_memory=True).numpy()
    pf = torch.empty((B, 20, 13), dtype=torch.float64, pin_memory=True).numpy()
    for t in range(6):
        xref, fsteps = sc.inputs()
        if t == 3:                                          # robot 0: one table row per step (16 rows, same QP)
            cnt = fsteps[0, :, 0].astype(int)
            tab = np.full((20, 13), np.nan); tab[:, 0] = 0.0
            tab[:cnt.sum()] = np.repeat(fsteps[0], cnt, axis=0)
            tab[:cnt.sum(), 0] = 1.0
            fsteps = fsteps.copy(); fsteps[0] = tab
            assert cnt.sum() == N and fsteps[0, 7, 0] != 0.0
        px[...] = xref; pf[...] = fsteps
        staged.run(t, xref, fsteps)
        inplace.run(t, px, pf)
        xs, xi = staged.solution(), inplace.solution()
        a, b = staged.info(), inplace.info()
        assert (a["status"] == 1).all()
        np.testing.assert_array_equal(xi, xs)
        np.testing.assert_array_equal(inplace.forces(), staged.forces())
        for key in ("status", "sweeps", "iters", "obj", "contact", "active", "y"):
            np.testing.assert_array_equal(b[key], a[key], err_msg=key)
        sc.advance(xs[:, :12] + xref[:, :, 1])
    staged.close(); inplace.close()


def test_engine_against_an_independent_solver():
    """The engine against HiGHS' QP solver (bundled in scipy; see test_oracle_independent.py) on the QP the reference would have
    built for seeded closed-loop robots of every gait: the engine's point is feasible for that QP, its objective is never above
    HiGHS' and within HiGHS' accuracy of it.  No code of this repository's oracle solver is involved in the comparison."""
    from oracle import mpc_build
    from test_oracle_independent import hc, solve_highs
    B, T = 12, 3
    eng = mpcqp.Engine(batch=B)
    sc = Scenario(B, gaits=["trot", "pace", "bound", "walk"], seed=4242)
    par = mpc_build.Params(n_steps=N)
    compared, abstained = 0, []
    for t in range(T):
        xref, fsteps = sc.inputs()
        eng.run(t, xref, fsteps)
        x, info = eng.solution(), eng.info(with_y=False)
        assert (info["status"] == 1).all()
        for b in range(B):
            Pd, A, l, u, _ = mpc_build.build_qp(xref[b], fsteps[b], par, first_tick=(t == 0))
            Ax = A @ x[b]
            assert (Ax >= l - 1e-8).all() and (Ax <= u + 1e-8).all(), "tick %d robot %d: infeasible for the reference QP" % (t, b)
            obj = 0.5 * float((Pd * x[b] * x[b]).sum())
            assert abs(obj - info["obj"][b]) <= 1e-12 * max(1.0, abs(obj))
            # HiGHS' QP solver gives up on some instances (kSolveError); it is a witness that may abstain, at two accuracies
            for tol, gate in ((1e-10, 5e-6), (1e-7, 1e-4)):
                xh, status = solve_highs(Pd, A, l, u, tol)
                if status == hc.HighsModelStatus.kOptimal:
                    break
            else:
                abstained.append((t, b, sc.kinds[b]))
                continue
            ref = 0.5 * float((Pd * xh * xh).sum())
            slack = 10.0 * tol * float(np.abs(Pd * xh).sum())                 # what HiGHS' own infeasibility can buy it
            assert obj <= ref + 1e-9 * abs(ref) + slack + 1e-13, (t, b, obj, ref)
            assert ref - obj <= gate * abs(ref) + 1e-12, (t, b, obj, ref)
            compared += 1
        sc.advance(x[:, :12] + xref[:, :, 1])
    assert compared >= (3 * B * T) // 4, "HiGHS abstained on %s" % abstained
    eng.close()


def test_full_batch_properties():
    """BASELINE configs[1] size (4096 robots): EVERY robot of every tick passes the vectorised KKT certificate (tests/batch_kkt.py,
    pinned to the oracle's sparse certificate by tests/test_batch_kkt.py), the oracle's own certificate on a sample, warm-start
    invariance, friction / unilateral / fz_max feasibility everywhere, objective consistent with x."""
    B = 4096
    eng = mpcqp.Engine(batch=B)
    sc = Scenario(B, gaits="trot", seed=20260)
    w = np.concatenate([np.tile(np.array(eng.params.w_state[:]), N), np.full(12 * N, eng.params.w_force)])
    for t in range(5):
        xref, fsteps = sc.inputs()
        eng.run(t, xref, fsteps)
        x, info = eng.solution(), eng.info()
        assert (info["status"] == 1).all()
        f = x[:, 12 * N:].reshape(B, N, 4, 3)
        mu = eng.params.mu
        assert (np.abs(f[..., 0]) <= mu * f[..., 2] + 1e-8).all() and (np.abs(f[..., 1]) <= mu * f[..., 2] + 1e-8).all()
        assert (f[..., 2] >= -1e-9).all() and (f[..., 2] <= 25 + 1e-8).all()
        assert (f[~info["contact"]] == 0).all()
        np.testing.assert_allclose(info["obj"], 0.5 * (x * x * w).sum(axis=1), rtol=1e-12)
        cert = batch_kkt.certificate(xref, fsteps, x, info["y"], first_tick=(t == 0))
        batch_kkt.assert_batch_certified(cert, "tick %d" % t)
        np.testing.assert_array_equal(cert["contact"], info["contact"])
        if t == 4:
            for b in range(0, B, 173):
                assert_certified(certify(xref[b], fsteps[b], x[b], info["y"][b]), "robot %d" % b)
            fresh = mpcqp.Engine(batch=B, warm_start=0)
            fresh.run(t, xref, fsteps)
            assert np.abs(fresh.solution() - x).max() <= 1e-7
            fresh.close()
        sc.advance(x[:, :12] + xref[:, :, 1])
    eng.close()


@pytest.mark.parametrize("noise", [(0.0, 0.0, 0.0, 0.0), (1e-3, 5e-3, 1e-2, 2e-2)], ids=["noiseless", "hash-noise"])
def test_device_closed_loop_matches_host_loop(noise):
    """SURVEY 8f rows f1 + f2: planner + closed-loop integration inside the solve kernel against the
    host loop (scenario.py, itself pinned to the reference planner) driving the same engine tick by tick:
    generated xref / fsteps, forces and robot states must agree every tick."""
    B, T = 48, 14
    kw = dict(gaits=["trot", "pace", "bound", "walk", "static"], seed=31, noise=noise, noise_kind="hash")
    host_sc, dev_sc = Scenario(B, **kw), Scenario(B, **kw)
    host, dev = mpcqp.Engine(batch=B), mpcqp.Engine(batch=B)
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
    host.close(); dev.close()


def test_device_closed_loop_with_a_32_step_gait_period():
    """dt = 0.01 (main.py:20-23): the gait period is T_gait / dt = 32 steps, two 64-bit words of contact flags per robot.  Device
    planner + closed loop against the host twin (pinned to the reference planner at this dt by
    tests/test_planner.py::test_trot_planner_matches_reference_at_dt_001) over more than one gait period, every robot certified."""
    from oracle import mpc_build
    B, T = 36, 40
    kw = dict(n_steps=32, dt=0.01, T_gait=0.32, gaits=["trot", "pace", "bound", "walk"], seed=58, noise=(1e-3, 5e-3, 1e-2, 2e-2),
              noise_kind="hash")
    host_sc, dev_sc = Scenario(B, **kw), Scenario(B, **kw)
    assert host_sc.period == 32 and dev_sc.seq_bits().shape == (B, 2)
    ekw = dict(n_steps=32, dt=0.01, T_gait=0.32)
    host, dev = mpcqp.Engine(batch=B, **ekw), mpcqp.Engine(batch=B, **ekw)
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
        info = dev.info()
        assert (info["status"] == 1).all()
        np.testing.assert_allclose(dev.forces(), host.forces(), rtol=0, atol=1e-6)
        if t % 13 == 0:
            cert = batch_kkt.certificate(xd_ref, fd, dev.solution(), info["y"], mpc_build.Params(n_steps=32, dt=0.01, T_gait=0.32), first_tick=(t == 0))
            batch_kkt.assert_batch_certified(cert, "tick %d" % t)
        host_sc.advance(xh[:, :12] + xref[:, :, 1])
        st = dev.scenario_state()
        np.testing.assert_allclose(st["state"], host_sc.state, rtol=0, atol=1e-9)
        np.testing.assert_allclose(st["frame"], host_sc.frame, rtol=0, atol=1e-9)
    # several ticks in one call (the overlapped index ranges of mpcqp_scenario_run) continue the same trajectory
    dev.scenario_run(5)
    for t in range(T, T + 5):
        xref, fsteps = host_sc.inputs()
        host.run(t, xref, fsteps)
        host_sc.advance(host.solution()[:, :12] + xref[:, :, 1])
    np.testing.assert_allclose(dev.scenario_state()["state"], host_sc.state, rtol=0, atol=1e-8)
    host.close(); dev.close()


def test_device_planner_follows_changing_commands():
    """SURVEY 8f row f1, the rest of the planner: joystick commands that change between ticks, including vz / roll-rate /
    pitch-rate (getRefStates' state machine, FootstepPlanner.py:128-152) and the `reduced` support polygon
    (FootstepPlanner.py:330-332).  Device planner vs the host twin (pinned to the reference planner by
    tests/test_planner.py::test_command_state_machine_and_reduced_polygon_match_reference): inputs, forces, states."""
    B, T = 40, 36
    kw = dict(gaits=["trot", "pace", "bound", "walk"], seed=77, noise=(1e-3, 5e-3, 1e-2, 2e-2), noise_kind="hash")
    host_sc, dev_sc = Scenario(B, **kw), Scenario(B, **kw)
    host, dev = mpcqp.Engine(batch=B), mpcqp.Engine(batch=B)
    dev.scenario_init(dev_sc)
    rng = np.random.default_rng(5)
    base = host_sc.v_ref.copy()
    flags = set()
    for t in range(T):
        if t % 6 == 0:                                   # new commands every 6 ticks; vz toggles above / below the dead band
            v = base.copy()
            phase = (t // 6) % 4
            v[:, 2] = [0.0, 0.09, 0.02, -0.07][phase] * np.sign(rng.standard_normal(B))
            v[:, 3] = rng.uniform(-0.2, 0.2, B) * (phase > 0)
            v[:, 4] = rng.uniform(-0.2, 0.2, B) * (phase > 0)
            red = phase >= 2
            host_sc.set_v_ref(v); host_sc.reduced = red
            dev.scenario_set_commands(v, red)
        xref, fsteps = host_sc.inputs()
        flags |= set(int(f) for f in host_sc.cmd_flag)
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
    assert flags == {0, 1, 2}
    host.close(); dev.close()


def test_mixed_gait_sweep_on_device():
    """BASELINE configs[2] shape (65 536 robots, trot / pace / bound / walk, per-instance contact masks),
    run as a device-resident closed loop: everything solved, forces feasible, contact masks follow the
    gait tables, and a sample passes the oracle certificate on the inputs the device generated."""
    B, T = 65536, 6
    sc = Scenario(B, gaits=["trot", "pace", "bound", "walk"], seed=99, noise_kind="hash")
    eng = mpcqp.Engine(batch=B)
    eng.scenario_init(sc)
    eng.scenario_run(T - 1)
    eng.scenario_run(1, emit_inputs=True)
    info, x = eng.info(), eng.solution()
    assert (info["status"] == 1).all()
    f = x[:, 12 * N:].reshape(B, N, 4, 3)
    mu = eng.params.mu
    assert (np.abs(f[..., 0]) <= mu * f[..., 2] + 1e-8).all() and (np.abs(f[..., 1]) <= mu * f[..., 2] + 1e-8).all()
    assert (f[..., 2] >= -1e-9).all() and (f[..., 2] <= 25 + 1e-8).all() and (f[~info["contact"]] == 0).all()
    idx = (T - 1 + sc.phase[:, None] + np.arange(N)[None, :]) % 16
    expect = np.take_along_axis(sc.seq, idx[:, :, None], axis=1) == 1.0
    np.testing.assert_array_equal(info["contact"], expect)
    xref, fsteps = eng.last_inputs()
    cert = batch_kkt.certificate(xref, fsteps, x, info["y"])                  # all 65 536 robots
    batch_kkt.assert_batch_certified(cert, "mixed sweep")
    for b in range(0, B, 4099):
        assert_certified(certify(xref[b], fsteps[b], x[b], info["y"][b]), "robot %d" % b)
    eng.close()
