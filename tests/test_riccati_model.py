"""The stage-wise (Riccati) equality-constrained solve (numpy model of mpcqp_riccati.cuh) against the dense
Woodbury-form solve of the same model and against the golden optima the reference run stored."""
import glob
import os

import numpy as np
import pytest

import kernel_model as km

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "solve_*.npz")))


def _sig_from_forces(p, f, contact):
    """The face every foot-step of an optimal f sits on."""
    N, mu, tol = p.N, p.mu, 1e-7
    sig = np.zeros((N, 4, 3), np.int8)
    for k in range(N):
        for j in range(4):
            if not contact[k, j]:
                continue
            fx, fy, fz = f[k, j]
            if fz < tol:
                sig[k, j] = [0, 0, 1]
                continue
            sx = 1 if fx - mu * fz > -tol else (-1 if -fx - mu * fz > -tol else 0)
            sy = 1 if fy - mu * fz > -tol else (-1 if -fy - mu * fz > -tol else 0)
            sig[k, j] = [sx, sy, 2 if fz > p.fz_max - tol else 0]
    return sig


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(q)[6:-4] for q in GOLD])
def test_riccati_on_the_optimal_face_reproduces_golden(path):
    g = np.load(path)
    N = g["x"].shape[1] // 24
    p = km.ModelParams(n_steps=N)
    worst = 0.0
    for t in range(0, len(g["k"]), 3):
        xref, fsteps = g["xref"][t], g["fsteps"][t]
        contact, Bv = km.decode(p, xref, fsteps, first_tick=(g["k"][t] == 0))
        fstar = g["x"][t][12 * N:].reshape(N, 4, 3)
        sig = _sig_from_forces(p, fstar, contact)
        f, grad, Xs = km.riccati_solve(p, xref, Bv, contact, sig)
        worst = max(worst, np.abs(f - fstar).max())
        xs = (Xs - xref[:, 1:].T).reshape(-1)
        np.testing.assert_allclose(xs, g["x"][t][:12 * N], rtol=0, atol=1e-8)
    assert worst < 1e-7, worst


def test_riccati_matches_dense_polish_on_arbitrary_faces():
    g = np.load([q for q in GOLD if q.endswith("solve_trot.npz")][0])
    p = km.ModelParams()
    eng = km.Engine(p)
    rng = np.random.default_rng(3)
    for t in (1, 7):
        xref, fsteps = g["xref"][t], g["fsteps"][t]
        contact, Bv = km.decode(p, xref, fsteps)
        _, _, gam, _ = km.free_response(p, xref)
        gg = eng._back(Bv, contact, gam)
        for trial in range(4):
            sig = np.stack([rng.integers(-1, 2, (p.N, 4)), rng.integers(-1, 2, (p.N, 4)), rng.integers(0, 3, (p.N, 4))], axis=2).astype(np.int8)
            sig = sig * contact[:, :, None]
            ok, fq, y, nsig = eng._polish(Bv, contact, gg, sig)
            f, grad, _ = km.riccati_solve(p, xref, Bv, contact, sig)
            np.testing.assert_allclose(f, fq, rtol=0, atol=1e-7)
            gd = eng._hess_apply(Bv, contact, fq) + gg
            np.testing.assert_allclose(grad, gd * contact[:, :, None], rtol=0, atol=1e-8)


def test_square_root_free_stage_equals_the_cholesky_form():
    """The kernel factors Pvv = U D U' and H = D^-1 + U'EU = W Delta W' (no square roots on the pivot chain); the rows it
    forms must be those of the Cholesky form the recursion is written in: Pt[:, v] = P[:, v] (I - Gamma Pvv) and P[:, v] Gamma
    with Gamma = (I + E Pvv)^-1 E."""
    rng = np.random.default_rng(2)
    for trial in range(5):
        A = rng.normal(size=(12, 12))
        P = A @ A.T + 0.1 * np.eye(12)
        Ppp, Ppv, Pvv = P[:6, :6], P[:6, 6:], P[6:, 6:]
        pp, pv = rng.normal(size=6), rng.normal(size=6)
        F = rng.normal(size=(6, 4 if trial % 2 else 7)) * (10.0 ** rng.integers(-1, 3))
        E = F @ F.T                                         # PSD, rank-deficient every other trial
        tr, kr = km.riccati_stage_ldl(Ppp, Ppv, Pvv, pp, pv, E)
        Gam = np.linalg.solve(np.eye(6) + E @ Pvv, E)
        rows = np.vstack([Ppv, Pvv, pv[None, :]])
        np.testing.assert_allclose(kr, rows @ Gam, rtol=1e-9, atol=1e-9 * np.abs(rows @ Gam).max())
        np.testing.assert_allclose(tr, rows @ (np.eye(6) - Gam @ Pvv), rtol=1e-9, atol=1e-9 * np.abs(rows).max())


def test_one_row_change_on_the_kept_factorisation_equals_a_full_resolve():
    """Round-2 groundwork: when the guard rejects a sweep because ONE friction row of ONE foot-step is on the
    wrong side (99.7 % of the rejected first sweeps in steady trot), the optimum on the corrected face follows from the kept
    factorisation by one vector-only pass and a scalar step -- no new 6x6 factorisations.  Checked here on the numpy model in both
    directions (a row that must be released, a row that must become active) against solving the corrected face from scratch."""
    p = km.ModelParams()
    mu = p.mu
    checked = {"release": 0, "activate": 0}
    for name, t in (("trot", 1), ("trot", 7), ("aggressive", 2), ("pace", 3), ("bound", 2)):
        g = np.load([q for q in GOLD if q.endswith("solve_%s.npz" % name)][0])
        xref, fsteps = g["xref"][t], g["fsteps"][t]
        contact, Bv = km.decode(p, xref, fsteps)
        fstar = g["x"][t][12 * p.N:].reshape(p.N, 4, 3)
        sig_opt = _sig_from_forces(p, fstar, contact)
        f_ref, grad_ref, X_ref = km.riccati_solve(p, xref, Bv, contact, sig_opt)
        for ks in range(p.N):
            for js in range(4):
                if not contact[ks, js] or sig_opt[ks, js, 2] != 0:
                    continue
                for axis in (0, 1):
                    wrong = sig_opt.copy()
                    unit = np.eye(3)[axis]
                    if sig_opt[ks, js, axis] == 0:
                        # the first sweep wrongly held the row (+1 side) active: release it
                        wrong[ks, js, axis] = 1
                        f0, g0, X0, kept = km.riccati_solve(p, xref, Bv, contact, wrong, keep=True)
                        f1, g1, X1 = km.riccati_change_one_row(p, Bv, contact, kept, f0, g0, X0, ks, js, release=unit)
                        checked["release"] += 1
                    else:
                        # the first sweep wrongly left the row free: activate it
                        sgn = float(sig_opt[ks, js, axis])
                        wrong[ks, js, axis] = 0
                        c = sgn * unit - np.array([0.0, 0.0, mu])
                        f0, g0, X0, kept = km.riccati_solve(p, xref, Bv, contact, wrong, keep=True)
                        f1, g1, X1 = km.riccati_change_one_row(p, Bv, contact, kept, f0, g0, X0, ks, js, activate=(c, 0.0))
                        checked["activate"] += 1
                    np.testing.assert_allclose(f1, f_ref, rtol=0, atol=1e-8)
                    np.testing.assert_allclose(X1, X_ref, rtol=0, atol=1e-9)
                    live = contact[:, :, None] & np.ones(3, bool)
                    live[ks, js] = False                      # at the changed foot-step the two faces define grad on different rows
                    np.testing.assert_allclose(g1[live], grad_ref[live], rtol=0, atol=1e-8)
    assert checked["release"] > 50 and checked["activate"] > 3, checked
