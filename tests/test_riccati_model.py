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
