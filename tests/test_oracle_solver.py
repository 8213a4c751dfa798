"""The restated OSQP algorithm + polish + certificate on problems with known answers and on the
golden QPs (whose stored solutions were KKT-certified when generated)."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import kkt, mpc_build
from oracle.osqp_port import OSQP

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "solve_*.npz")))


def test_box_qp_known_answer():
    # min 1/2 (x0^2 + x1^2) - x0 - 3 x1   s.t. 0 <= x <= 2  ->  x = (1, 2), multiplier of the upper bound on x1 = 1
    P = sp.identity(2, format="csc")
    q = np.array([-1.0, -3.0])
    A = sp.identity(2, format="csc")
    l, u = np.zeros(2), np.full(2, 2.0)
    s = OSQP()
    s.setup(P=P, q=q, A=A, l=l, u=u, eps_abs=1e-9, eps_rel=1e-9)
    r = s.solve()
    assert r.info.status == "solved"
    np.testing.assert_allclose(r.x, [1.0, 2.0], atol=1e-6)
    xp, yp, low, upp = kkt.polish(P, q, A, l, u, r.x, r.y, r.z)
    np.testing.assert_allclose(xp, [1.0, 2.0], atol=1e-12)
    np.testing.assert_allclose(yp, [0.0, 1.0], atol=1e-9)
    assert kkt.is_certified(kkt.certificate(P, q, A, l, u, xp, yp))


def test_equality_qp_known_answer():
    # min 1/2 |x|^2  s.t. x0 + x1 = 1 -> x = (.5, .5), y = -.5
    P = sp.identity(2, format="csc")
    A = sp.csc_matrix(np.array([[1.0, 1.0]]))
    s = OSQP()
    s.setup(P=P, q=np.zeros(2), A=A, l=np.array([1.0]), u=np.array([1.0]), eps_abs=1e-9, eps_rel=1e-9)
    r = s.solve()
    np.testing.assert_allclose(r.x, [0.5, 0.5], atol=1e-7)
    np.testing.assert_allclose(r.y, [-0.5], atol=1e-6)


def test_certificate_rejects_wrong_points():
    P = sp.identity(2, format="csc")
    q = np.array([-1.0, -3.0])
    A = sp.identity(2, format="csc")
    l, u = np.zeros(2), np.full(2, 2.0)
    assert not kkt.is_certified(kkt.certificate(P, q, A, l, u, np.array([1.0, 1.9]), np.array([0.0, 1.0])))
    assert not kkt.is_certified(kkt.certificate(P, q, A, l, u, np.array([1.0, 2.1]), np.array([0.0, 0.9])))
    assert not kkt.is_certified(kkt.certificate(P, q, A, l, u, np.array([1.0, 2.0]), np.array([0.0, -1.0])))


@pytest.mark.parametrize("name", ["trot", "aggressive"])
def test_port_reproduces_golden(name):
    """Cold solve of two golden ticks: ADMM at 1e-8 lands within 1e-2 N (error ~ residual / 1e-5), the polished point on the
    stored optimum to 1e-9 N and passes the certificate."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "solve_%s.npz" % name))
    N = 16
    for t in (0, 3):
        Pd, A, l, u, _ = mpc_build.build_qp(g["xref"][t], g["fsteps"][t], first_tick=(t == 0))
        P, q = sp.diags(Pd).tocsc(), np.zeros(24 * N)
        s = OSQP()
        s.setup(P=P, q=q, A=A, l=l, u=u, eps_abs=1e-8, eps_rel=1e-8)
        xp, yp, cert, raw = kkt.solve_certified(P, q, A, l, u, s)
        assert np.abs(raw.x - g["x"][t]).max() < 1e-2        # eps 1e-8 residuals do NOT pin forces to 1e-4 N
        assert np.abs(xp - g["x"][t]).max() < 1e-9
        assert kkt.is_certified(cert)
        assert abs(cert["obj"] - g["obj"][t]) <= 1e-12 * max(1.0, abs(g["obj"][t]))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[6:-4] for p in GOLD])
def test_golden_solutions_are_optimal(path):
    """Independent of any solver: every stored x satisfies the KKT conditions of the QP MPC.py built."""
    g = np.load(path)
    N = g["x"].shape[1] // 24
    for t in range(len(g["ML_data"])):
        A = sp.csc_matrix((g["ML_data"][t], g["ML_indices"], g["ML_indptr"]), shape=(44 * N, 24 * N))
        P, q = sp.diags(g["P_data"]).tocsc(), np.zeros(24 * N)
        l, u = g["NK_inf"][t], g["NK"][t]
        x = g["x"][t]
        low, upp = kkt.active_sets(A, l, u, x, np.zeros(44 * N))
        Ax = A @ x
        tight_u = np.abs(Ax - u) <= 1e-9
        tight_l = np.abs(Ax - l) <= 1e-9
        y = kkt.sign_feasible_multipliers(P, q, A, l, u, x, np.zeros(44 * N), tight_l & ~tight_u, tight_u, force=True)
        cert = kkt.certificate(P, q, A, l, u, x, y)
        assert cert["prim"] <= 1e-9 and cert["stat"] <= 1e-10 and cert["bad_sign"] <= 1e-15, (t, cert)
        assert g["cert_stat"][t] <= 1e-11 and g["cert_prim"][t] <= 1e-9
