"""Shared helpers for the parity tests: certify an engine solution with the oracle's KKT check on
the reference-layout QP (oracle.mpc_build restates MPC.py's matrices; pinned by tests/golden)."""
import numpy as np
import scipy.sparse as sp

from oracle import kkt, mpc_build

FORCE_TOL = 1e-4        # N, per component      (BASELINE.json north_star)
OBJ_RTOL = 1e-6         # relative objective    (BASELINE.json north_star)


def certify(xref, fsteps, x, y_cone, first_tick=False, params=None):
    """KKT certificate of the engine's (x, y) on the QP the reference would have built."""
    p = params or mpc_build.Params()
    N = p.n_steps
    Pd, A, l, u, contact = mpc_build.build_qp(xref, fsteps, p, first_tick=first_tick)
    P = sp.diags(Pd).tocsc()
    q = np.zeros(24 * N)
    y = kkt.lift_multipliers(P, q, A, l, u, x, y_cone)
    cert = kkt.certificate(P, q, A, l, u, x, y)
    cert["contact"] = contact
    Ax = A @ x
    rows = slice(24 * N, 44 * N)
    cert["active"] = (np.abs(Ax[rows] - u[rows]) <= 1e-9) | (np.abs(Ax[rows] - l[rows]) <= 1e-9)
    return cert


def assert_certified(cert, where=""):
    assert cert["prim"] <= 1e-8, (where, cert)
    assert cert["stat"] <= 1e-10, (where, cert)
    assert cert["comp"] <= 1e-8, (where, cert)
    assert cert["bad_sign"] <= 1e-9, (where, cert)
