"""The vectorised whole-batch certificate (tests/batch_kkt.py) against the oracle's sparse one (tests/common.certify) on the
golden fixtures: same verdicts, same numbers -- so that the full-batch GPU tests may certify EVERY robot with it."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

import batch_kkt
from oracle import kkt, mpc_build

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "solve_*.npz")) + glob.glob(os.path.join(HERE, "golden", "horizon_*.npz")))


def _ineq_multipliers(g, t, p):
    """Sign-feasible multipliers of the 20N pyramid rows at the golden optimum: non-negative least squares on the stationarity of
    the reference-layout QP, restricted to the rows that hold with equality (at the apex of a pyramid five rows meet in a point of
    R^3, so the multipliers are not unique and a plain least-squares solution need not have the right signs)."""
    from scipy.optimize import nnls
    n = p.n_steps
    Pd, A, l, u, _ = mpc_build.build_qp(g["xref"][t], g["fsteps"][t], p, first_tick=(g["k"][t] == 0))
    P, q = sp.diags(Pd).tocsc(), np.zeros(24 * n)
    x = g["x"][t]
    A = sp.csr_matrix(A)
    Ax = A @ x
    eq = np.flatnonzero(u - l <= 0.0)
    up = np.flatnonzero((u - l > 0.0) & (np.abs(Ax - u) <= 1e-9))          # multiplier >= 0
    lo = np.flatnonzero((u - l > 0.0) & (np.abs(Ax - l) <= 1e-9))          # multiplier <= 0
    M = np.hstack([A[eq].T.toarray(), -A[eq].T.toarray(), A[up].T.toarray(), -A[lo].T.toarray()])
    z, res = nnls(M, -(Pd * x), maxiter=20 * M.shape[1])
    assert res <= 1e-10, res
    y = np.zeros(A.shape[0])
    y[eq] = z[:len(eq)] - z[len(eq):2 * len(eq)]
    y[up] = z[2 * len(eq):2 * len(eq) + len(up)]
    y[lo] = -z[2 * len(eq) + len(up):]
    return y[24 * n:], kkt.certificate(P, q, A, l, u, x, y)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(q)[:-4] for q in GOLD])
def test_batch_certificate_agrees_with_the_oracle(path):
    g = np.load(path)
    n = g["x"].shape[1] // 24
    p = mpc_build.Params(dt=float(g["dt"]) if "dt" in g else 0.02, n_steps=n, T_gait=float(g["T_gait"]) if "T_gait" in g else 0.32)
    for first in ((False,) if n > 32 else (True, False)):              # the sign-feasible multipliers cost an NNLS: seconds at N >= 48
        ts = [t for t in range(min(len(g["k"]), 6)) if (g["k"][t] == 0) == first][:1 if n > 32 else 2]
        if not ts:
            continue
        ys, certs = zip(*[_ineq_multipliers(g, t, p) for t in ts])
        xref, fsteps, x, y = g["xref"][ts], g["fsteps"][ts], g["x"][ts], np.stack(ys)
        c = batch_kkt.certificate(xref, fsteps, x, y, p, first_tick=first)
        batch_kkt.assert_batch_certified(c, os.path.basename(path))
        for i, oc in enumerate(certs):
            assert abs(c["obj"][i] - oc["obj"]) <= 1e-12 * max(1.0, abs(oc["obj"]))
            assert c["stat"][i] <= 1e-10 and oc["stat"] <= 1e-10
        _, _, _, _, contact = mpc_build.build_qp(xref[0], fsteps[0], p, first_tick=first)
        np.testing.assert_array_equal(c["contact"][0], contact.astype(bool))
        # what the oracle rejects, this must reject: a force nudged by 1e-3 N breaks stationarity, a flipped multiplier its sign
        xb = x.copy()
        j = 12 * n + int(np.flatnonzero(np.abs(x[0, 12 * n:]) > 1e-3)[0])
        xb[0, j] += 1e-3
        cb = batch_kkt.certificate(xref, fsteps, xb, y, p, first_tick=first)
        assert cb["stat"][0] > 1e-9 or cb["dyn"][0] > 1e-9
        if np.abs(y[0]).max() > 1e-6:
            yb = y.copy()
            stance = np.repeat(c["contact"][0].reshape(-1), 5)                # multipliers of swing feet (pinned to 0) do not count
            k = int(np.argmax(np.abs(y[0]) * (np.arange(y.shape[1]) % 5 != 4) * stance))
            if abs(yb[0, k]) > 1e-6:
                yb[0, k] = -yb[0, k]
                cb = batch_kkt.certificate(xref, fsteps, x, yb, p, first_tick=first)
                assert cb["bad_sign"][0] > 1e-7 or cb["stat"][0] > 1e-7
