"""Exact active-set polish and KKT certificate for  min 1/2 x'Px + q'x  s.t.  l <= Ax <= u.
TEST INFRASTRUCTURE ONLY (see oracle/osqp_port.py header): never imported by the product path.

The reference runs OSQP at eps 1e-7 with polish off (MPC.py:414-416) and never checks the status.
For a parity gate of 1e-4 N on a QP whose smallest curvature is 1e-5 (force weight, MPC.py:282-284)
an eps-1e-8 ADMM point is not sharp enough (error ~ residual / 1e-5), so the oracle sharpens the
ADMM point on its active set and then PROVES the result with the KKT conditions.  The QP is strictly
convex, hence a point passing `certificate` is the unique optimum regardless of which solver found it.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

INF = 1e20


def active_sets(A, l, u, x, y, z=None):
    """OSQP's polish rule: lower-active if z - l < -y, upper-active if u - z < y."""
    Ax = A @ x if z is None else z
    low = (Ax - l < -y) & (l > -INF)
    upp = (u - Ax < y) & (u < INF)
    eq = (u - l) <= 0.0
    low = low & ~eq
    upp = upp | eq
    low = low & ~upp
    return low, upp


def polish(P, q, A, l, u, x, y, z=None, delta=1e-7, refine=60, tol=1e-14):
    """Solve the equality-constrained QP on the guessed active set.

    Regularised KKT [[P + delta I, Aa'], [Aa, -delta I]] factored once, iterative refinement against
    the un-regularised system (the scheme of OSQP's own polish step) until the correction stalls.
    All-zero rows (MPC.py stores the swing-pin coefficient explicitly even when it is 0) and
    linearly dependent active rows (pyramid apex) make the un-regularised system singular but
    consistent; the refinement converges to a minimum-norm-ish multiplier, which is all we need
    because only x is compared and the multipliers are re-derived by `certificate`.
    """
    A = sp.csc_matrix(A)
    n = A.shape[1]
    low, upp = active_sets(A, l, u, x, y, z)
    rows = np.flatnonzero(low | upp)
    b = np.where(low, l, u)[rows]
    Aa = A.tocsr()[rows].tocsc()
    ma = len(rows)
    K = sp.bmat([[P, Aa.T], [Aa, None]], format="csc") if ma else sp.csc_matrix(P)
    Kreg = (K + sp.diags(np.concatenate([np.full(n, delta), np.full(ma, -delta)]))).tocsc()
    solve = spla.factorized(Kreg)
    rhs = np.concatenate([-q, b])
    sol = solve(rhs)
    for _ in range(refine):
        r = rhs - K @ sol
        d = solve(r)
        sol = sol + d
        if np.abs(d).max() <= tol * max(1.0, np.abs(sol).max()):
            break
    xp = sol[:n]
    yp = np.zeros(A.shape[0])
    yp[rows] = sol[n:]
    yp = sign_feasible_multipliers(P, q, A, l, u, xp, yp, low, upp)
    return xp, yp, low, upp


def sign_feasible_multipliers(P, q, A, l, u, x, y, low, upp, force=False):
    """When active rows are linearly dependent (a swing foot is pinned by three equalities AND sits on
    all five pyramid rows; a stance foot unloaded to the apex) the multipliers are not unique and the
    linear solve may return a sign-infeasible choice although a feasible one exists.  Re-derive them:
    least squares on stationarity over the active rows with y >= 0 on upper-active inequality rows,
    y <= 0 on lower-active ones and free sign on equalities (bounded-variable least squares)."""
    eq = (u - l) <= 0.0
    ok = np.all(y[upp & ~eq] >= 0.0) and np.all(y[low] <= 0.0)
    if ok and not force:
        return y
    from scipy.optimize import lsq_linear
    rows = np.flatnonzero(low | upp)
    At = sp.csc_matrix(A).tocsr()[rows].T.toarray()          # n x ma
    lo = np.where(eq[rows], -np.inf, np.where(upp[rows], 0.0, -np.inf))
    hi = np.where(eq[rows], np.inf, np.where(upp[rows], np.inf, 0.0))
    sol = lsq_linear(At, -(P @ x + q), bounds=(lo, hi), method="bvls", tol=1e-15, max_iter=2000)
    out = np.zeros_like(y)
    out[rows] = sol.x
    return out


def certificate(P, q, A, l, u, x, y):
    """KKT residuals of (x, y): primal feasibility, stationarity, dual sign / complementarity, gap."""
    Ax = A @ x
    prim = max(0.0, float(np.max(np.maximum(l - Ax, 0.0))), float(np.max(np.maximum(Ax - u, 0.0))))
    stat = float(np.abs(P @ x + q + A.T @ y).max())
    yp, ym = np.maximum(y, 0.0), np.minimum(y, 0.0)
    fin_u, fin_l = u < INF, l > -INF
    # a positive multiplier needs a finite upper bound that is attained, a negative one a lower bound
    comp = 0.0
    if np.any(fin_u):
        comp = max(comp, float(np.abs(yp[fin_u] * (u[fin_u] - Ax[fin_u])).max()))
    if np.any(fin_l):
        comp = max(comp, float(np.abs(ym[fin_l] * (Ax[fin_l] - l[fin_l])).max()))
    bad_sign = 0.0
    if np.any(~fin_u):
        bad_sign = max(bad_sign, float(yp[~fin_u].max(initial=0.0)))
    if np.any(~fin_l):
        bad_sign = max(bad_sign, float((-ym[~fin_l]).max(initial=0.0)))
    obj = float(0.5 * x @ (P @ x) + q @ x)
    # Lagrange dual value at y: -1/2 x'Px - sup-function of [l,u] at y (with x the stationarity point)
    dual = float(-0.5 * x @ (P @ x) - (np.where(fin_u, u, 0.0) @ yp + np.where(fin_l, l, 0.0) @ ym))
    return dict(prim=prim, stat=stat, comp=comp, bad_sign=bad_sign, obj=obj, gap=obj - dual)


def is_certified(cert, prim=1e-9, stat=1e-11, comp=1e-9, sign=1e-9):
    return (cert["prim"] <= prim and cert["stat"] <= stat and cert["comp"] <= comp
            and cert["bad_sign"] <= sign)


def solve_certified(P, q, A, l, u, solver, max_rounds=6, run=None):
    """Run `solver` (an osqp_port.OSQP already set up / updated / warm-started), polish, certify.
    If the guessed active set is wrong the certificate fails; tighten and retry (the guard the
    survey calls for: polish after too few iterations can be off by 0.4 N on contact-switch ticks)."""
    res = None
    for rnd in range(max_rounds):
        res = run() if run is not None else solver.solve()
        xp, yp, low, upp = polish(P, q, A, l, u, res.x, res.y, res.z)
        cert = certificate(P, q, A, l, u, xp, yp)
        if is_certified(cert):
            return xp, yp, cert, res
        solver.update_settings(eps_abs=solver.settings["eps_abs"] * 0.1,
                               eps_rel=solver.settings["eps_rel"] * 0.1)
    raise RuntimeError("oracle could not certify a solution: %r" % (cert,))


def lift_multipliers(P, q, A, l, u, x, y_ineq):
    """Complete a multiplier vector known only on the inequality rows (u > l) to all rows: the
    equality rows' multipliers are free in sign, so take the least-squares solution of stationarity
        P x + q + A_eq' nu + A_in' y_in = 0.
    Used to run `certificate` on a solver that works on a condensed form of the same QP (the CUDA
    engine eliminates the states and the swing-foot forces) -- if the lifted pair passes, x is the
    optimum of the reference's QP."""
    A = sp.csr_matrix(A)
    eq = (u - l) <= 0.0
    y = np.zeros(A.shape[0])
    y[~eq] = np.asarray(y_ineq, dtype=np.float64).ravel()
    r = -(P @ x + q + A.T @ y)
    Aeq_t = A[np.flatnonzero(eq)].T.toarray()
    nu, *_ = np.linalg.lstsq(Aeq_t, r, rcond=None)
    y[eq] = nu
    return y
