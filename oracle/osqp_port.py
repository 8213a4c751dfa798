"""CPU restatement of the OSQP algorithm the reference drives from MPC.py -- TEST INFRASTRUCTURE ONLY.

Nothing in the product path may import this module: it is the checker for the CUDA engine
(`tests/`, `__graft_entry__.smoke()`, and `bench.py`'s cpu_baseline / --impl reference legs).

Why a restatement: the reference's solve lives in the third-party `osqp` package (C core + QDLDL),
which is NOT vendored under /root/reference, is version-unpinned there (README: `pip3 install --user
osqp`; era => 0.6.x) and is not installable in this image (no wheel, no network).  Call sites this
port serves, in the reference:
    MPC.py:73        osqp.OSQP()
    MPC.py:414       prob.setup(P, q, A, l, u, verbose=False)
    MPC.py:415-416   prob.update_settings(eps_abs=..), prob.update_settings(eps_rel=..)
    MPC.py:419       prob.update(Ax=ML.data, l=.., u=..)
    MPC.py:420       prob.warm_start(x=initx)
    MPC.py:427-428   sol = prob.solve(); sol.x

Algorithm (Stellato, Banjac, Goulart, Bemporad, Boyd: "OSQP: an operator splitting solver for
quadratic programs", Math. Prog. Comp. 2020, and the 0.6 defaults), restated from the paper:
    minimise 1/2 x'Px + q'x   subject to   l <= Ax <= u
  * modified Ruiz equilibration of the KKT matrix (`scaling` = 10 passes) plus cost scaling;
  * each iteration solves  [[P + sigma I, A'], [A, -diag(1/rho_i)]] [xt; nu] = [sigma x - q; z - y/rho]
    with sigma = 1e-6, rho = 0.1 and rho_i = 1e3 rho on equality rows, then
        zt = z + (nu - y)/rho,  x+ = alpha xt + (1-alpha) x,
        z+ = clip(alpha zt + (1-alpha) z + y/rho, l, u),  y+ = y + rho (alpha zt + (1-alpha) z - z+)
    with relaxation alpha = 1.6;
  * termination test every `check_termination` = 25 iterations on UNSCALED infinity-norm residuals;
  * adaptive rho (OSQP picks its interval from wall-clock timings, which makes the iterate path
    non-deterministic; this port uses a fixed interval -- irrelevant to the unique optimum);
  * warm_start(x) sets x and z = Ax and keeps the previous y; polish is off in the reference.

PARITY STATUS: *unpinned by the reference* -- the reference holds no golden vectors or asserting
tests for this boundary and the real package cannot be run here.  Results are instead certified by
their KKT residuals (see oracle/kkt.py): the QP is strictly convex (P diagonal > 0, MPC.py:255-275)
so a KKT-certified point is THE optimum any correct OSQP run converges to.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

OSQP_INFTY = 1e30
MIN_SCALING = 1e-4
MAX_SCALING = 1e4
RHO_MIN = 1e-6
RHO_MAX = 1e6
RHO_TOL = 1e-4
RHO_EQ_OVER_RHO_INEQ = 1e3

DEFAULTS = dict(rho=0.1, sigma=1e-6, alpha=1.6, scaling=10, max_iter=4000,
                eps_abs=1e-3, eps_rel=1e-3, check_termination=25,
                adaptive_rho=True, adaptive_rho_interval=100, adaptive_rho_tolerance=5.0,
                verbose=False, polish=False, warm_start=True)


class _Info:
    pass


class _Result:
    pass


def _limit_scaling(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.minimum(v, MAX_SCALING)


class OSQP:
    """Same call surface as `osqp.OSQP` for the five methods MPC.py uses."""

    def __init__(self):
        self.settings = dict(DEFAULTS)
        self._is_setup = False

    # ------------------------------------------------------------------ setup / update
    def setup(self, P=None, q=None, A=None, l=None, u=None, **settings):
        self.settings.update({k: v for k, v in settings.items() if k in self.settings})
        self.P0 = sp.csc_matrix(P).astype(np.float64)
        # OSQP keeps the upper triangle of P; MPC.py only ever passes a diagonal P.
        self.P0 = sp.triu(self.P0, format="csc")
        self.P0 = (self.P0 + sp.triu(self.P0, 1).T).tocsc()
        self.A0 = sp.csc_matrix(A).astype(np.float64).copy()
        self.A0.sort_indices()
        self.n = self.A0.shape[1]
        self.m = self.A0.shape[0]
        self.q0 = np.zeros(self.n) if q is None else np.asarray(q, dtype=np.float64).ravel().copy()
        self.l0 = np.maximum(np.asarray(l, dtype=np.float64).ravel(), -OSQP_INFTY)
        self.u0 = np.minimum(np.asarray(u, dtype=np.float64).ravel(), OSQP_INFTY)
        self.x = np.zeros(self.n)
        self.z = np.zeros(self.m)
        self.y = np.zeros(self.m)
        self.rho = float(self.settings["rho"])
        self._scale()
        self._set_rho_vec()
        self._factor()
        self._is_setup = True
        return 0

    def update_settings(self, **kw):
        for k, v in kw.items():
            if k not in self.settings:
                raise ValueError("unknown setting %s" % k)
            self.settings[k] = v
        if "rho" in kw:
            self.rho = float(kw["rho"])
            self._set_rho_vec()
            self._factor()
        return 0

    def update(self, q=None, l=None, u=None, Px=None, Ax=None):
        # Unscaled copies are kept, so "unscale, overwrite, rescale" (what OSQP 0.6 does on a matrix
        # update) is simply "overwrite, rescale".
        rescale = False
        if Px is not None:
            raise NotImplementedError("MPC.py never updates P")
        if Ax is not None:
            Ax = np.asarray(Ax, dtype=np.float64).ravel()
            if Ax.shape[0] != self.A0.data.shape[0]:
                raise ValueError("Ax has the wrong number of stored entries")
            self.A0.data[:] = Ax
            rescale = True
        if q is not None:
            self.q0 = np.asarray(q, dtype=np.float64).ravel().copy()
        if l is not None:
            self.l0 = np.maximum(np.asarray(l, dtype=np.float64).ravel(), -OSQP_INFTY)
        if u is not None:
            self.u0 = np.minimum(np.asarray(u, dtype=np.float64).ravel(), OSQP_INFTY)
        if np.any(self.l0 > self.u0):
            raise ValueError("lower bound above upper bound")
        if rescale:
            # iterates are stored scaled: carry them through the change of scaling
            x_un, z_un, y_un = self.D * self.x, self.z / self.E, self.E * self.y / self.c
            self._scale()
            self.x, self.z, self.y = x_un / self.D, self.E * z_un, self.c * y_un / self.E
            self._set_rho_vec()
            self._factor()
        else:
            self.q = self.c * self.D * self.q0
            self.l = self._scale_bound(self.l0)
            self.u = self._scale_bound(self.u0)
            if self._set_rho_vec():
                self._factor()
        return 0

    def warm_start(self, x=None, y=None):
        if x is not None:
            self.x = np.asarray(x, dtype=np.float64).ravel() / self.D
            self.z = self.A @ self.x
        if y is not None:
            self.y = self.c * np.asarray(y, dtype=np.float64).ravel() / self.E
        return 0

    # ------------------------------------------------------------------ scaling
    def _scale_bound(self, b):
        fin = np.abs(b) < OSQP_INFTY
        return np.where(fin, self.E * b, b)

    def _scale(self):
        n, m = self.n, self.m
        P, A, q = self.P0.copy(), self.A0.copy(), self.q0.copy()
        D, E, c = np.ones(n), np.ones(m), 1.0
        for _ in range(int(self.settings["scaling"])):
            # infinity norms of the columns of the KKT matrix [[P, A'], [A, 0]]
            absP, absA = abs(P), abs(A)
            col_P = np.asarray(absP.max(axis=0).todense()).ravel() if P.nnz else np.zeros(n)
            col_A = np.asarray(absA.max(axis=0).todense()).ravel() if A.nnz else np.zeros(n)
            row_A = np.asarray(absA.max(axis=1).todense()).ravel() if A.nnz else np.zeros(m)
            d = 1.0 / np.sqrt(_limit_scaling(np.maximum(col_P, col_A)))
            e = 1.0 / np.sqrt(_limit_scaling(row_A))
            Dm, Em = sp.diags(d), sp.diags(e)
            P = (Dm @ P @ Dm).tocsc()
            A = (Em @ A @ Dm).tocsc()
            q = d * q
            D, E = D * d, E * e
            # cost scaling
            col_P = np.asarray(abs(P).max(axis=0).todense()).ravel() if P.nnz else np.zeros(n)
            mean_P = _limit_scaling(np.array([col_P.mean()]))[0]
            norm_q = _limit_scaling(np.array([np.abs(q).max() if n else 0.0]))[0]
            ci = 1.0 / max(mean_P, norm_q)
            P, q, c = P * ci, q * ci, c * ci
        self.P, self.A, self.q = P.tocsc(), A.tocsc(), q
        self.D, self.E, self.c = D, E, c
        self.l = self._scale_bound(self.l0)
        self.u = self._scale_bound(self.u0)

    # ------------------------------------------------------------------ linear system
    def _set_rho_vec(self):
        lo_inf = self.l0 <= -OSQP_INFTY * MIN_SCALING
        up_inf = self.u0 >= OSQP_INFTY * MIN_SCALING
        kind = np.where(lo_inf & up_inf, -1, np.where(np.abs(self.u0 - self.l0) < RHO_TOL, 1, 0))
        changed = not hasattr(self, "_kind") or np.any(kind != self._kind)
        self._kind = kind
        self.rho_vec = np.where(kind == -1, RHO_MIN,
                                np.where(kind == 1, RHO_EQ_OVER_RHO_INEQ * self.rho, self.rho))
        self.rho_vec = np.clip(self.rho_vec, RHO_MIN, RHO_MAX * RHO_EQ_OVER_RHO_INEQ)
        return changed

    def _factor(self):
        sigma = self.settings["sigma"]
        K = sp.bmat([[self.P + sigma * sp.identity(self.n), self.A.T],
                     [self.A, -sp.diags(1.0 / self.rho_vec)]], format="csc")
        # quasi-definite: any sparse LU is a valid stand-in for QDLDL's LDL'
        self._solve = spla.factorized(K)

    # ------------------------------------------------------------------ residuals (unscaled)
    def _residuals(self):
        Dinv, Einv, c = 1.0 / self.D, 1.0 / self.E, self.c
        Ax = self.A @ self.x
        Px = self.P @ self.x
        Aty = self.A.T @ self.y
        pri = np.abs(Einv * (Ax - self.z)).max() if self.m else 0.0
        dua = np.abs(Dinv * (Px + self.q + Aty)).max() / c
        n_Ax = np.abs(Einv * Ax).max() if self.m else 0.0
        n_z = np.abs(Einv * self.z).max() if self.m else 0.0
        n_Px = np.abs(Dinv * Px).max() / c
        n_Aty = np.abs(Dinv * Aty).max() / c
        n_q = np.abs(Dinv * self.q).max() / c
        return pri, dua, max(n_Ax, n_z), max(n_Px, n_Aty, n_q)

    def _new_rho(self, pri, dua, n_pri, n_dua):
        pri_n = pri / (n_pri + 1e-10)
        dua_n = dua / (n_dua + 1e-10)
        return float(np.clip(self.rho * np.sqrt(pri_n / (dua_n + 1e-10)), RHO_MIN, RHO_MAX))

    # ------------------------------------------------------------------ solve
    def solve(self):
        s = self.settings
        sigma, alpha = s["sigma"], s["alpha"]
        x, z, y = self.x, self.z, self.y
        status, it, rho_updates = "maximum iterations reached", 0, 0
        pri = dua = np.inf
        for it in range(1, int(s["max_iter"]) + 1):
            rhs = np.concatenate([sigma * x - self.q, z - y / self.rho_vec])
            sol = self._solve(rhs)
            xt, nu = sol[:self.n], sol[self.n:]
            zt = z + (nu - y) / self.rho_vec
            x = alpha * xt + (1.0 - alpha) * x
            zr = alpha * zt + (1.0 - alpha) * z
            z_new = np.clip(zr + y / self.rho_vec, self.l, self.u)
            y = y + self.rho_vec * (zr - z_new)
            z = z_new
            self.x, self.z, self.y = x, z, y
            check = s["check_termination"] and it % int(s["check_termination"]) == 0
            adapt = s["adaptive_rho"] and s["adaptive_rho_interval"] and it % int(s["adaptive_rho_interval"]) == 0
            if check or adapt:
                pri, dua, n_pri, n_dua = self._residuals()
                if check:
                    if pri <= s["eps_abs"] + s["eps_rel"] * n_pri and dua <= s["eps_abs"] + s["eps_rel"] * n_dua:
                        status = "solved"
                        break
                if adapt:
                    rho_new = self._new_rho(pri, dua, n_pri, n_dua)
                    tol = s["adaptive_rho_tolerance"]
                    if rho_new > self.rho * tol or rho_new < self.rho / tol:
                        self.rho = rho_new
                        self._set_rho_vec()
                        self._factor()
                        rho_updates += 1
        res = _Result()
        res.x = self.D * x
        res.y = self.E * y / self.c
        res.z = z / self.E
        info = _Info()
        info.iter, info.status, info.rho_updates = it, status, rho_updates
        info.status_val = 1 if status == "solved" else 2
        info.pri_res, info.dua_res = pri, dua
        info.obj_val = float(0.5 * res.x @ (self.P0 @ res.x) + self.q0 @ res.x)
        info.rho_estimate = self.rho
        res.info = info
        return res
