"""Numpy model of the algorithm the CUDA engine runs (one instance at a time) -- TEST INFRASTRUCTURE.

It exists so the engine's mathematics (condensing in impulse space, Woodbury-form ADMM, active-set
polish with KKT guard) can be checked against the oracle on CPU, and so a GPU mismatch can be
bisected phase by phase.  The product never imports it (there is no CPU fallback).  Symbols follow
DESIGN.md section 3; reference lines are those of mpc-tsid_b200/csrc/mpcqp_kernels.cu.
"""
import numpy as np


class ModelParams:
    def __init__(self, n_steps=16, dt=0.02, mass=2.50000279, mu=0.9, fz_max=25.0, gravity=9.81,
                 w_force=1e-5, rho=5e-5, sigma=1e-6, alpha=1.6, check_every=5, min_iter=10,
                 max_iter=1000, max_polish=8, max_as=0, stable_need=1, bail_rule=0):
        self.max_as = max_as
        self.stable_need, self.bail_rule = stable_need, bail_rule
        self.N, self.dt, self.mass, self.mu, self.fz_max, self.gravity = n_steps, dt, mass, mu, fz_max, gravity
        self.gI = np.array([[3.09249e-2, -8.00101e-7, 1.865287e-5],
                            [-8.00101e-7, 5.106100e-2, 1.245813e-4],
                            [1.865287e-5, 1.245813e-4, 6.939757e-2]])
        w = np.zeros(12)
        w[0:3] = [0.1, 0.1, 1.0]
        w[3:6] = 0.11
        w[6:9] = 2.0 * np.sqrt(w[0:3])
        w[9:12] = 0.05 * np.sqrt(w[3:6])
        self.w_state, self.w_force = w, w_force
        self.footholds = np.array([[0.19, 0.19, -0.19, -0.19], [0.15005, -0.15005, 0.15005, -0.15005], [0, 0, 0, 0.0]])
        self.rho, self.sigma, self.alpha = rho, sigma, alpha
        self.check_every, self.min_iter, self.max_iter, self.max_polish = check_every, min_iter, max_iter, max_polish
        self.Minv = self._minv()

    def gram(self):
        """M_c[k,l] = dt^2 Qp_c C2[k,l] + Qv_c C0[k,l]; returns (6, N, N)."""
        N, dt = self.N, self.dt
        k = np.arange(N)
        C0 = np.zeros((N, N))
        C2 = np.zeros((N, N))
        for a in range(N):
            for b in range(N):
                i = np.arange(max(a, b), N)
                C0[a, b] = len(i)
                C2[a, b] = np.sum((i - a) * (i - b))
        return np.stack([dt * dt * self.w_state[c] * C2 + self.w_state[6 + c] * C0 for c in range(6)])

    def _minv(self):
        M = self.gram().astype(np.longdouble)
        out = np.zeros((6, self.N, self.N))
        for c in range(6):
            # Gauss-Jordan in extended precision (numpy has no longdouble LAPACK)
            n = self.N
            aug = np.concatenate([M[c], np.eye(n, dtype=np.longdouble)], axis=1)
            for p in range(n):
                aug[p] = aug[p] / aug[p, p]
                for r in range(n):
                    if r != p:
                        aug[r] = aug[r] - aug[r, p] * aug[p]
            out[c] = np.asarray(aug[:, n:], dtype=np.float64)
        return out


def decode(p, xref, fsteps, first_tick=False):
    """contact (N,4), Bv (N,4,6,3)."""
    N = p.N
    contact = np.zeros((N, 4), dtype=bool)
    feet = np.zeros((N, 4, 3))
    k = 0
    for r in range(fsteps.shape[0]):
        cnt = fsteps[r, 0]
        if cnt == 0 or k >= N:
            break
        cnt = int(cnt)
        for kk in range(k, min(k + cnt, N)):
            for j in range(4):
                x = fsteps[r, 1 + 3 * j]
                contact[kk, j] = not (np.isnan(x) or x == 0.0)
                v = fsteps[r, 1 + 3 * j:4 + 3 * j]
                feet[kk, j] = np.where(np.isnan(v), 0.0, v)
        k += cnt
    Bv = np.zeros((N, 4, 6, 3))
    for k in range(N):
        c, s = np.cos(xref[5, k]), np.sin(xref[5, k])
        R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
        Iinv = np.linalg.inv(R @ p.gI)
        for j in range(4):
            ft = p.footholds[:, j] if first_tick else feet[k, j]
            r = ft - xref[0:3, k]
            skew = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
            Bv[k, j, 0:3] = (p.dt / p.mass) * np.eye(3)
            Bv[k, j, 3:6] = p.dt * Iinv @ skew
    return contact, Bv


def free_response(p, xref):
    """ebar_p, ebar_v (N,6) for states 1..N, gamma_bar (N,6), constant c0."""
    N, dt = p.N, p.dt
    p0, v0 = xref[0:6, 0], xref[6:12, 0]
    gv = np.zeros(6)
    gv[2] = -p.gravity * dt
    ep = np.zeros((N, 6))
    ev = np.zeros((N, 6))
    for i in range(N):
        s = i + 1
        ep[i] = p0 + s * dt * v0 + dt * gv * (s * (s - 1) / 2.0) - xref[0:6, s]
        ev[i] = v0 + s * gv - xref[6:12, s]
    Qp, Qv = p.w_state[0:6], p.w_state[6:12]
    gam = np.zeros((N, 6))
    for k in range(N):
        for i in range(k, N):
            gam[k] += (i - k) * dt * Qp * ep[i] + Qv * ev[i]
    c0 = 0.5 * np.sum(Qp * ep * ep + Qv * ev * ev)
    return ep, ev, gam, c0


class Engine:
    """Per-instance state carried across ticks: forces f (N,4,3) and row multipliers y (N,4,5)."""

    def __init__(self, p: ModelParams):
        self.p = p
        N = p.N
        self.f = np.zeros((N, 4, 3))
        self.y = np.zeros((N, 4, 5))
        mu = p.mu
        self.C = np.array([[1, 0, -mu], [-1, 0, -mu], [0, 1, -mu], [0, -1, -mu], [0, 0, -1.0]])
        self.lo = np.array([-np.inf, -np.inf, -np.inf, -np.inf, -p.fz_max])
        self.hi = np.zeros(5)
        self.stats = {}
        self.sig = np.zeros((N, 4, 3), np.int8)

    # ---- linear algebra in impulse space
    def _W(self, Bv, contact, dinv):
        """W = Minv + blockdiag_k( sum_j Bv_kj diag(dinv_kj) Bv_kj' ); index = 6k + c."""
        p, N = self.p, self.p.N
        W = np.zeros((6 * N, 6 * N))
        for c in range(6):
            W[c::6, c::6] = p.Minv[c]
        for k in range(N):
            T = np.zeros((6, 6))
            for j in range(4):
                if contact[k, j]:
                    T += Bv[k, j] @ (dinv[k, j][:, None] * Bv[k, j].T)
            W[6 * k:6 * k + 6, 6 * k:6 * k + 6] += T
        return W

    def _impulse(self, Bv, contact, t):
        N = self.p.N
        s = np.zeros((N, 6))
        for k in range(N):
            for j in range(4):
                if contact[k, j]:
                    s[k] += Bv[k, j] @ t[k, j]
        return s

    def _back(self, Bv, contact, v):
        N = self.p.N
        out = np.zeros((N, 4, 3))
        for k in range(N):
            for j in range(4):
                if contact[k, j]:
                    out[k, j] = Bv[k, j].T @ v[k]
        return out

    def _hess_apply(self, Bv, contact, f):
        """H f = w f + Bv' M (Bv f)."""
        p, N = self.p, self.p.N
        M = p.gram()
        s = self._impulse(Bv, contact, f)
        ms = np.zeros((N, 6))
        for c in range(6):
            ms[:, c] = M[c] @ s[:, c]
        return p.w_force * f * contact[:, :, None] + self._back(Bv, contact, ms)

    # ---- one tick
    def solve(self, xref, fsteps, first_tick=False, warm=True):
        p, N = self.p, self.p.N
        mu, rho, sigma, alpha = p.mu, p.rho, p.sigma, p.alpha
        contact, Bv = decode(p, xref, fsteps, first_tick)
        ep, ev, gam, c0 = free_response(p, xref)
        g = self._back(Bv, contact, gam)
        cm = contact[:, :, None]
        # warm start: shift by one step (MPC.py:403-406), wrap the first block to the back
        if warm and not first_tick:
            f = np.roll(self.f, -1, axis=0) * cm
            y = np.roll(self.y, -1, axis=0) * cm
        else:
            f = np.zeros((N, 4, 3))
            y = np.zeros((N, 4, 5))
        C = self.C
        z = np.clip(f @ C.T, self.lo, self.hi) * cm
        d = np.array([p.w_force + sigma + 2 * rho, p.w_force + sigma + 2 * rho,
                      p.w_force + sigma + rho * (4 * mu * mu + 1)])
        dinv = np.broadcast_to(1.0 / d, (N, 4, 3)).copy()
        # stage A: primal-dual active-set sweeps from the shifted previous active set
        n_as = 0
        result = None
        status = 0
        if p.max_as > 0:
            sig = (np.roll(self.sig, -1, axis=0) if (warm and not first_tick) else np.zeros((N, 4, 3), np.int8)) * contact[:, :, None]
            seen = []
            prev_changed = None
            for n_as in range(1, p.max_as + 1):
                ok, fp, yp, nsig = self._polish(Bv, contact, g, sig)
                if ok:
                    result, status, self.sig = (fp, yp), 1, sig
                    break
                if any(np.array_equal(nsig, s_) for s_ in seen):
                    break
                changed = int((nsig != sig).any(axis=2).sum())
                if p.bail_rule and prev_changed is not None and changed >= prev_changed:
                    break
                if p.bail_rule >= 2 and changed > p.bail_rule:
                    break
                prev_changed = changed
                seen.append(sig)
                sig = nsig
        if result is not None:
            self.stats = dict(iters=0, polishes=n_as, status=status, n_as=n_as, n_chol=n_as)
            return self._finish(xref, Bv, contact, result)
        W = self._W(Bv, contact, dinv)
        Winv = np.linalg.inv(W)

        def kinv(r):
            t = r * dinv * cm
            s = self._impulse(Bv, contact, t)
            v = (Winv @ s.reshape(-1)).reshape(N, 6)
            return t - dinv * self._back(Bv, contact, v)

        it, n_polish, status = 0, 0, 0
        prev_sig = None
        stable = 0
        while True:
            it += 1
            rhs = (sigma * f - g + (rho * z - y) @ C) * cm
            ft = kinv(rhs)
            zt = ft @ C.T
            f_new = alpha * ft + (1 - alpha) * f
            zr = alpha * zt + (1 - alpha) * z
            z_new = np.clip(zr + y / rho, self.lo, self.hi) * cm
            y = (y + rho * (zr - z_new)) * cm
            f, z = f_new, z_new
            if it >= p.min_iter and it % p.check_every == 0:
                sig = self._signature(f, y, z, contact)
                stable = stable + 1 if (prev_sig is not None and np.array_equal(sig, prev_sig)) else 0
                prev_sig = sig
                if stable >= p.stable_need:
                    n_polish += 1
                    ok, fp, yp, _ = self._polish(Bv, contact, g, sig)
                    stable = 0
                    if ok:
                        result = (fp, yp)
                        status = 1
                        self.sig = sig
                        break
                    if n_polish >= p.max_polish:
                        pass
            if it >= p.max_iter:
                status = 2
                result = (f, y)
                self.sig = self._signature(f, y, z, contact)
                break
        self.stats = dict(iters=it, polishes=n_polish + n_as, status=status, n_as=n_as, n_chol=n_polish + n_as + 1)
        return self._finish(xref, Bv, contact, result)

    def _finish(self, xref, Bv, contact, result):
        p, N = self.p, self.p.N
        fo, yo = result
        self.f, self.y = fo.copy(), yo.copy()
        # states by forward simulation; objective
        s = self._impulse(Bv, contact, fo)
        X = np.zeros((N, 12))
        xk = xref[:, 0].copy()
        gv = np.zeros(12)
        gv[8] = -p.gravity * p.dt
        for k in range(N):
            nx = xk.copy()
            nx[0:6] += p.dt * xk[6:12]
            nx[6:12] += s[k]
            nx += gv
            X[k] = nx
            xk = nx
        err = X - xref[:, 1:].T
        obj = 0.5 * np.sum(p.w_state * err * err) + 0.5 * p.w_force * np.sum(fo * fo)
        x_full = np.concatenate([err.reshape(-1), fo.reshape(-1)])
        return dict(f=fo, y=yo, x=x_full, f_applied=fo[0].reshape(-1).copy(), obj=obj, contact=contact,
                    x_robot=X.T.copy(), **self.stats)

    def _signature(self, f, y, z, contact):
        """Per foot: sx, sy in {-1,0,1}, tz in {0 free, 1 apex, 2 top}; from OSQP's polish rule."""
        p = self.p
        N = p.N
        low = (z - self.lo < -y)
        upp = (self.hi - z < y)
        sig = np.zeros((N, 4, 3), dtype=np.int8)
        sig[:, :, 0] = np.where(upp[:, :, 0], 1, 0) - np.where(upp[:, :, 1], 1, 0)
        sig[:, :, 1] = np.where(upp[:, :, 2], 1, 0) - np.where(upp[:, :, 3], 1, 0)
        apex = upp[:, :, 4] | (upp[:, :, 0] & upp[:, :, 1]) | (upp[:, :, 2] & upp[:, :, 3])
        sig[:, :, 2] = np.where(apex, 1, np.where(low[:, :, 4], 2, 0))
        return sig * contact[:, :, None]

    def _polish(self, Bv, contact, g, sig):
        """Equality-constrained solve on the guessed active set, in Woodbury form, then KKT guard."""
        p, N, mu = self.p, self.p.N, self.p.mu
        w = p.w_force
        # per foot: f = pfix + Z q, Z columns (ex if sx==0), (ey if sy==0), (sx mu, sy mu, 1) if fz free
        Z = np.zeros((N, 4, 3, 3))
        pf = np.zeros((N, 4, 3))
        zz = np.zeros((N, 4, 3))          # Z'Z diagonal
        for k in range(N):
            for j in range(4):
                if not contact[k, j]:
                    continue
                sx, sy, tz = sig[k, j]
                if tz == 1:
                    continue
                if sx == 0:
                    Z[k, j, 0, 0] = 1.0
                if sy == 0:
                    Z[k, j, 1, 1] = 1.0
                if tz == 0:
                    Z[k, j, :, 2] = [sx * mu, sy * mu, 1.0]
                else:
                    pf[k, j] = [sx * mu * p.fz_max, sy * mu * p.fz_max, p.fz_max]
                zz[k, j] = np.sum(Z[k, j] * Z[k, j], axis=0)
        act = zz > 0
        dinv = np.where(act, 1.0 / np.where(act, w * zz, 1.0), 0.0)
        BZ = np.einsum('kjab,kjbc->kjac', Bv, Z)
        W = self._W(BZ, contact, dinv)
        L = np.linalg.cholesky(W)
        # gradient at pf: H pf + g
        grad0 = self._hess_apply(Bv, contact, pf) + g
        r = -np.einsum('kjba,kjb->kja', Z, grad0)
        t = r * dinv
        s = self._impulse(BZ, contact, t)
        v = np.linalg.solve(L.T, np.linalg.solve(L, s.reshape(-1))).reshape(N, 6)
        q = t - dinv * self._back(BZ, contact, v)
        # one step of iterative refinement on the reduced system
        fq = pf + np.einsum('kjab,kjb->kja', Z, q)
        grad = self._hess_apply(Bv, contact, fq) + g
        r2 = -np.einsum('kjba,kjb->kja', Z, grad)
        t2 = r2 * dinv
        s2 = self._impulse(BZ, contact, t2)
        v2 = np.linalg.solve(L.T, np.linalg.solve(L, s2.reshape(-1))).reshape(N, 6)
        q = q + t2 - dinv * self._back(BZ, contact, v2)
        fq = (pf + np.einsum('kjab,kjb->kja', Z, q)) * contact[:, :, None]
        grad = self._hess_apply(Bv, contact, fq) + g
        # multipliers + KKT guard
        y = np.zeros((N, 4, 5))
        ok = True
        ftol, ytol = 1e-9, 1e-12
        for k in range(N):
            for j in range(4):
                if not contact[k, j]:
                    continue
                sx, sy, tz = sig[k, j]
                gx, gy, gz = grad[k, j]
                if tz == 1:
                    # apex: need y >= 0 with C'y = -grad
                    qx, qy, qz = -gx, -gy, -gz
                    y[k, j, 0], y[k, j, 1] = max(qx, 0), max(-qx, 0)
                    y[k, j, 2], y[k, j, 3] = max(qy, 0), max(-qy, 0)
                    y4 = -qz - mu * (abs(qx) + abs(qy))
                    y[k, j, 4] = y4
                    if y4 < -ytol:
                        ok = False
                    continue
                yx = -sx * gx if sx != 0 else 0.0
                yy = -sy * gy if sy != 0 else 0.0
                if sx != 0:
                    y[k, j, 0 if sx > 0 else 1] = yx
                if sy != 0:
                    y[k, j, 2 if sy > 0 else 3] = yy
                y4 = gz - mu * (yx + yy)
                if tz == 2:
                    y[k, j, 4] = y4
                    if y4 > ytol:
                        ok = False
                if yx < -ytol or yy < -ytol:
                    ok = False
                # primal feasibility of the rows assumed inactive
                cf = self.C @ fq[k, j]
                if np.any(cf > self.hi + ftol) or cf[4] < self.lo[4] - ftol:
                    ok = False
        return ok, fq, y, self._next_signature(sig, fq, y, grad, contact)

    def _next_signature(self, sig, f, y, grad, contact):
        """One primal-dual active-set update: release rows whose multiplier has the wrong sign, add
        rows the polished point violates."""
        p, N, mu = self.p, self.p.N, self.p.mu
        out = sig.copy()
        tol = 1e-9
        for k in range(N):
            for j in range(4):
                if not contact[k, j]:
                    continue
                sx, sy, tz = sig[k, j]
                fx, fy, fz = f[k, j]
                if tz == 1:
                    if y[k, j, 4] < 0:
                        # leave the apex along the steepest feasible descent direction
                        qx, qy = -grad[k, j, 0], -grad[k, j, 1]
                        out[k, j] = [0, 0, 0]
                        if -grad[k, j, 2] + mu * abs(qx) <= 0:
                            out[k, j, 0] = 0
                    continue
                nsx, nsy, ntz = sx, sy, tz
                if sx != 0 and y[k, j, 0 if sx > 0 else 1] < 0:
                    nsx = 0
                elif sx == 0:
                    if fx - mu * fz > tol:
                        nsx = 1
                    elif -fx - mu * fz > tol:
                        nsx = -1
                if sy != 0 and y[k, j, 2 if sy > 0 else 3] < 0:
                    nsy = 0
                elif sy == 0:
                    if fy - mu * fz > tol:
                        nsy = 1
                    elif -fy - mu * fz > tol:
                        nsy = -1
                if tz == 2 and y[k, j, 4] > 0:
                    ntz = 0
                elif tz == 0:
                    if fz > p.fz_max + tol:
                        ntz = 2
                    elif fz < -tol:
                        ntz = 1
                out[k, j] = [nsx, nsy, ntz]
        return out


# --------------------------------------------------------------------------------------------------
# Stage-wise (Riccati) form of the same equality-constrained solve: model of mpcqp_riccati.cuh.
# State x = [p (6); v (6)], x+ = A x + [0; u + g], u = ubar + w, w = Bq q, stage input cost 1/2 q'Rq.
# Backward:  L L' = Pvv+,  C = L' E L,  G = I + C,  Gamma = L^-T (I - G^-1) L^-1  (= (I + E Pvv)^-1 E),
#            Pt = P+ - P+[:,v] Gamma P+[v,:],  P_k = Q + A' Pt A,  p_k = -Q xref_k + A'(Pt[:,v] beta + pt).
# Forward:   z = A x + b, lam = P+[v,:] z + p+[v], w = -Gamma lam, x+ = z + [0; w], lamhat = lam + Pvv+ w.
# Per foot:  q = -R^-1 Bq' lamhat, f = pf + Z q, grad = w_f f + Bv' lamhat.
# --------------------------------------------------------------------------------------------------
def riccati_faces(p, contact, sig):
    N, mu = p.N, p.mu
    Z = np.zeros((N, 4, 3, 3))
    pf = np.zeros((N, 4, 3))
    zz = np.zeros((N, 4, 3))
    for k in range(N):
        for j in range(4):
            if not contact[k, j]:
                continue
            sx, sy, tz = sig[k, j]
            if tz == 1:
                continue
            if sx == 0:
                Z[k, j, 0, 0] = 1.0
            if sy == 0:
                Z[k, j, 1, 1] = 1.0
            if tz == 0:
                Z[k, j, :, 2] = [sx * mu, sy * mu, 1.0]
            else:
                pf[k, j] = [sx * mu * p.fz_max, sy * mu * p.fz_max, p.fz_max]
            zz[k, j] = np.sum(Z[k, j] * Z[k, j], axis=0)
    act = zz > 0
    dinv = np.where(act, 1.0 / np.where(act, p.w_force * zz, 1.0), 0.0)
    return Z, pf, dinv


def riccati_solve(p, xref, Bv, contact, sig, keep=False):
    """-> f (N,4,3), grad (N,4,3) = H f + g of the condensed problem, X (N,12) states 1..N.
    keep=True also returns the per-stage blocks a one-direction update on this factorisation needs (riccati_one_direction)."""
    N, dt = p.N, p.dt
    Z, pf, dinv = riccati_faces(p, contact, sig)
    Qp, Qv = p.w_state[0:6], p.w_state[6:12]
    gv = np.zeros(6)
    gv[2] = -p.gravity * dt
    E = np.zeros((N, 6, 6))
    beta = np.zeros((N, 6))
    for k in range(N):
        beta[k] = gv
        for j in range(4):
            if contact[k, j]:
                Bq = Bv[k, j] @ Z[k, j]
                E[k] += Bq @ (dinv[k, j][:, None] * Bq.T)
                beta[k] += Bv[k, j] @ pf[k, j]
    I6 = np.eye(6)
    # terminal cost-to-go at state N
    Ppp, Ppv, Pvv = np.diag(Qp), np.zeros((6, 6)), np.diag(Qv)
    pp, pv = -Qp * xref[0:6, N], -Qv * xref[6:12, N]
    store = [None] * N
    for k in range(N - 1, -1, -1):
        L = np.linalg.cholesky(Pvv)
        Linv = np.linalg.inv(L)
        C = L.T @ E[k] @ L
        G = I6 + C
        M = np.linalg.cholesky(G)
        Minv = np.linalg.inv(M)
        X = L @ Minv.T                     # Ptvv = X X'
        Y = Ppv @ Linv.T                   # Ppv L^-T
        Y2 = Y @ Minv.T
        Gam = Linv.T @ (I6 - Minv.T @ Minv) @ Linv
        Ptvv = X @ X.T
        Ptpv = Y2 @ X.T
        store[k] = (Ppv.copy(), Pvv.copy(), pv.copy(), Gam, Ptpv, Ptvv)
        Ptpp = Ppp - Y @ Y.T + Y2 @ Y2.T
        gp = Gam @ pv
        ptp = pp - Ppv @ gp
        ptv = pv - Pvv @ gp
        if k == 0:
            break
        hp = Ptpv @ beta[k] + ptp          # [Pt[:,v] beta + pt]
        hv = Ptvv @ beta[k] + ptv
        nPpp = Ptpp + np.diag(Qp)
        nPpv = dt * Ptpp + Ptpv
        nPvv = dt * dt * Ptpp + dt * (Ptpv + Ptpv.T) + Ptvv + np.diag(Qv)
        npp = hp - Qp * xref[0:6, k]
        npv = dt * hp + hv - Qv * xref[6:12, k]
        Ppp, Ppv, Pvv, pp, pv = nPpp, nPpv, 0.5 * (nPvv + nPvv.T), npp, npv
    xp, xv = xref[0:6, 0].copy(), xref[6:12, 0].copy()
    Xs = np.zeros((N, 12))
    lam = np.zeros((N, 6))
    for k in range(N):
        Ppv_, Pvv_, pv_, Gam = store[k][:4]
        zp = xp + dt * xv
        zv = xv + beta[k]
        l0 = Ppv_.T @ zp + Pvv_ @ zv + pv_
        w = -Gam @ l0
        xp, xv = zp, zv + w
        lam[k] = l0 + Pvv_ @ w
        Xs[k, 0:6], Xs[k, 6:12] = xp, xv
    f = np.zeros((N, 4, 3))
    grad = np.zeros((N, 4, 3))
    for k in range(N):
        for j in range(4):
            if contact[k, j]:
                Bq = Bv[k, j] @ Z[k, j]
                q = -dinv[k, j] * (Bq.T @ lam[k])
                f[k, j] = pf[k, j] + Z[k, j] @ q
                grad[k, j] = p.w_force * f[k, j] + Bv[k, j].T @ lam[k]
    if keep:
        return f, grad, Xs, dict(store=store, Z=Z, dinv=dinv)
    return f, grad, Xs


def riccati_one_direction(p, Bv, contact, kept, ks, js, e):
    """Response of the equality-constrained optimum to moving foot-step (ks, js) by a unit along direction e (3-vector in
    force space), every face and every 6x6 factorisation of the kept solve unchanged: the impulse offset of stage ks grows by
    Bv e and nothing else does, so only the *vector* half of the recursion runs -- backward from ks to 1 (no matrix
    recursion, no factorisation), forward over the whole horizon.  -> df (N,4,3) on the kept faces (the e itself not
    included), dgrad (N,4,3), dX (N,12).  Round-2 plan of DESIGN.md 9.1: a second sweep at a quarter of the cost."""
    N, dt = p.N, p.dt
    store, Z, dinv = kept["store"], kept["Z"], kept["dinv"]
    dbeta = Bv[ks, js] @ e
    # backward, vector only: above ks nothing changes
    dpv_in = [np.zeros(6) for _ in range(N)]                  # change of the p-vector (v half) each stage's forward step reads
    dpp_in = [np.zeros(6) for _ in range(N)]
    pp, pv = np.zeros(6), np.zeros(6)
    for k in range(ks, 0, -1):
        Ppv, Pvv, _, Gam, Ptpv, Ptvv = store[k]
        gp = Gam @ pv
        ptp, ptv = pp - Ppv @ gp, pv - Pvv @ gp
        b = dbeta if k == ks else np.zeros(6)
        hp, hv = Ptpv @ b + ptp, Ptvv @ b + ptv
        pp, pv = hp, dt * hp + hv
        dpp_in[k - 1], dpv_in[k - 1] = pp, pv
    # forward from a fixed initial state
    xp, xv = np.zeros(6), np.zeros(6)
    dX = np.zeros((N, 12))
    dlam = np.zeros((N, 6))
    for k in range(N):
        Ppv, Pvv, _, Gam = store[k][:4]
        zp = xp + dt * xv
        zv = xv + (dbeta if k == ks else 0.0)
        l0 = Ppv.T @ zp + Pvv @ zv + dpv_in[k]
        w = -Gam @ l0
        xp, xv = zp, zv + w
        dlam[k] = l0 + Pvv @ w
        dX[k, 0:6], dX[k, 6:12] = xp, xv
    df = np.zeros((N, 4, 3))
    dgrad = np.zeros((N, 4, 3))
    for k in range(N):
        for j in range(4):
            if contact[k, j]:
                Bq = Bv[k, j] @ Z[k, j]
                df[k, j] = Z[k, j] @ (-dinv[k, j] * (Bq.T @ dlam[k]))
                dgrad[k, j] = p.w_force * df[k, j] + Bv[k, j].T @ dlam[k]
    return df, dgrad, dX


def riccati_change_one_row(p, Bv, contact, kept, f, grad, X, ks, js, release=None, activate=None):
    """The optimum on a face that differs from the kept one in ONE row of foot-step (ks, js), from the kept solve plus one
    riccati_one_direction.  release = d (3,): a force direction that becomes free (its component orthogonal to the kept face is
    used; the step length follows from stationarity along it).  activate = (c (3,), b): a row c'f = b that becomes active (the
    step runs along the component of c inside the kept face; its length follows from feasibility)."""
    Zk = kept["Z"][ks, js]
    zz = np.sum(Zk * Zk, axis=0)
    proj = lambda v: Zk @ (np.where(zz > 0, (Zk.T @ v) / np.where(zz > 0, zz, 1.0), 0.0))      # Z'Z is diagonal
    if release is not None:
        e = release - proj(release)
        df, dgrad, dX = riccati_one_direction(p, Bv, contact, kept, ks, js, e)
        g0 = e @ grad[ks, js]
        g1 = e @ (dgrad[ks, js] + p.w_force * e)             # grad(step)[ks, js] = grad + step (dgrad + w_f e)
        step = -g0 / g1
    else:
        c, b = activate
        e = proj(c)
        df, dgrad, dX = riccati_one_direction(p, Bv, contact, kept, ks, js, e)
        step = (b - c @ f[ks, js]) / (c @ (df[ks, js] + e))
    f2, g2 = f + step * df, grad + step * dgrad
    f2[ks, js] += step * e
    g2[ks, js] += step * p.w_force * e
    return f2, g2, X + step * dX


def riccati_stage_ldl(Ppp, Ppv, Pvv, pp, pv, E):
    """One backward stage in the square-root-free form the kernel uses (mpcqp_riccati.cuh: ldl6_regs + the row products):
    Pvv = U D U', H = D^-1 + U'EU = W Delta W'; for every row rho of [Ppv; Pvv; pv'] with yt = rho U^-T and
    e = ((yt D^-1) W^-T Delta^-1) W^-1:  rho (I - Gamma Pvv) = e U',  rho Gamma = ((yt - e) D^-1) U^-1.
    Returns (Pt[:, v] rows (13 x 6), [Ppv; Pvv; pv'] Gamma rows (13 x 6))."""
    def ldl(a):
        a = a.copy()
        n = a.shape[0]
        U = np.eye(n)
        d = np.zeros(n)
        for j in range(n):
            d[j] = a[j, j]
            U[j + 1:, j] = a[j + 1:, j] / d[j]
            a[j + 1:, j + 1:] -= np.outer(U[j + 1:, j], a[j + 1:, j])
        return U, d
    U, d = ldl(Pvv)
    H = np.diag(1.0 / d) + U.T @ E @ U
    W, dl = ldl(H)
    rows = np.vstack([Ppv, Pvv, pv[None, :]])
    yt = rows @ np.linalg.inv(U).T
    e = ((yt / d) @ np.linalg.inv(W).T / dl) @ np.linalg.inv(W)
    return e @ U.T, ((yt - e) / d) @ np.linalg.inv(U)
