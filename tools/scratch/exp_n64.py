"""CPU experiment: why do the N = 64 cold-start instances defeat stage A, and what converges on them.
Dense condensed form (H, g) of one robot + the kernel's guard / signature rules in numpy."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from scenario import Scenario
import kernel_model as km

MU, FZ = 0.9, 25.0


def condensed(p, xref, fsteps, first_tick):
    """H (n x n), g (n), idx -> (k, j) of stance foot-steps; f ordered foot-step major, 3 each."""
    N = p.N
    contact, Bv = km.decode(p, xref, fsteps, first_tick)
    ep, ev, gam, c0 = km.free_response(p, xref)
    M = p.gram()
    idx = [(k, j) for k in range(N) for j in range(4) if contact[k, j]]
    n = 3 * len(idx)
    Bm = np.zeros((6 * N, n))          # impulse index 6k + c
    for i, (k, j) in enumerate(idx):
        Bm[6 * k:6 * k + 6, 3 * i:3 * i + 3] = Bv[k, j]
    Mf = np.zeros((6 * N, 6 * N))
    for c in range(6):
        Mf[c::6, c::6] = M[c]
    H = p.w_force * np.eye(n) + Bm.T @ Mf @ Bm
    g = Bm.T @ gam.reshape(-1)
    return H, g, idx, c0


def sig_unpack(s):
    tz = s // 9; r = s - 9 * tz; return r % 3 - 1, r // 3 - 1, tz


def sig_pack(sx, sy, tz):
    return (sx + 1) + 3 * (sy + 1) + 9 * tz


def face(sig):
    """Z (3 x 3 cols, zero cols absent), pf."""
    sx, sy, tz = sig_unpack(sig)
    Z = np.zeros((3, 3)); pf = np.zeros(3)
    if tz == 1:
        return Z, pf
    if sx == 0: Z[0, 0] = 1
    if sy == 0: Z[1, 1] = 1
    if tz == 0: Z[:, 2] = [sx * MU, sy * MU, 1.0]
    else: pf = np.array([sx * MU * FZ, sy * MU * FZ, FZ])
    return Z, pf


def face_solve(H, g, sigs):
    n = H.shape[0]
    cols = []; pf = np.zeros(n)
    for i, s in enumerate(sigs):
        Z, p0 = face(s)
        pf[3 * i:3 * i + 3] = p0
        for c in range(3):
            if Z[:, c].any():
                v = np.zeros(n); v[3 * i:3 * i + 3] = Z[:, c]; cols.append(v)
    if cols:
        Zm = np.array(cols).T
        q = np.linalg.solve(Zm.T @ H @ Zm, -Zm.T @ (H @ pf + g))
        f = pf + Zm @ q
    else:
        f = pf
    return f, H @ f + g


def guard(sig, f, grad, ytol=1e-12, ftol=1e-9):
    """kkt_guard of mpcqp_foot.cuh: ok, nsig, y."""
    mu = MU
    sx, sy, tz = sig_unpack(sig)
    ok = True; nsx, nsy, ntz = sx, sy, tz
    y = np.zeros(5)
    if tz == 1:
        qx, qy, qz = -grad
        y[0], y[1], y[2], y[3] = max(qx, 0), max(-qx, 0), max(qy, 0), max(-qy, 0)
        y[4] = -qz - mu * (abs(qx) + abs(qy))
        if y[4] < -ytol:
            ok = False
            nsx = 1 if qx > ytol else (-1 if qx < -ytol else 0)
            nsy = 1 if qy > ytol else (-1 if qy < -ytol else 0)
            ntz = 0
    else:
        yx = -sx * grad[0] if sx != 0 else 0.0
        yy = -sy * grad[1] if sy != 0 else 0.0
        if sx > 0: y[0] = yx
        elif sx < 0: y[1] = yx
        if sy > 0: y[2] = yy
        elif sy < 0: y[3] = yy
        y4 = grad[2] - mu * (yx + yy)
        if tz == 2:
            y[4] = y4
            if y4 > ytol: ok = False; ntz = 0
        if sx != 0 and yx < -ytol: ok = False; nsx = 0
        if sy != 0 and yy < -ytol: ok = False; nsy = 0
        if sx == 0:
            if f[0] - mu * f[2] > ftol: ok = False; nsx = 1
            elif -f[0] - mu * f[2] > ftol: ok = False; nsx = -1
        if sy == 0:
            if f[1] - mu * f[2] > ftol: ok = False; nsy = 1
            elif -f[1] - mu * f[2] > ftol: ok = False; nsy = -1
        if tz == 0:
            if f[2] > FZ + ftol: ok = False; ntz = 2
            elif f[2] < -ftol: ok = False; ntz = 1
    return ok, sig_pack(nsx, nsy, ntz), y


def sweep(H, g, sigs, ytol=1e-12, ftol=1e-9):
    f, grad = face_solve(H, g, sigs)
    oks, ns = [], []
    for i, s in enumerate(sigs):
        ok, n_, _ = guard(s, f[3 * i:3 * i + 3], grad[3 * i:3 * i + 3], ytol, ftol)
        oks.append(ok); ns.append(n_)
    return f, grad, np.array(oks), np.array(ns)


def pdas(H, g, sigs, max_sweeps=60, ytol=1e-12, verbose=False):
    """stage A as in riccati_kernel: full update unless seen; then single changes in index order."""
    sigs = np.array(sigs)
    seen = set()
    careful = False
    for s in range(max_sweeps):
        f, grad, oks, ns = sweep(H, g, sigs, ytol)
        if oks.all():
            return True, s + 1, sigs, f
        if verbose:
            print("   sweep %d: %d feet fail, obj %.9f" % (s, (~oks).sum(), 0.5 * f @ H @ f + g @ f))
        seen.add(sigs.tobytes())
        if not careful:
            if ns.tobytes() in seen: careful = True
            else: sigs = ns; continue
        found = False
        for t in np.flatnonzero(ns != sigs):
            cand = sigs.copy(); cand[t] = ns[t]
            if cand.tobytes() not in seen:
                sigs = cand; found = True; break
        if not found:
            return False, s + 1, sigs, f
    return False, max_sweeps, sigs, f


def instances(N=64, B=8, seed=3, vlo=0.7, vhi=1.0, gaits=("trot",), tick=0):
    v_ref = np.zeros((B, 6)); v_ref[:, 0] = np.linspace(vlo, vhi, B)
    sc = Scenario(B, n_steps=N, gaits=list(gaits), seed=seed, v_ref=v_ref)
    xref, fsteps = sc.inputs()
    return xref, fsteps


if __name__ == "__main__":
    N = 64
    p = km.ModelParams(n_steps=N)
    xref, fsteps = instances(N)
    for b in range(8):
        H, g, idx, c0 = condensed(p, xref[b], fsteps[b], True)
        ev = np.linalg.eigvalsh(H)
        n = len(idx)
        for ytol in (1e-12, 1e-10, 1e-9):
            ok, ns, sg, f = pdas(H, g, np.full(n, 4), max_sweeps=80, ytol=ytol)
            print("robot %d n_a %d cond %.2e ytol %.0e: pdas ok %s sweeps %d" % (b, 3 * n, ev[-1] / ev[0], ytol, ok, ns))


# ---------------------------------------------------------------------------------------------------
# Stage C candidate: monotone projected active-set iteration from a feasible point
# ---------------------------------------------------------------------------------------------------
def clip_foot(f):
    fz = min(max(f[2], 0.0), FZ)
    return np.array([min(max(f[0], -MU * fz), MU * fz), min(max(f[1], -MU * fz), MU * fz), fz])


def clip_all(f):
    out = f.copy()
    for i in range(len(f) // 3):
        out[3 * i:3 * i + 3] = clip_foot(f[3 * i:3 * i + 3])
    return out


def activity(f, tol=1e-10):
    """signature of the rows that hold with equality at a feasible f"""
    n = len(f) // 3
    sg = np.zeros(n, dtype=np.int64)
    for i in range(n):
        fx, fy, fz = f[3 * i:3 * i + 3]
        if fz <= tol:
            sg[i] = sig_pack(0, 0, 1); continue
        sx = 1 if fx - MU * fz >= -tol else (-1 if -fx - MU * fz >= -tol else 0)
        sy = 1 if fy - MU * fz >= -tol else (-1 if -fy - MU * fz >= -tol else 0)
        tz = 2 if fz >= FZ - tol else 0
        sg[i] = sig_pack(sx, sy, tz)
    return sg


def phi(H, g, f):
    return 0.5 * f @ H @ f + g @ f


def ratio_alpha(f, d, tol=0.0):
    """largest alpha in [0, 1] with f + alpha d feasible (f feasible)."""
    a = 1.0
    for i in range(len(f) // 3):
        x, dx = f[3 * i:3 * i + 3], d[3 * i:3 * i + 3]
        rows = [(x[0] - MU * x[2], dx[0] - MU * dx[2]), (-x[0] - MU * x[2], -dx[0] - MU * dx[2]),
                (x[1] - MU * x[2], dx[1] - MU * dx[2]), (-x[1] - MU * x[2], -dx[1] - MU * dx[2]),
                (-x[2], -dx[2]), (x[2] - FZ, dx[2])]
        for c, dc in rows:
            if dc > 1e-300 and c + a * dc > tol:
                a = min(a, max((tol - c) / dc, 0.0))
    return a


def stage_c(H, g, f0, max_sweeps=200, ytol=1e-12, ftol=1e-9, verbose=False, release_all=True):
    f = clip_all(f0)
    sigs = activity(f)
    ph = phi(H, g, f)
    nev = 0
    for s in range(max_sweeps):
        fh, grad, oks, ns = sweep(H, g, sigs, ytol, ftol)
        # primal feasibility of the face minimiser
        feas = np.abs(clip_all(fh) - fh).max() <= ftol
        if feas:
            if oks.all():
                return True, s + 1, nev, sigs, fh
            f, ph = fh, phi(H, g, fh)
            # release rows with the wrong multiplier sign (guard proposals only release here)
            if release_all:
                sigs = ns.copy()
            else:
                t = np.flatnonzero(ns != sigs)[0]
                sigs = sigs.copy(); sigs[t] = ns[t]
            if verbose: print("   C %d: at face minimiser, phi %.12f, release %d" % (s, ph, (ns != sigs).sum()))
            continue
        d = fh - f
        a = 1.0
        acc = False
        amax = ratio_alpha(f, d)
        for _ in range(12):
            ft = clip_all(f + a * d); nev += 1
            pt = phi(H, g, ft)
            if pt < ph - 1e-14 * abs(ph):
                acc = True; break
            a *= 0.5
            if a <= amax: break
        if not acc:
            a = amax
            ft = f + a * d; pt = phi(H, g, ft)
        if verbose: print("   C %d: path step alpha %.3g (amax %.3g) phi %.12f -> %.12f" % (s, a, amax, ph, pt))
        if a == 0.0 and not release_all:
            return False, s + 1, nev, sigs, f
        f, ph = ft, pt
        sigs = activity(f)
    return False, max_sweeps, nev, sigs, f


def run_c():
    N = 64
    p = km.ModelParams(n_steps=N)
    xref, fsteps = instances(N)
    for b in range(8):
        H, g, idx, c0 = condensed(p, xref[b], fsteps[b], True)
        n = len(idx)
        ok, ns, sg, f = pdas(H, g, np.full(n, 4), max_sweeps=12)
        ok80, ns80, sg80, f80 = pdas(H, g, np.full(n, 4), max_sweeps=200)
        okc, nsc, nev, sgc, fc = stage_c(H, g, f, verbose=(b == 4))
        print("robot %d: pdas12 ok %s; pdas200 ok %s in %d; C from pdas12 iterate: ok %s sweeps %d evals %d  |f-f*| %.2e" % (
            b, ok, ok80, ns80, okc, nsc, nev, np.abs(fc - f80).max() if ok80 else np.nan))
        okc, nsc, nev, sgc, fc = stage_c(H, g, np.zeros(3 * n))
        print("         C from zero: ok %s sweeps %d evals %d" % (okc, nsc, nev))


if __name__ == "__main__":
    run_c()
