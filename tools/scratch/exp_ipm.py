"""CPU experiment 3: primal-dual interior point (one Riccati-shaped solve per iteration) as the robust fallback, followed
by active-set identification + guarded sweeps."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/tools/scratch")
from exp_n64 import *
from exp_stagec import testset, stage_c2

C6 = np.array([[1, 0, -MU], [-1, 0, -MU], [0, 1, -MU], [0, -1, -MU], [0, 0, -1.0], [0, 0, 1.0]])
H6 = np.array([0, 0, 0, 0, 0, FZ])


def ipm(H, g, sigma=0.2, tau=0.995, max_it=60, mu_stop=1e-9, verbose=False, try_every=1, ytol=1e-11, adaptive=True):
    n = H.shape[0]; nf = n // 3
    f = np.zeros(n); f[2::3] = 2.0
    G = np.zeros((6 * nf, n))
    for i in range(nf):
        G[6 * i:6 * i + 6, 3 * i:3 * i + 3] = C6
    h = np.tile(H6, nf)
    s = h - G @ f
    y = 1.0 / s
    sweeps_extra = 0
    sig_prev = None
    for it in range(1, max_it + 1):
        mu = (y @ s) / len(s)
        D = y / s
        if adaptive and it > 1:
            pass
        K = H + G.T @ (D[:, None] * G)
        rhs = -g - G.T @ (y - D * h + sigma * mu / s)
        fp = np.linalg.solve(K, rhs)
        df = fp - f
        ds = (h - G @ fp) - s
        dy = sigma * mu / s - D * ds - y
        ap = min(1.0, tau * np.min(np.where(ds < 0, -s / np.where(ds < 0, ds, -1), np.inf)))
        ad = min(1.0, tau * np.min(np.where(dy < 0, -y / np.where(dy < 0, dy, -1), np.inf)))
        a = min(ap, ad)
        f = f + a * df; s = s + a * ds; y = y + a * dy
        if adaptive:
            sigma = min(0.5, max(0.05, (1 - a) ** 2 * 4 + 0.05)) if True else sigma
        mu = (y @ s) / len(s)
        rd = np.abs(H @ f + g + G.T @ y).max()
        if verbose:
            print("  ipm %d: alpha %.3f mu %.3e rd %.3e" % (it, a, mu, rd))
        if mu < 1e-4 and it % try_every == 0:
            # identify: row active iff y > s
            act = (y > s).reshape(nf, 6)
            sg = np.zeros(nf, dtype=np.int64)
            for i in range(nf):
                a6 = act[i]
                apex = a6[4] or (a6[0] and a6[1]) or (a6[2] and a6[3])
                sx = (1 if a6[0] else 0) - (1 if a6[1] else 0)
                sy = (1 if a6[2] else 0) - (1 if a6[3] else 0)
                sg[i] = sig_pack(sx, sy, 1 if apex else (2 if a6[5] else 0))
            if sig_prev is not None and np.array_equal(sg, sig_prev):
                continue
            sig_prev = sg
            sweeps_extra += 1
            fh, grad, oks, ns = sweep(H, g, sg, ytol)
            if oks.all():
                return True, it, sweeps_extra, sg, fh, f
        if mu < mu_stop:
            break
    return False, it, sweeps_extra, sig_prev, f, f


if __name__ == "__main__":
    ts = testset()
    # plus the 8 hard ones
    p64 = km.ModelParams(n_steps=64)
    xref, fsteps = instances(64)
    for b in range(8):
        ts.append((64, p64, xref[b], fsteps[b]))
    stats = {}
    t0 = time.time()
    for ii, (N, p, xr, fs) in enumerate(ts):
        H, g, idx, c0 = condensed(p, xr, fs, True)
        n = len(idx)
        ok, it, sw, sg, fh, fi = ipm(H, g, verbose=(ii == len(ts) - 4))
        extra = 0
        if not ok:
            ok2, extra, sg2, f2 = pdas(H, g, sg, max_sweeps=40, ytol=1e-11)
            ok = ok2
        stats.setdefault(N, []).append((ok, it, sw, extra))
    for N, rows in stats.items():
        r = np.array(rows, dtype=float)
        print("N %d: %d instances; solved %d; ipm iterations mean %.1f max %d; polish attempts mean %.1f max %d; pdas-after mean %.1f max %d" % (
            N, len(r), r[:, 0].sum(), r[:, 1].mean(), r[:, 1].max(), r[:, 2].mean(), r[:, 2].max(), r[:, 3].mean(), r[:, 3].max()))
    print("time %.1f s" % (time.time() - t0))
