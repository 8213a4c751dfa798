import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/tools/scratch")
from exp_n64 import *
from exp_stagec import testset
from exp_ipm import C6, H6


def ipm2(H, g, mu0=1.0, fz0=2.0, sig_rule="adapt", tau=0.995, max_it=80, mu_id=1e-8, ytol=1e-11, mehrotra=False, split=False):
    n = H.shape[0]; nf = n // 3
    f = np.zeros(n); f[2::3] = fz0
    G = np.zeros((6 * nf, n))
    for i in range(nf):
        G[6 * i:6 * i + 6, 3 * i:3 * i + 3] = C6
    h = np.tile(H6, nf)
    s = h - G @ f
    y = mu0 / s
    sigma = 0.3
    nsolve = 0
    for it in range(1, max_it + 1):
        mu = (y @ s) / len(s)
        D = y / s
        K = H + G.T @ (D[:, None] * G)
        def solve(sig_mu, corr=None):
            c = y - D * h + sig_mu / s
            if corr is not None:
                c = c + corr / s
            fp = np.linalg.solve(K, -g - G.T @ c)
            ds = (h - G @ fp) - s
            dy = sig_mu / s - D * ds - y
            if corr is not None:
                dy = dy + corr / s
            return fp - f, ds, dy
        def steps(ds, dy):
            ap = min(1.0, tau * np.min(np.where(ds < 0, -s / np.where(ds < 0, ds, -1), np.inf)))
            ad = min(1.0, tau * np.min(np.where(dy < 0, -y / np.where(dy < 0, dy, -1), np.inf)))
            return ap, ad
        if mehrotra:
            df, ds, dy = solve(0.0); nsolve += 1
            ap, ad = steps(ds, dy)
            mu_aff = ((s + ap * ds) @ (y + ad * dy)) / len(s)
            sigma = (mu_aff / mu) ** 3
            df, ds, dy = solve(sigma * mu, corr=-ds * dy); nsolve += 1
        else:
            df, ds, dy = solve(sigma * mu); nsolve += 1
        ap, ad = steps(ds, dy)
        if not split:
            ap = ad = min(ap, ad)
        f = f + ap * df; s = s + ap * ds; y = y + ad * dy
        a = min(ap, ad)
        if sig_rule == "adapt":
            sigma = 0.05 if a > 0.9 else (0.15 if a > 0.6 else (0.3 if a > 0.3 else 0.5))
        else:
            sigma = float(sig_rule)
        mu = (y @ s) / len(s)
        if mu < mu_id:
            break
    act = (y > s).reshape(nf, 6)
    sg = np.zeros(nf, dtype=np.int64)
    for i in range(nf):
        a6 = act[i]
        apex = a6[4] or (a6[0] and a6[1]) or (a6[2] and a6[3])
        sx = (1 if a6[0] else 0) - (1 if a6[1] else 0)
        sy = (1 if a6[2] else 0) - (1 if a6[3] else 0)
        sg[i] = sig_pack(sx, sy, 1 if apex else (2 if a6[5] else 0))
    ok, nsw, sg2, f2 = pdas(H, g, sg, max_sweeps=40, ytol=ytol)
    return ok, nsolve, nsw


if __name__ == "__main__":
    ts = testset()
    p64 = km.ModelParams(n_steps=64)
    xref, fsteps = instances(64)
    for b in range(8):
        ts.append((64, p64, xref[b], fsteps[b]))
    ts = [t for t in ts if t[0] == 64] + [t for t in ts if t[0] == 16][:16]
    probs = [(N,) + tuple(condensed(p, xr, fs, True)[:2]) for (N, p, xr, fs) in ts]
    for kw in (dict(), dict(mu0=10.0), dict(mu0=100.0, fz0=5.0), dict(sig_rule=0.2), dict(sig_rule=0.1), dict(mehrotra=True),
               dict(split=True), dict(split=True, mu0=10.0, fz0=5.0), dict(mu_id=1e-6), dict(mu_id=1e-10), dict(split=True, mu_id=1e-6)):
        res = {}
        for (N, H, g) in probs:
            res.setdefault(N, []).append(ipm2(H, g, **kw))
        msg = []
        for N, rows in res.items():
            r = np.array(rows, dtype=float)
            msg.append("N%d: solved %d/%d solves mean %.1f max %d, pdas mean %.1f max %d" % (N, r[:, 0].sum(), len(r), r[:, 1].mean(), r[:, 1].max(), r[:, 2].mean(), r[:, 2].max()))
        print(kw, " | ".join(msg))
