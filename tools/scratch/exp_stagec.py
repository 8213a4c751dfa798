"""CPU experiment 2: variants of the monotone fallback (stage C) on a wider instance set."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/tools/scratch")
from exp_n64 import *


def active_rows(sig):
    sx, sy, tz = sig_unpack(sig)
    if tz == 1:
        return [True] * 6
    return [sx > 0, sx < 0, sy > 0, sy < 0, False, tz == 2]


def ratio_alpha2(f, d, sigs):
    a = 1.0
    blk = -1
    for i in range(len(f) // 3):
        x, dx = f[3 * i:3 * i + 3], d[3 * i:3 * i + 3]
        act = active_rows(sigs[i])
        rows = [(x[0] - MU * x[2], dx[0] - MU * dx[2]), (-x[0] - MU * x[2], -dx[0] - MU * dx[2]),
                (x[1] - MU * x[2], dx[1] - MU * dx[2]), (-x[1] - MU * x[2], -dx[1] - MU * dx[2]),
                (-x[2], -dx[2]), (x[2] - FZ, dx[2])]
        for r, (c, dc) in enumerate(rows):
            if act[r]:
                continue
            if dc > 0 and c + a * dc > 0:
                a = min(a, max(-c / dc, 0.0)); blk = i
    return a, blk


def stage_c2(H, g, f0, sig0=None, max_sweeps=300, ytol=1e-12, ftol=1e-9, verbose=False, path=True, kmax=8):
    """monotone: face minimiser -> feasible? accept + guard/release : path search (clip retraction, delta-phi from the
    difference) with the ratio-test step as the guaranteed fallback."""
    f = clip_all(f0)
    sigs = activity(f)
    gf = H @ f + g
    nev = 0
    single = False
    last_release = None
    for s in range(max_sweeps):
        fh, grad, oks, ns = sweep(H, g, sigs, ytol, ftol)
        feas = np.abs(clip_all(fh) - fh).max() <= ftol
        if feas:
            if oks.all():
                return True, s + 1, nev, sigs, fh
            f, gf = fh, grad
            cand = np.flatnonzero(ns != sigs)
            if single:
                # most negative multiplier first: approximate by order of cand, skipping those tried at this point
                t = cand[0]
                for c in cand:
                    if last_release is None or c not in last_release:
                        t = c; break
                last_release = (last_release or []) + [t]
                sigs = sigs.copy(); sigs[t] = ns[t]
            else:
                sigs = ns.copy()
                last_release = None
            if verbose: print("   C %d: face minimiser, %d candidates, single %s" % (s, len(cand), single))
            continue
        d = fh - f
        amax, blk = ratio_alpha2(f, d, sigs)
        best = None
        if path:
            a = 1.0
            for _ in range(kmax):
                ft = clip_all(f + a * d); nev += 1
                dl = ft - f
                dphi = gf @ dl + 0.5 * dl @ H @ dl
                if dphi < 0 and (best is None or dphi < best[0]):
                    best = (dphi, a, ft)
                    break
                a *= 0.5
                if a <= amax: break
        dl = amax * d
        dphi_r = gf @ dl + 0.5 * dl @ H @ dl
        if best is None or dphi_r < best[0]:
            best = (dphi_r, amax, f + dl)
        dphi, a, ft = best
        if verbose: print("   C %d: step alpha %.3g (amax %.3g) dphi %.3e" % (s, a, amax, dphi))
        if a == 0.0 or np.abs(ft - f).max() == 0.0:
            # null step: a row outside the working set blocks at zero distance -> add it (degenerate), or switch to single release
            if last_release is not None and not single:
                single = True
                sigs = activity(f)          # back to the face we sat on
                last_release = []
                continue
            sigs = activity(ft + 0.0)
            if blk >= 0:
                pass
        gf = gf + H @ (ft - f)
        f = ft
        sigs = activity(f)
        single = False if a > 0 else single
    return False, max_sweeps, nev, sigs, f


def testset():
    out = []
    rng = np.random.default_rng(7)
    for N, B in ((16, 48), (32, 32), (64, 24)):
        p = km.ModelParams(n_steps=N)
        v_ref = np.zeros((B, 6))
        v_ref[:, 0] = rng.uniform(-0.8, 1.5, B); v_ref[:, 1] = rng.uniform(-0.5, 0.5, B); v_ref[:, 5] = rng.uniform(-0.8, 0.8, B)
        sc = Scenario(B, n_steps=N, gaits=["trot", "pace", "bound", "walk"], seed=100 + N, v_ref=v_ref)
        xref, fsteps = sc.inputs()
        # aggressive initial states
        xref[:, 6:12, 0] += rng.normal(0, 0.3, (B, 6))
        xref[:, 3:5, 0] += rng.normal(0, 0.1, (B, 2))
        for b in range(B):
            out.append((N, p, xref[b], fsteps[b]))
    return out


if __name__ == "__main__":
    ts = testset()
    stats = {}
    t0 = time.time()
    for (N, p, xr, fs) in ts:
        H, g, idx, c0 = condensed(p, xr, fs, True)
        n = len(idx)
        okA, nA, sgA, fA = pdas(H, g, np.full(n, 4), max_sweeps=300)
        ok12, n12, sg12, f12 = pdas(H, g, np.full(n, 4), max_sweeps=12)
        # PDAS until cycle detection, then C
        okC, nC, nev, sgC, fC = stage_c2(H, g, f12) if not ok12 else (True, 0, 0, sg12, f12)
        okR, nR, nevR, sgR, fR = stage_c2(H, g, f12, path=False) if not ok12 else (True, 0, 0, sg12, f12)
        err = np.abs(fC - fA).max() if okA and okC else np.nan
        stats.setdefault(N, []).append((okA, nA, ok12, okC, nC, nev, okR, nR, err))
    for N, rows in stats.items():
        r = np.array(rows, dtype=float)
        print("N %d: %d instances; pdas300 solved %d (max sweeps %d); pdas12 solved %d; C solved %d/%d (mean %.1f max %d sweeps, evals mean %.1f); "
              "ratio-only solved %d (mean %.1f max %d); max |fC - fA| %.2e" % (
                  N, len(r), r[:, 0].sum(), r[:, 1].max(), r[:, 2].sum(), r[:, 3].sum(), len(r), r[r[:, 2] == 0, 4].mean() if (r[:, 2] == 0).any() else 0,
                  r[:, 4].max(), r[:, 5].mean(), r[:, 6].sum(), r[r[:, 2] == 0, 7].mean() if (r[:, 2] == 0).any() else 0, r[:, 7].max(), np.nanmax(r[:, 8])))
    print("time %.1f s" % (time.time() - t0))
