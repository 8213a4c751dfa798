"""Why do pace / bound / walk need ~1.8 sweeps per solve?  For robots that needed 2+ sweeps: which steps' active sets differ from
the warm-start guess (previous tick shifted by one step), which rows (fx+, fx-, fy+, fy-, fz) flip, and at which gait phase."""
import os, sys
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp
from scenario import Scenario
B = 2048
for gait in os.environ.get("GAITS", "pace,bound,walk,trot").split(","):
    sc = Scenario(B, gaits=gait, seed=20260)
    eng = mpcqp.Engine(batch=B)
    prev_act = None; hk = np.zeros(16, int); nchg = []; rows = np.zeros((5, 2), int); ph2 = np.zeros(16); phn = np.zeros(16); sw = []
    newcontact = np.zeros(2, int)
    for t in range(60):
        xref, fsteps = sc.inputs()
        eng.run(t, xref, fsteps); x = eng.solution(); info = eng.info(with_y=False)
        con = info["contact"]                                   # (B, N, 4)
        act = info["active"] & con[..., None]                   # (B, N, 4, 5)
        if t >= 25 and prev_act is not None:
            guess = np.concatenate([prev_act[:, 1:], prev_act[:, :1]], axis=1) & con[..., None]
            d = guess != act
            two = info["sweeps"] >= 2
            sw.append(info["sweeps"].mean())
            ph = (t + sc.phase) % 16
            for p in range(16):
                phn[p] += (ph == p).sum(); ph2[p] += (two & (ph == p)).sum()
            dd = d[two]
            hk += dd.any(axis=(2, 3)).sum(axis=0)
            nchg.append(dd.any(axis=(2, 3)).sum(axis=1).mean() if two.any() else 0)
            rows[:, 0] += (dd & act[two]).sum(axis=(0, 1, 2))       # rows that became active
            rows[:, 1] += (dd & ~act[two]).sum(axis=(0, 1, 2))      # rows that were released
            # is the changed foot-step the first step of a stance phase (touch-down)?
            first_of_stance = con & ~np.concatenate([con[:, :1], con[:, :-1]], axis=1)
            chg_fs = dd.any(axis=3)
            newcontact[0] += (chg_fs & first_of_stance[two]).sum(); newcontact[1] += chg_fs.sum()
        prev_act = info["active"]
        sc.advance(x[:, :12] + xref[:, :, 1])
    print("== %s: sweeps/solve %.3f, changed steps per 2+-sweep robot %.2f" % (gait, np.mean(sw), np.mean(nchg)))
    print("   changed step histogram k=0..15:", hk)
    print("   rows became active / released (fx+ fx- fy+ fy- fz):", rows[:, 0], rows[:, 1])
    print("   P(2+ | phase):", np.round(ph2 / np.maximum(phn, 1), 2))
    print("   changed foot-steps that are the first step of a stance phase: %d of %d" % tuple(newcontact))
    eng.close()
