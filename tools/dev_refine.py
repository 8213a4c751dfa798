import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo/tests")
import mpcqp
from scenario import Scenario
B, T = 1024, 12
for gaits in (["trot"], ["trot", "pace", "bound", "walk", "static"]):
    a = mpcqp.Engine(batch=B, refine=1); b = mpcqp.Engine(batch=B, refine=0)
    sc = Scenario(B, gaits=gaits, seed=77)
    worst = 0; sa = sb = 0; fa = fb = 0
    for t in range(T):
        xr, fs = sc.inputs()
        a.run(t, xr, fs); b.run(t, xr, fs)
        xa, xb = a.solution(), b.solution(); ia, ib = a.info(False), b.info(False)
        worst = max(worst, np.abs(xa - xb).max()); sa += ia["sweeps"].mean(); sb += ib["sweeps"].mean()
        fa += (ia["iters"] > 0).sum(); fb += (ib["iters"] > 0).sum()
        assert (ia["status"] == 1).all() and (ib["status"] == 1).all()
        sc.advance(xa[:, :12] + xr[:, :, 1])
    print(gaits, "max |x_refine1 - x_refine0| %.2e" % worst, "sweeps", sa / T, sb / T, "fallbacks", fa, fb)
