"""Closed loop of B robots on the engine, tick by tick against the numpy kernel model and the oracle certificate: python dev_closed_loop.py [B] [T] [gait,gait,...] [mode]"""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo/tests")
import mpcqp
from scenario import Scenario
from common import certify
from kernel_model import ModelParams, Engine as Model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10
gaits = sys.argv[3].split(",") if len(sys.argv) > 3 else ["trot"]
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 3
eng = mpcqp.Engine(batch=B, mode=mode)
sc = Scenario(B, gaits=gaits, seed=3)
models = [Model(ModelParams(max_as=6)) for _ in range(min(B, 4))]
worst = dict(prim=0, stat=0, comp=0, bad_sign=0); wm = 0
for t in range(T):
    xr, fs = sc.inputs()
    eng.run(t, xr, fs)
    f0 = eng.forces(); x = eng.solution(); info = eng.info()
    nfb = eng.fallback_count()
    for b in range(min(B, 8)):
        c = certify(xr[b], fs[b], x[b], info["y"][b], first_tick=(t == 0))
        for k_ in worst: worst[k_] = max(worst[k_], c[k_])
        if not np.array_equal(c["contact"].astype(bool), info["contact"][b]): print("  contact mask mismatch", t, b)
        if not np.array_equal(c["active"].reshape(16,4,5), info["active"][b]): print("  active mask mismatch", t, b, np.argwhere(c["active"].reshape(16,4,5) != info["active"][b])[:5])
    for b in range(len(models)):
        out = models[b].solve(xr[b], fs[b], first_tick=(t == 0))
        wm = max(wm, np.abs(out["x"] - x[b]).max())
    print("tick", t, "status", np.bincount(info["status"], minlength=4), "sweeps mean %.2f max %d" % (info["sweeps"].mean(), info["sweeps"].max()),
          "iters max", info["iters"].max(), "fallback", nfb, "f0[0]", np.round(f0[0][:6], 3), "obj0 %.6f" % info["obj"][0])
    xn = x[:, :12] + xr[:, :, 1]
    sc.advance(xn)
print("worst cert", worst, "worst |x - model|", wm)
