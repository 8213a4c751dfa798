"""Sweeps per solve against the guard's tolerances (dual_tol, feas_tol): how many second sweeps are numerical near-ties?"""
import sys
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp
from scenario import Scenario
def run(B, gaits, **kw):
    sc = Scenario(B, gaits=gaits, seed=4242, noise_kind="hash")
    eng = mpcqp.Engine(batch=B, **kw)
    eng.scenario_init(sc)
    eng.scenario_run(25)
    sw = []
    hist = np.zeros(20, int)
    for t in range(10):
        eng.scenario_run(1)
        info = eng.info(with_y=False)
        sw.append(info["sweeps"].mean()); hist += np.bincount(info["sweeps"], minlength=20)[:20]
    print("%-20s %s: sweeps/solve %.4f  hist %s fallback %.4f unsolved %d" % ("/".join(gaits) if not isinstance(gaits, str) else gaits, kw, np.mean(sw), hist[:8], (info["iters"] > 0).mean(), (info["status"] != 1).sum()), flush=True)
    eng.close()
for gaits in ("trot", ["trot", "pace", "bound", "walk"], ["pace"], ["bound"], ["walk"]):
    for kw in ({}, {"dual_tol": 1e-10}, {"dual_tol": 1e-9}, {"dual_tol": 1e-8}, {"dual_tol": 1e-9, "feas_tol": 1e-8}):
        run(8192, gaits, **kw)
