"""Stage-wise ADMM stage (mode 15) against the dense engine (mode 3): cold starts and mixed gaits; N = 64 closed loop."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import mpcqp
from scenario import Scenario
from common import certify, assert_certified
from oracle import mpc_build
B = 64
sc = Scenario(B, gaits=["trot", "pace", "bound", "walk", "static"], seed=5)
dense = mpcqp.Engine(batch=B, mode=3)
sw = mpcqp.Engine(batch=B, mode=15, max_sweeps=3)       # few sweeps -> many robots reach the stage-wise ADMM stage
for t in range(6):
    xref, fsteps = sc.inputs()
    dense.run(t, xref, fsteps); xd = dense.solution(); idn = dense.info()
    t0 = time.perf_counter(); sw.run(t, xref, fsteps); xs = sw.solution(); dt = time.perf_counter() - t0
    isw = sw.info()
    print("tick %d  max|df| %.2e  status dense %s sw %s  admm robots %d  iters mean %.0f max %d  sweeps %.2f  masks %s  %.1f ms" % (
        t, np.abs(xd - xs).max(), np.bincount(idn["status"], minlength=4), np.bincount(isw["status"], minlength=4), (isw["iters"] > 0).sum(),
        isw["iters"][isw["iters"] > 0].mean() if (isw["iters"] > 0).any() else 0, isw["iters"].max(), isw["sweeps"].mean(),
        np.array_equal(idn["active"], isw["active"]), dt * 1e3))
    sc.advance(xd[:, :12] + xref[:, :, 1])
dense.close(); sw.close()
n = 64
sc = Scenario(16, n_steps=n, gaits=["trot"], seed=64)
eng = mpcqp.Engine(batch=16, n_steps=n)
par = mpc_build.Params(n_steps=n)
for t in range(5):
    xref, fsteps = sc.inputs()
    t0 = time.perf_counter(); eng.run(t, xref, fsteps); x = eng.solution(); dt = time.perf_counter() - t0
    info = eng.info()
    for b in (0, 8, 10):
        if info["status"][b] == 1:
            assert_certified(certify(xref[b], fsteps[b], x[b], info["y"][b], first_tick=(t == 0), params=par), "N64 %d %d" % (t, b))
    print("N=64 tick %d status %s sweeps %s iters %s  %.1f ms" % (t, info["status"], info["sweeps"], info["iters"], dt * 1e3))
    sc.advance(x[:, :12] + xref[:, :, 1])
