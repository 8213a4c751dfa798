import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import mpcqp
from scenario import Scenario
from common import certify, assert_certified
from oracle import mpc_build
n = 64
par = mpc_build.Params(n_steps=n)
for dual_tol in (1e-12, 1e-11, 1e-10):
    sc = Scenario(16, n_steps=n, gaits=["trot"], seed=64)
    eng = mpcqp.Engine(batch=16, n_steps=n, dual_tol=dual_tol, max_iter=200)
    for t in range(4):
        xref, fsteps = sc.inputs()
        t0 = time.perf_counter(); eng.run(t, xref, fsteps); x = eng.solution(); dt = time.perf_counter() - t0
        info = eng.info()
        worst = 0
        for b in range(16):
            if info["status"][b] == 1 and (b % 5 == 0 or info["sweeps"][b] > 20):
                c = certify(xref[b], fsteps[b], x[b], info["y"][b], first_tick=(t == 0), params=par)
                worst = max(worst, c["stat"], c["bad_sign"], c["prim"])
        print("dual_tol %.0e tick %d status %s sweeps %s iters %s  worst cert %.1e  %.1f ms" % (dual_tol, t, info["status"], info["sweeps"], info["iters"], worst, dt * 1e3))
        sc.advance(x[:, :12] + xref[:, :, 1])
    eng.close()
