"""Pure active-set (no ADMM stage) convergence from cold start, and N = 64."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import mpcqp
from scenario import Scenario
from common import certify, assert_certified
from oracle import mpc_build
for N, B, ms in ((16, 256, 60), (32, 128, 60), (64, 64, 60)):
    sc = Scenario(B, n_steps=N, gaits=["trot", "pace", "bound", "walk", "static"], seed=11)
    eng = mpcqp.Engine(batch=B, n_steps=N, mode=5, max_sweeps=ms)
    par = mpc_build.Params(n_steps=N)
    for t in range(6):
        xref, fsteps = sc.inputs()
        eng.run(t, xref, fsteps)
        x, info = eng.solution(), eng.info()
        ns = (info["status"] != 1).sum()
        msg = ""
        if N == 64 or t < 2:
            for b in range(0, B, max(B // 4, 1)):
                if info["status"][b] == 1:
                    cert = certify(xref[b], fsteps[b], x[b], info["y"][b], first_tick=(t == 0), params=par)
                    assert_certified(cert, "N=%d tick %d robot %d" % (N, t, b))
            msg = "certified sample ok"
        print("N %d tick %d: unsolved %d of %d, sweeps mean %.2f max %d  %s" % (N, t, ns, B, info["sweeps"].mean(), info["sweeps"].max(), msg))
        sc.advance(x[:, :12] + xref[:, :, 1])
    eng.close()
