"""N = 32 / 64 closed loop on the device: per-tick sweeps / fallback statistics under parameter variants."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import torch, mpcqp
from scenario import Scenario
def run(N, B, ticks=40, **kw):
    sc = Scenario(B, n_steps=N, gaits=["trot"], seed=4242, noise_kind="hash")
    eng = mpcqp.Engine(batch=B, n_steps=N, **kw)
    eng.scenario_init(sc)
    rows = []
    for t in range(ticks):
        t0 = time.perf_counter()
        eng.scenario_run(1); eng.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        info = eng.info(with_y=False)
        rows.append((ms, info["sweeps"].mean(), (info["iters"] > 0).mean(), info["iters"][info["iters"] > 0].mean() if (info["iters"] > 0).any() else 0, (info["status"] != 1).sum(), info["sweeps"].max()))
    r = np.array(rows)
    print("N %d B %d %s" % (N, B, kw))
    for a, b in ((0, 5), (5, 15), (15, 25), (25, 40)):
        q = r[a:b]
        print("   ticks %2d..%2d: ms %.2f sweeps %.2f fallback %.3f ipm-iters %.1f unsolved %d max-sweeps %d" % (a, b - 1, q[:, 0].mean(), q[:, 1].mean(), q[:, 2].mean(), q[:, 3].mean(), q[:, 4].sum(), q[:, 5].max()))
    eng.close()
if __name__ == "__main__":
    run(64, 2048)
    run(64, 2048, dual_tol=1e-10)
    run(64, 2048, max_sweeps=40)
    run(64, 2048, mode=5, max_sweeps=60)
    run(32, 4096)
