import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import torch, mpcqp
from scenario import Scenario
def run(N, B, ticks=60, settle=25, **kw):
    sc = Scenario(B, n_steps=N, gaits=["trot"], seed=4242, noise_kind="hash")
    eng = mpcqp.Engine(batch=B, n_steps=N, **kw)
    eng.scenario_init(sc)
    eng.scenario_run(settle); eng.synchronize()
    t0 = time.perf_counter(); eng.scenario_run(ticks - settle); eng.synchronize()
    ms = (time.perf_counter() - t0) / (ticks - settle) * 1e3
    info = eng.info(with_y=False)
    print("N %d B %d %s: %.2f ms/tick  sweeps %.2f fallback %.3f unsolved %d" % (N, B, kw, ms, info["sweeps"].mean(), (info["iters"] > 0).mean(), (info["status"] != 1).sum()), flush=True)
    eng.close()
for ms in (4, 6, 8, 10, 12, 16):
    run(64, 2048, max_sweeps=ms)
for ms in (6, 8, 12, 16):
    run(32, 4096, max_sweeps=ms)
for ms in (6, 8, 12, 16):
    run(16, 4096, max_sweeps=ms)
