"""N = 64, 2048 trot robots under commands up to 1 m/s (configs[3]): tick time against max_sweeps of the active-set stage."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import torch, mpcqp
from scenario import Scenario
B, N = 2048, 64
for ms in (16, 10, 8, 6, 4, 3):
    sc = Scenario(B, n_steps=N, gaits=["trot"], seed=4242, noise_kind="hash")
    eng = mpcqp.Engine(batch=B, n_steps=N, max_sweeps=ms)
    eng.scenario_init(sc)
    eng.scenario_run(25)
    eng.synchronize()
    stream = torch.cuda.ExternalStream(eng.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    eng.scenario_run(20)
    eng.join()
    e1.record(stream); eng.synchronize()
    info = eng.info(with_y=False)
    print("max_sweeps %2d: %.3f ms per tick, %.3f M solves/s, fallback %.3f, sweeps %.2f, unsolved %d" % (
        ms, e0.elapsed_time(e1) / 20, B * 20 / e0.elapsed_time(e1) / 1e3, (info["iters"] > 0).mean(), info["sweeps"].mean(), (info["status"] != 1).sum()), flush=True)
    eng.close()
