import sys, numpy as np, torch
sys.path.insert(0, "/root/repo/mpc-tsid_b200")
import mpcqp
from scenario import Scenario
B, T, n = 1184, 30, 32
eng = mpcqp.Engine(batch=B, n_steps=n)
sc = Scenario(B, n_steps=n, gaits="trot", seed=5)
stream = torch.cuda.ExternalStream(eng.stream)
for t in range(T):
    xr, fs = sc.inputs()
    dx, df = torch.from_numpy(xr).cuda(), torch.from_numpy(fs).cuda(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); eng.run_device(t, dx.data_ptr(), df.data_ptr()); e1.record(stream); eng.synchronize()
    x = eng.solution(); info = eng.info(False)
    if t % 5 == 0 or t == T - 1:
        print("N=32 tick %2d  %.3f ms  -> %.0f solves/s  sweeps %.2f fallbacks %d unsolved %d" % (t, e0.elapsed_time(e1), B / e0.elapsed_time(e1) * 1e3, info["sweeps"].mean(), (info["iters"] > 0).sum(), (info["status"] != 1).sum()))
    sc.advance(x[:, :12] + xr[:, :, 1])
