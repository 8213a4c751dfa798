"""Device-resident tick time at several batch sizes and CTA-per-SM caps (MPCQP_RIC_CTAS)."""
import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import torch, mpcqp
from scenario import Scenario
def run(B, ctas, N=16, ticks=60, settle=25, gaits="trot"):
    if ctas: os.environ["MPCQP_RIC_CTAS"] = str(ctas)
    else: os.environ.pop("MPCQP_RIC_CTAS", None)
    eng = mpcqp.Engine(batch=B, n_steps=N)
    sc = Scenario(B, n_steps=N, gaits=gaits, seed=20260)
    T = settle + ticks
    hx = np.empty((T, B, 12, N + 1)); hf = np.empty((T, B, 20, 13))
    for t in range(T):
        xr, fs = sc.inputs(); hx[t], hf[t] = xr, fs
        eng.run(t, xr, fs); x = eng.solution()
        sc.advance(x[:, :12] + xr[:, :, 1])
    dx, df = torch.from_numpy(hx).cuda(), torch.from_numpy(hf).cuda()
    stream = torch.cuda.ExternalStream(eng.stream)
    eng.reset_warm_start()
    for t in range(settle): eng.run_device(t, dx[t].data_ptr(), df[t].data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.synchronize()
    e0.record(stream)
    for t in range(settle, T): eng.run_device(t, dx[t].data_ptr(), df[t].data_ptr())
    e1.record(stream); eng.synchronize()
    ms = e0.elapsed_time(e1) / ticks
    info = eng.info(with_y=False)
    print("B %6d N %d ctas/SM %s: %.4f ms/tick  %.2f M solves/s  sweeps %.3f unsolved %d" % (B, N, ctas or "max", ms, B / ms / 1e3, info["sweeps"].mean(), (info["status"] != 1).sum()), flush=True)
    eng.close()
if __name__ == "__main__":
    for B in (4096, 16384):
        for c in (0, 8, 7, 6):
            run(B, c)
