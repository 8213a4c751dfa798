"""Device-resident tick time with consecutive ticks issued as independent index ranges (mpcqp_set_overlap)."""
import os, sys
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import torch, mpcqp
from scenario import Scenario
def run(B, N=16, ticks=60, settle=25, gaits="trot", variants=(1, 2, 3, 4)):
    eng = mpcqp.Engine(batch=B, n_steps=N)
    sc = Scenario(B, n_steps=N, gaits=gaits, seed=20260)
    T = settle + ticks
    hx = np.empty((T, B, 12, N + 1)); hf = np.empty((T, B, 20, 13))
    for t in range(T):
        xr, fs = sc.inputs(); hx[t], hf[t] = xr, fs
        eng.run(t, xr, fs); x = eng.solution()
        sc.advance(x[:, :12] + xr[:, :, 1])
    xlast = x.copy()
    dx, df = torch.from_numpy(hx).cuda(), torch.from_numpy(hf).cuda()
    stream = torch.cuda.ExternalStream(eng.stream)
    for R in variants:
        eng.set_overlap(R)
        best = 1e9
        for rep in range(3):
            eng.reset_warm_start()
            for t in range(settle): eng.run_device(t, dx[t].data_ptr(), df[t].data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.synchronize()
            e0.record(stream)
            for t in range(settle, T): eng.run_device(t, dx[t].data_ptr(), df[t].data_ptr())
            eng.join()
            e1.record(stream); eng.synchronize()
            best = min(best, e0.elapsed_time(e1) / ticks)
        same = np.array_equal(eng.solution(), xlast)
        info = eng.info(with_y=False)
        print("B %6d N %d ranges %d: %.4f ms/tick  %.2f M solves/s  sweeps %.3f unsolved %d identical %s" % (
            B, N, R, best, B / best / 1e3, info["sweeps"].mean(), (info["status"] != 1).sum(), same), flush=True)
    eng.close()
if __name__ == "__main__":
    run(4096, variants=(1, 4, 6, 8))
    run(8192, variants=(1, 2, 4, 8))
    run(4096, gaits=["trot", "pace", "bound", "walk"], variants=(1, 4, 8))
    run(4096, N=32, ticks=30, variants=(1, 2, 4, 8))
