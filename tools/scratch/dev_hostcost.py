"""Host time to enqueue a tick (no synchronisation) against the device time per tick, by number of index ranges."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import torch, mpcqp
from scenario import Scenario
B, N, T, settle = 4096, 16, 225, 25
eng = mpcqp.Engine(batch=B)
sc = Scenario(B, gaits="trot", seed=20260)
hx = np.empty((T, B, 12, N + 1)); hf = np.empty((T, B, 20, 13))
for t in range(T):
    xr, fs = sc.inputs(); hx[t], hf[t] = xr, fs
    eng.run(t, xr, fs); x = eng.solution(); sc.advance(x[:, :12] + xr[:, :, 1])
dx, df = torch.from_numpy(hx).cuda(), torch.from_numpy(hf).cuda()
px = [a.data_ptr() for a in dx]; pf = [a.data_ptr() for a in df]
for R in (1, 2, 4, 8):
    eng.set_overlap(R)
    eng.reset_warm_start()
    for t in range(settle): eng.run_device(t, px[t], pf[t])
    eng.synchronize()
    t0 = time.perf_counter()
    for t in range(settle, T): eng.run_device(t, px[t], pf[t])
    t1 = time.perf_counter()
    eng.synchronize()
    t2 = time.perf_counter()
    print("ranges %d: host enqueue %.1f us per tick, until everything finished %.1f us per tick" % (R, 1e6 * (t1 - t0) / (T - settle), 1e6 * (t2 - t0) / (T - settle)), flush=True)
# the ctypes call itself
lib, h = eng.lib, eng._h
t0 = time.perf_counter()
for i in range(20000): lib.mpcqp_launch_count(h)
print("ctypes call overhead %.2f us" % (1e6 * (time.perf_counter() - t0) / 20000))
