"""End-to-end (host buffers) tick time for different chunk sizes: MPCQP_CHUNK=<n> python dev_e2e.py"""
import sys, os, time
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp, torch
from scenario import Scenario
B = 4096
sc = Scenario(B, gaits="trot", seed=20260)
eng = mpcqp.Engine(batch=B)
T = 45
hx = torch.empty((T, B, 12, 17), dtype=torch.float64, pin_memory=True).numpy()
hf = torch.empty((T, B, 20, 13), dtype=torch.float64, pin_memory=True).numpy()
for t in range(T):
    xr, fs = sc.inputs(); hx[t], hf[t] = xr, fs
    eng.run(t, hx[t], hf[t]); x = eng.solution(); sc.advance(x[:, :12] + xr[:, :, 1])
out = torch.empty((B, 12), dtype=torch.float64, pin_memory=True).numpy()
best = 1e9
for rep in range(3):
    eng.reset_warm_start()
    for t in range(25):
        eng.run(t, hx[t], hf[t]); eng.forces(out=out)
    t0 = time.perf_counter()
    for t in range(25, T):
        eng.run(t, hx[t], hf[t]); eng.forces(out=out)
    best = min(best, (time.perf_counter() - t0) / (T - 25))
print("MPCQP_CHUNK=%s: e2e %.3f ms per tick -> %.2f M solves/s" % (os.environ.get("MPCQP_CHUNK", "default"), best * 1e3, B / best * 1e-6))
