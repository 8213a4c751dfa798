"""The asynchronous host loop on its own (no torch.distributed): python dev_async2.py <device> <ranges>"""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import torch, mpcqp
from scenario import Scenario
dev, R = int(sys.argv[1]), int(sys.argv[2])
torch.cuda.set_device(dev)
B, N, T, W = 4096, 16, 80, 25
eng = mpcqp.Engine(batch=B, device=dev)
sc = Scenario(B, gaits="trot", seed=20260 + dev)
hx = torch.empty((T, B, 12, N + 1), dtype=torch.float64, pin_memory=True).numpy()
hf = torch.empty((T, B, 20, 13), dtype=torch.float64, pin_memory=True).numpy()
for t in range(T):
    xr, fs = sc.inputs(); hx[t], hf[t] = xr, fs
    eng.run(t, hx[t], hf[t]); x = eng.solution(); sc.advance(x[:, :12] + xr[:, :, 1])
out = torch.empty((B, 12), dtype=torch.float64, pin_memory=True).numpy()
for rep in range(3):
    eng.set_overlap(R)
    eng.reset_warm_start()
    for t in range(W):
        eng.run(t, hx[t], hf[t]); eng.forces(out=out)
    t0 = time.perf_counter()
    tr = tw = 0.0
    for i in range(W, T):
        a = time.perf_counter()
        eng.run(i, hx[i], hf[i]); eng.result_async(i & 1)
        b = time.perf_counter()
        if i > W: eng.result_wait((i - 1) & 1, out)
        c = time.perf_counter()
        tr += b - a; tw += c - b
    eng.result_wait((T - 1) & 1, out)
    dt = (time.perf_counter() - t0) / (T - W)
    print("dev %d ranges %d: %.3f ms per tick (%.2f M solves/s); host: issue %.1f us, wait %.1f us per tick" % (dev, R, dt * 1e3, B / dt / 1e6, 1e6 * tr / (T - W), 1e6 * tw / (T - W)), flush=True)
