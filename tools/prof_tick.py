"""Profiling driver: 30 closed-loop ticks of 4096 trot robots through the host API (2 riccati_kernel launches per tick: the
host-input path solves in two chunks), then the same 30 ticks replayed device-resident (1 launch per tick).  Under
`ncu -k regex:riccati_kernel --launch-skip 85 --launch-count 1` the capture is tick 25 of the replay: a steady-state,
full-batch launch of the headline configuration.  Usage: python tools/prof_tick.py [batch] [n_steps]"""
import sys
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import torch, mpcqp
from scenario import Scenario
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
T = 30
sc = Scenario(B, n_steps=N, gaits="trot", seed=20260)
eng = mpcqp.Engine(batch=B, n_steps=N)
hx, hf = np.empty((T, B, 12, N + 1)), np.empty((T, B, 20, 13))
for t in range(T):
    xr, fs = sc.inputs(); hx[t], hf[t] = xr, fs
    eng.run(t, xr, fs); x = eng.solution(); sc.advance(x[:, :12] + xr[:, :, 1])
dx, df = torch.from_numpy(hx).cuda(), torch.from_numpy(hf).cuda()
eng.reset_warm_start()
for t in range(T):
    eng.run_device(t, dx[t].data_ptr(), df[t].data_ptr())
eng.synchronize()
info = eng.info(with_y=False)
print("ok: sweeps/solve %.3f, unsolved %d, launches %d" % (info["sweeps"].mean(), (info["status"] != 1).sum(), eng.launches))
