"""Profiling driver: mixed gaits, 16384 robots, host-staged inputs replayed from the device (MPCQP_LIB selects the build)."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import torch, mpcqp
from scenario import Scenario
B, N, T = 16384, 16, 30
sc = Scenario(B, gaits=["trot", "pace", "bound", "walk"], seed=20260)
eng = mpcqp.Engine(batch=B)
hx, hf = np.empty((T, B, 12, N + 1)), np.empty((T, B, 20, 13))
for t in range(T):
    xr, fs = sc.inputs(); hx[t], hf[t] = xr, fs
    eng.run(t, xr, fs); x = eng.solution(); sc.advance(x[:, :12] + xr[:, :, 1])
dx, df = torch.from_numpy(hx).cuda(), torch.from_numpy(hf).cuda()
eng.reset_warm_start()
for t in range(T):
    eng.run_device(t, dx[t].data_ptr(), df[t].data_ptr())
eng.synchronize()
print("ok", eng.info(with_y=False)["sweeps"].mean())
