"""A/B two builds of libmpcqp on the bench workload: steady-state and cold-window tick times."""
import sys, os, numpy as np, torch, ctypes
sys.path.insert(0, "/root/repo/mpc-tsid_b200")
lib = sys.argv[1]
import mpcqp
mpcqp._LIB_PATH = os.path.join(os.path.dirname(mpcqp._LIB_PATH), lib)
from scenario import Scenario
B, T = 4096, 60
eng = mpcqp.Engine(batch=B)
sc = Scenario(B, gaits="trot", seed=20260)
xs, fs_ = [], []
for t in range(T):
    xr, fs = sc.inputs(); eng.run(t, xr, fs); x = eng.solution(); sc.advance(x[:, :12] + xr[:, :, 1])
    xs.append(torch.from_numpy(xr).cuda()); fs_.append(torch.from_numpy(fs).cuda())
stream = torch.cuda.ExternalStream(eng.stream)
best = None
for rep in range(3):
    eng.reset_warm_start(); eng.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(T + 1)]
    for t in range(T):
        ev[t].record(stream); eng.run_device(t, xs[t].data_ptr(), fs_[t].data_ptr())
    ev[T].record(stream); eng.synchronize()
    ms = np.array([ev[t].elapsed_time(ev[t + 1]) for t in range(T)])
    best = ms if best is None else np.minimum(best, ms)
print("%-18s cold ticks 0-9: %.3f ms mean | steady ticks 25-59: mean %.4f p50 %.4f max %.4f ms -> %.2f M solves/s" % (
    lib, best[:10].mean(), best[25:].mean(), np.median(best[25:]), best[25:].max(), B / best[25:].mean() / 1e3))
