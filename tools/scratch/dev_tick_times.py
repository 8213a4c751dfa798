"""Per-tick device time vs fallback statistics on the bench workload (B=4096 trot closed loop)."""
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo/mpc-tsid_b200")
import mpcqp
from scenario import Scenario
B, T = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 40
kw = {}
for a in sys.argv[2:]:
    k, v = a.split("="); kw[k] = float(v) if "." in v or "e" in v else int(v)
eng = mpcqp.Engine(batch=B, **kw)
sc = Scenario(B, gaits="trot", seed=20260)
stream = torch.cuda.ExternalStream(eng.stream)
rows = []
for t in range(T):
    xr, fs = sc.inputs()
    dx, df = torch.from_numpy(xr).cuda(), torch.from_numpy(fs).cuda()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); eng.run_device(t, dx.data_ptr(), df.data_ptr()); e1.record(stream)
    eng.synchronize()
    x = eng.solution(); info = eng.info(False)
    it = info["iters"]
    rows.append((e0.elapsed_time(e1), (it > 0).sum(), it.max(), info["sweeps"].mean(), (info["status"] != 1).sum()))
    sc.advance(x[:, :12] + xr[:, :, 1])
for t, r in enumerate(rows):
    if t < 3 or t % 4 == 0: print("tick %2d  %.3f ms  fallbacks %4d  max iters %4d  sweeps %.2f unsolved %d" % ((t,) + r))
ms = np.array([r[0] for r in rows[5:]])
print("steady: mean %.3f  p50 %.3f  p99 %.3f  max %.3f ms;  corr(ms, max iters) %.2f" % (ms.mean(), np.median(ms), np.percentile(ms, 99), ms.max(),
      np.corrcoef(ms, np.array([r[2] for r in rows[5:]]))[0, 1]))
