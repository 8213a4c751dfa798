"""Fallback statistics per tick for a gait mix (closed loop, host inputs)."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp
from scenario import Scenario
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 40
gaits = sys.argv[3].split(",") if len(sys.argv) > 3 else ["trot", "pace", "bound", "walk"]
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 7
kw = {}
if len(sys.argv) > 5: kw["max_sweeps"] = int(sys.argv[5])
sc = Scenario(B, gaits=gaits, seed=99)
eng = mpcqp.Engine(batch=B, mode=mode, **kw)
import torch
for t in range(T):
    xref, fsteps = sc.inputs()
    dx, df = torch.from_numpy(xref).cuda(), torch.from_numpy(fsteps).cuda()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.run_device(t, dx.data_ptr(), df.data_ptr()); eng.synchronize()
    dt = time.perf_counter() - t0
    x = eng.solution(); info = eng.info(with_y=False)
    fb = info["iters"] > 0
    if t < 6 or t % 5 == 0:
        print("tick %2d  %.3f ms  fallback %5d (%.2f%%)  admm iters mean over fallbacks %.0f max %d  sweeps mean %.2f  unsolved %d" % (
            t, dt * 1e3, fb.sum(), 100.0 * fb.mean(), info["iters"][fb].mean() if fb.any() else 0, info["iters"].max(), info["sweeps"].mean(), (info["status"] != 1).sum()))
    sc.advance(x[:, :12] + xref[:, :, 1])
fb = info["iters"] > 0
print("sweeps histogram, solved in stage A:", np.bincount(info["sweeps"][~fb], minlength=8)[:10])
print("sweeps histogram, fallback robots (stage A sweeps + polish attempts):", np.bincount(info["sweeps"][fb], minlength=12)[:14])
g = np.array(sc.kinds)
if g is not None:
    for name in sorted(set(g)):
        m = g == name
        print("  %-6s robots %5d  fallback %.2f%%  sweeps mean %.2f" % (name, m.sum(), 100.0 * fb[m].mean(), info["sweeps"][m].mean()))
