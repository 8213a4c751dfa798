"""Stage-wise (Riccati) active-set stage against the dense one: same closed loop, tick by tick; then timing."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp
from scenario import Scenario

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 12
N = int(sys.argv[3]) if len(sys.argv) > 3 else 16
gaits = ["trot", "pace", "bound", "walk", "static"]
kw = dict(n_steps=N) if N != 16 else {}
sc = Scenario(B, gaits=gaits, seed=5, **kw)
dense = mpcqp.Engine(batch=B, n_steps=N, mode=3)
ric = mpcqp.Engine(batch=B, n_steps=N, mode=7)
for t in range(T):
    xref, fsteps = sc.inputs()
    dense.run(t, xref, fsteps); xd = dense.solution(); idn = dense.info()
    ric.run(t, xref, fsteps); xr = ric.solution(); irc = ric.info()
    df = np.abs(xd[:, 12 * N:] - xr[:, 12 * N:]).max(axis=1)
    dx = np.abs(xd[:, :12 * N] - xr[:, :12 * N]).max()
    print("tick %2d  max|df| %.2e  max|dx| %.2e  status dense %s ric %s  sweeps dense %.2f ric %.2f  admm-iters dense %d ric %d  obj rel %.1e  masks equal %s %s" % (
        t, df.max(), dx, np.bincount(idn["status"], minlength=4), np.bincount(irc["status"], minlength=4), idn["sweeps"].mean(), irc["sweeps"].mean(),
        idn["iters"].sum(), irc["iters"].sum(), np.abs(idn["obj"] - irc["obj"]).max() / np.abs(idn["obj"]).max(),
        np.array_equal(idn["contact"], irc["contact"]), np.array_equal(idn["active"], irc["active"])))
    sc.advance(xd[:, :12 * N].reshape(B, N, 12)[:, 0, :] + xref[:, :, 1] if False else xd[:, :12] + xref[:, :, 1])
dense.close(); ric.close()

# timing: trot, device-resident replay
import torch
Bt = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
for mode in (3, 7):
    sc = Scenario(Bt, gaits="trot", seed=20260, **kw)
    eng = mpcqp.Engine(batch=Bt, n_steps=N, mode=mode)
    Tt = 40
    xs, fs = [], []
    for t in range(Tt):
        xref, fsteps = sc.inputs()
        xs.append(torch.from_numpy(xref).cuda()); fs.append(torch.from_numpy(fsteps).cuda())
        eng.run_device(t, xs[-1].data_ptr(), fs[-1].data_ptr())
        x = eng.solution()
        sc.advance(x[:, :12] + xref[:, :, 1])
    info = eng.info(with_y=False)
    eng.reset_warm_start()
    for t in range(25):
        eng.run_device(t, xs[t].data_ptr(), fs[t].data_ptr())
    eng.synchronize()
    stream = torch.cuda.ExternalStream(eng.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(25, Tt):
        eng.run_device(t, xs[t].data_ptr(), fs[t].data_ptr())
    e1.record(stream)
    eng.synchronize()
    ms = e0.elapsed_time(e1) / (Tt - 25)
    print("mode %d  B %d N %d: %.3f ms per tick -> %.2f M solves/s   (sweeps %.3f, fallback %d)" % (
        mode, Bt, N, ms, Bt / ms * 1e-3, info["sweeps"].mean(), (info["iters"] > 0).sum()))
    eng.close()
