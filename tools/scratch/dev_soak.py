"""Soak: many closed-loop ticks, all gaits, checking statuses / finiteness / feasibility every tick."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp
from scenario import Scenario
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
N = int(sys.argv[3]) if len(sys.argv) > 3 else 16
sc = Scenario(B, n_steps=N, gaits=["trot", "pace", "bound", "walk"], seed=777)
eng = mpcqp.Engine(batch=B, n_steps=N)
mu = eng.params.mu
worst = dict(status2=0, status0=0, max_sweeps=0, max_iters=0, max_ms=0.0)
t0 = time.perf_counter()
for t in range(T):
    xref, fsteps = sc.inputs()
    s0 = time.perf_counter(); eng.run(t, xref, fsteps); x = eng.solution(); ms = (time.perf_counter() - s0) * 1e3
    info = eng.info(with_y=False)
    f = x[:, 12 * N:].reshape(B, N, 4, 3)
    viol = np.maximum.reduce([np.abs(f[..., 0]) - mu * f[..., 2], np.abs(f[..., 1]) - mu * f[..., 2], -f[..., 2], f[..., 2] - 25.0]).max(axis=(1, 2))
    bad = (~np.isfinite(x).all(axis=1)) | (viol > 1e-7)
    if bad.any():
        b = np.flatnonzero(bad)
        print("tick %d: %d bad robots; first %s status %s sweeps %s iters %s viol %s gait %s" % (t, len(b), b[:5], info["status"][b[:5]], info["sweeps"][b[:5]], info["iters"][b[:5]], viol[b[:5]], [sc.kinds[i] for i in b[:5]]))
        worst.setdefault("bad", 0); worst["bad"] += len(b)
    worst["status2"] += int((info["status"] == 2).sum()); worst["status0"] += int((info["status"] == 0).sum())
    worst["max_sweeps"] = max(worst["max_sweeps"], int(info["sweeps"].max())); worst["max_iters"] = max(worst["max_iters"], int(info["iters"].max()))
    if t > 5: worst["max_ms"] = max(worst["max_ms"], ms)
    sc.advance(x[:, :12] + xref[:, :, 1])
print("soak B=%d N=%d ticks=%d: %s  total %.1f s" % (B, N, T, worst, time.perf_counter() - t0))
