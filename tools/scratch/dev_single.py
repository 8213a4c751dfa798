"""BASELINE configs[0] through the drop-in: ONE Solo robot, trot, N = 16, 1000 closed-loop ticks, one QP per tick through
MPC_Wrapper.solve + get_latest_result (the two calls processing.py:142-145 makes).  Wall time per tick, p50 / p99."""
import sys, time, types
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
from MPC_Wrapper import MPC_Wrapper
from scenario import Scenario
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
sc = Scenario(1, gaits="trot", seed=20260)
w = MPC_Wrapper(0.02, 16, 20, 0.32)
planner = types.SimpleNamespace(xref=None, fsteps=None)
lat = np.zeros(T)
for t in range(T):
    xref, fsteps = sc.inputs()
    planner.xref, planner.fsteps = xref[0], fsteps[0]
    t0 = time.perf_counter()
    w.solve(20 * t, planner)
    f = w.get_latest_result()
    lat[t] = time.perf_counter() - t0
    sc.advance((w.mpc.x[:12] + xref[0, :, 1])[None])
s = lat[20:] * 1e6
print("single robot through MPC_Wrapper, %d ticks: p50 %.1f us  p99 %.1f us  max %.1f us  mean %.1f us  (%.0f solves/s)" % (
    T, np.percentile(s, 50), np.percentile(s, 99), s.max(), s.mean(), 1e6 / s.mean()))
