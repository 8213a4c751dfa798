import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import torch, mpcqp, MPC, MPC_Wrapper
from scenario import Scenario
B, N, T = 4096, 16, 60
sc = Scenario(B, gaits="trot", seed=1)
eng = mpcqp.Engine(batch=B)
h_x = torch.empty((T, B, 12, N + 1), dtype=torch.float64, pin_memory=True); h_f = torch.empty((T, B, 20, 13), dtype=torch.float64, pin_memory=True)
hx, hf = h_x.numpy(), h_f.numpy()
for t in range(T):
    xr, fs = sc.inputs(); hx[t], hf[t] = xr, fs
    eng.run(t, hx[t], hf[t]); x = eng.solution(); sc.advance(x[:, :12] + xr[:, :, 1])
class P: pass
def timeit(name, fn, n=30):
    fn(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    print("%-40s %.3f ms" % (name, (time.perf_counter() - t0) / n * 1e3), flush=True)
out = torch.empty((B, 12), dtype=torch.float64, pin_memory=True).numpy()
def a(i=0):
    eng.run(20 + i, hx[20 + i], hf[20 + i]); eng.forces(out=out)
timeit("engine run + forces(pinned)", a)
def a2(i=0):
    eng.run(20 + i, hx[20 + i], hf[20 + i]); eng.forces()
timeit("engine run + forces(pageable)", a2)
def a3(i=0):
    eng.run(20 + i, hx[20 + i], hf[20 + i]); eng.step_result()
timeit("engine run + step_result(pageable)", a3)
m = MPC.MPC(0.02, 16, 0.32)
def b(i=0):
    m.run(20 + i, hx[20 + i], hf[20 + i]); return m.f_applied
timeit("MPC.run + f_applied", b)
def b2(i=0):
    m.run(20 + i, hx[20 + i], hf[20 + i]); m._engine.synchronize()
timeit("MPC.run + sync only", b2)
w = MPC_Wrapper.MPC_Wrapper(0.02, 16, 20, 0.32)
pl = P()
def c(i=0):
    pl.xref, pl.fsteps = hx[20 + i], hf[20 + i]; w.solve(20 * (20 + i), pl); return w.get_latest_result()
timeit("wrapper solve + get_latest_result", c)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(20): c(i)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
