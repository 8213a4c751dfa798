"""What the bus gives: pinned host -> device copy rates at the sizes of one tick, the device time of one tick with
device-resident inputs and with host inputs, and the host-side cost of one run() call."""
import sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp
from scenario import Scenario
for mb in (1, 3.4, 6.7, 10.1, 15.2, 64, 256):
    n = int(mb * 1e6 / 8)
    h = torch.empty(n, dtype=torch.float64, pin_memory=True); d = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("H2D %6.1f MB: %.3f ms  %.1f GB/s" % (mb, ms, mb / ms))
B = 4096
sc = Scenario(B, gaits="trot", seed=20260)
eng = mpcqp.Engine(batch=B)
T = 60
hx = torch.empty((T, B, 12, 17), dtype=torch.float64, pin_memory=True); hf = torch.empty((T, B, 20, 13), dtype=torch.float64, pin_memory=True)
for t in range(T):
    xr, fs = sc.inputs(); hx[t], hf[t] = torch.from_numpy(xr), torch.from_numpy(fs)
    eng.run(t, xr, fs); x = eng.solution(); sc.advance(x[:, :12] + xr[:, :, 1])
dx, df = hx.cuda(), hf.cuda()
stream = torch.cuda.ExternalStream(eng.stream)
for name in ("device", "host"):
    eng.reset_warm_start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(T)]
    for t in range(T):
        ev[t][0].record(stream)
        if name == "device": eng.run_device(t, dx[t].data_ptr(), df[t].data_ptr())
        else: eng.run(t, hx[t].numpy(), hf[t].numpy())
        ev[t][1].record(stream)
        eng.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in ev[25:]])
    print("%s inputs: device time per tick p50 %.3f ms  min %.3f  max %.3f" % (name, np.median(ms), ms.min(), ms.max()))
# host-side cost of one run() call (no GPU wait)
t0 = time.perf_counter()
for t in range(25, T): eng.run(t, hx[t].numpy(), hf[t].numpy())
t1 = time.perf_counter(); eng.synchronize()
print("host time of run() per call: %.1f us" % ((t1 - t0) / (T - 25) * 1e6))
