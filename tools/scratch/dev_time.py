"""Device-resident timing of one engine configuration: python dev_time.py <lib.so> <mode> [B] [N]"""
import sys, os
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp
if sys.argv[1] != "-":
    mpcqp._LIB_PATH = os.path.abspath(sys.argv[1])
from scenario import Scenario
import torch
mode = int(sys.argv[2]); Bt = int(sys.argv[3]) if len(sys.argv) > 3 else 4096; N = int(sys.argv[4]) if len(sys.argv) > 4 else 16
kw = dict(n_steps=N) if N != 16 else {}
if os.environ.get("VREF_SCALE"):
    rng = np.random.default_rng(1); sc_ = float(os.environ["VREF_SCALE"])
    v = np.zeros((Bt, 6)); v[:, 0] = rng.uniform(-0.5, 1.0, Bt) * sc_; v[:, 1] = rng.uniform(-0.3, 0.3, Bt) * sc_; v[:, 5] = rng.uniform(-0.4, 0.4, Bt) * sc_
    kw["v_ref"] = v
if os.environ.get("NOISE") == "0":
    kw["noise"] = (0.0, 0.0, 0.0, 0.0)
sc = Scenario(Bt, gaits=os.environ.get("GAITS", "trot").split(","), seed=20260, **kw)
eng = mpcqp.Engine(batch=Bt, n_steps=N, mode=mode, **({"max_iter": int(os.environ["MAX_ITER"])} if os.environ.get("MAX_ITER") else {}))
Tt = 40
xs, fs = [], []
for t in range(Tt):
    xref, fsteps = sc.inputs()
    xs.append(torch.from_numpy(xref).cuda()); fs.append(torch.from_numpy(fsteps).cuda())
    eng.run_device(t, xs[-1].data_ptr(), fs[-1].data_ptr())
    x = eng.solution()
    sc.advance(x[:, :12] + xref[:, :, 1])
info = eng.info(with_y=False)
best = 1e9
for rep in range(3):
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
    best = min(best, e0.elapsed_time(e1) / (Tt - 25))
print("%s mode %d  B %d N %d: %.3f ms per tick -> %.2f M solves/s   (sweeps %.3f, fallback %d, unsolved %d)" % (
    os.path.basename(sys.argv[1]), mode, Bt, N, best, Bt / best * 1e-3, info["sweeps"].mean(), (info["iters"] > 0).sum(), (info["status"] != 1).sum()))
