"""A/B of several builds in one box visit with the closed-loop inputs generated once:
python tools/scratch/dev_ab4.py libA.so libB.so ...   (names under mpc-tsid_b200/; the first one generates the inputs)"""
import os, subprocess, sys
gen = r"""
import sys, numpy as np
sys.path.insert(0, '/root/repo/mpc-tsid_b200'); sys.path.insert(0, '/root/repo')
import mpcqp
from scenario import Scenario
for B, gaits, tag in ((4096, 'trot', 'trot4k'), (8192, 'trot', 'trot8k'), (8192, ['trot', 'pace', 'bound', 'walk'], 'mixed8k')):
    T = 55; eng = mpcqp.Engine(batch=B); sc = Scenario(B, gaits=gaits, seed=20260)
    hx = np.empty((T, B, 12, 17)); hf = np.empty((T, B, 20, 13))
    for t in range(T):
        xr, fs = sc.inputs(); hx[t], hf[t] = xr, fs
        eng.run(t, xr, fs); x = eng.solution(); sc.advance(x[:, :12] + xr[:, :, 1])
    np.save('/tmp/ab_%s_x.npy' % tag, hx); np.save('/tmp/ab_%s_f.npy' % tag, hf); eng.close()
"""
rep = r"""
import sys, numpy as np
sys.path.insert(0, '/root/repo/mpc-tsid_b200'); sys.path.insert(0, '/root/repo')
import torch, mpcqp
out = []
for B, tag in ((4096, 'trot4k'), (8192, 'trot8k'), (8192, 'mixed8k')):
    hx = np.load('/tmp/ab_%s_x.npy' % tag); hf = np.load('/tmp/ab_%s_f.npy' % tag); T = hx.shape[0]
    dx, df = torch.from_numpy(hx).cuda(), torch.from_numpy(hf).cuda()
    eng = mpcqp.Engine(batch=B); stream = torch.cuda.ExternalStream(eng.stream)
    best = 1e9
    for p in range(3):
        eng.reset_warm_start()
        for t in range(25): eng.run_device(t, dx[t].data_ptr(), df[t].data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.synchronize(); e0.record(stream)
        for t in range(25, T): eng.run_device(t, dx[t].data_ptr(), df[t].data_ptr())
        e1.record(stream); eng.synchronize()
        best = min(best, e0.elapsed_time(e1) / (T - 25))
    info = eng.info(with_y=False)
    out.append('%s %.4f ms %.2f M/s (sweeps %.3f, unsolved %d)' % (tag, best, B / best / 1e3, info['sweeps'].mean(), (info['status'] != 1).sum()))
    eng.close()
print(' | '.join(out))
"""
libs = sys.argv[1:]
env = dict(os.environ, MPCQP_LIB="/root/repo/mpc-tsid_b200/" + libs[0])
r = subprocess.run([sys.executable, "-c", gen], env=env, capture_output=True, text=True)
if r.returncode: print(r.stderr[-2000:]); sys.exit(1)
for it in range(2):
    for lib in libs:
        env = dict(os.environ, MPCQP_LIB="/root/repo/mpc-tsid_b200/" + lib)
        r = subprocess.run([sys.executable, "-c", rep], env=env, capture_output=True, text=True)
        print("== %-22s %s" % (lib, r.stdout.strip() or r.stderr[-800:]), flush=True)
