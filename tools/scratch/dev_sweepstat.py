"""Second sweeps in steady trot: rate by gait phase, persistence from tick to tick, which step of the horizon changed."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp
from scenario import Scenario
B = 4096
sc = Scenario(B, gaits="trot", seed=20260)
eng = mpcqp.Engine(batch=B)
hist = np.zeros((16, 4)); prev = None; both = 0; tot2 = 0; cnt = 0
for t in range(60):
    xref, fsteps = sc.inputs()
    eng.run(t, xref, fsteps); x = eng.solution(); info = eng.info(with_y=False)
    if t >= 25:
        ph = (t + sc.phase) % 16
        s2 = info["sweeps"] >= 2
        for p in range(16):
            hist[p, 0] += (ph == p).sum(); hist[p, 1] += (s2 & (ph == p)).sum()
        if prev is not None:
            both += (s2 & prev).sum(); tot2 += s2.sum(); cnt += B
        prev = s2
    sc.advance(x[:, :12] + xref[:, :, 1])
print("P(2+ sweeps | gait phase):", np.round(hist[:, 1] / hist[:, 0], 3))
print("P(2+) = %.4f, P(2+ | 2+ at previous tick) = %.3f" % (tot2 / cnt, both / max(tot2, 1)))
# first-row step count of fsteps as a predictor
print("fsteps[0,0] values:", np.unique(fsteps[:, 0, 0]))
# where along the horizon does the warm-start guess differ from the final active set, for robots that needed 2+ sweeps?
sc = Scenario(B, gaits="trot", seed=20260)
eng2 = mpcqp.Engine(batch=B)
prev_act = None; hk = np.zeros(16, int); nchg = []
for t in range(50):
    xref, fsteps = sc.inputs()
    eng2.run(t, xref, fsteps); x = eng2.solution(); info = eng2.info(with_y=False)
    act = info["active"] & info["contact"][..., None]
    if t >= 25 and prev_act is not None:
        guess = np.concatenate([prev_act[:, 1:], prev_act[:, :1]], axis=1) & info["contact"][..., None]
        diff = (guess != act).any(axis=(2, 3))                     # (B, N) steps whose active set changed
        two = info["sweeps"] >= 2
        for b in np.flatnonzero(two):
            ks = np.flatnonzero(diff[b])
            if len(ks):
                hk[ks.max()] += 1; nchg.append(len(ks))
    prev_act = info["active"]
    sc.advance(x[:, :12] + xref[:, :, 1])
print("highest changed step among 2+-sweep robots (histogram over k = 0..15):", hk)
print("mean number of changed steps:", np.mean(nchg))
