import sys
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp
from scenario import Scenario
n = 64
for ms in (40, 200):
    sc = Scenario(16, n_steps=n, gaits=["trot"], seed=64)
    eng = mpcqp.Engine(batch=16, n_steps=n, max_sweeps=ms)
    for t in range(5):
        xref, fsteps = sc.inputs()
        eng.run(t, xref, fsteps)
        x, info = eng.solution(), eng.info()
        print("max_sweeps %d tick %d status %s sweeps %s" % (ms, t, info["status"], info["sweeps"]))
        sc.advance(x[:, :12] + xref[:, :, 1])
    eng.close()
