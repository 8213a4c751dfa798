"""Small all-mode smoke (compute-sanitizer is closed on this pool; bounds are checked by the parity suite): 6 robots, mixed gaits, 4 ticks,
default mode and the dense + ADMM modes, plus 3 ticks of the device-resident closed loop.
python tools/scratch/dev_sanitize.py"""
import sys
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp
from scenario import Scenario
for mode, n in ((13, 16), (-13, 16), (7, 16), (3, 16), (2, 16), (13, 24), (-13, 64)):
    B = 6
    sc = Scenario(B, n_steps=n, gaits=["trot", "walk", "pace"], seed=5)
    eng = mpcqp.Engine(batch=B, n_steps=n, mode=abs(mode), **({'max_sweeps': 0} if mode < 0 else {}))       # mode < 0: every robot through ipm_kernel
    for t in range(4):
        xref, fsteps = sc.inputs()
        eng.run(t, xref, fsteps)
        x = eng.solution(); info = eng.info()
        assert np.isfinite(x).all() and (info["status"] == 1).all(), (mode, n, t, info["status"])
        sc.advance(x[:, :12] + xref[:, :, 1])
    eng.close()
    print("mode %d N %d ok" % (mode, n), flush=True)
sc = Scenario(8, gaits=["trot", "bound"], seed=9, noise_kind="hash")
eng = mpcqp.Engine(batch=8)
eng.scenario_init(sc)
eng.scenario_run(3)
st = eng.scenario_state()
assert np.isfinite(st["state"]).all()
print("device closed loop ok")
