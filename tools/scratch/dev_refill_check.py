"""Warp-level sweeps per robot pair (profiling build): does a finished half really take a new robot while its neighbour sweeps on?"""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp
mpcqp._LIB_PATH = os.path.join(os.path.dirname(mpcqp._LIB_PATH), "libmpcqp_prof.so")
from scenario import Scenario
B = 16384
lib = mpcqp.load()
buf = (ctypes.c_ulonglong * 64)()
for gaits in (["trot"], ["trot", "pace", "bound", "walk"]):
    sc = Scenario(B, gaits=gaits, seed=4242, noise_kind="hash")
    eng = mpcqp.Engine(batch=B)
    eng.set_overlap(1)
    eng.scenario_init(sc)
    eng.scenario_run(25)
    eng.synchronize()
    lib.mpcqp_debug_profile(buf)
    for t in range(3):
        eng.scenario_run(1)
        info = eng.info(with_y=False)
        lib.mpcqp_debug_profile(buf)
        v = np.array(buf[:], dtype=np.float64)[16:]
        s = info["sweeps"]
        print("/".join(gaits), "mean sweeps %.3f  E[max of index pairs] %.3f  warp sweeps per pair of robots %.3f  fetches by half A %d (of %d)  cycles per warp sweep: %s finish %.0f" % (
            s.mean(), np.maximum(s[0::2], s[1::2]).mean(), v[0] / (B / 2), v[12], B // 2, " ".join("%.0f" % (v[1 + i] / max(v[0], 1)) for i in range(11)), v[16] / max(v[12], 1)), flush=True)
    eng.close()
