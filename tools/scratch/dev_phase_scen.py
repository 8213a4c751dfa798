"""Per-robot phase cycles in the device closed loop (profiling build): planner (load), decode, sweeps, outputs + integration."""
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
    eng.scenario_run(3)
    eng.synchronize()
    lib.mpcqp_debug_profile(buf)
    v = np.array(buf[:], dtype=np.float64)[16:]
    ns, ni = max(v[0], 1), max(v[12], 1)
    print("/".join(gaits), "per pair of robots (cycles): planner/load %.0f decode %.0f sweeps %.0f finish %.0f | per warp sweep %.0f (assemble %.0f, feet+guard incl. core %.0f)" % (
        v[13] / ni, v[14] / ni, v[15] / ni, v[16] / ni, (v[1] + v[11]) / ns, v[1] / ns, v[11] / ns))
    eng.close()
