"""Per-phase cycle breakdown of the solve kernel (needs libmpcqp_prof.so built with -DMPCQP_PROFILE)."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200")
import mpcqp
mpcqp._LIB_PATH = os.path.join(os.path.dirname(mpcqp._LIB_PATH), "libmpcqp_prof.so")
from scenario import Scenario
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
eng = mpcqp.Engine(batch=B)
lib = mpcqp.load()
sc = Scenario(B, gaits="trot", seed=20260)
buf = (ctypes.c_ulonglong * 16)()
names = ["assemble", "factor+invert", "grad(H f)", "rhs+sync", "tri-solve", "back+guard", "-", "build", "solve-stage", "finish"]
for t in range(T):
    xr, fs = sc.inputs()
    eng.run(t, xr, fs); x = eng.solution(); info = eng.info(with_y=False)
    lib.mpcqp_debug_profile(buf)
    v = np.array(buf[:], dtype=np.float64)
    ns = max(v[15], 1)
    print("tick %d sweeps/inst %.2f fallback %d | per sweep cycles: %s | per instance A: %.0f (n=%d)  B: %.0f (n=%d)" % (
        t, info["sweeps"].mean(), (info["iters"] > 0).sum(),
        "  ".join("%s %.0f" % (n, v[i] / ns) for i, n in enumerate(names)), v[12] / max(v[10], 1), v[10], v[13] / max(v[11], 1), v[11]))
    sc.advance(x[:, :12] + xr[:, :, 1])
