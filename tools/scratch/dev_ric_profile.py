"""Per-phase cycle breakdown of the stage-wise kernel (needs libmpcqp_prof.so built with -DMPCQP_PROFILE)."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200")
import mpcqp
mpcqp._LIB_PATH = os.path.join(os.path.dirname(mpcqp._LIB_PATH), "libmpcqp_prof.so")
from scenario import Scenario
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 24
N = int(sys.argv[3]) if len(sys.argv) > 3 else 16
eng = mpcqp.Engine(batch=B, n_steps=N, mode=7)
lib = mpcqp.load()
sc = Scenario(B, gaits="trot", seed=20260, **({} if N == 16 else dict(n_steps=N)))
buf = (ctypes.c_ulonglong * 64)()
sw = ["E+beta+term", "chol Pvv", "T,Y,av", "G", "chol G", "X,Y2,U,cv", "Pt,KpT,G2", "gain+P_k", "forward", "costate", "feet+guard"]
for t in range(T):
    xr, fs = sc.inputs()
    eng.run(t, xr, fs); x = eng.solution(); info = eng.info(with_y=False)
    lib.mpcqp_debug_profile(buf)
    v = np.array(buf[:], dtype=np.float64)[16:]
    ns, ni = max(v[0], 1), max(v[12], 1)
    if t >= T - 3:
        print("tick %d sweeps/inst %.2f | per sweep: %s | per stage backward %.0f | per instance: load %.0f decode %.0f sweeps %.0f finish %.0f" % (
            t, info["sweeps"].mean(), "  ".join("%s %.0f" % (n, v[1 + i] / ns) for i, n in enumerate(sw)), sum(v[2:9]) / ns / N,
            v[13] / ni, v[14] / ni, v[15] / ni, v[16] / ni) + "  stage-top loads %.0f per stage" % (v[17] / ns / N))
    sc.advance(x[:, :12] + xr[:, :, 1])
