import sys, glob
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import mpcqp
for path in sorted(glob.glob("/root/repo/tests/golden/solve_*.npz")):
    g = np.load(path)
    n = g["x"].shape[1] // 24
    eng = mpcqp.Engine(batch=1, n_steps=n, mode=13, max_sweeps=0)
    out = []
    for t in range(len(g["k"])):
        eng.run(g["k"][t], g["xref"][t][None], g["fsteps"][t][None])
        x, info = eng.solution()[0], eng.info()
        err = np.abs(x[12 * n:] - g["x"][t][12 * n:]).max()
        out.append("%d/%d/%d/%.0e" % (info["status"][0], info["iters"][0], info["sweeps"][0], err))
    print(path.split("/")[-1], " ".join(out))
    eng.close()
