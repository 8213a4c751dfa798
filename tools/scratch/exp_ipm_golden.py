import sys, glob
import numpy as np
sys.path.insert(0, "/root/repo/tools/scratch")
from exp_n64 import *
from exp_ipm2 import ipm2
for path in sorted(glob.glob("/root/repo/tests/golden/solve_*.npz")):
    g = np.load(path)
    n = g["x"].shape[1] // 24
    p = km.ModelParams(n_steps=n)
    out = []
    for t in range(len(g["k"])):
        H, gg, idx, c0 = condensed(p, g["xref"][t], g["fsteps"][t], g["k"][t] == 0)
        if len(idx) == 0:
            out.append("empty"); continue
        ok, ns, nsw = ipm2(H, gg, split=True)
        out.append("%s/%d/%d" % ("ok" if ok else "FAIL", ns, nsw))
    print(path.split("/")[-1], " ".join(out))
