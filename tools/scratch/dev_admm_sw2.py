"""Pure ADMM stages against each other: dense (mode 2) vs stage-wise (mode 15 with max_sweeps = 0)."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp, os
if os.environ.get("MPCQP_LIB"): mpcqp._LIB_PATH = os.environ["MPCQP_LIB"]
from scenario import Scenario
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sc = Scenario(B, gaits=["trot", "walk"], seed=5)
dense = mpcqp.Engine(batch=B, mode=2)
sw = mpcqp.Engine(batch=B, mode=15, max_sweeps=0)
for t in range(4):
    xref, fsteps = sc.inputs()
    dense.run(t, xref, fsteps); xd = dense.solution(); idn = dense.info()
    sw.run(t, xref, fsteps); xs = sw.solution(); isw = sw.info()
    print("tick %d  max|df| %.2e nan %d status dense %s sw %s\n   iters dense %s\n   iters sw    %s\n   sweeps dense %s sw %s" % (
        t, np.nanmax(np.abs(xd - xs)), np.isnan(xs).sum(), idn["status"], isw["status"], idn["iters"], isw["iters"], idn["sweeps"], isw["sweeps"]))
    sc.advance(xd[:, :12] + xref[:, :, 1])
