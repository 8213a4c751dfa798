import sys
import numpy as np
sys.path.insert(0, "/root/repo/tools/scratch")
from exp_n64 import *
from exp_stagec import testset
def trace(H, g, n, max_sweeps=40):
    sigs = np.full(n, 4); seen = set(); out = []
    for s in range(max_sweeps):
        f, grad, oks, ns = sweep(H, g, sigs)
        out.append(int((~oks).sum()))
        if oks.all(): return True, out
        seen.add(sigs.tobytes())
        if ns.tobytes() in seen: return False, out
        sigs = ns
    return False, out
ts = [t for t in testset() if t[0] == 64]
p64 = km.ModelParams(n_steps=64)
xref, fsteps = instances(64)
for b in range(8): ts.append((64, p64, xref[b], fsteps[b]))
for (N, p, xr, fs) in ts:
    H, g, idx, c0 = condensed(p, xr, fs, True)
    ok, tr = trace(H, g, len(idx))
    print("ok" if ok else "CYCLE", len(tr), tr)
