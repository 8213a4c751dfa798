"""Two measurements behind BASELINE.md's notes on the CPU arm (run in the build container, needs /root/reference):

1. the reference's own build half: the UNMODIFIED /root/reference/MPC.py (construct_gait, update_ML, update_NK, warm-start shift,
   retrieve_result) timed per tick with a null solver in place of `osqp`, next to the plain-C restatement the bench's CPU arm uses;
2. why the restated OSQP needs ~1100 iterations per solve at eps 1e-8 where SURVEY.md 7.3's scratch solver needed 50-75: OSQP's
   cost scaling treats ||q||_inf = 0 as 1 (the QP has q = 0, MPC.py:286-288), which pins the cost scale c at 1; the scratch solver
   scaled the cost by 1 / mean column norm of P (c ~ 6e3).  Same tick, same algorithm, the two rules side by side.
usage: python tools/baseline_experiments.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-tsid_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(1, "/root/reference")
np.int = int
import scipy.sparse, scipy.sparse.csc      # noqa: E401,F401
import MPC as RefMPC                        # /root/reference/MPC.py
from scenario import Scenario
from oracle import c_port, osqp_port

assert RefMPC.__file__.startswith("/root/reference")


class _NullSolver:
    """stands in for osqp.OSQP(): accepts the five calls MPC.py makes, returns the warm start"""
    def __init__(self): self.x = None
    def setup(self, P=None, q=None, A=None, l=None, u=None, **kw): self.x = np.zeros(P.shape[0])
    def update_settings(self, **kw): pass
    def update(self, **kw): pass
    def warm_start(self, x=None, **kw): self.x = np.array(x)
    def solve(self):
        class R: pass
        r = R(); r.x = self.x; return r


def build_half(N, ticks=200):
    sc = Scenario(1, n_steps=N, gaits="trot", seed=3)
    mpc = RefMPC.MPC(0.02, N, 0.32)
    mpc.prob = _NullSolver()
    m = c_port.MPC(n_steps=N)
    inputs = []
    for t in range(ticks):
        xr, fs = sc.inputs(); inputs.append((xr[0].copy(), fs[0].copy()))
        sc.advance(xr[:, :, 1])
    t0 = time.perf_counter()
    for t, (xr, fs) in enumerate(inputs):
        mpc.run(t, xr, fs.copy())
    ref_ms = (time.perf_counter() - t0) / ticks * 1e3
    t0 = time.perf_counter()
    for t, (xr, fs) in enumerate(inputs):
        m.build(xr, fs, first_tick=(t == 0))
    c_ms = (time.perf_counter() - t0) / ticks * 1e3
    m.close()
    return ref_ms, c_ms


def cost_scaling_rules():
    g = np.load(os.path.join(ROOT, "tests", "golden", "solve_trot.npz"))
    from oracle import mpc_build
    import scipy.sparse as sp
    rows = []
    for t in (3, 8, 14):
        Pd, A, l, u, _ = mpc_build.build_qp(g["xref"][t], g["fsteps"][t])
        P, q = sp.diags(Pd).tocsc(), np.zeros(len(Pd))
        out = {}
        for rule in ("osqp", "ignore-zero-q"):
            if rule == "ignore-zero-q":
                keep = osqp_port._limit_scaling
                osqp_port._limit_scaling = lambda v: np.minimum(np.where(v < osqp_port.MIN_SCALING, (osqp_port.MIN_SCALING if v.size == 1 else 1.0), v), osqp_port.MAX_SCALING) if v.size == 1 else keep(v)
            s = osqp_port.OSQP()
            s.setup(P=P, q=q, A=A, l=l, u=u, eps_abs=1e-8, eps_rel=1e-8)
            s.warm_start(x=g["warm_x"][t])
            r = s.solve()
            out[rule] = (r.info.iter, float(s.c), float(np.abs(r.x - g["x"][t]).max()))
            if rule == "ignore-zero-q":
                osqp_port._limit_scaling = keep
        rows.append((t, out))
    return rows


if __name__ == "__main__":
    for N in (16, 32, 64):
        ref_ms, c_ms = build_half(N)
        print("build half, N = %d: reference MPC.py %.3f ms per tick (one core, numpy %s); plain-C restatement %.4f ms" % (N, ref_ms, np.__version__, c_ms))
    for t, out in cost_scaling_rules():
        print("solve_trot tick %d: " % t + "; ".join("%s rule: %d iterations, cost scale c = %.3g, |x - x*| = %.1e" % (k, *v) for k, v in out.items()))
