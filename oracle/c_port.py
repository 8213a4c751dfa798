"""ctypes binding of oracle/mpc_osqp.c -- TEST INFRASTRUCTURE ONLY (see that file's header).

The plain-C restatement of MPC.py's per-tick path (build + OSQP algorithm + extraction).  Used by
tests/ (against the numpy/scipy restatement, the golden fixtures and the CUDA engine), by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.  Never imported by
the product path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmpc_oracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        build()
    lib = C.CDLL(_LIB)
    vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p
    lib.mpc_oracle_create.restype = vp
    lib.mpc_oracle_create.argtypes = [C.c_int, C.c_double, C.c_double]
    lib.mpc_oracle_destroy.argtypes = [vp]
    lib.mpc_oracle_nnz.argtypes = [vp]
    lib.mpc_oracle_build.argtypes = [vp, dp, dp, C.c_int, dp, dp, dp, ip, ip]
    lib.mpc_oracle_run.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp, dp]
    lib.mpc_oracle_replay_mt.restype = C.c_double
    lib.mpc_oracle_replay_mt.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, dp, dp, dp, dp]
    lib.mpc_oracle_solve_qp.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, dp, dp, dp, C.c_double, dp, dp]
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class MPC:
    """One robot, tick after tick: the C restatement of MPC.run (MPC.py:460-514) at eps_abs = eps_rel = eps."""

    def __init__(self, n_steps=16, dt=0.02, eps=1e-8):
        self.lib = load()
        self.N = int(n_steps)
        self._h = self.lib.mpc_oracle_create(self.N, float(dt), float(eps))
        if not self._h:
            raise ValueError("bad horizon")
        self.nnz = self.lib.mpc_oracle_nnz(self._h)

    def close(self):
        if self._h:
            self.lib.mpc_oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build(self, xref, fsteps, first_tick=False):
        """(Ap, Ai, Ax, l, u) of the reference's QP for these inputs (MPC.py:98-234, 316-378)."""
        N = self.N
        xref, fsteps = _f64(xref), _f64(fsteps)
        Ax, l, u = np.empty(self.nnz), np.empty(44 * N), np.empty(44 * N)
        Ap, Ai = np.empty(24 * N + 1, dtype=np.int32), np.empty(self.nnz, dtype=np.int32)
        self.lib.mpc_oracle_build(self._h, _p(xref), _p(fsteps), int(first_tick), _p(Ax), _p(l), _p(u), _p(Ap), _p(Ai))
        return Ap, Ai, Ax, l, u

    def run(self, first_tick, xref, fsteps):
        """-> dict(x (24N), y (44N), f (12), iter, status, rho_updates)."""
        N = self.N
        xref, fsteps = _f64(xref), _f64(fsteps)
        x, y, f, info = np.empty(24 * N), np.empty(44 * N), np.empty(12), np.empty(4)
        rc = self.lib.mpc_oracle_run(self._h, int(bool(first_tick)), _p(xref), _p(fsteps), _p(x), _p(y), _p(f), _p(info))
        if rc:
            raise RuntimeError("mpc_oracle_run failed (%d)" % rc)
        return dict(x=x, y=y, f=f, iter=int(info[0]), status=int(info[1]), rho_updates=int(info[2]), factorizations=int(info[3]))


def set_cost_scaling_variant(ignore_zero_q):
    """False (default): OSQP's published rule (a zero q counts as 1: cost scale 1 for the reference's QP).  True: leave a zero q
    out (cost scale 1 / mean column norm of P).  Process-wide."""
    load().mpc_oracle_set_cost_scaling_variant(1 if ignore_zero_q else 0)


def replay_mt(xref, fsteps, warm, n_steps=16, dt=0.02, eps=1e-8):
    """`threads` robots, one per host thread, each replaying its own recorded input sequence.
    xref (threads, T, 12, N+1), fsteps (threads, T, 20, 13).  -> (seconds of the slowest thread over its
    timed ticks, forces (threads, T, 12), mean iterations per timed solve)."""
    lib = load()
    xref, fsteps = _f64(xref), _f64(fsteps)
    threads, T = xref.shape[0], xref.shape[1]
    out = np.zeros((threads, T, 12))
    it = C.c_double(0.0)
    sec = lib.mpc_oracle_replay_mt(threads, int(n_steps), float(dt), float(eps), T, int(warm), _p(xref), _p(fsteps), _p(out), C.byref(it))
    if sec < 0:
        raise RuntimeError("mpc_oracle_replay_mt failed")
    return sec, out, it.value


def solve_qp(Pdiag, q, A, l, u, eps=1e-8):
    """Generic QP with diagonal P through the C restatement of the OSQP algorithm -> (x, y, iterations)."""
    import scipy.sparse as sp
    lib = load()
    A = sp.csc_matrix(A).astype(np.float64)
    A.sort_indices()
    m, n = A.shape
    Ap, Ai = np.ascontiguousarray(A.indptr, dtype=np.int32), np.ascontiguousarray(A.indices, dtype=np.int32)
    Ax, Pd, q, l, u = _f64(A.data), _f64(Pdiag), _f64(q), _f64(l), _f64(u)
    x, y = np.empty(n), np.empty(m)
    it = lib.mpc_oracle_solve_qp(n, m, _p(Ap), _p(Ai), _p(Ax), _p(Pd), _p(q), _p(l), _p(u), float(eps), _p(x), _p(y))
    return x, y, it
