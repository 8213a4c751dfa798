"""Drop-in for the reference's `MPC` class (MPC.py:10-82, 460-514) backed by libmpcqp.so on a B200.

Same constructor and call surface as the reference --
    mpc = MPC(dt, n_steps, T_gait)
    mpc.run(k, xref, fsteps)            # -> 0
    mpc.f_applied, mpc.x, mpc.x_robot, mpc.q_next, mpc.v_next, mpc.q_w, mpc.P, mpc.n_steps, mpc.h_ref
-- plus a batch dimension: pass xref (B, 12, N+1) and fsteps (B, 20, 13) and every result gains a
leading axis of size B.  `run(k, T_gait=..., fsteps=..., xref=...)` (the north-star's paraphrase of the
signature) is accepted by keyword.  Differences on purpose:
  * the caller's fsteps is NOT modified (the reference overwrites its NaNs, MPC.py:327);
  * `status` / `info` report per-instance solver status (the reference never looks, MPC.py:427);
  * there is no CPU path: construction fails without the CUDA extension or without a B200.
"""
import numpy as np

import mpcqp


class MPC:
    def __init__(self, dt, n_steps, T_gait, batch=None, device=0, **solver_options):
        self.dt = dt                                    # MPC.py:25
        self.n_steps = int(n_steps)                     # MPC.py:42
        self.T_gait = T_gait                            # MPC.py:45
        self._device = device
        self._opts = dict(solver_options)
        p = mpcqp.default_params(dt=float(dt), n_steps=self.n_steps, T_gait=float(T_gait), **self._opts)
        self.mass = p.mass                              # MPC.py:28
        self.gI = np.array(p.gI[:]).reshape(3, 3)       # MPC.py:35-37
        self.mu = p.mu                                  # MPC.py:39
        self.footholds = np.array(p.footholds[:]).reshape(3, 4)   # MPC.py:67-70
        self._w = np.concatenate([np.tile(np.array(p.w_state[:]), self.n_steps),
                                  np.full(12 * self.n_steps, p.w_force)])
        self.xref = np.zeros((12, 1 + self.n_steps))    # MPC.py:49
        self.x = np.zeros((12 * self.n_steps * 2,))     # MPC.py:52
        self.q = np.array([[0.0, 0.0, 0.2027682, 0.0, 0.0, 0.0]]).transpose()   # MPC.py:55
        self.q_w = self.q.copy()                        # MPC.py:58
        self.v = np.zeros((6, 1))                       # MPC.py:61
        self.h_ref = self.q[2, 0]                       # MPC.py:64
        self.f_applied = np.zeros((12,))
        self.x_robot = np.zeros((12, self.n_steps))
        self.q_next = np.zeros((6, 1))
        self.v_next = np.zeros((6, 1))
        self.status = None
        self._engine = None
        self._batched = None
        if batch is not None:
            self._make_engine(int(batch))

    # Logger.py:411-418 reads mpc.P.data: same diagonal, same order (MPC.py:236-288)
    @property
    def P(self):
        import scipy.sparse as sp
        return sp.diags(self._w).tocsc()

    @property
    def Q(self):
        return np.zeros(24 * self.n_steps)

    def _make_engine(self, batch):
        if self._engine is not None:
            self._engine.close()
        self._engine = mpcqp.Engine(batch=batch, n_steps=self.n_steps, device=self._device, dt=float(self.dt),
                                    T_gait=float(self.T_gait), **self._opts)
        self._qw = np.tile(self.q_w[:, 0], (batch, 1))

    def run(self, k, xref=None, fsteps=None, T_gait=None):
        if xref is None or fsteps is None:
            raise TypeError("run(k, xref, fsteps): xref and fsteps are required")
        xref = np.asarray(xref, dtype=np.float64)
        fsteps = np.asarray(fsteps, dtype=np.float64)
        batched = xref.ndim == 3
        xb = xref if batched else xref[None]
        fb = fsteps if batched else fsteps[None]
        if xb.shape[1:] != (12, self.n_steps + 1) or fb.shape[1:] != (20, 13) or fb.shape[0] != xb.shape[0]:
            raise ValueError("xref must be ([B,] 12, %d) and fsteps ([B,] 20, 13)" % (self.n_steps + 1))
        B = xb.shape[0]
        if self._engine is None or self._engine.B != B:
            self._make_engine(B)
        self._batched = batched
        eng = self._engine
        eng.run(float(k), xb, fb)
        x = eng.solution()
        f0 = eng.forces()
        N = self.n_steps
        # retrieve_result, MPC.py:432-450
        x_robot = x[:, :12 * N].reshape(B, N, 12).transpose(0, 2, 1) + xb[:, :, 1:]
        # dead-reckoned world pose, MPC.py:503-510
        c, s = np.cos(self._qw[:, 5]), np.sin(self._qw[:, 5])
        qn = x_robot[:, 0:6, 0]
        self._qw[:, 0] += c * qn[:, 0] - s * qn[:, 1]
        self._qw[:, 1] += s * qn[:, 0] + c * qn[:, 1]
        self._qw[:, 2] = qn[:, 2]
        self._qw[:, 3:5] = qn[:, 3:5]
        self._qw[:, 5] += qn[:, 5]
        self.xref = xref
        self.x0 = xref[..., 0:1]
        if batched:
            self.x, self.f_applied, self.x_robot = x, f0, x_robot
            self.q_next, self.v_next = x_robot[:, 0:6, 0:1], x_robot[:, 6:12, 0:1]
            self.q_w = self._qw[:, :, None].copy()
        else:
            self.x, self.f_applied, self.x_robot = x[0], f0[0], x_robot[0]
            self.q_next, self.v_next = x_robot[0, 0:6, 0:1], x_robot[0, 6:12, 0:1]
            self.q_w = self._qw[0].reshape(6, 1).copy()
            if k > 0:                                     # MPC.py:478-482
                self.q[0:6, 0] = xref[0:6, 0]
                self.v[0:6, 0] = xref[6:12, 0]
        self.status = None
        return 0

    def run_async(self, k, xref, fsteps, slot):
        """Enqueue one tick and the copy of its forces into the engine's pinned slot; returns without waiting.
        Used by MPC_Wrapper's asynchronous mode; the caller owns xref / fsteps until the result has been collected."""
        xref = np.asarray(xref, dtype=np.float64)
        fsteps = np.asarray(fsteps, dtype=np.float64)
        batched = xref.ndim == 3
        xb = xref if batched else xref[None]
        fb = fsteps if batched else fsteps[None]
        if xb.shape[1:] != (12, self.n_steps + 1) or fb.shape[1:] != (20, 13) or fb.shape[0] != xb.shape[0]:
            raise ValueError("xref must be ([B,] 12, %d) and fsteps ([B,] 20, 13)" % (self.n_steps + 1))
        if self._engine is None or self._engine.B != xb.shape[0]:
            self._make_engine(xb.shape[0])
        self._batched = batched
        self._engine.run(float(k), np.array(xb), np.array(fb))      # private copies: the planner reuses its arrays
        self._engine.result_async(slot)
        return 0

    @property
    def info(self):
        """Per-instance solver diagnostics of the last run (status, sweeps, iters, obj, masks, y)."""
        return self._engine.info()

    def reset(self):
        if self._engine is not None:
            self._engine.reset_warm_start()
