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
  * results travel when they are read: `run` enqueues the tick and returns; `f_applied` / `q_next` / `v_next` / `q_w` cost one
    small copy (24 doubles per robot, what MPC_Wrapper.py:103-114 needs), `x` / `x_robot` the full 24N per robot;
  * there is no CPU path: construction fails without the CUDA extension or without a B200.
"""
import warnings

import numpy as np

import mpcqp


class MPC:
    _SLOTS = 8                                          # page-locked result slots per engine (batched results are views of them)

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
        self.q = np.array([[0.0, 0.0, 0.2027682, 0.0, 0.0, 0.0]]).transpose()   # MPC.py:55
        self._qw0 = self.q[:, 0].copy()                 # MPC.py:58: the world pose every robot starts from
        self.v = np.zeros((6, 1))                       # MPC.py:61
        self.h_ref = self.q[2, 0]                       # MPC.py:64
        self._engine = None
        self._batched = False
        self._B = 1
        self._ran = False
        self._step = None                               # (f_applied, x_next) of the last tick once fetched
        self._full = None                               # full solution of the last tick once fetched
        self._status = None
        self._xb = None
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
            self._settle()
            self._engine.close()
        self._engine = mpcqp.Engine(batch=batch, n_steps=self.n_steps, device=self._device, dt=float(self.dt),
                                    T_gait=float(self.T_gait), **self._opts)
        # a new batch size is a new set of robots: they start from the initial world pose (MPC.py:58)
        if not np.array_equal(self._qw0, [0.0, 0.0, 0.2027682, 0.0, 0.0, 0.0]):
            self._engine.world_pose(np.tile(self._qw0, (batch, 1)))
        self._B = batch
        self._slots = [(self._engine.pinned((batch, 12)), self._engine.pinned((batch, 12))) for _ in range(self._SLOTS)]
        self._tick = 0
        self._ran = False
        self._step = self._full = self._status = None

    def _check_inputs(self, xref, fsteps):
        xref = np.asarray(xref, dtype=np.float64)
        fsteps = np.asarray(fsteps, dtype=np.float64)
        batched = xref.ndim == 3
        xb = xref if batched else xref[None]
        fb = fsteps if batched else fsteps[None]
        if xb.shape[1:] != (12, self.n_steps + 1) or fb.shape[1:] != (20, 13) or fb.shape[0] != xb.shape[0]:
            raise ValueError("xref must be ([B,] 12, %d) and fsteps ([B,] 20, 13)" % (self.n_steps + 1))
        if self._engine is None or self._engine.B != xb.shape[0]:
            self._make_engine(xb.shape[0])
        return xref, batched, xb, fb

    def run(self, k, xref=None, fsteps=None, T_gait=None):
        if xref is None or fsteps is None:
            raise TypeError("run(k, xref, fsteps): xref and fsteps are required")
        xref, batched, xb, fb = self._check_inputs(xref, fsteps)
        self._batched = batched
        self._engine.run(float(k), xb, fb)              # asynchronous: build + solve + extract on the device
        self._ran = True
        self._step = self._full = self._status = None
        self._xb = xb
        self.xref = xref                                # MPC.py:486 (aliased there too)
        self.x0 = xref[..., 0:1]
        if not batched and k > 0:                       # MPC.py:478-482
            self.q[0:6, 0] = xref[0:6, 0]
            self.v[0:6, 0] = xref[6:12, 0]
        return 0

    # ---- results, fetched when read ---------------------------------------------------------------------------
    def _settle(self):
        """Forces + first predicted state of the last tick: one small copy into a page-locked slot.  (The dead reckoning of
        MPC.py:503-510 runs on the device at the end of every solve; q_w is fetched when it is read.)"""
        if not self._ran or self._step is not None:
            return
        # result slots are used in rotation: a batched result is a VIEW of its slot and stays valid for _SLOTS ticks
        slot = self._slots[self._tick % len(self._slots)]
        f0, xn = self._engine.step_result(slot[0], slot[1])
        if not self._batched:
            f0, xn = f0.copy(), xn.copy()               # one robot: hand out private arrays like the reference does
        self._step = (f0, xn)
        self._tick += 1

    def _fetch_full(self):
        if self._full is None and self._ran:
            self._settle()
            self._full = self._engine.solution()
        return self._full

    @property
    def f_applied(self):                                # MPC.py:441
        if not self._ran:
            return np.zeros((12,))
        self._settle()
        return self._step[0] if self._batched else self._step[0][0]

    @property
    def q_next(self):                                   # MPC.py:448-449
        if not self._ran:
            return np.zeros((6, 1))
        self._settle()
        xn = self._step[1]
        return xn[:, 0:6, None] if self._batched else xn[0, 0:6, None]

    @property
    def v_next(self):                                   # MPC.py:450
        if not self._ran:
            return np.zeros((6, 1))
        self._settle()
        xn = self._step[1]
        return xn[:, 6:12, None] if self._batched else xn[0, 6:12, None]

    @property
    def q_w(self):                                      # MPC.py:58, 503-510
        if self._engine is None:
            return self._qw0.reshape(6, 1).copy()
        qw = self._engine.world_pose()
        return qw[:, :, None] if self._batched else qw[0].reshape(6, 1)

    @q_w.setter
    def q_w(self, value):
        """Re-anchor the dead reckoning (the reference's attribute is plain data; (6, 1) or (B, 6, 1))."""
        v = np.asarray(value, dtype=np.float64)
        if self._engine is None:
            self._qw0 = v.reshape(-1)[:6].copy()
        else:
            self._engine.world_pose(np.broadcast_to(v.reshape(-1, 6), (self._B, 6)))

    @property
    def x(self):                                        # MPC.py:52, 428
        if not self._ran:
            return np.zeros((24 * self.n_steps,))
        x = self._fetch_full()
        return x if self._batched else x[0]

    @property
    def x_robot(self):                                  # MPC.py:437-447
        if not self._ran:
            return np.zeros((12, self.n_steps))
        x, N = self._fetch_full(), self.n_steps
        xr = x[:, :12 * N].reshape(self._B, N, 12).transpose(0, 2, 1) + self._xb[:, :, 1:]
        return xr if self._batched else xr[0]

    @property
    def status(self):
        """Per-instance solver status of the last run (mpcqp.STATUS: 1 solved, 2 max_iter, 3 bad_input); warns once per
        tick when a robot was not solved -- its forces are feasible (or zero) but not certified optimal."""
        if not self._ran:
            return None
        if self._status is None:
            self._status = self._engine.status()
            bad = int((self._status != 1).sum())
            if bad:
                warnings.warn("%d of %d MPC instances were not solved (status %s)" % (
                    bad, self._B, sorted(set(int(s) for s in self._status[self._status != 1]))), RuntimeWarning)
        return self._status if self._batched else int(self._status[0])

    def run_async(self, k, xref, fsteps, slot):
        """Enqueue one tick and the copy of its forces into the engine's pinned slot; returns without waiting.
        Used by MPC_Wrapper's asynchronous mode; the caller owns xref / fsteps until the result has been collected."""
        xref, batched, xb, fb = self._check_inputs(xref, fsteps)
        self._settle()
        self._batched = batched
        self._engine.run(float(k), np.array(xb), np.array(fb))      # private copies: the planner reuses its arrays
        self._engine.result_async(slot)
        self._ran = True
        self._step = self._full = self._status = None
        self._xb = xb
        return 0

    @property
    def info(self):
        """Per-instance solver diagnostics of the last run (status, sweeps, iters, obj, masks, y)."""
        return self._engine.info()

    def reset(self):
        if self._engine is not None:
            self._engine.reset_warm_start()
