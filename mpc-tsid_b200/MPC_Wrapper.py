"""Drop-in for the reference's `MPC_Wrapper` (MPC_Wrapper.py:20-114).

    wrapper = MPC_Wrapper(dt, n_steps, k_mpc, T_gait, multiprocessing=False)
    wrapper.solve(k, fstep_planner)         # reads fstep_planner.xref and fstep_planner.fsteps
    f = wrapper.get_latest_result()         # 12 forces; [0, 0, 8] * 4 on the very first call

A planner whose xref / fsteps carry a leading batch axis makes everything batched.

`multiprocessing=True` is the reference's asynchronous protocol (MPC_Wrapper.py:116-260: a second process solves while
the control loop goes on and `get_latest_result` returns the newest forces that have ARRIVED, never blocking).  The
reference's implementation is dead code (MPC_Wrapper.py:48-51 raises before reaching it; wrong arity at :199).  Here
it is a stream pipeline instead of a process: `solve` enqueues the tick and the copy of its forces into one of two
pinned slots and returns; `get_latest_result` picks up the newest slot whose copy has landed.
"""
import numpy as np

import MPC


class MPC_Wrapper:
    def __init__(self, dt, n_steps, k_mpc, T_gait, multiprocessing=False, **solver_options):
        self.f_applied = np.zeros((12,))
        self.not_first_iter = False
        self.k_mpc = k_mpc                                  # MPC_Wrapper.py:26
        self.multiprocessing = multiprocessing
        self.mpc = MPC.MPC(dt, n_steps, T_gait, **solver_options)
        self._issued = 0                                    # asynchronous mode: ticks enqueued so far
        self._collected = 0                                 # ... and the newest one whose forces were picked up

    def solve(self, k, fstep_planner):
        if self.multiprocessing:
            self.run_MPC_asynchronous(k, fstep_planner)     # MPC_Wrapper.py:52
        else:
            self.run_MPC_synchronous(k, fstep_planner)      # MPC_Wrapper.py:55
        return 0

    def get_latest_result(self):
        if self.not_first_iter:
            if self.multiprocessing:
                self._collect(block=False)                   # MPC_Wrapper.py:64-70: take a new result if there is one
                return self.f_applied
            return self.mpc.f_applied                        # MPC_Wrapper.py:74
        self.not_first_iter = True                           # MPC_Wrapper.py:76-78
        first = np.array([0.0, 0.0, 8.0] * 4)
        return np.tile(first, (self.mpc._B, 1)) if self.mpc._batched else first

    def run_MPC_synchronous(self, k, fstep_planner):
        self.mpc.run((k / self.k_mpc), fstep_planner.xref, fstep_planner.fsteps)    # MPC_Wrapper.py:103
        self.f_applied = self.mpc.f_applied                  # MPC_Wrapper.py:114

    def run_MPC_asynchronous(self, k, fstep_planner):
        """MPC_Wrapper.py:116-141: hand the inputs over and return; the result is picked up later."""
        self.mpc.run_async((k / self.k_mpc), fstep_planner.xref, fstep_planner.fsteps, self._issued & 1)
        self._issued += 1

    def _collect(self, block):
        eng = self.mpc._engine
        while self._collected < self._issued:
            nxt = self._collected                            # oldest tick not collected yet lives in slot nxt & 1
            if self._issued - nxt > 2:                       # its slot was reused by a newer tick: skip it
                self._collected += 1
                continue
            if not block and not eng.result_ready(nxt & 1):
                break
            f = eng.result_wait(nxt & 1)
            self.f_applied = f if self.mpc._batched else f[0]
            self._collected += 1

    def wait(self):
        """Block until every enqueued tick has been solved and collected (asynchronous mode)."""
        if self.multiprocessing:
            self._collect(block=True)
        return self.f_applied

    def stop_parallel_loop(self):                            # MPC_Wrapper.py:262-268
        return 0
