"""Drop-in for the reference's `MPC_Wrapper` (MPC_Wrapper.py:20-114), synchronous path.

    wrapper = MPC_Wrapper(dt, n_steps, k_mpc, T_gait, multiprocessing=False)
    wrapper.solve(k, fstep_planner)         # reads fstep_planner.xref and fstep_planner.fsteps
    f = wrapper.get_latest_result()         # 12 forces; [0, 0, 8] * 4 on the very first call

A planner whose xref / fsteps carry a leading batch axis makes everything batched.  The reference's
asynchronous path is dead code (MPC_Wrapper.py:48-51 raises before reaching it; wrong arity at :199);
asking for it here raises NotImplementedError.
"""
import numpy as np

import MPC


class MPC_Wrapper:
    def __init__(self, dt, n_steps, k_mpc, T_gait, multiprocessing=False, **solver_options):
        self.f_applied = np.zeros((12,))
        self.not_first_iter = False
        self.k_mpc = k_mpc                                  # MPC_Wrapper.py:26
        self.multiprocessing = multiprocessing
        if multiprocessing:
            raise NotImplementedError("Asynchronous MPC is not up to date (as in the reference, MPC_Wrapper.py:48-51)")
        self.mpc = MPC.MPC(dt, n_steps, T_gait, **solver_options)

    def solve(self, k, fstep_planner):
        self.run_MPC_synchronous(k, fstep_planner)
        return 0

    def get_latest_result(self):
        if self.not_first_iter:
            return self.mpc.f_applied                        # MPC_Wrapper.py:74
        self.not_first_iter = True                           # MPC_Wrapper.py:76-78
        first = np.array([0.0, 0.0, 8.0] * 4)
        fa = np.asarray(self.mpc.f_applied)
        return np.tile(first, (fa.shape[0], 1)) if fa.ndim == 2 else first

    def run_MPC_synchronous(self, k, fstep_planner):
        self.mpc.run((k / self.k_mpc), fstep_planner.xref, fstep_planner.fsteps)    # MPC_Wrapper.py:103
        self.f_applied = self.mpc.f_applied                  # MPC_Wrapper.py:114
