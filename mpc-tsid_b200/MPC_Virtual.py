"""Drop-in for the reference's solver selector (MPC_Virtual.py:6-35): the plug point main.py:46 uses."""
import MPC_Wrapper


class MPC_Virtual():
    def __init__(self, mpc_type, dt_mpc, n_steps, k_mpc, T_gait, **solver_options):
        if not mpc_type:
            raise NotImplementedError("only the QP-based MPC (mpc_type=True) exists, as in the reference")
        self.solver = MPC_Wrapper.MPC_Wrapper(dt_mpc, n_steps, k_mpc, T_gait, multiprocessing=False, **solver_options)
        self.solve = self.solver.solve
        self.get_latest_result = self.solver.get_latest_result
