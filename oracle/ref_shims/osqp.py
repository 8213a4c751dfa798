"""Stand-in for the `osqp` package so the UNMODIFIED reference MPC.py can run end to end in this
container (test infrastructure; used only by tests/golden/make_golden.py and oracle self-checks).

`OSQP` is oracle/osqp_port.OSQP -- the restated algorithm -- with two additions: every call MPC.py
makes is recorded (so fixtures can capture exactly what crossed the MPC.py -> osqp boundary), and
`solve()` runs at the north-star tolerance eps_abs = eps_rel = 1e-8, then polishes on the active
set and KKT-certifies the point (oracle/kkt.py) before handing it back to MPC.py.
"""
import numpy as np

from oracle import kkt
from oracle.osqp_port import OSQP as _Port

ORACLE_EPS = 1e-8


class OSQP(_Port):
    def __init__(self):
        super().__init__()
        self.calls = []          # [(name, kwargs-copy)]
        self.last_cert = None
        self.last_raw = None
        self.last_y = None

    def setup(self, **kw):
        self.calls.append(("setup", {k: (v.copy() if hasattr(v, "copy") else v) for k, v in kw.items()}))
        return super().setup(**kw)

    def update_settings(self, **kw):
        self.calls.append(("update_settings", dict(kw)))
        return super().update_settings(**kw)

    def update(self, **kw):
        self.calls.append(("update", {k: np.array(v, copy=True) for k, v in kw.items()}))
        return super().update(**kw)

    def warm_start(self, **kw):
        self.calls.append(("warm_start", {k: np.array(v, copy=True) for k, v in kw.items()}))
        return super().warm_start(**kw)

    def solve(self):
        self.calls.append(("solve", {}))
        self.settings["eps_abs"] = ORACLE_EPS
        self.settings["eps_rel"] = ORACLE_EPS
        P = self.P0
        xp, yp, cert, raw = kkt.solve_certified(P, self.q0, self.A0, np.where(self.l0 <= -1e30, -np.inf, self.l0),
                                                np.where(self.u0 >= 1e30, np.inf, self.u0), self,
                                                run=lambda: _Port.solve(self))
        self.last_cert, self.last_raw, self.last_y = cert, raw, yp
        raw.x_admm = raw.x
        raw.x = xp
        raw.y = yp
        return raw
