"""The numpy model of the engine's algorithm (condensed impulse-space QP, Woodbury solves, active-set
sweeps, ADMM + guarded polish) against the golden solutions -- checks the mathematics on CPU."""
import os

import numpy as np
import pytest

from common import FORCE_TOL, OBJ_RTOL
from kernel_model import Engine, ModelParams

HERE = os.path.dirname(__file__)


@pytest.mark.parametrize("name,max_as", [("trot", 6), ("trot", 0), ("bound", 6), ("aggressive", 6), ("motionless", 6)])
def test_model_matches_golden(name, max_as):
    g = np.load(os.path.join(HERE, "golden", "solve_%s.npz" % name))
    eng = Engine(ModelParams(max_as=max_as))
    T = min(len(g["k"]), 6)
    for t in range(T):
        out = eng.solve(g["xref"][t], g["fsteps"][t], first_tick=(t == 0))
        assert out["status"] == 1
        assert np.abs(out["x"][192:] - g["x"][t][192:]).max() <= FORCE_TOL * 1e-3      # forces, N
        assert np.abs(out["x"][:192] - g["x"][t][:192]).max() <= 1e-9                  # states
        assert abs(out["obj"] - g["obj"][t]) <= OBJ_RTOL * 1e-3 * abs(g["obj"][t])
        np.testing.assert_allclose(out["f_applied"], g["f_applied"][t], atol=FORCE_TOL * 1e-3)
