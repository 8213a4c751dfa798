"""Import shim (test infrastructure) standing in for the reference's utils.py, which cannot be
imported as shipped (it imports a missing ContactSequencer, pybullet and pinocchio).  Only the
one helper the MPC hot path uses is provided; it restates utils.py:179-185."""
import numpy as np


def getSkew(a):
    # [a]x such that [a]x b = a x b   (reference: utils.py:179-185)
    return np.array([[0.0, -a[2], a[1]],
                     [a[2], 0.0, -a[0]],
                     [-a[1], a[0], 0.0]], dtype=a.dtype)
