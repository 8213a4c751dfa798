"""Import shim (test infrastructure): FootstepPlanner.py imports pybullet but never calls it."""
