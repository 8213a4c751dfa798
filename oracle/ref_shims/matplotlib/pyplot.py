"""Import shim (test infrastructure) for `from matplotlib import pyplot as plt` in the reference."""


def __getattr__(name):
    def _unavailable(*a, **k):
        raise RuntimeError("matplotlib is not installed; plotting is outside the hot path")
    return _unavailable
