"""No-op stand-in for `matplotlib` (absent from this image).  The reference imports it at module
import time (`train.py:15-17`, `visualization.py:12-15`, `a2c.py:12,566,788`); plotting has no effect
on any simulation or training result, so every call is swallowed."""
import sys
import types


class _Null:
    """Absorbs any attribute access, call, indexing or iteration."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Null()

    def __call__(self, *a, **k):
        return _Null()

    def __getitem__(self, key):
        return _Null()

    def __setitem__(self, key, value):
        pass

    def __iter__(self):
        return iter(())

    def __len__(self):
        return 0

    def __bool__(self):
        return False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _NullModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Null()


def use(*a, **k):
    return None


def get_backend():
    return "agg"


def _subplots(nrows=1, ncols=1, *a, **k):
    import numpy as np

    fig = _Null()
    n = int(nrows) * int(ncols)
    if n == 1:
        return fig, _Null()
    axes = np.empty((nrows, ncols), dtype=object)
    for i in range(nrows):
        for j in range(ncols):
            axes[i, j] = _Null()
    return fig, (axes.reshape(-1) if 1 in (nrows, ncols) else axes)


pyplot = _NullModule(__name__ + ".pyplot")
pyplot.subplots = _subplots
patches = _NullModule(__name__ + ".patches")
sys.modules[__name__ + ".pyplot"] = pyplot
sys.modules[__name__ + ".patches"] = patches
__version__ = "0.0-fjsp-b200-standin"
