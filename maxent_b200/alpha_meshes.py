"""Alpha meshes: 1-D float64 ``ndarray`` subclasses holding the alpha values in DESCENDING order
(the alpha loop warm-starts from large to small alpha).  Public surface of python/alpha_meshes.py."""
import numpy as np


class BaseAlphaMesh(np.ndarray):
    def __new__(cls, alpha_min=0.0001, alpha_max=20, n_points=20, *args, **kwargs):
        return super(BaseAlphaMesh, cls).__new__(cls, shape=(int(n_points),), dtype=np.float64)

    def __init__(self, alpha_min=0.0001, alpha_max=20, n_points=20):
        if n_points > 1:
            if alpha_min > alpha_max:
                raise Exception('alpha_min must be smaller than alpha_max')
            if alpha_min <= 0 or alpha_max <= 0:
                raise Exception('All alpha values must be positive')
        self.alpha_min, self.alpha_max, self.n_points = alpha_min, alpha_max, n_points
        vals = self._values(alpha_min, alpha_max, int(n_points))
        if vals is not None:
            self[...] = vals

    def _values(self, lo, hi, n):
        return None

    def __array_finalize__(self, parent):
        for name in ("alpha_min", "alpha_max", "n_points"):
            if parent is not None and hasattr(parent, name):
                setattr(self, name, getattr(parent, name))


class DataAlphaMesh(BaseAlphaMesh):
    """User-supplied alpha values, sorted descending (python/alpha_meshes.py:58-65)."""

    def __new__(cls, data):
        return super(DataAlphaMesh, cls).__new__(cls, np.min(data), np.max(data), len(data))

    def __init__(self, data):
        self._data = np.sort(np.asarray(data, dtype=np.float64))[::-1]
        super(DataAlphaMesh, self).__init__(np.min(data), np.max(data), len(data))

    def _values(self, lo, hi, n):
        return self._data


class LogAlphaMesh(BaseAlphaMesh):
    """n points equidistant in log10(alpha) between alpha_max and alpha_min (python/alpha_meshes.py:81-85)."""

    def _values(self, lo, hi, n):
        return np.logspace(np.log10(lo), np.log10(hi), n)[::-1]


class LinearAlphaMesh(BaseAlphaMesh):
    """n equidistant points between alpha_max and alpha_min (python/alpha_meshes.py:101-103)."""

    def _values(self, lo, hi, n):
        return np.linspace(lo, hi, n)[::-1]
