"""Default models D(omega).  As in the reference (python/default_models.py:61-63,88-93) the stored
``.D`` already contains the integration weight delta omega, i.e. it lives in the same space as H."""
import numpy as np


class BaseDefaultModel(object):
    def __init__(self, omega):
        self.omega = omega
        self._D = None

    @property
    def D(self):
        return self._D

    def parameter_change(self):
        self._fill_values()

    def _fill_values(self):
        raise NotImplementedError("Use a subclass of BaseDefaultModel")

    def __len__(self):
        return len(self._D)


class FlatDefaultModel(BaseDefaultModel):
    """Constant spectral density normalised to one: D_i = delta_i / sum(delta)."""

    def __init__(self, omega):
        super(FlatDefaultModel, self).__init__(omega)
        self._fill_values()

    def _fill_values(self):
        delta = self.omega.delta
        self._D = np.ones(np.shape(self.omega)) / np.sum(delta) * delta


class DataDefaultModel(BaseDefaultModel):
    """Default model given on a grid ``omega_in``; interpolated linearly onto ``omega`` if the grids differ."""

    def __init__(self, default, omega_in, omega=None):
        super(DataDefaultModel, self).__init__(omega_in if omega is None else omega)
        self.omega_in = omega_in
        self.default = default
        self._fill_values()

    def _fill_values(self):
        same = len(self.omega_in) == len(self.omega) and np.all(np.asarray(self.omega_in) == np.asarray(self.omega))
        dens = np.asarray(self.default) if same else np.interp(self.omega, self.omega_in, self.default)
        self._D = dens * self.omega.delta


class FileDefaultModel(DataDefaultModel):
    """Two-column text file: omega, D(omega).  (The reference's constructor is broken --
    python/default_models.py:110-115 refers to an undefined ``cls`` -- this one works.)"""

    def __init__(self, filename, omega=None):
        from .omega_meshes import DataOmegaMesh
        data = np.loadtxt(filename)
        super(FileDefaultModel, self).__init__(data[:, 1], DataOmegaMesh(data[:, 0]), omega)
