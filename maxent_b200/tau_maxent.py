"""``TauMaxEnt``: continuation of one G(tau) -- the user-facing front end of the reference
(python/tau_maxent.py:34-356) on top of the fused ``MaxEntLoop``.

All attributes of the loop (``omega, alpha_mesh, D, K, G, err, reduce_singular_space, A_init,
scale_alpha, analyzers, probability, minimizer, cost_function, logtaker, run, ...``) are reachable on the
``TauMaxEnt`` object itself (attribute shadowing as in python/tau_maxent.py:70-77).  Data enter through
``set_G_tau_data / set_G_tau_file``, error bars through ``set_error`` or a covariance matrix through
``set_cov / set_cov_file`` (eigen-decomposition and rotation of kernel and data into the diagonal
basis, python/tau_maxent.py:253-288).  The TRIQS Green-function setters (``set_G_tau``, ``set_G_iw``)
need the ``triqs`` package, which is not a dependency of this build."""
import copy

import numpy as np

from .default_models import FlatDefaultModel
from .kernels import TauKernel
from .maxent_loop import MaxEntLoop
from .omega_meshes import HyperbolicOmegaMesh


class TauMaxEnt(object):
    maxent_loop = None           # must exist on the class for the attribute shadowing below

    def __init__(self, cov_threshold=1.e-14, **kwargs):
        self.maxent_loop = MaxEntLoop(**kwargs)
        omega = HyperbolicOmegaMesh()
        self.D = FlatDefaultModel(omega)
        self.K = TauKernel([0, 1], omega)          # placeholder kernel until data are set
        self.omega = omega
        self.cov_threshold = cov_threshold

    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, 'maxent_loop'), name)

    def __setattr__(self, name, value):
        if hasattr(self.maxent_loop, name):
            setattr(self.maxent_loop, name, value)
        else:
            object.__setattr__(self, name, value)

    # ---- data -----------------------------------------------------------------------------------------
    def set_G_tau(self, G_tau, re=True, tau_new=None):
        raise NotImplementedError("set_G_tau needs TRIQS Green-function objects; use set_G_tau_data")

    def set_G_iw(self, G_iw, np_tau=-1, **kwargs):
        raise NotImplementedError("set_G_iw needs TRIQS Green-function objects; use set_G_tau_data")

    def set_G_tau_data(self, tau, G_tau):
        """tau grid and G(tau) as arrays of equal length."""
        assert len(tau) == len(G_tau), "tau and G_tau don't have the same dimension"
        self.tau = tau
        self.G = G_tau
        self._transform(self._T, G_original_basis=True)

    def set_G_tau_file(self, filename, tau_col=0, G_col=1, err_col=None):
        """Text file with columns tau, G(tau) [, error] (0-based column numbers)."""
        dat = np.loadtxt(filename)
        self.tau = dat[:, tau_col]
        self.G = dat[:, G_col]
        if err_col is not None:
            self.err = dat[:, err_col]
            self._transform(None, G_original_basis=True)     # a diagonal error replaces any covariance rotation
        else:
            self._transform(self._T, G_original_basis=True)

    def set_error(self, error):
        """Scalar error or one value per tau point."""
        if not np.all(np.isreal(error)):
            raise Exception('complex error supplied, only real accepted')
        error = np.real(error)
        try:
            if len(error) != len(self.G):
                raise Exception('Supply scalar error or with length of G_tau.')
            self.err = error
        except TypeError:
            self.err = error * np.ones(np.shape(self.G))
        self._transform(None)

    def set_cov(self, cov):
        """Symmetric covariance matrix of G(tau): rotate into its eigenbasis, errors = sqrt(eigenvalues),
        eigenvalues below ``cov_threshold`` dropped."""
        self.cov = cov
        assert np.max(np.abs(cov - cov.transpose())) < 1.e-10, 'Supplied covariance matrix is not symmetric.'
        e, v = np.linalg.eigh(cov)
        if np.any(e < 0):
            self.logtaker.error_message(
                "Eigenvalues of the covariance matrix are not all positive; they will be ignored. "
                "Smallest negative value: {}", np.min(e))
        keep = e >= self.cov_threshold
        e, v = e[keep], v[:, keep]
        self.err = None
        if hasattr(self.cost_function, "_G_orig"):
            self.G = self.cost_function._G_orig
        self._transform(v.conjugate().transpose())
        self.err = np.sqrt(e)

    def set_cov_file(self, filename):
        self.set_cov(np.loadtxt(filename))

    def _transform(self, T_, G_original_basis=False):
        """Rotate G and K from the left so that they carry the absolute rotation ``T_``
        (python/tau_maxent.py:303-325)."""
        if G_original_basis:
            self.cost_function._G_orig = copy.deepcopy(self.G)
        T_from = None if G_original_basis else self._T
        if T_ is None:
            step = 1 if T_from is None else T_from.conjugate().transpose()
        else:
            step = T_ if T_from is None else np.dot(T_, T_from.conjugate().transpose())
        self.G = np.dot(step, self.G)
        self.K.transform(T_)
        self.K = self.K                      # notify the cost function

    # ---- tau <-> data_variable ------------------------------------------------------------------------
    def get_tau(self):
        return self.maxent_loop.get_data_variable()

    def set_tau(self, tau, update_K=True, update_chi2=True, update_Q=True, update_H_of_v=True):
        self.maxent_loop.set_data_variable(tau, update_K=update_K, update_chi2=update_chi2, update_Q=update_Q,
                                           update_H_of_v=update_H_of_v)

    tau = property(get_tau, set_tau)

    @property
    def _T(self):
        return self.K._T
