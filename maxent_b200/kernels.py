"""Continuation kernels K(data variable, omega) and their singular value decomposition.

Mirror of the kernel objects of the reference (python/kernels.py): ``KernelSVD`` (:36-122), ``Kernel``
(:125-180), ``DataKernel`` (:183-207), ``TauKernel`` (:210-280).  The numbers are produced on the device
through the C ABI: ``mx_tau_kernel`` fills K (TauKernel._fill_values, python/kernels.py:244-266) and
``mx_svd_jacobi`` replaces ``np.linalg.svd`` (KernelSVD.svd, python/kernels.py:53-64).  The host keeps
numpy copies because the reference's public attributes (``K``, ``K_delta``, ``U``, ``S``, ``V``) are arrays.

There is no CPU fallback: building a TauKernel or asking for the SVD without a CUDA device raises
``MaxEntLibraryError``.

One documented deviation: ``reduce_singular_space(threshold)`` keeps ``S >= max(threshold,
rank_floor * S[0])`` with ``rank_floor = 5e-16``.  The reference's cut is purely absolute
(python/kernels.py:117); below ``rank_floor * S[0]`` the computed singular triplets are rounding noise
whose count depends on the SVD implementation (LAPACK gesdd keeps 76 / 996 / 2000 of them at
1000x400 / 2000x1000 / 10000x2000), and the optimum A_alpha does not depend on them.
Pass ``rank_floor=0`` to ``reduce_singular_space`` to get the purely absolute cut."""
import numpy as np

RANK_FLOOR = 5.e-16


def _device_svd(K):
    from . import engine
    return engine.svd_jacobi_host(K)


class KernelSVD(object):
    """A matrix with a lazily computed thin SVD  K = U diag(S) V^T  (V stored as n_omega x k)."""

    def __init__(self, K=None):
        self._K = K
        self._U = self._S = self._V = None
        self._last_threshold = None
        self._svd_version = 0          # bumped whenever U/S/V change (the engine caches on it)

    def _drop_svd(self):
        self._U = self._S = self._V = None
        self._svd_version += 1

    def fused_matrix(self):
        """The real matrix the fused path works with (and whose SVD ``svd`` returns): K itself for real kernels; complex
        kernels stack real and imaginary rows (IOmegaKernel)."""
        return np.asarray(self.K, dtype=np.float64)

    def svd(self):
        """Perform the SVD if not yet done; returns (U, S, V)."""
        K = self.fused_matrix()                 # materialise first: a dirty kernel may (re)compute its SVD on the way
        if self._U is None:
            self._U, self._S, self._V = _device_svd(K)
            self._svd_version += 1
        return self._U, self._S, self._V

    @property
    def U(self):
        return self.svd()[0]

    @property
    def S(self):
        return self.svd()[1]

    @property
    def V(self):
        return self.svd()[2]

    @property
    def K(self):
        return self._K

    def reduce_singular_space(self, threshold=1.e-14, rank_floor=RANK_FLOOR):
        """Drop all singular triplets with S < threshold (absolute, as in the reference), but never keep
        anything below the numerical rank ``rank_floor * S[0]`` (see the module docstring)."""
        if self._last_threshold is not None and (threshold is None or threshold < self._last_threshold):
            self._drop_svd()                     # a lower cut needs the dropped vectors back
        self._last_threshold = threshold
        U, S, V = self.svd()
        thr = -np.inf if threshold is None else threshold
        if len(S):
            thr = max(thr, rank_floor * S[0])
        keep = np.where(S >= thr)[0]
        if len(keep) != len(S):
            self._U, self._S, self._V = U[:, keep], S[keep], V[:, keep]
            self._svd_version += 1
        return self


class Kernel(KernelSVD):
    """Kernel of the analytic continuation on an omega mesh; may carry a rotation T of the data space
    (covariance whitening, TauMaxEnt.set_cov)."""

    def __init__(self):
        super(Kernel, self).__init__()
        self.omega = None
        self._T = None
        self._K_delta = None


    @property
    def K_delta(self):
        """K times delta omega, in the ORIGINAL (unrotated) data basis: G_rec = K_delta A."""
        return self._K_delta

    @property
    def data_variable(self):
        raise NotImplementedError("Use a subclass of Kernel")

    def parameter_change(self):
        """To be called after the data variable or the omega mesh changed."""
        self._fill_values()

    def _fill_values(self):
        raise NotImplementedError("Use a subclass of Kernel")

    def transform(self, T_):
        """Left-multiply the kernel (and U) so that it carries the ABSOLUTE rotation ``T_`` with respect to
        the unrotated kernel (python/kernels.py:160-180).  ``None`` undoes the current rotation."""
        if T_ is None:
            if self._T is None:
                return
            step = self._T.conjugate().transpose()
        elif self._T is not None:
            step = np.dot(T_, self._T.conjugate().transpose())
        else:
            step = T_
        self._T = T_
        self._U = np.dot(step, self.U)          # S and V are those of the unrotated kernel (reference quirk)
        self._K = np.dot(step, self._K)
        self._svd_version += 1


class DataKernel(Kernel):
    """Kernel given verbatim as a matrix K[len(data_variable), len(omega)]."""

    def __init__(self, data_variable, omega, K):
        super(DataKernel, self).__init__()
        self._data_variable = data_variable
        self.omega = omega
        self._K = np.array(K, dtype=np.float64)
        self._K_delta = self._K * np.asarray(self.omega.delta)[None, :]

    def _fill_values(self):
        self._K_delta = self._K * np.asarray(self.omega.delta)[None, :]

    @property
    def data_variable(self):
        return self._data_variable


class TauKernel(Kernel):
    r"""K(tau, omega) = -exp(-tau omega) / (1 + exp(-beta omega)), evaluated in the overflow-free form of
    python/kernels.py:259-264 by the device kernel ``mx_tau_kernel``.  ``beta`` defaults to ``tau[-1]``."""

    def __init__(self, tau, omega, beta=None):
        super(TauKernel, self).__init__()
        self.tau = tau
        self.omega = omega
        self.beta = beta
        self._dirty = True
        self._fill_values()

    def _fill_values(self):
        """Mark the values stale; they are (re)computed on the device at the next access, so that setting
        tau, omega and beta one after the other costs one kernel fill, not three."""
        self._drop_svd()
        self._dirty = True

    def _ensure(self):
        if not self._dirty:
            return
        from . import engine
        self._dirty = False
        beta = self.tau[-1] if self.beta is None else self.beta
        self._K = engine.tau_kernel_host(np.asarray(self.tau, dtype=np.float64),
                                         np.asarray(self.omega, dtype=np.float64), float(beta))
        self._K_delta = self._K * np.asarray(self.omega.delta)[None, :]
        T, self._T = self._T, None
        self.transform(T)                        # re-apply a covariance rotation (python/kernels.py:268-271)

    @property
    def K(self):
        self._ensure()
        return self._K

    @property
    def K_delta(self):
        self._ensure()
        return self._K_delta

    def transform(self, T_):
        if T_ is None and self._T is None:
            return                               # nothing to do; do not force a kernel fill
        self._ensure()
        super(TauKernel, self).transform(T_)

    @property
    def data_variable(self):
        return self.tau

    @data_variable.setter
    def data_variable(self, value):
        self.tau = value


class IOmegaKernel(Kernel):
    """Matsubara-frequency kernel K(i omega_n, omega) = 1 / (i omega_n - omega)  (python/kernels.py:283-346):
    G(i omega_n) = int d omega K A(omega); ``iomega`` is the REAL array of Matsubara frequencies, ``K`` and ``K_delta``
    are complex like the reference's.

    On the fused path the spectral function is real (Normal / PlusMinus / Bryan cost functions): the misfit
    chi2 = sum_n |G_n - (K H)_n|^2 / sigma_n^2  (ComplexChi2.f, python/functions.py:401-404, for a real H) is the
    NormalChi2 of the real system with the rows [Re K; Im K], [Re G; Im G], [sigma; sigma] -- ``fused_matrix`` returns
    that stacked real matrix and ``U, S, V`` are ITS singular triplets (V real, U of 2 n rows), not those of the complex
    matrix.  The complex-A formalism of the reference (ComplexPlusMinusEntropy / ComplexPlusMinusH_of_v) is not on the
    fused path."""

    is_complex = True

    def __init__(self, iomega, omega, beta=None):
        super(IOmegaKernel, self).__init__()
        self.iomega = np.asarray(iomega, dtype=np.float64)
        self.omega = omega
        self.beta = beta
        self._fill_values()

    def _fill_values(self):
        self._drop_svd()
        oomega, iiomega = np.meshgrid(np.asarray(self.omega, dtype=np.float64), self.iomega)
        self._K = 1.0 / (1.0j * iiomega - oomega)
        self._K_delta = self._K * np.asarray(self.omega.delta)[None, :]      # trapezoid weights folded in
        if self._T is not None:
            raise NotImplementedError("a covariance rotation of complex Matsubara data is not on the fused path")

    def fused_matrix(self):
        return np.ascontiguousarray(np.vstack([self._K.real, self._K.imag]))

    @staticmethod
    def stack(x):
        """[Re x; Im x] of a data-space vector (data, or an error bar repeated for both parts)."""
        x = np.asarray(x)
        return np.concatenate([x.real, x.imag]) if np.iscomplexobj(x) else np.concatenate([x, x])

    def transform(self, T_):
        if T_ is not None:
            raise NotImplementedError("a covariance rotation of complex Matsubara data is not on the fused path")

    @property
    def data_variable(self):
        return self.iomega

    @data_variable.setter
    def data_variable(self, value):
        self.iomega = np.asarray(value, dtype=np.float64)
        self._fill_values()


class PreblurKernel(Kernel):
    """Kernel of the preblur formalism (python/kernels.py:349-413): K_pb = K diag(delta omega) B, so that
    G = K_pb H for a hidden image H (which carries a delta omega) and A = B H.  Wraps another kernel (e.g. a
    TauKernel); always use it together with ``PreblurA_of_H(b, omega)``.  ``K_delta`` is the inner kernel's."""

    def __init__(self, K, b):
        KernelSVD.__init__(self)
        self._T = None
        self.kernel = K
        self._b = b
        self._fill_values()

    def parameter_change(self):
        self.kernel.parameter_change()
        self._fill_values()

    def _fill_values(self):
        from . import engine
        from .preblur import get_preblur
        self._drop_svd()
        self._B = get_preblur(self.omega, self._b)
        dB = np.asarray(self.omega.delta)[:, None] * self._B
        self._K = engine.matmul_host(np.asarray(self.kernel.K, dtype=np.float64), dB)       # device GEMM
        self._K_delta = self.kernel.K_delta

    def transform(self, T):
        self.kernel.transform(T)
        self._fill_values()
        self._T = self.kernel._T

    def get_omega(self):
        return self.kernel.omega

    def set_omega(self, omega):
        self.kernel.omega = omega

    omega = property(get_omega, set_omega)

    @property
    def b(self):
        return self._b

    @property
    def data_variable(self):
        return self.kernel.data_variable

    @data_variable.setter
    def data_variable(self, value):
        self.kernel.data_variable = value
