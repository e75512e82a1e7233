"""Ingredients of the cost function: misfit chi2(H), entropy S(H), parametrisation H(v), output map A(H).

In the reference (python/functions.py) these are numpy classes evaluated thousands of times per
spectrum by the Python minimizer.  Here they are *descriptors*: they carry the problem data
(K, G, err, D, omega) with the reference's attribute names and select the variant of the fused
device kernel (``MX_VARIANT_*``); the arithmetic itself lives in csrc/mx_sweep2.cuh:

* ``NormalChi2``        chi2 = sum_i (sum_j K_ij H_j - G_i)^2 / err_i^2      (python/functions.py:336-377)
* ``NormalEntropy``     S = sum (H - D - H log(H/D))                          (python/functions.py:491-520)
* ``PlusMinusEntropy``  S = S_n(H+) + S_n(H-), H+- = (sqrt(H^2+4D^2) +- H)/2  (python/functions.py:523-564)
* ``NormalH_of_v``      H = D exp(V v)                                        (python/functions.py:720-755)
* ``PlusMinusH_of_v``   H = D (exp(V v) - exp(-V v))                          (python/functions.py:758-796)
* ``IdentityA_of_H``    A = H / delta omega                                   (python/functions.py:937-964)

Evaluating one of them on the host (``chi2(H).f()`` ...) is deliberately not offered: there is no CPU
implementation of the hot path in this package.  Variants that the fused path does not cover
(complex chi2 / entropies, NoExp / Identity H(v)) raise ``NotImplementedError`` on construction."""
import numpy as np


def safelog(A):
    """log(A) with |A| <= 1e-100 replaced by 1e-100 IN the argument (python/functions.py:53-56); the device
    code applies the same clamp.  Host utility for users' post-processing only."""
    A[np.abs(A) <= 1.e-100] = 1.e-100
    return np.log(A)


def view_real(A):
    """Complex array -> real array with a trailing axis of length 2."""
    return np.ascontiguousarray(A).view(float).reshape(A.shape + (2,))


def view_complex(A):
    return np.ascontiguousarray(A).view(complex).reshape(A.shape[:-1])


class _NotEvaluable(object):
    _what = "function"

    def __call__(self, x):
        raise NotImplementedError(
            "%s is a descriptor of the fused device kernel; evaluate it through MaxEntLoop.run / the C ABI "
            "(there is no host implementation of the hot path)" % type(self).__name__)


class GenericFunction(_NotEvaluable):
    pass


DoublyDerivableFunction = GenericFunction
CachedFunction = GenericFunction


def cached(func):
    """Kept for import compatibility (python/functions.py:71-83); descriptors have nothing to cache."""
    return func


class InvertibleFunction(GenericFunction):
    pass


class NullFunction(GenericFunction):
    """Placeholder that evaluates to zero (python/functions.py:228-241)."""

    def __call__(self, x):
        return self

    def f(self):
        return 0.0

    def d(self):
        return 0.0

    def dd(self):
        return 0.0


# ---- chi2 -----------------------------------------------------------------------------------------
class Chi2(GenericFunction):
    """Misfit descriptor holding K, G, err (setters named as in python/functions.py:244-333)."""

    def __init__(self, K=None, G=None, err=None):
        self._K, self._G, self._err = K, G, err

    def get_K(self):
        return self._K

    def set_K(self, K, update_chi2=True):
        self._K = K

    K = property(get_K, set_K)

    def get_G(self):
        return self._G

    def set_G(self, G, update_chi2=True):
        self._G = G

    G = property(get_G, set_G)

    def get_err(self):
        return self._err

    def set_err(self, err, update_chi2=True):
        self._err = err

    err = property(get_err, set_err)

    def get_omega(self):
        return None if self._K is None else self._K.omega

    omega = property(get_omega)

    def get_data_variable(self):
        return None if self._K is None else self._K.data_variable

    data_variable = property(get_data_variable)

    def parameter_change(self):
        """The reference re-runs an O(n_tau n_omega^2) einsum here on every setter (python/functions.py:372-377);
        the singular-space formulation needs no K^T W K, so this is a no-op."""

    @property
    def input_size(self):
        return (len(self._K.omega),)

    @property
    def axes_preference(self):
        return (0,)


class NormalChi2(Chi2):
    variant_tag = "normal"


class ComplexChi2(Chi2):
    """chi2 = sum_i |G_i - sum_j K_ij H_j|^2 / sigma_i^2 for complex data and kernel (python/functions.py:380-437).  With
    the real spectral functions of the fused path this is NormalChi2 on the stacked rows [Re; Im] (see
    kernels.IOmegaKernel); the complex-A entropies of the reference are not on the fused path."""
    variant_tag = "normal"


# ---- entropies ------------------------------------------------------------------------------------
class Entropy(GenericFunction):
    """Entropy descriptor holding the default model D (python/functions.py:440-488)."""

    def __init__(self, D=None):
        self._D = D

    def get_D(self):
        return self._D

    def set_D(self, D):
        self._D = D

    D = property(get_D, set_D)

    @property
    def omega(self):
        return None if self._D is None else self._D.omega

    def parameter_change(self):
        pass


class NormalEntropy(Entropy):
    variant_tag = "normal"


class PlusMinusEntropy(Entropy):
    variant_tag = "plusminus"


def _unsupported(name, why):
    class _U(GenericFunction):
        def __init__(self, *args, **kwargs):
            raise NotImplementedError("%s is outside the fused B200 path (%s)" % (name, why))
    _U.__name__ = name
    return _U


ComplexPlusMinusEntropy = _unsupported("ComplexPlusMinusEntropy", "complex spectral functions")
AbsoluteEntropy = _unsupported("AbsoluteEntropy", "not named by the hot path")
ShiftedAbsoluteEntropy = _unsupported("ShiftedAbsoluteEntropy", "not named by the hot path")


# ---- H(v) -----------------------------------------------------------------------------------------
class GenericH_of_v(InvertibleFunction):
    """Parametrisation descriptor holding D and K (python/functions.py:656-717)."""

    def __init__(self, D=None, K=None):
        self._D, self._K = D, K

    def get_D(self):
        return self._D

    def set_D(self, D):
        self._D = D

    D = property(get_D, set_D)

    def get_K(self):
        return self._K

    def set_K(self, K):
        self._K = K

    K = property(get_K, set_K)

    @property
    def omega(self):
        return None if self._D is None else self._D.omega

    def parameter_change(self):
        pass


class NormalH_of_v(GenericH_of_v):
    variant_tag = "normal"


class PlusMinusH_of_v(GenericH_of_v):
    variant_tag = "plusminus"


ComplexPlusMinusH_of_v = _unsupported("ComplexPlusMinusH_of_v", "complex spectral functions")
NoExpH_of_v = _unsupported("NoExpH_of_v", "not named by the hot path")
IdentityH_of_v = _unsupported("IdentityH_of_v", "not named by the hot path")


# ---- A(H) -----------------------------------------------------------------------------------------
class GenericA_of_H(GenericFunction):
    def __init__(self, omega=None):
        self._omega = omega

    def get_omega(self):
        return self._omega

    def set_omega(self, omega):
        self._omega = omega

    omega = property(get_omega, set_omega)

    def parameter_change(self):
        pass


class IdentityA_of_H(GenericA_of_H):
    """A = H / delta omega (python/functions.py:947-952); applied by the sweep kernel when it writes A."""


class PreblurA_of_H(GenericA_of_H):
    """A = B H (python/functions.py:967-1023): the output map of the preblur formalism.  The sweep kernel works on
    the hidden image H; the blur is applied to its H_alpha on the device when the results are collected
    (maxent_loop.run_jobs).  Use together with ``PreblurKernel``."""

    def __init__(self, b, omega):
        self._omega = omega
        self._b = b
        self.parameter_change()

    def parameter_change(self):
        from .preblur import get_preblur
        self._B = get_preblur(self._omega, self._b)

    def set_omega(self, omega):
        self._omega = omega
        self.parameter_change()

    omega = property(GenericA_of_H.get_omega, set_omega)

    def get_b(self):
        return self._b

    def set_b(self, b, update_A_of_H=True):
        self._b = b
        if update_A_of_H:
            self.parameter_change()

    b = property(get_b, set_b)
