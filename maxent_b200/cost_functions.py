"""Cost functions Q_alpha(v) = 1/2 eta chi2(H(v)) - alpha S(H(v)) as descriptors of the fused device kernel.

Attribute surface of the reference's ``CostFunction`` (python/cost_functions/cost_function.py:43-268):
``chi2 / S / H_of_v / A_of_H`` components, forwarded ``K / G / err / D / omega / data_variable``
properties with ``set_*`` methods, ``chi2_factor``, ``set_alpha``.  Which device variant runs is decided
by ``variant()``:

* ``MaxEntCostFunction`` + Normal entropy / parametrisation  -> MX_VARIANT_NORMAL
  (f = V^T diag(H) (K^T W r + alpha log(H/D)), J = V^T diag(H) (K^T W K + alpha/H) diag(H) V;
  python/cost_functions/maxent_cost_function.py:68-165 with d_dv=False, dA_projection=2)
* ``MaxEntCostFunction`` + PlusMinus entropy / parametrisation -> MX_VARIANT_PLUSMINUS
* ``BryanCostFunction`` (f = g + alpha v, J = Gamma Z; python/cost_functions/bryan_cost_function.py:84-128)
  -> MX_VARIANT_BRYAN

Anything else (``d_dv=True``, ``dA_projection != 2``, foreign component classes) is refused with
``NotImplementedError`` when the loop runs: there is no generic host minimiser to fall back to."""
import numpy as np

from .functions import (NormalChi2, ComplexChi2, NormalEntropy, PlusMinusEntropy, NormalH_of_v, PlusMinusH_of_v,
                        IdentityA_of_H, PreblurA_of_H)


def _same_grid(a, b):
    if a is b:
        return True
    if a is None or b is None:
        return False
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all(a == b))


class CostFunction(object):

    def __init__(self, chi2=None, S=None, H_of_v=None, A_of_H=None, chi2_factor=1.0):
        self._chi2 = NormalChi2() if chi2 is None else chi2
        self._S = NormalEntropy() if S is None else S
        self._H_of_v = NormalH_of_v() if H_of_v is None else H_of_v
        self._A_of_H = IdentityA_of_H(getattr(self._chi2, "omega", None)) if A_of_H is None else A_of_H
        self.chi2_factor = chi2_factor
        self._alpha = None

    def set_alpha(self, alpha):
        self._alpha = alpha

    def parameter_change(self):
        pass

    def __call__(self, v):
        raise NotImplementedError("Q(v) is evaluated inside the fused device kernel only; run MaxEntLoop.run")

    # ---- which device kernel ---------------------------------------------------------------------
    def variant(self):
        raise NotImplementedError("Please use a subclass of CostFunction.")

    def _entropy_variant(self):
        s_tag = getattr(self._S, "variant_tag", None)
        h_tag = getattr(self._H_of_v, "variant_tag", None)
        if not isinstance(self._chi2, (NormalChi2, ComplexChi2)) or not isinstance(self._A_of_H, (IdentityA_of_H, PreblurA_of_H)):
            raise NotImplementedError("only NormalChi2 / ComplexChi2 with IdentityA_of_H or PreblurA_of_H run on the fused path")
        if s_tag is None or s_tag != h_tag:
            raise NotImplementedError("entropy %s with parametrisation %s is not a fused variant (use Normal+Normal "
                                      "or PlusMinus+PlusMinus)" % (type(self._S).__name__, type(self._H_of_v).__name__))
        return s_tag

    # ---- forwarded problem data (names of python/cost_functions/cost_function.py:128-261) ----------
    def get_K(self):
        return self._chi2.K

    def set_K(self, K, update_chi2=True, update_H_of_v=True, update_Q=True):
        self._chi2.set_K(K)
        self._H_of_v.set_K(K)

    K = property(get_K, set_K)

    def get_G(self):
        return self._chi2.G

    def set_G(self, G, update_chi2=True, update_Q=True):
        self._chi2.set_G(G)

    G = property(get_G, set_G)

    def get_err(self):
        return self._chi2.err

    def set_err(self, err, update_chi2=True, update_Q=True):
        self._chi2.set_err(err)

    err = property(get_err, set_err)

    def get_omega(self):
        return self._chi2.K.omega

    def set_omega(self, omega, update_K=True, update_chi2=True, update_D=True, update_S=True,
                  update_H_of_v=True, update_A_of_H=True, update_Q=True):
        K, D = self._chi2.K, self._S.D
        if K is not None:
            same = _same_grid(K.omega, omega)
            K.omega = omega
            if update_K and not same:               # an unchanged mesh keeps the kernel values and its SVD
                K.parameter_change()
        if D is not None:
            same = _same_grid(D.omega, omega)
            D.omega = omega
            if update_D and not same:
                D.parameter_change()
        self._A_of_H.set_omega(omega)

    omega = property(get_omega, set_omega)

    def get_data_variable(self):
        return self._chi2.K.data_variable

    def set_data_variable(self, data_variable, update_K=True, update_chi2=True, update_Q=True, update_H_of_v=True):
        K = self._chi2.K
        same = _same_grid(K.data_variable, data_variable)
        K.data_variable = data_variable
        if update_K and not same:                   # re-setting the same tau grid keeps kernel and SVD
            K.parameter_change()

    data_variable = property(get_data_variable, set_data_variable)

    def get_D(self):
        return self._S.D

    def set_D(self, D, update_S=True, update_H_of_v=True, update_Q=True, update_A_of_H=True):
        self._S.set_D(D)
        self._H_of_v.set_D(D)
        self._A_of_H.set_omega(D.omega)

    D = property(get_D, set_D)

    def get_chi2(self):
        return self._chi2

    def set_chi2(self, chi2, update_Q=True):
        self._chi2 = chi2

    chi2 = property(get_chi2, set_chi2)

    def get_S(self):
        return self._S

    def set_S(self, S, update_Q=True):
        self._S = S

    S = property(get_S, set_S)

    def get_H_of_v(self):
        return self._H_of_v

    def set_H_of_v(self, H_of_v, update_Q=True):
        self._H_of_v = H_of_v

    H_of_v = property(get_H_of_v, set_H_of_v)

    def get_A_of_H(self):
        return self._A_of_H

    def set_A_of_H(self, A_of_H, update_Q=True):
        self._A_of_H = A_of_H

    A_of_H = property(get_A_of_H, set_A_of_H)

    @property
    def G_orig(self):
        """G in the original (unrotated) basis when a covariance rotation is active (tau_maxent.py:303-306)."""
        return getattr(self, "_G_orig", self.G)


class MaxEntCostFunction(CostFunction):
    """General MaxEnt cost function.  The fused path implements the reference's default projection
    (``d_dv=False, dA_projection=2``: gradient and Gauss-Newton Hessian projected with dH/dv)."""

    def __init__(self, d_dv=False, dA_projection=2, **kwargs):
        self.d_dv = d_dv
        self.dA_projection = dA_projection
        super(MaxEntCostFunction, self).__init__(**kwargs)

    def variant(self):
        if self.d_dv or self.dA_projection != 2:
            raise NotImplementedError("MaxEntCostFunction(d_dv=%r, dA_projection=%r) is not on the fused path; "
                                      "only the default d_dv=False, dA_projection=2 is" % (self.d_dv, self.dA_projection))
        return self._entropy_variant()


class BryanCostFunction(CostFunction):
    """Bryan's singular-space equations (normal chi2 and entropy only, as in the reference)."""

    def __init__(self, **kwargs):
        super(BryanCostFunction, self).__init__(**kwargs)

    def variant(self):
        if self._entropy_variant() != "normal":
            raise NotImplementedError("BryanCostFunction needs NormalEntropy and NormalH_of_v")
        return "bryan"
