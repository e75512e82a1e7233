"""Levenberg-Marquardt minimiser parameters and convergence criteria.

The iteration itself (python/minimizers/levenberg_minimizer.py:123-248) runs inside the fused sweep
kernel (csrc/mx_sweep2.cuh: ``lm_run`` replays the reference's damping search verbatim); this module
holds the user-facing objects with the reference's constructor arguments and translates them into the
``MxLMParams`` struct of the C ABI.  Convergence criteria are small expression trees
(python/minimizers/convergence_methods.py:24-122); the device evaluates

    max|dQ/dv| < conv_max_derivative   OR   | |Q0 - Q1| / Q1 | < conv_rel_change

    OR   |Q0 - Q1| < conv_abs_change

so any and/or tree over ``MaxDerivativeConvergenceMethod`` / ``RelativeFunctionChangeConvergenceMethod`` /
``FunctionChangeConvergenceMethod`` / ``NullConvergenceMethod`` maps onto it (the reference's ``&`` is an ``or``
too -- convergence_methods.py:61).  A negative threshold switches a criterion off."""
import numpy as np


class ConvergenceMethod(object):
    """Base class; combine with ``&`` / ``|``."""

    def __and__(self, other):
        return AndConvergenceMethod(self, other)

    def __or__(self, other):
        return OrConvergenceMethod(self, other)

    def __call__(self, function, v, **kwargs):
        raise NotImplementedError('Convergence is tested on the device; see thresholds()')

    def thresholds(self):
        """(conv_max_derivative, conv_rel_change, conv_abs_change) for the device test; -1 disables a criterion."""
        raise NotImplementedError('Please use a subclass of ConvergenceMethod.')


class _Pair(ConvergenceMethod):
    def __init__(self, one, two):
        self.one, self.two = one, two

    def thresholds(self):
        a, b = self.one.thresholds(), self.two.thresholds()
        return tuple(max(x, y) for x, y in zip(a, b))    # both conjunctions accept when EITHER side does


class AndConvergenceMethod(_Pair):
    """Behaves as 'or', like the reference (python/minimizers/convergence_methods.py:61)."""


class OrConvergenceMethod(_Pair):
    pass


class MaxDerivativeConvergenceMethod(ConvergenceMethod):
    def __init__(self, convergence_criterion):
        self.convergence_criterion = convergence_criterion

    def thresholds(self):
        return float(self.convergence_criterion), -1.0, -1.0


class RelativeFunctionChangeConvergenceMethod(ConvergenceMethod):
    def __init__(self, convergence_criterion):
        self.convergence_criterion = convergence_criterion

    def thresholds(self):
        return -1.0, float(self.convergence_criterion), -1.0


class NullConvergenceMethod(ConvergenceMethod):
    """Everything counts as converged."""

    def thresholds(self):
        return float(np.inf), -1.0, -1.0


class FunctionChangeConvergenceMethod(ConvergenceMethod):
    def __init__(self, convergence_criterion):
        self.convergence_criterion = convergence_criterion

    def thresholds(self):
        return -1.0, -1.0, float(self.convergence_criterion)


class Minimizer(object):
    def minimize(self, function, v0):
        raise NotImplementedError("Use a subclass of Minimizer")


class LevenbergMinimizer(Minimizer):
    """Parameters of the reference's LevenbergMinimizer (python/minimizers/levenberg_minimizer.py:92-121).
    ``marquardt=True`` damps with mu * diag(J) instead of mu * 1 (:181-185); ``J_squared`` (off by default) is not
    implemented by the fused kernel."""

    def __init__(self, convergence=None, maxiter=1000, miniter=0, J_squared=False, marquardt=False,
                 mu0=1.e-18, nu=1.3, max_mu=1.e20, verbose_callback=None):
        self.convergence = (OrConvergenceMethod(MaxDerivativeConvergenceMethod(1.e-4),
                                                RelativeFunctionChangeConvergenceMethod(1.e-16))
                            if convergence is None else convergence)
        self.maxiter, self.miniter = maxiter, miniter
        self.J_squared, self.marquardt = J_squared, marquardt
        self.mu0, self.nu, self.max_mu = mu0, nu, max_mu
        self.verbose_callback = verbose_callback
        self.n_iter = 0              # iterations in total
        self.n_iter_last = 0         # iterations of the last alpha
        self.converged = False

    def lm_params(self):
        """-> engine.LMParams for the C ABI."""
        from .engine import LMParams
        if self.J_squared:
            raise NotImplementedError("LevenbergMinimizer(J_squared=True) is not on the fused path")
        cd, cr, ca = self.convergence.thresholds()
        return LMParams(maxiter=self.maxiter, miniter=self.miniter, mu0=self.mu0, nu=self.nu, max_mu=self.max_mu,
                        conv_max_derivative=cd, conv_rel_change=cr, conv_abs_change=ca, marquardt=self.marquardt)

    def minimize(self, function, v0):
        raise NotImplementedError("LevenbergMinimizer.minimize runs inside the fused device kernel (MaxEntLoop.run); "
                                  "there is no host implementation")
