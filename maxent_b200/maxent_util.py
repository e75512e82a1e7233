"""Small numerical helpers of the reference's public namespace (python/maxent_util.py).

``numder`` / ``check_der`` are host utilities for users who write their own functions of ``x`` (finite-difference
Jacobian and a comparison against an analytic derivative, python/maxent_util.py:170-237); the two converters to
TRIQS Green functions need TRIQS containers and raise, as in a reference build without TRIQS."""
import itertools

import numpy as np

from .triqs_support import require_triqs


def numder(fun, x, delta=1.e-6):
    """Central-difference Jacobian of ``fun`` at ``x``: result[..., i] = (fun(x + delta e_i) - fun(x - delta e_i)) / (2 delta),
    with the shape of ``fun(x)`` (``(1,)`` for a scalar function) followed by the shape of ``x``."""
    x = np.asarray(x)
    jac = None
    for idx in itertools.product(*[range(n) for n in x.shape]):
        # the lower point is reached from the upper one ((x_i + delta) - 2 delta), as in the reference: same rounding
        hi = np.array(x, dtype=float)
        hi[idx] += delta
        lo = hi.copy()
        lo[idx] -= 2 * delta
        up, down = np.asarray(fun(hi)), np.asarray(fun(lo))
        if jac is None:
            lead = up.shape if up.ndim else (1,)
            jac = np.empty(lead + x.shape, dtype=up.dtype)
        jac[(Ellipsis,) + idx] = (up - down) / (2.0 * delta)
    return jac


def check_der(f, d, around, renorm=False, prec=1.e-8, name=''):
    """True if the analytic derivative ``d(around)`` agrees with ``numder(f, around)`` to ``prec`` (absolute; relative
    to ``f(around)`` if ``renorm is True``, to ``renorm`` if it is a number); prints the reference's message if not."""
    err = np.abs(numder(f, around) - d(around))
    if renorm is True:
        err = err / np.abs(f(around))
    elif renorm is not False:
        err = err / np.abs(renorm)
    worst = np.max(err)
    if worst > prec:
        print('numerical derivative does not fit analytic derivative: {} - difference {}'.format(name, worst))
        return False
    return True


@require_triqs
def get_G_w_from_A_w(A_w, w_points, np_interp_A=None, np_omega=2000, w_min=-10, w_max=10, broadening_factor=1.0):
    """G(omega) as a TRIQS ``GfReFreq`` from A(omega) (python/maxent_util.py:43-132)."""


@require_triqs
def get_G_tau_from_A_w(A_w, w_points, beta, np_tau):
    """G(tau) as a TRIQS ``GfImTime`` from A(omega) (python/maxent_util.py:135-167).  Without TRIQS the same numbers
    are ``TauKernel(tau, omega, beta).K_delta @ A_w``."""
