"""Self-energy continuation helpers (python/sigma_continuator.py).

The reference builds an auxiliary Green function G_aux(i omega) from a TRIQS self-energy container, continues it with
MaxEnt and maps A_aux(omega) back to Sigma(omega).  Every step before and after the continuation works on TRIQS block
Green functions, which this package does not depend on, so the three classes exist with the reference's names and
constructor signatures and raise ``NotImplementedError`` -- exactly what a reference build with ``USE_TRIQS=OFF`` does
(python/sigma_continuator.py:32-33,144-145,196-197).  The continuation itself is the ordinary ``TauMaxEnt`` /
``ElementwiseMaxEnt`` hot path once G_aux has been tabulated on a tau grid."""
from .triqs_support import require_triqs


class SigmaContinuator(object):
    @require_triqs
    def __init__(self):
        pass

    @require_triqs
    def set_S_iw(self, S_iw):
        pass

    @require_triqs
    def set_Gaux_w_from_Aaux_w(self, Aaux_w, w_points, *args, **kwargs):
        pass

    @require_triqs
    def set_Gaux_w(self, Gaux_w):
        pass


class DirectSigmaContinuator(SigmaContinuator):
    """G_aux = Sigma - Sigma(i inf), normalised  (python/sigma_continuator.py:130-179)."""

    @require_triqs
    def __init__(self, S_iw):
        pass


class InversionSigmaContinuator(SigmaContinuator):
    """G_aux = 1 / (i omega + C - Sigma)  (python/sigma_continuator.py:182-228)."""

    @require_triqs
    def __init__(self, S_iw, constant_shift=0):
        pass
