"""Probability of a MaxEnt solution given alpha.

``NormalLogProbability`` selects the device evaluation of (python/probabilities.py:76-85)

    log p = -1/2 log det(d2Q/dH2) + 1/2 log det(-d2S/dH2) + (N_omega/2) log alpha - Q - log alpha

which the sweep kernel computes in its Sylvester-reduced form
log p = -1/2 log det(1 + eta Xi Z Xi / alpha) - Q - log alpha with an n_sv x n_sv Cholesky
(csrc/mx_sweep2.cuh, ``logdet_prob``)."""


class Probability(object):
    def __call__(self, Q):
        raise NotImplementedError("probabilities are evaluated by the fused device kernel (MaxEntLoop.run)")


class NormalLogProbability(Probability):
    """Descriptor: ask the sweep for log p at every alpha (``MxProblem.want_probability``)."""
