"""Blur matrix of the preblur formalism (python/preblur.py:33-58): G = K B H with a hidden image H.

    B_ij = exp(-(w_i - w_j)^2 / (2 b^2)) / sqrt(2 pi b^2),  then normalised along rows and along columns with
    the integration weights of the omega mesh.

Problem set-up on the host (an n_omega x n_omega table, computed once per (mesh, b)); the products with the kernel
and with the hidden images are done on the device (kernels.PreblurKernel, maxent_loop)."""
import numpy as np


def get_preblur(omega, b):
    w = np.asarray(omega, dtype=np.float64)
    delta = np.asarray(omega.delta, dtype=np.float64)
    B = np.exp(-np.subtract.outer(w, w) ** 2 / (2.0 * b * b)) / np.sqrt(2.0 * np.pi * b * b)
    B = B / np.dot(delta, B)[:, None]
    B = B / np.dot(B, delta)[None, :]
    return B


def preblur_scan(tm, b_values, K=None):
    """The b scan of the preblur workflow (doc/guide/preblur_example.py:46-75) as ONE launch of the alpha sweep.

    ``tm`` is a configured TauMaxEnt (data, error, omega, alpha mesh set); ``K`` the unblurred kernel (default: the
    kernel ``tm`` holds).  For every b the reference builds ``PreblurKernel(K, b)`` / ``PreblurA_of_H(b, omega)`` and
    calls ``run()`` in a loop; here the b values become jobs of one batch -- every b brings its own kernel SVD and its own
    V' (a whitening group of the launch).  Returns {b: MaxEntResult}; ``tm`` is left with the plain kernel.  (The
    reference example warm-starts each b from the previous result through A_init; the scan starts every b from the
    default model instead, which changes the Levenberg path but not the optimum.)"""
    from .kernels import PreblurKernel
    from .functions import PreblurA_of_H, IdentityA_of_H
    from .maxent_result import MaxEntResult
    ml = tm.maxent_loop
    K0 = tm.K if K is None else K
    jobs, out = [], {}
    for b in b_values:
        tm.A_of_H = PreblurA_of_H(b=b, omega=tm.omega)
        tm.K = PreblurKernel(K=K0, b=b)
        job = ml.snapshot()
        job["result"] = out[b] = MaxEntResult()        # one result object per b, like the reference's loop
        jobs.append(job)
    tm.A_of_H = IdentityA_of_H(tm.omega)
    tm.K = K0
    ml.run_jobs(jobs)
    return out
