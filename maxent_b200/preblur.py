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
