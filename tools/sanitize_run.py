#!/usr/bin/env python
"""Tiny full-shape run for compute-sanitizer: two spectra at n_tau=2000, n_omega=1000 (NT = 7 instantiation of the
sweep kernel), 6 alphas, probability on; the kernel SVD comes from torch so that the run stays short under the tool.

    compute-sanitizer --tool racecheck python tools/sanitize_run.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxent_b200 import batched

job = batched.BatchedTauMaxEnt(reduce_singular_space=1e-11, probability="normal", svd="torch")
G = batched.synthetic_bootstrap_batch(2000, 1000, 2, seed=5)
job.set_kernel_tau(np.linspace(0.0, 40.0, 2000), batched.hyperbolic_omega(-10.0, 10.0, 1000), beta=40.0)
job.set_alpha_mesh_log(1.0, 2000.0, 6)
job.set_error(1.e-4)
out = job.run(G)
print("n_sv", out.n_sv, "LM iterations", int(out.n_iter.sum()), "converged", bool(out.converged.all()),
      "picks", out.alpha_index[:, :2].tolist())
