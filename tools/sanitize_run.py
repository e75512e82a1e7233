#!/usr/bin/env python
"""Tiny full-shape run for compute-sanitizer: two spectra at n_tau=2000, n_omega=1000 (NT = 7 instantiation of the
sweep kernel), 6 alphas, probability on; the kernel SVD comes from torch so that the run stays short under the tool.

    compute-sanitizer --tool racecheck python tools/sanitize_run.py
    compute-sanitizer --tool memcheck python tools/sanitize_run.py --wide      # n_sv = 100: a wide instantiation (16 tiles)
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maxent_b200 import batched

if "--wide" in sys.argv:
    from maxent_b200 import engine
    from oracle import maxent_oracle as mo          # inputs only
    rng = np.random.RandomState(7)
    n_sv, n_tau, n_om = 100, 184, 140
    om = mo.linear_omega_mesh(-4, 4, n_om)
    U, _ = np.linalg.qr(rng.randn(n_tau, n_om))
    V, _ = np.linalg.qr(rng.randn(n_om, n_om))
    K = (U * np.concatenate([np.logspace(0, -6, n_sv), 1e-14 * np.ones(n_om - n_sv)])) @ V.T
    delta = mo.omega_delta(om)
    G = (K * delta[None, :]) @ np.exp(-(om - 0.5) ** 2) + 1e-4 * rng.randn(3, n_tau)
    for variant in ("normal", "plusminus", "bryan"):
        prob = engine.SharedProblem(K, 1e-4, mo.flat_default_model(om), delta, variant=variant, reduce_singular_space=1e-9)
        res = engine.run_sweep(prob, G, mo.log_alpha_mesh(5.0, 500, 3) * n_tau, probability=True)
        print(variant, "n_sv", prob.n_sv, "LM iterations", int(res.n_iter.sum()), "converged", bool((res.status & 1).all()))
    sys.exit(0)

job = batched.BatchedTauMaxEnt(reduce_singular_space=1e-11, probability="normal", svd="torch")
G = batched.synthetic_bootstrap_batch(2000, 1000, 2, seed=5)
job.set_kernel_tau(np.linspace(0.0, 40.0, 2000), batched.hyperbolic_omega(-10.0, 10.0, 1000), beta=40.0)
job.set_alpha_mesh_log(1.0, 2000.0, 6)
job.set_error(1.e-4)
out = job.run(G)
print("n_sv", out.n_sv, "LM iterations", int(out.n_iter.sum()), "converged", bool(out.converged.all()),
      "picks", out.alpha_index[:, :2].tolist())
