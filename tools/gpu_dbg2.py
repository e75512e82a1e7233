import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maxent_b200 import engine
from oracle import maxent_oracle as mo
np.set_printoptions(linewidth=250, precision=2)
g = dict(np.load("tests/golden/g5b_config1_default_cut.npz"))
K = mo.tau_kernel(g["tau"], g["omega"], None)
D = mo.flat_default_model(g["omega"])
for svd, thr in (("jacobi", 1e-14), ("torch", 1e-14), ("jacobi", 1e-12), ("jacobi", 3e-14)):
    prob = engine.SharedProblem(K, g["err"], D, mo.omega_delta(g["omega"]), reduce_singular_space=thr, svd=svd)
    res = engine.run_sweep(prob, g["G"], g["ref_alpha"])
    chi2 = res.chi2[0].cpu().numpy(); A = res.A[0].cpu().numpy()
    print("svd", svd, "thr", thr, "n_sv", prob.n_sv, "cfg", prob.config)
    print("  rel chi2", np.abs(chi2 / g["ref_chi2"] - 1)[30:])
    print("  rel A   ", (np.max(np.abs(A - g["ref_A"]), axis=1) / np.max(np.abs(g["ref_A"]), axis=1))[30:])
    print("  n_iter", res.n_iter[0].cpu().numpy()[30:], "conv", res.status[0].cpu().numpy()[30:])
