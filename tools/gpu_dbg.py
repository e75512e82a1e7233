import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maxent_b200 import engine
from oracle import maxent_oracle as mo
np.set_printoptions(linewidth=250)
g = dict(np.load("tests/golden/g2_synth_200x100.npz"))
K = mo.tau_kernel(g["tau"], g["omega"], None)
D = mo.flat_default_model(g["omega"])
for eng in (1, 2):
    prob = engine.SharedProblem(K, g["err"], D, mo.omega_delta(g["omega"]), reduce_singular_space=1e-11, engine=eng)
    res = engine.run_sweep(prob, g["G"], g["ref_alpha"], lm=engine.LMParams(maxiter=int(os.environ.get("MAXITER", "3"))))
    print("engine", eng)
    print(" solves", res.n_solve[0].cpu().numpy())
    print(" qevals", res.n_qeval[0].cpu().numpy())
    print(" chi2", res.chi2[0].cpu().numpy()[:6])
    print(" S", res.S[0].cpu().numpy()[:6])
