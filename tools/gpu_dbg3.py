import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import maxent_b200 as mb
from maxent_b200 import engine
from tests import gpu_common as gc
from oracle import maxent_oracle as mo
np.set_printoptions(linewidth=200, precision=3)
g = gc.load_golden("g1_semicircular_prob.npz")
prob, res = gc.run_fixture(g)
print("engine path  chi2 rel", np.abs(res.chi2[0].cpu().numpy() / g["ref_chi2"] - 1), "n_sv", prob.n_sv)
print("   S tail", prob.S.cpu().numpy()[-6:], " QtQ err", float((prob.Q.T @ prob.Q - torch.eye(prob.n_sv, device='cuda', dtype=torch.float64)).abs().max()))
for mode in ("data", "hyper"):
    tm = mb.TauMaxEnt(probability='normal')
    tm.set_verbosity(mb.VerbosityFlags.Quiet)
    tm.set_G_tau_data(g["tau"], g["G"])
    tm.omega = mb.DataOmegaMesh(g["omega"]) if mode == "data" else mb.HyperbolicOmegaMesh(-10, 10, 200)
    tm.alpha_mesh = mb.DataAlphaMesh(g["alpha_mesh"])
    tm.set_error(1e-3)
    r = tm.run()
    p = tm.maxent_loop.shared_problem()
    print(mode, "chi2 rel", np.abs(r.chi2 / g["ref_chi2"] - 1), "n_sv", p.n_sv, "prob err", np.abs(r.probability - g["known_probability"]))
    print("   S tail", tm.K.S[-6:], " QtQ err", float((p.Q.T @ p.Q - torch.eye(p.n_sv, device='cuda', dtype=torch.float64)).abs().max()))
    Kref = mo.tau_kernel(g["tau"], g["omega"], None)
    print("   K diff", np.max(np.abs(tm.K.K - Kref)), "A rel", gc.rel_A(r.A, g["ref_A"]))
    print("   c0 check: G proj", float(torch.as_tensor(g["G"]).norm()))
