#!/usr/bin/env python
"""Wall time of the single-spectrum BASELINE configs through the drop-in API (TauMaxEnt / ElementwiseMaxEnt),
next to the reference's own times measured in SURVEY.md section 6 / BASELINE.md (8 host cores).

    python tools/config_latency.py
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import maxent_b200 as mb
from oracle import maxent_oracle as mo          # synthetic inputs only


def timed(fn):
    torch.cuda.synchronize(); t0 = time.time(); r = fn(); torch.cuda.synchronize(); return r, time.time() - t0


out = {}
# warm up CUDA context + library
mb.TauKernel(np.linspace(0, 1, 8), mb.LinearOmegaMesh(-1, 1, 8)).K

# config 1: n_tau=1000, n_omega=400, 60 alphas, LineFit (reference: 7.2 s default cut / 5.7 s cut 1e-11)
pr = mo.synthetic_problem(1000, 400, mu=1.0, seed=1234)
def cfg1(cut):
    tm = mb.TauMaxEnt(reduce_singular_space=cut)
    tm.set_verbosity(mb.VerbosityFlags.Quiet)
    tm.set_G_tau_data(pr["tau"], pr["G"][0])
    tm.omega = mb.HyperbolicOmegaMesh(-10, 10, 400)
    tm.alpha_mesh = mb.LogAlphaMesh(0.01, 2000, 60)
    tm.set_error(1e-4)
    res, t = timed(tm.run)
    res2, t2 = timed(tm.run)                      # kernel SVD and device problem cached
    return dict(first_run_s=round(t, 3), rerun_s=round(t2, 3), n_sv=len(tm.K.S),
                linefit=res.analyzer_results['LineFitAnalyzer']['alpha_index'],
                chi2curv=res.analyzer_results['Chi2CurvatureAnalyzer']['alpha_index'],
                entropy=res.analyzer_results['EntropyAnalyzer']['alpha_index'], lm_iterations=int(res.n_iter.sum()))
out["config1_cut1e-11"] = dict(cfg1(1e-11), reference_s=5.7, reference_picks=[23, 28, 42])
out["config1_default_cut"] = dict(cfg1(1e-14), reference_s=7.2)

# config 2: 2x2 matrix, PlusMinus off-diagonals (fixture of the reference's elementwise test)
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "g6_elementwise_2x2.npz"))
def cfg2():
    ew = mb.ElementwiseMaxEnt(use_hermiticity=True)
    ew.set_verbosity(mb.VerbosityFlags.Quiet)
    ew.set_G_tau_data(g["tau"], g["G"])
    ew.omega = mb.DataOmegaMesh(g["omega"])
    ew.alpha_mesh = mb.DataAlphaMesh(g["alpha_mesh"])
    ew.set_error(float(g["err"]))
    return ew.run()
_, t = timed(cfg2)
out["config2_elementwise_2x2"] = dict(run_s=round(t, 3))

# config 4: n_tau=10000, n_omega=2000, 100 alphas, probability (reference: 205.6 s run + 298 s setup)
pr4 = mo.synthetic_problem(10000, 2000, mu=1.0, seed=1234)
def cfg4():
    tm = mb.TauMaxEnt(probability='normal', reduce_singular_space=1e-11)
    tm.set_verbosity(mb.VerbosityFlags.Quiet)
    tm.set_G_tau_data(pr4["tau"], pr4["G"][0])
    tm.omega = mb.HyperbolicOmegaMesh(-10, 10, 2000)
    tm.alpha_mesh = mb.LogAlphaMesh(0.01, 2000, 100)
    tm.set_error(1e-4)
    res, t = timed(tm.run)
    res2, t2 = timed(tm.run)
    return dict(first_run_s=round(t, 3), rerun_s=round(t2, 3), n_sv=len(tm.K.S), lm_iterations=int(res.n_iter.sum()),
                picks=[res.analyzer_results[k]['alpha_index'] for k in ('LineFitAnalyzer', 'Chi2CurvatureAnalyzer',
                                                                         'EntropyAnalyzer', 'ClassicAnalyzer')])
out["config4_large_kernel"] = dict(cfg4(), reference_run_s=205.6, reference_setup_s=298.0, reference_picks=[38, 48, 92, 99])
print(json.dumps(out, indent=1))
