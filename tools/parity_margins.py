#!/usr/bin/env python
"""How much of the tiered tolerance the device results use on the golden fixtures: worst (deviation / tolerance) per
fixture and field -- the margin the parity tests have against a rounding-level change of the kernels.

    python tools/parity_margins.py
"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_common as gc

out = {}
for name in ["g1_semicircular_prob.npz", "g2_synth_200x100.npz", "g3_plusminus_offdiag.npz", "g4_bryan_200x100.npz",
             "g5_config1_cut1e-11.npz", "g5b_config1_default_cut.npz", "g15_low_temperature_wide.npz"]:
    g = gc.load_golden(name)
    prob, res = gc.run_fixture(g)
    tolA, tolc = gc.tolerances(g, "A"), gc.tolerances(g, "chi2")
    dA = gc.rel_A(res.A[0].cpu().numpy(), g["ref_A"])
    dc = np.abs(res.chi2[0].cpu().numpy() / g["ref_chi2"] - 1)
    out[name] = dict(n_sv=prob.n_sv, worst_A=round(float(np.max(dA / tolA)), 3), at_alpha=int(np.argmax(dA / tolA)),
                     worst_chi2=round(float(np.max(dc / tolc)), 3), at_alpha_chi2=int(np.argmax(dc / tolc)),
                     n_alpha=int(len(dA)), alphas_at_floor=int(np.sum(tolA == 1e-8)))
    print(name, json.dumps(out[name])); sys.stdout.flush()
