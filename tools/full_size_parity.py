#!/usr/bin/env python
"""Parity of the GPU path against the oracle at the full benchmark shape (n_tau=2000, n_omega=1000, 60 alphas,
cut 1e-11) on the first N spectra of the benchmark batch (N = host cores by default).  Prints one JSON summary.

    python tools/full_size_parity.py [N]
"""
import json, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import maxent_oracle as mo
from maxent_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 8)
dump = os.path.join(tempfile.mkdtemp(), "oracle.npz")
r = subprocess.run([sys.executable, "-m", "oracle.cpu_baseline", "--n-tau", "2000", "--n-omega", "1000", "--n-alpha", "60",
                    "--spectra", str(n), "--procs", str(n), "--thr", "1e-11", "--dump", dump], cwd=ROOT, check=True,
                   capture_output=True, text=True)
cpu = json.loads(r.stdout.strip().splitlines()[-1])
o = np.load(dump)
pr = mo.synthetic_problem(2000, 1000, mu=np.ones(1), noise=np.zeros((1, 2000)))
prob = engine.SharedProblem(pr["K"], pr["err"], mo.flat_default_model(pr["omega"]), pr["delta"], reduce_singular_space=1e-11)
res = engine.run_sweep(prob, o["G"], mo.log_alpha_mesh(0.01, 2000, 60) * 2000)
A = res.A.cpu().numpy(); idx = res.alpha_index.cpu().numpy(); chi2 = res.chi2.cpu().numpy()
dA = np.max(np.abs(A - o["A"]), axis=-1) / np.max(np.abs(o["A"]), axis=-1)          # [n, n_alpha]
dc = np.abs(chi2 / o["chi2"] - 1)
out = dict(spectra=n, n_sv=prob.n_sv,
           linefit_identical=int(np.sum(idx[:, 0] == o["linefit"])), chi2curv_identical=int(np.sum(idx[:, 1] == o["chi2curv"])),
           max_rel_A_at_picks=float(max(max(dA[b, idx[b, 0]], dA[b, idx[b, 1]]) for b in range(n))),
           max_rel_A_alpha_idx_0_36=float(dA[:, :37].max()), max_rel_chi2_alpha_idx_0_36=float(dc[:, :37].max()),
           max_rel_A_tail_37_59=float(dA[:, 37:].max()), max_rel_chi2_all=float(dc.max()),
           lm_iterations_gpu=res.n_iter.sum(1).cpu().numpy().tolist(), lm_iterations_oracle=o["n_iter"].sum(1).tolist(),
           oracle_wall_s=cpu["wall_s"], oracle_cores=cpu["cores"])
print(json.dumps(out))
