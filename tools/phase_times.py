#!/usr/bin/env python
"""Per-phase device time of the sweep kernel (MxSweepOut.phase_cycles, clock64 of thread 0 of every CTA) on a
slice of the benchmark batch.  MX_MAX_CTAS_PER_SM=1 gives the uncontended phase times (one CTA per SM).

    python tools/phase_times.py [spectra]
"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maxent_b200 import batched, engine

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 592
job = batched.BatchedTauMaxEnt(reduce_singular_space=1e-11)
G = batched.synthetic_bootstrap_batch(2000, 1000, B, seed=5)
job.set_kernel_tau(np.linspace(0.0, 40.0, 2000), batched.hyperbolic_omega(-10.0, 10.0, 1000), beta=40.0)
job.set_alpha_mesh_log(0.01, 2000.0, 60)
job.set_error(1.e-4)
prob = job.prepare()
Gd = G.cuda()
names = ["planner", "solver", "T-pass", "H-pass", "gradient", "J assembly", "accept/convergence/output", "replay"]
out = {}
for per_sm in ("2", "1"):
    os.environ["MX_MAX_CTAS_PER_SM"] = per_sm
    lm = engine.LMParams()
    n = B if per_sm == "2" else B // 2
    for _ in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = engine.run_sweep(prob, Gd[:n], job.alpha_effective(), want_v=False, analyze_results=False, phase_timers=True, lm=lm)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cyc = res.phase_cycles.double().cpu().numpy()
    it = float(res.n_iter.sum()); nb = float(res.n_batch.sum())
    us = cyc.sum(0) / 1965.0 / it                 # per LM iteration, 1965 MHz
    tpass_per_batch = cyc[:, 2].sum() / 1965.0 / nb
    tot = us.sum()
    out[per_sm] = dict(ms=ms, spectra=n, us_per_iteration=dict(zip(names, np.round(us, 2).tolist())), total_us=round(tot, 2),
                       batches_per_iteration=round(nb / it, 3), spectra_per_s=round(n / ms * 1e3, 1))
    out[per_sm]["tpass_us_per_batch"] = round(tpass_per_batch, 2)
    print("CTAs/SM", per_sm, json.dumps(out[per_sm]))
