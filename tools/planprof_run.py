import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MAXENT_B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "maxent_b200/libmaxent_b200_planprof.so")
import numpy as np, torch
from maxent_b200 import batched, engine
job = batched.BatchedTauMaxEnt(reduce_singular_space=1e-11)
G = batched.synthetic_bootstrap_batch(2000, 1000, 1, seed=5)
job.set_kernel_tau(np.linspace(0.0, 40.0, 2000), batched.hyperbolic_omega(-10.0, 10.0, 1000), beta=40.0)
job.set_alpha_mesh_log(0.01, 2000.0, 60); job.set_error(1.e-4)
prob = job.prepare(); Gd = G.cuda()
r = engine.run_sweep(prob, Gd, job.alpha_effective(), want_v=False, analyze_results=False, phase_timers=True)
torch.cuda.synchronize()
pf = r.phase_cycles.double().cpu().numpy().sum(0)
it = float(r.n_iter.sum()); nb = float(r.n_batch.sum())
code = int(r.n_trial.long().sum())
gen = dict(first=code % 1000, pump=(code // 1000) % 1000, probe=(code // 1000000) % 100, walk=code // 100000000)
print(json.dumps(dict(batches=nb, iterations=it, generic_by_phase=gen, fast_us_per_it=pf[0] / 1965 / it,
                      generic_plus_other_us_per_it=pf[6] / 1965 / it, replay_us_per_it=pf[7] / 1965 / it)))
