#!/usr/bin/env python
"""Trace of the reference's damping search (oracle port) per Levenberg iteration: how many pump steps, which
direction the probe chooses, how long the walk is.  Input to the speculation policy of the sweep kernel
(csrc/mx_sweep2.cuh planner).  CPU only; test/measurement tooling."""
import sys, os, collections
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import maxent_oracle as mo

def trace(prob, alpha, v, rec, nu=1.3, mu0=1e-18, max_mu=1e20, maxiter=1000):
    mu = mu0
    fv = mo.BoundQ(prob, alpha, v); Q1 = fv.f(); Q0 = np.nan
    solve = np.linalg.solve
    eps_nu = nu * np.finfo(float).eps
    prev_dir = 1
    for i in range(maxiter):
        f = fv.d(); J = fv.dd()
        with np.errstate(all='ignore'):
            if np.max(np.abs(f)) < 1e-4 or np.abs(np.abs(Q0 - Q1) / Q1) < 1e-16: break
        Id = np.eye(len(J)); Q0 = Q1
        old = np.seterr(all='ignore')
        jd = np.diag(J)
        uniq = set()
        def q(m):
            uniq.add(tuple(jd + m))
            return mo.BoundQ(prob, alpha, v - solve(J + m * Id, f)).f(), solve(J + m * Id, f)
        Q1, dv = q(mu)
        pump = 0
        while (Q1 > Q0 or np.isnan(Q1)) and mu < max_mu:
            mu *= nu; Q1, dv = q(mu); pump += 1
        Q2, dv2 = q(nu * mu)
        if Q2 < Q1: nuf = nu; mu *= nu; Q2 = Q1; dvnew = dv2; d = 1
        else: nuf = 1.0 / nu; mu /= nuf; dvnew = dv; d = 0
        Q1 = np.inf; walk = 0
        while Q2 < Q1 and mu < max_mu and mu > eps_nu:
            Q1 = Q2; dv = dvnew; mu *= nuf; Q2, dvnew = q(mu); walk += 1
        np.seterr(**old)
        rec.append((pump, d, walk, len(uniq), mu > eps_nu, prev_dir))
        prev_dir = d
        v -= dv
        fv = mo.BoundQ(prob, alpha, v); Q1 = fv.f()
    return v

n_tau, n_om = int(sys.argv[1]), int(sys.argv[2])
pr = mo.synthetic_problem(n_tau, n_om, mu=np.ones(1), noise=np.random.default_rng(5).standard_normal((1, n_tau)))
U, S, V = mo.kernel_svd(pr["K"], 1e-11)
D = mo.flat_default_model(pr["omega"])
prob = mo.Problem(pr["K"], pr["G"][0], pr["err"], D, pr["delta"], U, S, V, "normal", 1.0, fast_d2=True)
H0 = D * pr["delta"]; v = prob.v_of_H(H0)
rec = []
for a in mo.log_alpha_mesh(0.01, 2000, 60):
    v = trace(prob, a * n_tau, v, rec)
rec = np.array(rec)
print("iterations", len(rec), "mean uniq/iter", rec[:, 3].mean())
print("pump hist", collections.Counter(rec[:, 0].tolist()).most_common(8))
act = rec[rec[:, 4] == 1]
print("active-walk iterations", len(act), "of", len(rec))
print("dir up frac", act[:, 1].mean(), " dir same as previous", (act[:, 1] == act[:, 5]).mean())
for d in (0, 1):
    w = act[act[:, 1] == d][:, 2]
    print("dir", d, "walk hist", sorted(collections.Counter(w.tolist()).items()))
print("uniq hist", sorted(collections.Counter(rec[:, 3].tolist()).items()))
np.save("/tmp/lm_trace_%d_%d.npy" % (n_tau, n_om), rec)
