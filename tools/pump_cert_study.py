"""CPU study (oracle, build container): how early in a pass over omega can a "pump" trial of the Levenberg damping
search be certified as worse than Q0?

In the first `while` of levenberg_minimizer.py:203-206 the reference only asks whether Q1 > Q0 (or NaN).  Since
Q = chi2/2 - alpha S with chi2 >= 0 and every term of -S non-negative, the running sum alpha * (-S_partial) over the
omega rows processed so far is a rigorous lower bound of Q1: once it exceeds Q0 the answer is known and the rest of the
pass is dead work.  This script replays the oracle on one benchmark spectrum and records, for every group of 8
consecutive pump trials (one device batch), the first checkpoint (in omega rows) at which all 8 are certified.

    python tools/pump_cert_study.py [n_tau n_omega]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import maxent_oracle as mo

n_tau = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_omega = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
pr = mo.synthetic_problem(n_tau, n_omega, seed=1234)
K, G, err, omega = pr["K"], pr["G"][0] if pr["G"].ndim == 2 else pr["G"], pr["err"], pr["omega"]
delta = mo.omega_delta(omega)
D = mo.flat_default_model(omega)
U, S, V = mo.kernel_svd(K, 1e-11)
prob = mo.Problem(K, G, err, D, delta, U, S, V, "normal", 1.0, fast_d2=True)
alpha_mesh = mo.log_alpha_mesh(0.01, 2000.0, 60)
checkpoints = [64, 128, 256, 512, n_omega]
first_cp = {c: 0 for c in checkpoints}
first_cp["never"] = 0
trial_first = {c: 0 for c in checkpoints}
trial_first["never"] = 0
n_pump_trials = 0
margins = []


def cert_point(alpha, v, dv, Q0):
    """first checkpoint at which alpha * (-S_partial) > Q0 with a relative margin; None if never"""
    x = V @ (v - dv)
    with np.errstate(all="ignore"):
        H = D * np.exp(x)
        lg = np.where(np.exp(x) <= 1e-100, -230.25850929940458, x)
        terms = -(H - D - H * lg)            # >= 0 up to rounding
    cs = np.cumsum(np.where(np.isfinite(terms), terms, np.inf))
    thr = Q0 + 1e-6 * abs(Q0)
    for c in checkpoints:
        if alpha * cs[min(c, n_omega) - 1] > thr:
            return c
    return None


def minimize(alpha, v):
    global n_pump_trials
    mu = 1e-18
    nu = 1.3
    fv = mo.BoundQ(prob, alpha, v)
    Q1 = fv.f()
    Q0 = np.nan
    for i in range(1000):
        f = fv.d()
        J = fv.dd()
        with np.errstate(all="ignore"):
            if np.max(np.abs(f)) < 1e-4 or np.abs(np.abs(Q0 - Q1) / Q1) < 1e-16:
                break
        Id = np.eye(len(J))
        Q0 = Q1
        old = np.seterr(all="ignore")
        dv = np.linalg.solve(J + mu * Id, f)
        Q1 = mo.BoundQ(prob, alpha, v - dv).f()
        pump = []
        while (Q1 > Q0 or np.isnan(Q1)) and mu < 1e20:
            mu *= nu
            dv = np.linalg.solve(J + mu * Id, f)
            Q1 = mo.BoundQ(prob, alpha, v - dv).f()
            pump.append((cert_point(alpha, v, dv, Q0), Q1, Q0))
        # device batches: the first batch of an iteration holds mu0 and other walk points; once in PH_PUMP the
        # batches are 8 consecutive pump dampings
        for b in range(0, len(pump), 8):
            grp = pump[b:b + 8]
            n_pump_trials += len(grp)
            cps = [g[0] for g in grp]
            for c in cps:
                trial_first["never" if c is None else c] += 1
            if any(c is None for c in cps):
                first_cp["never"] += 1
            else:
                first_cp[max(cps)] += 1
        dv2 = np.linalg.solve(J + nu * mu * Id, f)
        Q2 = mo.BoundQ(prob, alpha, v - dv2).f()
        if Q2 < Q1:
            nuf = nu; mu *= nu; Q2 = Q1; dvnew = dv2
        else:
            nuf = 1.0 / nu; mu /= nuf; dvnew = dv
        Q1 = np.inf
        while Q2 < Q1 and mu < 1e20 and mu > nu * np.finfo(float).eps:
            Q1 = Q2; dv = dvnew; mu *= nuf
            dvnew = np.linalg.solve(J + mu * Id, f)
            Q2 = mo.BoundQ(prob, alpha, v - dvnew).f()
        np.seterr(**old)
        v -= dv
        fv = mo.BoundQ(prob, alpha, v)
        Q1 = fv.f()
    return v


H0 = D * delta
v = prob.v_of_H(H0.copy())
for a in alpha_mesh:
    v = minimize(a * len(G), v)
print(json.dumps(dict(n_tau=n_tau, n_omega=n_omega, n_sv=len(S), pump_trials=n_pump_trials,
                      batches_by_first_checkpoint={str(k): v for k, v in first_cp.items()},
                      trials_by_first_checkpoint={str(k): v for k, v in trial_first.items()})))
