#!/usr/bin/env python
"""Feasibility study for the next kernel step: does the Levenberg damping search keep the
reference's decisions when every solve (J + mu 1) dv = f of one LM iteration comes from ONE factorisation of J
(tridiagonal reduction or eigen-decomposition, O(s) / O(s^2) per mu) instead of one factorisation per mu?

CPU only; drives oracle/maxent_oracle.py (test infrastructure) with np.linalg.solve replaced by
  lu      the reference's own solver (np.linalg.solve)                      -> must reproduce the fixture bit for bit
  chol    Cholesky per mu, failure = NaN step (what the device kernel does today)
  tridiag Householder tridiagonalisation of J once per iteration, LDL^T of T + mu 1 per mu, back-transform
  eig     symmetric eigen-decomposition of J once per iteration
and compares analyzer picks, A(alpha) and iteration counts with the reference run stored in a golden fixture.

    python tools/solver_study.py [g2_synth_200x100.npz | g5_config1_cut1e-11.npz ...]
"""
import json
import os
import sys

import numpy as np
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import maxent_oracle as mo  # noqa: E402

_real_solve = np.linalg.solve
_cache = {}


def _key(J):
    return (J.shape, J.tobytes())


def solve_lu(A, f, J=None, mu=None):
    return _real_solve(A, f)


def solve_chol(A, f, J=None, mu=None):
    try:
        c = sla.cho_factor(A, lower=True, check_finite=False)
    except (np.linalg.LinAlgError, ValueError):
        return np.full_like(f, np.nan)
    return sla.cho_solve(c, f, check_finite=False)


def _tridiag_of(J):
    k = _key(J)
    if _cache.get("k") != k:
        T, Q = sla.hessenberg(0.5 * (J + J.T), calc_q=True)
        _cache.update(k=k, d=np.diag(T).copy(), e=np.diag(T, -1).copy(), Q=Q)
    return _cache["d"], _cache["e"], _cache["Q"]


def solve_tridiag(A, f, J=None, mu=None):
    d, e, Q = _tridiag_of(J)
    n = len(d)
    g = Q.T @ f
    # LDL^T of T + mu 1 without pivoting (the O(s) chain a warp would run); a non-positive pivot = failed step
    piv = np.empty(n)
    l = np.empty(n - 1)
    piv[0] = d[0] + mu
    for i in range(1, n):
        if not piv[i - 1] > 0.0:
            return np.full_like(f, np.nan)
        l[i - 1] = e[i - 1] / piv[i - 1]
        piv[i] = d[i] + mu - l[i - 1] * e[i - 1]
    if not piv[-1] > 0.0:
        return np.full_like(f, np.nan)
    y = g.copy()
    for i in range(1, n):
        y[i] -= l[i - 1] * y[i - 1]
    y /= piv
    for i in range(n - 2, -1, -1):
        y[i] -= l[i] * y[i + 1]
    return Q @ y


def solve_eig(A, f, J=None, mu=None):
    k = _key(J)
    if _cache.get("ke") != k:
        w, W = np.linalg.eigh(0.5 * (J + J.T))
        _cache.update(ke=k, w=w, W=W)
    w, W = _cache["w"], _cache["W"]
    den = w + mu
    if not np.all(den > 0.0):
        return np.full_like(f, np.nan)
    return W @ ((W.T @ f) / den)


SOLVERS = dict(lu=solve_lu, chol=solve_chol, tridiag=solve_tridiag, eig=solve_eig)


def run(fixture, solver):
    g = np.load(os.path.join(ROOT, "tests", "golden", fixture))
    if str(g["variant"]) == "bryan":
        raise SystemExit("the Bryan Hessian is not symmetric (the device solves a transformed system); "
                         "this study covers the Normal / PlusMinus cost functions")
    K = mo.tau_kernel(g["tau"], g["omega"], None)
    fn = SOLVERS[solver]
    # levenberg_minimize evaluates solve(J + mu * Id, f): intercept the sum by handing it a J that remembers itself
    state = {}

    class JProxy(np.ndarray):
        def __add__(self, other):
            out = np.ndarray.__add__(self, other).view(np.ndarray)
            state["J"], state["shift"] = self.view(np.ndarray), other
            return out

    orig_dd = mo.BoundQ.dd

    def dd(self):
        return np.asarray(orig_dd(self)).view(JProxy)

    def solve(A, f):
        J = state["J"]
        mu = float(state["shift"][0, 0])                 # Id = eye: the shift matrix is mu * 1
        return fn(A, f, J=J, mu=mu)

    mo.BoundQ.dd = dd
    np.linalg.solve = solve
    try:
        out = mo.maxent_loop(K, g["G"], g["err"], g["omega"], g["alpha_mesh"], variant=str(g["variant"]),
                             probability=False, reduce_singular_space=float(g["reduce_singular_space"]))
    finally:
        np.linalg.solve = _real_solve
        mo.BoundQ.dd = orig_dd
    dA = np.max(np.abs(out["A"] - g["ref_A"]), axis=1) / np.max(np.abs(g["ref_A"]), axis=1)
    tol = np.maximum(1e-8, 10 * g["noise_A"])
    picks = {n: int(out["analyzers"][n]["alpha_index"]) for n in ("LineFitAnalyzer", "Chi2CurvatureAnalyzer")}
    ref_picks = {n: int(g["ref_idx_" + n]) for n in picks}
    return dict(fixture=fixture, solver=solver, lm_iterations=int(out["n_iter"].sum()), solves=int(out["n_solve"]),
                picks=picks, picks_identical=picks == ref_picks, max_dA_over_tol=float(np.max(dA / tol)),
                max_dA_well_determined=float(np.max(dA[g["noise_A"] < 1e-9])) if np.any(g["noise_A"] < 1e-9) else None,
                bit_identical=bool(np.array_equal(out["A"], g["ref_A"])))


if __name__ == "__main__":
    fixtures = sys.argv[1:] or ["g2_synth_200x100.npz"]
    for fx in fixtures:
        for s in ("lu", "chol", "tridiag", "eig"):
            print(json.dumps(run(fx, s)))
