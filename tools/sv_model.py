"""Numpy model of the singular-space algorithm the CUDA kernel implements (design validation only).
Compares against golden reference outputs."""
import sys, os, time
import numpy as np
import scipy.linalg as sl
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import maxent_oracle as mo

def prep(K, G, err, thr):
    U, S, V = mo.kernel_svd(K, thr)
    err = np.asarray(err, float) * np.ones(len(G))
    w = 1.0 / err**2
    Gam0 = U.T @ (w[:, None] * U)
    R0 = np.linalg.cholesky(Gam0).T          # upper, Gam0 = R0^T R0
    p = U.T @ (w * G)
    gt = sl.solve_triangular(R0, p, trans='T', lower=False)
    q = sl.solve_triangular(R0, gt, lower=False)
    resid = G - U @ q
    c = np.sum(w * resid**2)
    R = R0 * S[None, :]
    return U, S, V, R, gt, c

def model_run(K, G, err, omega, alphas, thr, variant="normal", solver="chol", maxiter=1000):
    delta = mo.omega_delta(omega); D = mo.flat_default_model(omega)
    U, S, V, R, gt, c = prep(K, G, err, thr)
    Gam = R.T @ R
    s = len(S)
    cnt = dict(q=0, solve=0, fail=0)
    def Qeval(v, alpha, want=False):
        cnt['q'] += 1
        x = V @ v
        if variant == "plusminus":
            ep = np.exp(x); em = np.exp(-x)
            H = D * (ep - em); w = D * (ep + em)
            Hp = D * ep; Hm = D * em
            lp = np.where(Hp / D <= 1e-100, np.log(1e-100), x); lm = np.where(Hm / D <= 1e-100, np.log(1e-100), -x)
            Sv = np.sum(Hp - D - Hp * lp) + np.sum(Hm - D - Hm * lm)
        else:
            H = D * np.exp(x); w = H
            lg = np.where(H / D <= 1e-100, np.log(1e-100), x)
            Sv = np.sum(H - D - H * lg)
        y = V.T @ H
        rr = R @ y - gt
        chi2 = rr @ rr + c
        Qv = 0.5 * chi2 - alpha * Sv
        if want:
            return Qv, chi2, Sv, H, w, rr
        return Qv
    def solve(J, mu, f):
        cnt['solve'] += 1
        A = J + mu * np.eye(s)
        if solver == "chol":
            try:
                L = np.linalg.cholesky(A)
            except np.linalg.LinAlgError:
                cnt['fail'] += 1
                return None
            return sl.cho_solve((L, True), f)
        return np.linalg.solve(A, f)
    H0 = D * delta
    if variant == "plusminus":
        v = V.T @ np.log((H0 + np.sqrt(H0**2 + 4 * D**2)) / (2 * D))
    else:
        v = V.T @ np.log(H0 / D)
    out = dict(chi2=[], S=[], Q=[], A=[], n_iter=[], conv=[])
    eps = np.finfo(float).eps
    nu = 1.3; max_mu = 1e20
    for a in alphas:
        alpha = a * len(G)
        mu = 1e-18
        Q1, chi2, Sv, H, w, rr = Qeval(v, alpha, True)
        Q0 = np.nan; conv = False
        for it in range(maxiter):
            g = R.T @ rr
            Z = V.T @ (w[:, None] * V)
            if variant == "bryan":
                f = g + alpha * v; J = Gam @ Z
            else:
                f = Z @ (g + alpha * v); J = Z @ Gam @ Z + alpha * Z
            with np.errstate(all='ignore'):
                conv = bool(np.max(np.abs(f)) < 1e-4 or abs(abs(Q0 - Q1) / Q1) < 1e-16)
            if conv: break
            Q0 = Q1
            def trial(mu_):
                dv = solve(J, mu_, f)
                if dv is None: return None, np.nan
                with np.errstate(all='ignore'):
                    return dv, Qeval(v - dv, alpha)
            dv, Q1 = trial(mu)
            while (Q1 > Q0 or np.isnan(Q1)) and mu < max_mu:
                mu *= nu; dv, Q1 = trial(mu)
            dv2, Q2 = trial(nu * mu)
            if Q2 < Q1:
                nuf = nu; mu *= nu; Q2 = Q1; dvnew = dv2
            else:
                nuf = 1.0 / nu; mu /= nuf; dvnew = dv
            Q1 = np.inf
            while Q2 < Q1 and mu < max_mu and mu > nu * eps:
                Q1 = Q2; dv = dvnew; mu *= nuf
                dvnew, Q2 = trial(mu)
            v = v - dv
            Q1, chi2, Sv, H, w, rr = Qeval(v, alpha, True)
        out['chi2'].append(chi2); out['S'].append(Sv); out['Q'].append(Q1); out['A'].append(H / delta)
        out['n_iter'].append(it + 1); out['conv'].append(conv)
    for k in out: out[k] = np.array(out[k])
    out['cnt'] = cnt; out['n_sv'] = s
    return out

if __name__ == "__main__":
    name = sys.argv[1]; solver = sys.argv[2] if len(sys.argv) > 2 else "chol"
    g = dict(np.load(os.path.join("tests/golden", name)))
    K = mo.tau_kernel(g["tau"], g["omega"], None)
    t0 = time.time()
    out = model_run(K, g["G"], g["err"], g["omega"], g["alpha_mesh"], float(g["reduce_singular_space"]), str(g["variant"]), solver)
    print("time", time.time() - t0, "n_sv", out['n_sv'], out['cnt'], "iters", out['n_iter'].sum())
    rc = np.abs(out['chi2'] / g['ref_chi2'] - 1)
    rA = np.max(np.abs(out['A'] - g['ref_A']), axis=1) / np.max(np.abs(g['ref_A']), axis=1)
    rS = np.abs(out['S'] / g['ref_S'] - 1)
    np.set_printoptions(linewidth=200, precision=2)
    print("alpha_eff", g['ref_alpha'])
    print("rel chi2", rc); print("rel A   ", rA); print("rel S", rS)
    print("n_iter", out['n_iter'])
    an = mo.analyze_all(g['ref_alpha'], out['chi2'], out['S'], np.full(len(rc), np.nan), out['A'])
    for k in ("LineFitAnalyzer", "Chi2CurvatureAnalyzer", "EntropyAnalyzer"):
        print(k, an[k]['alpha_index'], int(g.get('ref_idx_' + k, -1)))
