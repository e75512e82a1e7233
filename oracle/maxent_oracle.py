"""CPU oracle for the MaxEnt alpha-sweep hot path.  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

This file is a plain-numpy *restatement* of the algorithm TRIQS/maxent (v1.2.0) runs for
``MaxEntLoop.run`` and everything below it.  It exists only so that

* ``tests/`` can compare the CUDA path against it,
* ``__graft_entry__.smoke()`` can check one small run against it,
* ``bench.py`` can time it as the host-CPU baseline (``cpu_baseline`` / ``--impl reference``).

Nothing under ``maxent_b200/`` may import it; the product path has no CPU fallback.

Parity pin: ``oracle/make_golden.py`` runs the real reference (staged from ``/root/reference`` by
``oracle/stage_reference.py``) and stores its inputs and outputs as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` requires this file to reproduce those chi2/S/Q/H/probability arrays and
analyzer picks *bit for bit*, and checks the reference's own known-answer numbers
(``test/python/tau_maxent.py:134-135``).

The formulation is deliberately the reference's own (full kernel GEMVs, dense n_omega x n_omega
Hessian, LU solves through ``np.linalg.solve``) -- NOT the singular-space formulation the CUDA
kernels use -- so that (a) the comparison is between two independent derivations and (b) the CPU
baseline does the work the reference does.  Citations are ``file:line`` under ``/root/reference``.
"""
import time

import numpy as np

# --------------------------------------------------------------------------------------------
# meshes, default model, kernel
# --------------------------------------------------------------------------------------------


def hyperbolic_omega_mesh(omega_min=-10.0, omega_max=10.0, n_points=100):
    """python/omega_meshes.py:215-222"""
    u = np.linspace(-1, 1, n_points)
    w = np.sign(u) * (np.sqrt(1 + u**2) - 1)
    return omega_min + (omega_max - omega_min) * (w - w[0]) / (w[-1] - w[0])


def linear_omega_mesh(omega_min=-10.0, omega_max=10.0, n_points=100):
    """python/omega_meshes.py:82-84"""
    return np.linspace(omega_min, omega_max, n_points)


def lorentzian_omega_mesh(omega_min=-10.0, omega_max=10.0, n_points=100, cut=0.01, smaller=False):
    """python/omega_meshes.py:124-137 (and 171-183 for the 'smaller' variant)"""
    u = np.linspace(0, 1, n_points + 1)
    temp = np.tan(np.pi * (u * (1. - 2 * cut) + cut - 0.5))
    t = (temp - temp[0]) / (temp[-1] - temp[0])
    w = omega_min + (omega_max - omega_min) * t
    w = (w[:-1] + w[1:]) / 2.0
    if not smaller:
        w = (w - w[0]) / (w[-1] - w[0]) * (omega_max - omega_min) + omega_min
    return w


def omega_delta(omega):
    """python/omega_meshes.py:54-62 : trapezoid weights"""
    omega = np.asarray(omega, dtype=float)
    delta = np.empty(len(omega))
    delta[1:-1] = (omega[2:] - omega[:-2]) / 2.0
    delta[0] = (omega[1] - omega[0]) / 2.0
    delta[-1] = (omega[-1] - omega[-2]) / 2.0
    return delta


def log_alpha_mesh(alpha_min=0.0001, alpha_max=20, n_points=20):
    """python/alpha_meshes.py:81-85 (descending)"""
    return np.logspace(np.log10(alpha_min), np.log10(alpha_max), n_points)[::-1].copy()


def linear_alpha_mesh(alpha_min=0.0001, alpha_max=20, n_points=20):
    """python/alpha_meshes.py:101-103"""
    return np.linspace(alpha_min, alpha_max, n_points)[::-1].copy()


def flat_default_model(omega):
    """python/default_models.py:61-63 : D already contains delta omega"""
    delta = omega_delta(omega)
    return np.ones(np.shape(omega)) / np.sum(delta) * delta


def data_default_model(default, omega_in, omega=None):
    """python/default_models.py:88-93"""
    if omega is None:
        omega = omega_in
    if len(omega_in) == len(omega) and np.all(np.asarray(omega_in) == np.asarray(omega)):
        D = np.asarray(default, dtype=float)
    else:
        D = np.interp(omega, omega_in, default)
    return D * omega_delta(omega)


def tau_kernel(tau, omega, beta=None):
    """python/kernels.py:244-263 : K(tau, omega), two branches by the sign of omega"""
    tau = np.asarray(tau, dtype=float)
    omega = np.asarray(omega, dtype=float)
    if beta is None:
        beta = tau[-1]
    oomega, ttau = np.meshgrid(omega, tau)
    L = oomega >= 0.0
    iL = np.where(L)
    nL = np.where(np.logical_not(L))
    K = np.empty(oomega.shape)
    K[iL] = -np.exp(-oomega[iL] * ttau[iL]) / (np.exp(-beta * oomega[iL]) + 1.0)
    K[nL] = -np.exp(oomega[nL] * (beta - ttau[nL])) / (1.0 + np.exp(beta * oomega[nL]))
    return K


def kernel_svd(K, threshold=1.e-14):
    """python/kernels.py:53-64,101-122 : thin SVD, V as n_omega x k, keep S >= threshold (absolute)"""
    U, S, Vt = np.linalg.svd(K, full_matrices=False)
    V = Vt.transpose()
    if threshold is None:
        return U, S, V
    L = np.where(S >= threshold)[0]
    return U[:, L], S[L], V[:, L]


def safelog(A):
    """python/functions.py:53-56 (clamps its argument IN PLACE)"""
    A[np.where(np.abs(A) <= 1.e-100)] = 1.e-100
    return np.log(A)


# --------------------------------------------------------------------------------------------
# cost function  (python/cost_functions/*.py on top of python/functions.py)
# --------------------------------------------------------------------------------------------

class Problem(object):
    """Everything one MaxEntLoop.run needs that does not depend on alpha or v."""

    def __init__(self, K, G, err, D, delta, U, S, V, variant="normal", chi2_factor=1.0, fast_d2=False):
        self.K = np.asarray(K, dtype=float)
        self.G = np.asarray(G, dtype=float)
        self.err = np.asarray(err, dtype=float) * np.ones(self.G.shape)
        self.D = np.asarray(D, dtype=float)
        self.delta = np.asarray(delta, dtype=float)
        self.U, self.S, self.V = U, S, V
        self.variant = variant
        self.chi2_factor = chi2_factor
        if variant not in ("normal", "plusminus", "bryan"):
            raise ValueError("unknown variant %r" % (variant,))
        # NormalChi2.parameter_change, python/functions.py:372-377 : d2 = 2 K^T W K (constant)
        if fast_d2:
            self.d2 = 2.0 * np.dot(self.K.T * (1. / self.err**2), self.K)
        else:
            self.d2 = 2 * np.einsum('il,ik,i->kl', np.conjugate(self.K), self.K, 1. / self.err**2)
        self.n_qeval = 0
        self.n_solve = 0

    # ---- H(v) --------------------------------------------------------------------------
    def H_of_v(self, v):
        x = np.dot(self.V, v)
        if self.variant == "plusminus":          # python/functions.py:778-781
            return self.D * (np.exp(x) - np.exp(-np.dot(self.V, v)))
        return self.D * np.exp(x)                # python/functions.py:739-741

    def dH_dv(self, v):
        if self.variant == "plusminus":          # python/functions.py:783-786
            return self.D[:, np.newaxis] * self.V * (
                np.exp(np.dot(self.V, v))[:, np.newaxis] + np.exp(-np.dot(self.V, v))[:, np.newaxis])
        return self.D[:, np.newaxis] * self.V * np.exp(np.dot(self.V, v))[:, np.newaxis]  # :743-746

    def v_of_H(self, H):
        if self.variant == "plusminus":          # python/functions.py:793-796
            return np.dot(self.V.transpose().conjugate(),
                          safelog((H + np.sqrt(H**2 + 4 * self.D**2)) / (2 * self.D)))
        return np.dot(self.V.transpose(), safelog(H / self.D))   # python/functions.py:753-755

    # ---- chi2 --------------------------------------------------------------------------
    def chi2_f(self, H):
        # python/functions.py:358-360 -- the reference uses the python builtin sum(), i.e. strictly
        # sequential accumulation; cumsum()[-1] is the same left-to-right sum in C.
        t = np.abs(np.dot(self.K, H) - self.G)**2 / self.err**2
        return np.cumsum(t)[-1]

    def chi2_d(self, H):
        # python/functions.py:362-365
        return np.dot(2 * (np.dot(self.K, H) - self.G) / self.err**2, np.conjugate(self.K))

    # ---- entropy -----------------------------------------------------------------------
    def _S_normal_f(self, A):
        return np.sum((A - self.D - A * safelog(A / self.D)))      # python/functions.py:508-510

    def _S_normal_d(self, A):
        return - (safelog(A) - safelog(self.D))                    # python/functions.py:512-514

    def _S_normal_dd(self, A):
        A[np.where(np.abs(A) <= 1.e-100)] = 1.e-100                # python/functions.py:516-520
        return -np.diag(1.0 / A)

    def _A_plus(self, A):
        return (np.sqrt(A**2.0 + 4.0 * self.D**2) + A) / 2.0       # python/functions.py:544-546

    def _A_minus(self, A):
        return (np.sqrt(A**2.0 + 4.0 * self.D**2) - A) / 2.0       # python/functions.py:548-550

    def S_f(self, H):
        if self.variant == "plusminus":                            # python/functions.py:552-555
            return self._S_normal_f(self._A_plus(H)) + self._S_normal_f(self._A_minus(H))
        return self._S_normal_f(H)

    def S_d(self, H):
        if self.variant == "plusminus":                            # python/functions.py:557-559
            return self._S_normal_d(self._A_plus(H))
        return self._S_normal_d(H)

    def S_dd(self, H):
        if self.variant == "plusminus":                            # python/functions.py:561-564
            return self._S_normal_dd(self._A_plus(H) + self._A_minus(H))
        return self._S_normal_dd(H)


class BoundQ(object):
    """One cost-function value bound to one v (what ``CostFunction.__call__`` returns,
    python/cost_functions/cost_function.py:73-85): H is computed once and shared."""

    def __init__(self, prob, alpha, v):
        self.p = prob
        self.alpha = alpha
        self.v = v.view()
        self.H = prob.H_of_v(v)
        self._f = self._chi2 = self._S = self._dH = self._d = self._ddH = self._dd = None

    def chi2(self):
        if self._chi2 is None:
            self._chi2 = self.p.chi2_f(self.H)
        return self._chi2

    def S(self):
        if self._S is None:
            self._S = self.p.S_f(self.H)
        return self._S

    def f(self):
        # python/cost_functions/maxent_cost_function.py:68-83 (bryan: bryan_cost_function.py:57-72)
        if self._f is None:
            self.p.n_qeval += 1
            self._f = 0.5 * self.chi2() * self.p.chi2_factor - self.alpha * self.S()
        return self._f

    def dH(self):
        # python/cost_functions/maxent_cost_function.py:85-93
        if self._dH is None:
            dchi2 = self.p.chi2_d(self.H)
            dS = self.p.S_d(self.H)
            self._dH = 0.5 * dchi2 * self.p.chi2_factor - self.alpha * dS
        return self._dH

    def d(self):
        if self._d is None:
            p = self.p
            if p.variant == "bryan":
                # python/cost_functions/bryan_cost_function.py:84-102
                dchi2 = 2 * (np.dot(p.K, self.H) - p.G) / p.err**2
                ret = p.S * np.dot(p.U.conjugate().transpose(), 0.5 * dchi2 * p.chi2_factor)
                self._d = -(-ret - self.alpha * self.v)
            else:
                # python/cost_functions/maxent_cost_function.py:95-121, dA_projection == 2
                dQ_dH = self.dH().reshape(-1)
                dH_dv = p.dH_dv(self.v)
                self._d = np.dot(dH_dv.transpose(), dQ_dH)
        return self._d

    def ddH(self):
        # python/cost_functions/maxent_cost_function.py:123-131
        if self._ddH is None:
            self._ddH = 0.5 * self.p.d2 * self.p.chi2_factor - self.alpha * self.p.S_dd(self.H)
        return self._ddH

    def dd(self):
        if self._dd is None:
            p = self.p
            if p.variant == "bryan":
                # python/cost_functions/bryan_cost_function.py:114-128
                ret = np.dot(p.V.conjugate().transpose(), p.d2)
                ret = np.einsum('ij,j,jk->ik', ret, self.H, p.V)
                self._dd = 0.5 * ret * p.chi2_factor
            else:
                # python/cost_functions/maxent_cost_function.py:133-165, dA_projection == 2
                ddQ = self.ddH()
                dH_dv = p.dH_dv(self.v)
                self._dd = np.dot(dH_dv.transpose(), np.dot(ddQ, dH_dv))
        return self._dd

    def log_probability(self):
        """python/probabilities.py:76-85 with the defaults of :61-67"""
        ddQ = self.ddH()
        _, pr = np.linalg.slogdet(ddQ)
        pr = -0.5 * pr
        pr += 1 / 2.0 * np.linalg.slogdet(-self.p.S_dd(self.H))[1]
        pr += (len(self.H) / 2.0) * np.log(self.alpha)
        pr -= self.f()
        pr += -np.log(self.alpha)
        return pr


# --------------------------------------------------------------------------------------------
# Levenberg minimiser  (python/minimizers/levenberg_minimizer.py:123-248)
# --------------------------------------------------------------------------------------------

def levenberg_minimize(prob, alpha, v, maxiter=1000, miniter=0, mu0=1.e-18, nu=1.3, max_mu=1.e20,
                       max_derivative=1.e-4, rel_change=1.e-16, abs_change=-1.0, marquardt=False):
    """Returns (v, converged, n_iter_last).  ``v`` is updated in place like the reference (:239).
    Convergence = MaxDerivative(1e-4) | RelativeFunctionChange(1e-16) [| FunctionChange(abs_change)]  (:103-106,
    python/minimizers/convergence_methods.py:64-122); ``marquardt`` damps with diag(J) (:181-185)."""
    converged = False
    mu = mu0
    func_val = BoundQ(prob, alpha, v)
    Q1 = func_val.f()
    Q0 = np.nan
    i = -1
    solve = np.linalg.solve
    for i in range(maxiter):
        f = func_val.d()
        J = func_val.dd()
        with np.errstate(all='ignore'):
            is1 = np.max(np.abs(f)) < max_derivative
            is2 = np.abs(np.abs(Q0 - Q1) / Q1) < rel_change
            is3 = np.abs(Q0 - Q1) < abs_change
        converged = bool(is1 or is2 or is3)
        if converged and i >= miniter:
            break
        Id = np.diag(np.diag(J)) if marquardt else np.eye(len(J))
        Q0 = Q1
        dv = solve(J + mu * Id, f); prob.n_solve += 1
        old = np.seterr(all='ignore')
        Q1 = BoundQ(prob, alpha, v - dv).f()
        while (Q1 > Q0 or np.isnan(Q1)) and mu < max_mu:
            mu *= nu
            dv = solve(J + mu * Id, f); prob.n_solve += 1
            Q1 = BoundQ(prob, alpha, v - dv).f()
        dv2 = solve(J + nu * mu * Id, f); prob.n_solve += 1
        Q2 = BoundQ(prob, alpha, v - dv2).f()
        if Q2 < Q1:
            nuf = nu
            mu *= nu
            Q2 = Q1
            dvnew = dv2
        else:
            nuf = 1.0 / nu
            mu /= nuf
            dvnew = dv
        Q1 = np.inf
        while (Q2 < Q1 and mu < max_mu and mu > nu * np.finfo(float).eps):
            Q1 = Q2
            dv = dvnew
            mu *= nuf
            dvnew = solve(J + mu * Id, f); prob.n_solve += 1
            Q2 = BoundQ(prob, alpha, v - dvnew).f()
        np.seterr(**old)
        v -= dv
        func_val = BoundQ(prob, alpha, v)
        Q1 = func_val.f()
    return v, converged, i + 1


# --------------------------------------------------------------------------------------------
# the alpha loop  (python/maxent_loop.py:144-302)
# --------------------------------------------------------------------------------------------

def maxent_loop(K, G, err, omega, alpha_mesh, D=None, variant="normal", probability=False,
                reduce_singular_space=1.e-14, scale_alpha="Ndata", A_init=None, G_threshold=1.e-10,
                chi2_factor=1.0, maxiter=1000, fast_d2=False, svd=None, analyzers=True, lm_options=None):
    """One ``MaxEntLoop.run``.  Returns a dict of arrays (the fields of MaxEntResultData,
    python/maxent_result.py:181-188) plus counters, or None if G is below threshold (:174-179)."""
    G = np.asarray(G, dtype=float)
    omega = np.asarray(omega, dtype=float)
    delta = omega_delta(omega)
    if D is None:
        D = flat_default_model(omega)
    if np.max(np.abs(G)) < G_threshold:
        return None
    if svd is None:
        U, S, V = kernel_svd(K, reduce_singular_space)               # :184
    else:
        U, S, V = svd
    prob = Problem(K, G, err, D, delta, U, S, V, variant, chi2_factor, fast_d2=fast_d2)
    # initial v (:196-203).  Quirk kept: D already holds delta, and is multiplied by delta again.
    H0 = np.empty(len(D))
    H0[:] = (D if A_init is None else np.asarray(A_init, dtype=float)) * delta
    v = prob.v_of_H(H0)
    if scale_alpha is None:
        scale = 1.0
    elif isinstance(scale_alpha, str):
        if scale_alpha.lower() != "ndata":
            raise Exception("Unknown value {} for scale_alpha".format(scale_alpha))
        scale = len(G)                                                # :216-220
    else:
        scale = scale_alpha
    n_a = len(alpha_mesh)
    out = dict(alpha=np.empty(n_a), chi2=np.empty(n_a), S=np.empty(n_a), Q=np.empty(n_a),
               probability=np.full(n_a, np.nan), v=np.empty((n_a, len(S))),
               H=np.empty((n_a, len(D))), A=np.empty((n_a, len(D))),
               n_iter=np.zeros(n_a, dtype=np.int64), converged=np.zeros(n_a, dtype=bool),
               omega=omega, n_sv=len(S))
    t0 = time.perf_counter()
    for ia, alpha in enumerate(alpha_mesh):                           # :241-266
        a_eff = alpha * scale
        v, conv, nit = levenberg_minimize(prob, a_eff, v, maxiter=maxiter, **(lm_options or {}))
        Qmin = BoundQ(prob, a_eff, v)
        out["alpha"][ia] = a_eff
        out["chi2"][ia] = Qmin.chi2()
        out["S"][ia] = Qmin.S()
        out["Q"][ia] = Qmin.f()
        out["v"][ia] = v
        out["H"][ia] = Qmin.H
        out["A"][ia] = Qmin.H / delta                                 # python/functions.py:947-952
        out["n_iter"][ia] = nit
        out["converged"][ia] = conv
        if probability:
            with np.errstate(all='ignore'):
                out["probability"][ia] = Qmin.log_probability()
    out["run_time"] = time.perf_counter() - t0
    out["n_qeval"] = prob.n_qeval
    out["n_solve"] = prob.n_solve
    if analyzers:
        out["analyzers"] = analyze_all(out["alpha"], out["chi2"], out["S"], out["probability"], out["A"])
    return out


# --------------------------------------------------------------------------------------------
# analyzers  (python/analyzers/*.py)
# --------------------------------------------------------------------------------------------

def fit_piecewise(logx, logy, p2_deg=0):
    """python/analyzers/linefit_analyzer.py:28-87"""
    chi2 = np.full(len(logx), np.nan)
    p1 = [0] * len(logx)
    p2 = [0] * len(logx)

    def denan(what, check=None):
        if check is None:
            check = what
        return what[np.logical_not(np.isnan(check))]
    for i in range(2, len(logx) - 2):
        chi2[i] = 0.0
        try:
            p1[i], residuals, _, _, _ = np.polyfit(denan(logx[:i], logy[:i]), denan(logy[:i]), deg=1, full=True)
            if len(residuals) > 0:
                chi2[i] += residuals[0]
            p2[i], residuals, _, _, _ = np.polyfit(denan(logx[i:], logy[i:]), denan(logy[i:]), deg=p2_deg, full=True)
            if len(residuals) > 0:
                chi2[i] += residuals[0]
        except TypeError:
            p1[i] = np.nan
            p2[i] = np.nan
            chi2[i] = np.nan
    i = np.nanargmin(chi2)
    X_x = ((p2[i][1] if p2_deg == 1 else p2[i][0]) - p1[i][1]) / \
        (p1[i][0] - (p2[i][0] if p2_deg == 1 else 0.0))
    idx = np.nanargmin(np.abs(logx - X_x))
    return int(idx), (p1[i], p2[i]), chi2


def curvature(x, y):
    """python/analyzers/chi2_curvature_analyzer.py:25-49"""
    n = len(x)
    der2 = np.full(n, np.nan)
    der1 = np.full(n, np.nan)
    for k in range(1, n - 1):
        der2[k] = (y[k + 1] - 2 * y[k] + y[k - 1]) / ((x[k + 1] - x[k]) * (x[k] - x[k - 1]))
        der1[k] = ((y[k + 1] - y[k]) / (x[k + 1] - x[k]) + (y[k] - y[k - 1]) / (x[k] - x[k - 1])) / 2
    return der2 / (1 + der1 * der1)**(3. / 2.), der1, der2


def analyze_linefit(alpha, chi2, A, linefit_deg=0):
    """python/analyzers/linefit_analyzer.py:151-183"""
    idx, params, _ = fit_piecewise(np.log(alpha), np.log(chi2), linefit_deg)
    return dict(alpha_index=idx, A_out=A[idx], linefit_params=params)


def analyze_chi2_curvature(alpha, chi2, A, gamma=0.2):
    """python/analyzers/chi2_curvature_analyzer.py:101-131"""
    c, _, _ = curvature(gamma * np.log10(alpha), np.log10(chi2))
    idx = int(np.nanargmax(c))
    return dict(alpha_index=idx, A_out=A[idx], curvature=c)


def analyze_entropy(alpha, S, A):
    """python/analyzers/entropy_analyzer.py:89-103"""
    dS = np.full(len(alpha), np.nan)
    dS[1:-1] = (S[2:] - S[:-2]) / (np.log(alpha[2:]) - np.log(alpha[:-2]))
    idx = int(np.nanargmin(dS**2))
    return dict(alpha_index=idx, A_out=A[idx], dS_dalpha=dS)


def analyze_classic(alpha, probability, A):
    """python/analyzers/classic_analyzer.py:67-82"""
    if np.all(np.isnan(probability)):
        return dict(alpha_index=-1, A_out=None)
    idx = int(np.nanargmax(probability))
    return dict(alpha_index=idx, A_out=A[idx])


def _get_delta(alpha):
    """python/analyzers/bryan_analyzer.py get_delta (trapezoid weights on the alpha mesh)"""
    d = np.empty(len(alpha))
    d[1:-1] = (alpha[2:] - alpha[:-2]) / 2.0
    d[0] = (alpha[1] - alpha[0]) / 2.0
    d[-1] = (alpha[-1] - alpha[-2]) / 2.0
    return d


def analyze_bryan(alpha, probability, A, average_by_integration=False):
    """python/analyzers/bryan_analyzer.py:123-154"""
    if np.all(np.isnan(probability)):
        return dict(A_out=None)
    A_out = np.zeros(A.shape[-1])
    prob = np.exp(probability - np.nanmax(probability))
    L = np.where(np.logical_not(np.isnan(prob)))
    if average_by_integration:
        trapz = getattr(np, "trapz", None) or np.trapezoid
        prob[L] /= trapz(prob[L], alpha[L])
        delta_alpha = np.full(len(prob), np.nan)
        delta_alpha[L] = _get_delta(alpha[L])
    else:
        prob[L] /= np.sum(prob[L])
    for i in range(len(alpha)):
        if np.isnan(prob[i]):
            continue
        if average_by_integration:
            A_out += prob[i] * A[i, :] * delta_alpha[i]
        else:
            A_out += prob[i] * A[i, :]
    return dict(A_out=A_out, weights=prob)


def analyze_all(alpha, chi2, S, probability, A, gamma=0.2, linefit_deg=0):
    res = {}
    with np.errstate(all='ignore'):
        for name, fn in (("LineFitAnalyzer", lambda: analyze_linefit(alpha, chi2, A, linefit_deg)),
                         ("Chi2CurvatureAnalyzer", lambda: analyze_chi2_curvature(alpha, chi2, A, gamma)),
                         ("EntropyAnalyzer", lambda: analyze_entropy(alpha, S, A)),
                         ("ClassicAnalyzer", lambda: analyze_classic(alpha, probability, A)),
                         ("BryanAnalyzer", lambda: analyze_bryan(alpha, probability, A))):
            try:
                res[name] = fn()
            except ValueError as e:                      # python/maxent_result.py:819-822
                res[name] = dict(error=str(e))
    return res


# --------------------------------------------------------------------------------------------
# synthetic benchmark problem  (SURVEY.md section 8(d) recipe; no reference counterpart)
# --------------------------------------------------------------------------------------------

def synthetic_problem(n_tau, n_omega, beta=40.0, mu=1.0, width=0.5, sigma=1.e-4, noise=None, seed=1234):
    """Gaussian A(omega) -> G(tau) through the TauKernel, plus sigma * noise."""
    tau = np.linspace(0, beta, n_tau)
    omega = hyperbolic_omega_mesh(-10, 10, n_omega)
    delta = omega_delta(omega)
    K = tau_kernel(tau, omega, beta)
    mu = np.atleast_1d(np.asarray(mu, dtype=float))
    A = np.exp(-(omega[None, :] - mu[:, None])**2 / (2 * width**2))
    trapz = getattr(np, "trapezoid", None) or np.trapz
    A /= trapz(A, omega, axis=1)[:, None]
    G_exact = np.dot(A, (K * delta[None, :]).T)
    if noise is None:
        rng = np.random.RandomState(seed)
        noise = rng.randn(*G_exact.shape)
    G = G_exact + sigma * np.asarray(noise).reshape(G_exact.shape)
    return dict(tau=tau, omega=omega, delta=delta, K=K, A_true=A, G=G, err=sigma, beta=beta)
