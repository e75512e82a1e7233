#!/usr/bin/env python
"""Generate ``tests/golden/*.npz`` by running the REAL reference (TRIQS/maxent 1.2.0, staged from
/root/reference by ``stage_reference.py``).  TEST INFRASTRUCTURE; runs only in the build container.

Each fixture holds the inputs (tau, G, err, omega, alpha mesh, options) and the reference's own
outputs (alpha, chi2, S, Q, probability, H, A, n_iter, analyzer picks).  Input data that comes from
the reference's test fixtures (``test/python/g_tau_semicircular.dat``, ``elementwise_g_tau.npz``)
is stored as arrays.

    python oracle/make_golden.py            # all cases (~1 min)
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from stage_reference import import_reference, REF  # noqa: E402

warnings.filterwarnings("ignore")
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def run_reference(tm_mod, tau, G, err, omega_pts, alpha_mesh, cost_function="normal", probability=None,
                  reduce_singular_space=1e-14, extra_analyzers=False, noise_floor=True, preblur_b=None,
                  make_minimizer=None):
    """TauMaxEnt run through the reference's public API (python/tau_maxent.py)."""
    m = tm_mod
    kw = dict(cost_function=cost_function, probability=probability,
              reduce_singular_space=reduce_singular_space)
    if make_minimizer is not None:
        kw["minimizer"] = make_minimizer()
    tm = m.TauMaxEnt(**kw)
    tm.set_verbosity(m.VerbosityFlags.Quiet)
    tm.set_G_tau_data(np.array(tau), np.array(G))
    tm.omega = m.DataOmegaMesh(np.array(omega_pts))
    tm.alpha_mesh = m.DataAlphaMesh(np.array(alpha_mesh))
    tm.set_error(err)
    if preblur_b is not None:            # doc/guide/preblur_example.py:46-52
        tm.A_of_H = m.PreblurA_of_H(b=preblur_b, omega=tm.omega)
        tm.K = m.PreblurKernel(K=tm.K, b=preblur_b)
    t0 = time.time()
    res = tm.run()
    wall = time.time() - t0
    out = dict(tau=np.array(tau), G=np.array(G), err=np.asarray(err, dtype=float), omega=np.array(omega_pts),
               alpha_mesh=np.array(alpha_mesh), variant=cost_function,
               use_probability=probability is not None, reduce_singular_space=reduce_singular_space,
               ref_alpha=np.array(res.alpha), ref_chi2=np.array(res.chi2), ref_S=np.array(res.S),
               ref_Q=np.array(res.Q), ref_probability=np.array(res.probability, dtype=float),
               ref_H=np.array(res.H), ref_A=np.array(res.A), ref_v=np.array(res.v),
               ref_n_sv=len(tm.K.S), ref_wall=wall,
               ref_K_S=np.array(tm.K.S), ref_D=np.array(tm.D.D))
    for name, ar in res.analyzer_results.items():
        if isinstance(ar, dict) or hasattr(ar, "keys"):
            if "alpha_index" in ar:
                out["ref_idx_" + name] = int(ar["alpha_index"])
            if "A_out" in ar and ar["A_out"] is not None:
                out["ref_Aout_" + name] = np.array(ar["A_out"])
    if noise_floor:
        # The reference's own reproducibility (SURVEY.md section 0.4 / 8(c) tier T4): re-run it with
        # G * (1 + 1e-15) and record, per alpha, how far its A / chi2 / S move.
        if make_minimizer is not None:
            kw["minimizer"] = make_minimizer()
        tm2 = m.TauMaxEnt(**kw)
        tm2.set_verbosity(m.VerbosityFlags.Quiet)
        tm2.set_G_tau_data(np.array(tau), np.array(G) * (1.0 + 1.e-15))
        tm2.omega = m.DataOmegaMesh(np.array(omega_pts))
        tm2.alpha_mesh = m.DataAlphaMesh(np.array(alpha_mesh))
        tm2.set_error(err)
        if preblur_b is not None:
            tm2.A_of_H = m.PreblurA_of_H(b=preblur_b, omega=tm2.omega)
            tm2.K = m.PreblurKernel(K=tm2.K, b=preblur_b)
        res2 = tm2.run()
        A1, A2 = np.array(res.A), np.array(res2.A)
        out["noise_A"] = np.max(np.abs(A1 - A2), axis=1) / np.max(np.abs(A1), axis=1)
        out["noise_chi2"] = np.abs(np.array(res2.chi2) / np.array(res.chi2) - 1.0)
        out["noise_S"] = np.abs(np.array(res2.S) / np.array(res.S) - 1.0)
    return out, tm, res


def synthetic(n_tau, n_omega, seed=1234, mu=1.0, sigma=1e-4, beta=40.0):
    """SURVEY.md 8(d) recipe, built with the reference's own mesh/kernel classes."""
    m = import_reference()
    tau = np.linspace(0, beta, n_tau)
    omega = m.HyperbolicOmegaMesh(-10, 10, n_omega)
    K = m.TauKernel(tau, omega, beta)
    A = np.exp(-(omega - mu)**2 / (2 * 0.5**2))
    A /= np.trapz(A, omega)
    G_exact = np.dot(K.K_delta, A)
    np.random.seed(seed)
    G = G_exact + sigma * np.random.randn(n_tau)
    return tau, np.array(G), np.array(omega)


def main():
    m = import_reference()
    os.makedirs(GOLD, exist_ok=True)
    tests = os.path.join(REF, "test", "python")

    # --- G1: the reference's own known-answer test (test/python/tau_maxent.py:31-45,134-135) ----
    dat = np.loadtxt(os.path.join(tests, "g_tau_semicircular.dat"))
    np.random.seed(9)
    tau, G0 = dat[:, 0], dat[:, 1]
    G = G0 + 1.e-3 * np.random.randn(len(G0))
    omega = m.HyperbolicOmegaMesh(omega_min=-10, omega_max=10, n_points=200)
    amesh = m.LogAlphaMesh(alpha_min=0.08, n_points=5)
    out, _, _ = run_reference(m, tau, G, 1.e-3, omega, amesh, probability="normal")
    known = np.array([-8476.52812836, -2343.02752796, -704.28318351, -280.26627323, -175.30592555])
    np.testing.assert_almost_equal(out["ref_probability"], known, 6)
    out["known_probability"] = known
    np.savez_compressed(os.path.join(GOLD, "g1_semicircular_prob.npz"), **out)
    print("g1", out["ref_n_sv"], out["ref_wall"])

    # --- G2: synthetic 200 x 100, 20 alphas, probability, truncated singular space --------------
    tau, G, om = synthetic(200, 100)
    amesh = m.LogAlphaMesh(0.01, 2000, 20)
    out, _, _ = run_reference(m, tau, G, 1.e-4, om, amesh, probability="normal", reduce_singular_space=1e-11)
    np.savez_compressed(os.path.join(GOLD, "g2_synth_200x100.npz"), **out)
    print("g2", out["ref_n_sv"], out["ref_wall"], out.get("ref_idx_LineFitAnalyzer"))

    # --- G3: plusminus, off-diagonal element of the elementwise fixture -------------------------
    with np.load(os.path.join(tests, "elementwise_g_tau.npz")) as data:
        tau_e = data["tau"]
        G_e = data["G_tau_noise"]
    om80 = m.HyperbolicOmegaMesh(omega_min=-10, omega_max=10, n_points=80)
    am8 = m.LogAlphaMesh(alpha_min=0.05, alpha_max=500, n_points=8)
    out, _, _ = run_reference(m, tau_e, G_e[0, 1], 1.e-3, om80, am8, cost_function="plusminus")
    np.savez_compressed(os.path.join(GOLD, "g3_plusminus_offdiag.npz"), **out)
    print("g3", out["ref_n_sv"], out["ref_wall"])

    # --- G4: Bryan cost function on the G2 data ---------------------------------------------------
    out, _, _ = run_reference(m, tau, G, 1.e-4, om, amesh, cost_function="bryan", reduce_singular_space=1e-11)
    np.savez_compressed(os.path.join(GOLD, "g4_bryan_200x100.npz"), **out)
    print("g4", out["ref_n_sv"], out["ref_wall"])

    # --- G5: BASELINE config 1 (n_tau=1000, n_omega=400, 60 alphas), cut 1e-11 ---------------------
    tau1, G1, om1 = synthetic(1000, 400)
    am60 = m.LogAlphaMesh(0.01, 2000, 60)
    out, _, _ = run_reference(m, tau1, G1, 1.e-4, om1, am60, reduce_singular_space=1e-11)
    for k in ("ref_H", "ref_v"):
        out.pop(k)                      # A is enough; keeps the fixture small
    np.savez_compressed(os.path.join(GOLD, "g5_config1_cut1e-11.npz"), **out)
    print("g5", out["ref_n_sv"], out["ref_wall"], out.get("ref_idx_LineFitAnalyzer"),
          out.get("ref_idx_Chi2CurvatureAnalyzer"), out.get("ref_idx_EntropyAnalyzer"))

    # --- G5b: same with the reference's default cut 1e-14 (n_sv = 76) -----------------------------
    out, _, _ = run_reference(m, tau1, G1, 1.e-4, om1, am60, reduce_singular_space=1e-14)
    for k in ("ref_H", "ref_v"):
        out.pop(k)
    np.savez_compressed(os.path.join(GOLD, "g5b_config1_default_cut.npz"), **out)
    print("g5b", out["ref_n_sv"], out["ref_wall"], out.get("ref_idx_LineFitAnalyzer"))

    elementwise_cases(m)
    preblur_case(m)
    marquardt_case(m)
    mesh_case(m)
    covariance_case(m)
    iomega_case(m)
    low_temperature_case(m)
    config4_case(m)


def preblur_case(m):
    """G7: preblur formalism (PreblurKernel + PreblurA_of_H, b = 0.3) on the G2 data."""
    tau, G, om = synthetic(200, 100)
    amesh = m.LogAlphaMesh(0.05, 2000, 14)
    out, tm, res = run_reference(m, tau, G, 1.e-4, om, amesh, reduce_singular_space=1e-11, preblur_b=0.3)
    out["preblur_b"] = 0.3
    out["ref_G_rec_last"] = np.array(res.G_rec)[-1]
    np.savez_compressed(os.path.join(GOLD, "g7_preblur_200x100.npz"), **out)
    print("g7", out["ref_n_sv"], out["ref_wall"], out.get("ref_idx_LineFitAnalyzer"), out["noise_A"])


def marquardt_case(m):
    """G8: LevenbergMinimizer(marquardt=True) with MaxDerivative(1e-4) | FunctionChange(1e-9) on the G2 data
    (python/minimizers/levenberg_minimizer.py:181-185, convergence_methods.py:100-110)."""
    tau, G, om = synthetic(200, 100)
    amesh = m.LogAlphaMesh(0.01, 2000, 20)
    mk = lambda: m.LevenbergMinimizer(marquardt=True, convergence=m.OrConvergenceMethod(
        m.MaxDerivativeConvergenceMethod(1.e-4), m.FunctionChangeConvergenceMethod(1.e-9)))
    out, tm, res = run_reference(m, tau, G, 1.e-4, om, amesh, reduce_singular_space=1e-11, make_minimizer=mk)
    out["lm_marquardt"] = True
    out["lm_abs_change"] = 1.e-9
    np.savez_compressed(os.path.join(GOLD, "g8_marquardt_200x100.npz"), **out)
    print("g8", out["ref_n_sv"], out["ref_wall"], out.get("ref_idx_LineFitAnalyzer"), out["noise_A"])


def mesh_case(m):
    """G0: the host-side objects of tier T0 as the real reference builds them: the four omega meshes
    with their integration weights (python/omega_meshes.py), the alpha meshes (python/alpha_meshes.py), the flat
    default model (python/default_models.py:61-63) and a small TauKernel with K_delta (python/kernels.py:244-266)."""
    out = {}
    for name in ("LinearOmegaMesh", "LorentzianOmegaMesh", "LorentzianSmallerOmegaMesh", "HyperbolicOmegaMesh"):
        for (lo, hi, n) in ((-10, 10, 10), (-7.5, 12.25, 57)):
            mesh = getattr(m, name)(omega_min=lo, omega_max=hi, n_points=n)
            key = "%s_%d" % (name, n)
            out[key] = np.array(mesh)
            out[key + "_delta"] = np.array(mesh.delta)
            out[key + "_flatD"] = np.array(m.FlatDefaultModel(omega=mesh).D)
    out["LogAlphaMesh"] = np.array(m.LogAlphaMesh(alpha_min=0.0001, alpha_max=20, n_points=20))
    out["LogAlphaMesh_60"] = np.array(m.LogAlphaMesh(0.01, 2000, 60))
    out["LinearAlphaMesh"] = np.array(m.LinearAlphaMesh(alpha_min=0.0001, alpha_max=20, n_points=20))
    tau = np.linspace(0, 7.5, 23)
    om = m.HyperbolicOmegaMesh(-6, 6, 31)
    K = m.TauKernel(tau, om, 7.5)
    out["kernel_tau"], out["kernel_omega"], out["kernel_K"], out["kernel_K_delta"] = tau, np.array(om), np.array(K.K), np.array(K.K_delta)
    # DataDefaultModel on a different grid (python/default_models.py:83-113) and the preblur matrix (python/preblur.py:33-58)
    om_in = np.linspace(-8, 8, 41)
    dens = np.exp(-om_in**2 / 3.0) + 0.01
    out["ddm_omega_in"], out["ddm_default"] = om_in, dens
    out["ddm_D"] = np.array(m.DataDefaultModel(dens, om_in, om).D)
    out["ddm_D_same_grid"] = np.array(m.DataDefaultModel(np.exp(-np.array(om)**2) + 0.1, om, om).D)
    out["preblur_B"] = np.array(m.get_preblur(om, 0.4))
    np.savez_compressed(os.path.join(GOLD, "g0_meshes.npz"), **out)
    print("g0", len(out))


def covariance_case(m):
    """G9: TauMaxEnt.set_cov with a full (correlated) covariance matrix (python/tau_maxent.py:253-288,
    python/kernels.py:160-180) on the 200 x 100 problem: C_ij = sigma^2 (delta_ij + 0.4 exp(-|i-j| / 2.5))."""
    n_tau, n_omega, sigma = 200, 100, 1.e-4
    tau, _, om = synthetic(n_tau, n_omega)
    K = m.TauKernel(tau, m.DataOmegaMesh(om), 40.0)
    A = np.exp(-(om - 1.0)**2 / (2 * 0.5**2))
    A /= np.trapz(A, om)
    i = np.arange(n_tau)
    C = sigma**2 * (np.eye(n_tau) + 0.4 * np.exp(-np.abs(i[:, None] - i[None, :]) / 2.5))
    G = np.dot(K.K_delta, A) + np.dot(np.linalg.cholesky(C), np.random.RandomState(77).randn(n_tau))
    amesh = np.array(m.LogAlphaMesh(0.05, 2000, 16))

    def run(Gin):
        tm = m.TauMaxEnt(reduce_singular_space=1e-11)
        tm.set_verbosity(m.VerbosityFlags.Quiet)
        tm.set_G_tau_data(np.array(tau), np.array(Gin))
        tm.omega = m.DataOmegaMesh(np.array(om))
        tm.alpha_mesh = m.DataAlphaMesh(amesh)
        tm.set_cov(np.array(C))
        return tm, tm.run()

    tm, res = run(G)
    _, res2 = run(G * (1.0 + 1.e-15))
    A1, A2 = np.array(res.A), np.array(res2.A)
    out = dict(tau=tau, G=G, cov=C, omega=np.array(om), alpha_mesh=amesh, variant="normal", use_probability=False,
               reduce_singular_space=1e-11, err=np.array(tm.err), ref_alpha=np.array(res.alpha),
               ref_chi2=np.array(res.chi2), ref_S=np.array(res.S), ref_Q=np.array(res.Q), ref_A=A1, ref_H=np.array(res.H),
               ref_probability=np.array(res.probability, dtype=float), ref_n_sv=len(tm.K.S),
               ref_G_rotated=np.array(tm.G), ref_K_rotated_row0=np.array(tm.K.K)[0],
               noise_A=np.max(np.abs(A1 - A2), axis=1) / np.max(np.abs(A1), axis=1),
               noise_chi2=np.abs(np.array(res2.chi2) / np.array(res.chi2) - 1.0),
               noise_S=np.abs(np.array(res2.S) / np.array(res.S) - 1.0))
    for name, ar in res.analyzer_results.items():
        if hasattr(ar, "keys") and "alpha_index" in ar:
            out["ref_idx_" + name] = int(ar["alpha_index"])
            if ar.get("A_out", None) is not None:
                out["ref_Aout_" + name] = np.array(ar["A_out"])
    np.savez_compressed(os.path.join(GOLD, "g9_covariance_200x100.npz"), **out)
    print("g9", out["ref_n_sv"], out.get("ref_idx_LineFitAnalyzer"), out["noise_A"])


def _matrix_run(m, cls, tau, G, scale=1.0, **kw):
    """One elementwise-type run of the real reference on a matrix-valued G(tau) (test/python/elementwise_maxent.py:101-147)."""
    ew = getattr(m, cls)(**kw)
    ew.set_verbosity(m.VerbosityFlags.Quiet)
    ew.set_G_tau_data(np.array(tau), np.array(G) * scale)
    ew.omega = m.HyperbolicOmegaMesh(omega_min=-10, omega_max=10, n_points=80)
    ew.alpha_mesh = m.LogAlphaMesh(alpha_min=0.05, alpha_max=500, n_points=8)
    ew.set_error(1.e-3)
    return ew, ew.run()


def _matrix_fixture(m, cls, tau, G, **kw):
    """Fixture of a matrix run + the reference's own reproducibility (re-run with G * (1 + 1e-15)), element by element."""
    ew, res = _matrix_run(m, cls, tau, G, **kw)
    _, res2 = _matrix_run(m, cls, tau, G, scale=1.0 + 1.e-15, **kw)
    A1, A2 = np.array(res.A), np.array(res2.A)
    c1, c2 = np.array(res.chi2), np.array(res2.chi2)
    with np.errstate(all="ignore"):
        nA = np.max(np.abs(A1 - A2), axis=-1) / np.max(np.abs(A1), axis=-1)
        nc = np.abs(c2 / c1 - 1.0)
    out = dict(tau=np.array(tau), G=np.array(G), err=1.e-3, omega=np.array(ew.omega), alpha_mesh=np.array(ew.alpha_mesh),
               ref_alpha=np.array(res.alpha), ref_chi2=c1, ref_S=np.array(res.S), ref_A=A1, ref_A_out=np.array(res.A_out),
               noise_A=nA, noise_chi2=nc, options=repr(sorted(kw.items())))
    n = np.array(G).shape[0]
    idx = np.full((n, n, 2), -1, dtype=np.int64)
    for i in range(n):
        for j in range(n):
            ar = res.analyzer_results[i][j]
            if isinstance(ar, (list, tuple)):                       # complex elements: [re, im]
                for c, a in enumerate(ar):
                    if a and "LineFitAnalyzer" in a:
                        idx[i, j, c] = int(a["LineFitAnalyzer"]["alpha_index"])
            elif ar and "LineFitAnalyzer" in ar:
                idx[i, j, 0] = int(ar["LineFitAnalyzer"]["alpha_index"])
    out["ref_idx_LineFitAnalyzer"] = idx
    return out


def elementwise_cases(m):
    """G6 (BASELINE config 2 = ElementwiseMaxEnt on the reference's 2x2 fixture, now with the reference's own noise
    floor), G10 (PoormanMaxEnt, python/elementwise_maxent.py:562-653) and G11 (complex matrix elements,
    use_complex=True, python/elementwise_maxent.py:244-268, for both front ends): the front ends of
    test/python/elementwise_maxent.py:101-147 and test/python/complex_elementwise_maxent.py:76-137."""
    tests = os.path.join(REF, "test", "python")
    with np.load(os.path.join(tests, "elementwise_g_tau.npz")) as data:
        tau_e = data["tau"]
        G_e = data["G_tau_noise"]
    out = _matrix_fixture(m, "ElementwiseMaxEnt", tau_e, G_e, use_hermiticity=True)
    for i in range(2):
        for j in range(2):
            if out["ref_idx_LineFitAnalyzer"][i, j, 0] >= 0:
                out["ref_idx_LineFitAnalyzer_%d%d" % (i, j)] = int(out["ref_idx_LineFitAnalyzer"][i, j, 0])
    np.savez_compressed(os.path.join(GOLD, "g6_elementwise_2x2.npz"), **out)
    print("g6", out["noise_A"].max(), out["ref_idx_LineFitAnalyzer"][..., 0].tolist())
    out = _matrix_fixture(m, "PoormanMaxEnt", tau_e, G_e, use_hermiticity=False)
    np.savez_compressed(os.path.join(GOLD, "g10_poorman_2x2.npz"), **out)
    print("g10", out["noise_A"].max(), out["ref_idx_LineFitAnalyzer"][..., 0].tolist())
    # complex Hermitian matrix: the real fixture rotated by diag(1, exp(0.3 i))  (the reference's own complex test
    # builds its data with TRIQS, test/python/complex_elementwise_maxent.py:41-87)
    ph = np.exp(0.3j)
    G_c = np.array(G_e, dtype=complex)
    G_c[0, 1] = G_e[0, 1] * np.conjugate(ph)
    G_c[1, 0] = G_e[1, 0] * ph
    out = _matrix_fixture(m, "ElementwiseMaxEnt", tau_e, G_c, use_hermiticity=False, use_complex=True)
    np.savez_compressed(os.path.join(GOLD, "g11_complex_elementwise_2x2.npz"), **out)
    print("g11", out["ref_A"].shape, out["noise_A"].max(), out["ref_idx_LineFitAnalyzer"].tolist())
    out = _matrix_fixture(m, "PoormanMaxEnt", tau_e, G_c, use_hermiticity=True, use_complex=True)
    np.savez_compressed(os.path.join(GOLD, "g12_complex_poorman_2x2.npz"), **out)
    print("g12", out["ref_A"].shape, out["noise_A"].max(), out["ref_idx_LineFitAnalyzer"].tolist())


def config4_case(m):
    """G13: BASELINE config 4 (n_tau = 10000, n_omega = 2000, 100 alphas, probability, cut 1e-11) -- one run of the real
    reference (~10 min on 8 cores) plus its reproducibility run; A is stored at the analyzer picks and every tenth alpha."""
    tau, G, om = synthetic(10000, 2000)
    amesh = m.LogAlphaMesh(0.01, 2000, 100)
    out, tm, res = run_reference(m, tau, G, 1.e-4, om, amesh, probability="normal", reduce_singular_space=1e-11,
                                 noise_floor=True)
    keep = sorted(set(list(range(0, 100, 10)) + [99] + [int(v) for k, v in out.items() if k.startswith("ref_idx_")]))
    out["A_rows"] = np.array(keep)
    out["ref_A"] = out["ref_A"][keep]
    for k in ("ref_H", "ref_v", "tau", "omega", "G"):
        out.pop(k)                                   # rebuilt from the recipe (seed 1234) in the test
    out["G_sha_check"] = float(np.sum(G))
    np.savez_compressed(os.path.join(GOLD, "g13_config4_10000x2000.npz"), **out)
    print("g13", out["ref_n_sv"], out["ref_wall"], {k: v for k, v in out.items() if k.startswith("ref_idx_")})


def low_temperature_case(m):
    """G15: a kernel that keeps more than 80 singular values -- beta = 1000, n_tau = 1000, n_omega = 500, the reference's
    default cut 1e-14 (83 values with LAPACK; 81 above the numerical rank floor) -- through TauMaxEnt, 30 alphas."""
    tau, G, om = synthetic(1000, 500, beta=1000.0)
    amesh = m.LogAlphaMesh(0.01, 2000, 30)
    out, tm, res = run_reference(m, tau, G, 1.e-4, om, amesh, reduce_singular_space=1e-14)
    for k in ("ref_H", "ref_v"):
        out.pop(k)
    np.savez_compressed(os.path.join(GOLD, "g15_low_temperature_wide.npz"), **out)
    print("g15", out["ref_n_sv"], out["ref_wall"], {k: v for k, v in out.items() if k.startswith("ref_idx_")})


def iomega_case(m):
    """G14: continuation of Matsubara data G(i omega_n) with a REAL spectral function.  The reference's IOmegaKernel
    (python/kernels.py:283-346) supplies the complex kernel; chi2 = sum |G - K H|^2 / sigma^2 (ComplexChi2.f,
    python/functions.py:401-404) for a real H is NormalChi2 on the stacked rows [Re; Im], which the reference runs
    through a DataKernel of that stacked matrix (MaxEntLoop driven as in test/python/maxent_loop.py:49-76)."""
    beta, n_iw, n_omega, sigma = 40.0, 200, 100, 1.e-4
    iomega = (2 * np.arange(n_iw) + 1) * np.pi / beta
    omega = m.HyperbolicOmegaMesh(omega_min=-10, omega_max=10, n_points=n_omega)
    K_iw = m.IOmegaKernel(iomega, omega, beta=beta)
    A = np.exp(-(omega - 1.0)**2 / (2 * 0.5**2))
    A /= np.trapz(A, omega)
    rng = np.random.RandomState(4711)
    G = np.dot(K_iw.K_delta, A) + sigma * (rng.randn(n_iw) + 1j * rng.randn(n_iw))
    Ks = np.vstack([K_iw.K.real, K_iw.K.imag])
    Gs = np.concatenate([G.real, G.imag])
    errs = sigma * np.ones(2 * n_iw)
    amesh = m.LogAlphaMesh(0.01, 2000, 20)

    def run(Gin):
        K = m.DataKernel(np.arange(2 * n_iw, dtype=float), omega, Ks)
        D = m.FlatDefaultModel(omega=omega)
        Q = m.MaxEntCostFunction(chi2=m.NormalChi2(K=K, G=Gin, err=errs), S=m.NormalEntropy(D=D), H_of_v=m.NormalH_of_v(D=D, K=K))
        lt = m.Logtaker()
        lt.verbose = m.VerbosityFlags.Quiet
        ml = m.MaxEntLoop(cost_function=Q, alpha_mesh=amesh, logtaker=lt, reduce_singular_space=1e-11,
                          scale_alpha=float(n_iw), probability="normal")
        return K, ml.run()

    K, res = run(Gs)
    _, res2 = run(Gs * (1.0 + 1.e-15))
    A1, A2 = np.array(res.A), np.array(res2.A)
    out = dict(iomega=iomega, beta=beta, G=G, err=sigma, omega=np.array(omega), alpha_mesh=np.array(amesh), variant="normal",
               use_probability=True, reduce_singular_space=1e-11, scale_alpha=float(n_iw),
               ref_alpha=np.array(res.alpha), ref_chi2=np.array(res.chi2), ref_S=np.array(res.S), ref_Q=np.array(res.Q),
               ref_A=A1, ref_H=np.array(res.H), ref_probability=np.array(res.probability, dtype=float), ref_n_sv=len(K.S),
               ref_K_row0=np.array(K_iw.K)[0], ref_K_delta_row0=np.array(K_iw.K_delta)[0],
               noise_A=np.max(np.abs(A1 - A2), axis=1) / np.max(np.abs(A1), axis=1),
               noise_chi2=np.abs(np.array(res2.chi2) / np.array(res.chi2) - 1.0),
               noise_S=np.abs(np.array(res2.S) / np.array(res.S) - 1.0))
    for name, ar in res.analyzer_results.items():
        if hasattr(ar, "keys") and "alpha_index" in ar:
            out["ref_idx_" + name] = int(ar["alpha_index"])
            if ar.get("A_out", None) is not None:
                out["ref_Aout_" + name] = np.array(ar["A_out"])
    np.savez_compressed(os.path.join(GOLD, "g14_iomega_200x100.npz"), **out)
    print("g14", out["ref_n_sv"], {k: v for k, v in out.items() if k.startswith("ref_idx_")}, out["noise_A"].max())


if __name__ == "__main__":
    if "--low-temperature-only" in sys.argv:
        low_temperature_case(import_reference())
    elif "--iomega-only" in sys.argv:
        iomega_case(import_reference())
    elif "--elementwise-only" in sys.argv:
        elementwise_cases(import_reference())
    elif "--config4-only" in sys.argv:
        config4_case(import_reference())
    elif "--covariance-only" in sys.argv:
        covariance_case(import_reference())
    elif "--meshes-only" in sys.argv:
        mesh_case(import_reference())
    elif "--marquardt-only" in sys.argv:
        marquardt_case(import_reference())
    elif "--preblur-only" in sys.argv:
        preblur_case(import_reference())
    else:
        main()
