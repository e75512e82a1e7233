"""GPU tests of the drop-in interface (TauMaxEnt / MaxEntLoop / ElementwiseMaxEnt ...): they read like the
reference's own end-to-end tests (test/python/tau_maxent.py, elementwise_maxent.py, cov.py, huge_alpha.py,
bryan_cost_function.py, pickle_maxent_result.py) and compare with the outputs of the REAL reference stored in
tests/golden/*.npz under the tolerances of tests/gpu_common.py."""
import pickle

import numpy as np
import pytest

import maxent_b200 as mb
from oracle import maxent_oracle as mo
from tests import gpu_common as gc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    import torch
    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"


class _ResView(object):
    """Adapts a MaxEntResult to the tensor-like fields gc.check_against_reference reads."""

    class _T(object):
        def __init__(self, a):
            self.a = np.asarray(a)

        def __getitem__(self, i):
            return self

        def cpu(self):
            return self

        def numpy(self):
            return self.a

    def __init__(self, res):
        self.A, self.chi2, self.S, self.Q = (self._T(x) for x in (res.A, res.chi2, res.S, res.Q))
        self.logp = self._T(res.probability)
        names = gc.AN_NAMES
        ar = res.analyzer_results
        idx = [ar[n].get('alpha_index', -1) if n in ar and not isinstance(ar[n], str) else -1 for n in names]
        n_om = res.A.shape[-1]
        Aout = [ar[n]['A_out'] if n in ar and not isinstance(ar[n], str) and 'A_out' in ar[n] else np.full(n_om, np.nan)
                for n in names]
        self.alpha_index, self.A_out = self._T(np.array(idx)), self._T(np.array(Aout))


def _tau_maxent_from_fixture(g, **kw):
    tm = mb.TauMaxEnt(cost_function=str(g["variant"]), probability='normal' if bool(g["use_probability"]) else None,
                      reduce_singular_space=float(g["reduce_singular_space"]), **kw)
    tm.set_verbosity(mb.VerbosityFlags.Quiet)
    tm.set_G_tau_data(g["tau"], g["G"])
    tm.omega = mb.DataOmegaMesh(g["omega"])
    tm.alpha_mesh = mb.DataAlphaMesh(g["alpha_mesh"])
    tm.set_error(float(g["err"]) if np.ndim(g["err"]) == 0 else g["err"])
    return tm


@pytest.mark.parametrize("name", ["g1_semicircular_prob.npz", "g2_synth_200x100.npz", "g3_plusminus_offdiag.npz",
                                  "g4_bryan_200x100.npz", "g5_config1_cut1e-11.npz"])
def test_tau_maxent_matches_reference_run(name):
    """The same user script as oracle/make_golden.py ran on the reference, on this package."""
    g = gc.load_golden(name)
    tm = _tau_maxent_from_fixture(g)
    res = tm.run()
    assert len(tm.K.S) == int(g["ref_n_sv"])
    np.testing.assert_allclose(res.alpha, g["ref_alpha"], rtol=1e-14)
    np.testing.assert_allclose(tm.D.D, g["ref_D"], rtol=1e-13)
    gc.check_against_reference(g, _ResView(res), rtol_chi2_S=2e-7)
    assert res.default_analyzer_name == 'LineFitAnalyzer'
    np.testing.assert_array_equal(res.A_out, res.analyzer_results['LineFitAnalyzer']['A_out'])
    np.testing.assert_allclose(res.H, res.A * tm.omega.delta[None, :], rtol=1e-15)
    assert res.v.shape == (len(g["ref_alpha"]), int(g["ref_n_sv"]))
    assert np.all(res.converged)


def test_tau_maxent_marquardt_minimizer():
    """The user-level spelling of the G8 fixture: TauMaxEnt(minimizer=LevenbergMinimizer(marquardt=True,
    convergence=MaxDerivative(1e-4) | FunctionChange(1e-9)))."""
    g = gc.load_golden("g8_marquardt_200x100.npz")
    tm = _tau_maxent_from_fixture(g)
    tm.minimizer = mb.LevenbergMinimizer(marquardt=True, convergence=mb.OrConvergenceMethod(
        mb.MaxDerivativeConvergenceMethod(1.e-4), mb.FunctionChangeConvergenceMethod(1.e-9)))
    res = tm.run()
    gc.check_against_reference(g, _ResView(res), rtol_chi2_S=2e-7)
    assert np.all(res.converged)


def test_full_covariance_matches_reference_run():
    """TauMaxEnt.set_cov with a correlated covariance matrix (python/tau_maxent.py:253-288) against the run of the
    real reference stored in tests/golden/g9_covariance_200x100.npz."""
    g = gc.load_golden("g9_covariance_200x100.npz")
    tm = mb.TauMaxEnt(reduce_singular_space=float(g["reduce_singular_space"]))
    tm.set_verbosity(mb.VerbosityFlags.Quiet)
    tm.set_G_tau_data(g["tau"], g["G"])
    tm.omega = mb.DataOmegaMesh(g["omega"])
    tm.alpha_mesh = mb.DataAlphaMesh(g["alpha_mesh"])
    tm.set_cov(g["cov"])
    np.testing.assert_allclose(tm.err, g["err"], rtol=1e-12)
    np.testing.assert_allclose(tm.G, g["ref_G_rotated"], rtol=0, atol=1e-13)
    res = tm.run()
    assert len(tm.K.S) == int(g["ref_n_sv"])
    np.testing.assert_allclose(res.alpha, g["ref_alpha"], rtol=1e-14)
    gc.check_against_reference(g, _ResView(res), rtol_chi2_S=2e-7)
    assert np.all(res.converged)


def test_tau_maxent_vs_hand_assembled_loop():
    """test/python/tau_maxent.py:31-135: TauMaxEnt and a MaxEntLoop assembled from its parts give the same
    result field by field, and the probabilities are the reference's literal numbers to 6 decimals."""
    g = gc.load_golden("g1_semicircular_prob.npz")
    tm = mb.TauMaxEnt(probability='normal')
    tm.set_verbosity(mb.VerbosityFlags.Quiet)
    tm.set_G_tau_data(g["tau"], g["G"])
    tm.alpha_mesh = mb.LogAlphaMesh(alpha_min=0.08, n_points=5)
    tm.omega = mb.HyperbolicOmegaMesh(omega_min=-10, omega_max=10, n_points=200)
    tm.set_error(1.e-3)
    assert np.max(np.abs(mb.TauKernel(tm.tau, tm.omega).K - tm.K.K)) < 1.e-14
    omega = mb.HyperbolicOmegaMesh(omega_min=-10, omega_max=10, n_points=200)
    K = mb.TauKernel(tm.tau, omega)
    D = mb.FlatDefaultModel(omega=omega)
    Q = mb.MaxEntCostFunction(chi2=mb.NormalChi2(K=K, G=tm.G, err=1.e-3 * np.ones(len(tm.G))),
                              S=mb.NormalEntropy(D=D), H_of_v=mb.NormalH_of_v(D=D, K=K))
    log = mb.Logtaker()
    log.verbose = mb.VerbosityFlags.Quiet
    ml = mb.MaxEntLoop(cost_function=Q, minimizer=mb.LevenbergMinimizer(), logtaker=log,
                       alpha_mesh=mb.LogAlphaMesh(alpha_min=0.08, n_points=5), probability='normal')
    for a, b in ((ml.G, tm.G), (ml.alpha_mesh, tm.alpha_mesh), (ml.data_variable, tm.tau), (ml.err, tm.err),
                 (ml.omega, tm.omega), (ml.D.D, tm.D.D), (ml.K.K, tm.K.K), (ml.K.S, tm.K.S)):
        np.testing.assert_almost_equal(np.asarray(a), np.asarray(b), 13)
    r1, r2 = ml.run(), tm.run()
    assert np.max(np.abs(r1.A_out - r2.A_out)) < 1.e-14
    for field in ('alpha', 'chi2', 'S', 'Q', 'A', 'H', 'v', 'probability', 'G', 'G_rec', 'omega'):
        np.testing.assert_almost_equal(getattr(r1, field), getattr(r2, field), 13)
    assert r1.matrix_structure is None and r2.effective_matrix_structure is None
    for key in r1.analyzer_results:
        np.testing.assert_almost_equal(r1.analyzer_results[key]['A_out'], r2.analyzer_results[key]['A_out'], 13)
    assert sorted(r1.analyzer_results) == sorted(['LineFitAnalyzer', 'Chi2CurvatureAnalyzer', 'EntropyAnalyzer',
                                                  'BryanAnalyzer', 'ClassicAnalyzer'])
    np.testing.assert_almost_equal(r2.probability, [-8476.52812836, -2343.02752796, -704.28318351,
                                                    -280.26627323, -175.30592555], 6)
    # G_rec = K_delta A reproduces the data within the error bars at small alpha
    assert np.max(np.abs(r2.G_rec[-1] - tm.G)) < 1e-2
    # analyzer extras
    lf = r2.analyzer_results['LineFitAnalyzer']
    assert len(lf['linefit_params']) == 2 and len(lf['linefit_params'][0]) == 2 and 'linefit' in lf['info']
    cv = r2.analyzer_results['Chi2CurvatureAnalyzer']['curvature']
    assert cv.shape == (5,) and np.isnan(cv[0]) and np.isnan(cv[-1]) and not np.any(np.isnan(cv[1:-1]))
    assert int(np.nanargmax(cv)) == r2.analyzer_results['Chi2CurvatureAnalyzer']['alpha_index']
    dS = r2.analyzer_results['EntropyAnalyzer']['dS_dalpha']
    ref = (r2.S[2:] - r2.S[:-2]) / (np.log(r2.alpha[2:]) - np.log(r2.alpha[:-2]))
    np.testing.assert_allclose(dS[1:-1], ref, rtol=1e-12)
    # pickling the array twin (test/python/pickle_maxent_result.py)
    again = pickle.loads(pickle.dumps(r2.data))
    np.testing.assert_array_equal(again.A_out, r2.A_out)
    np.testing.assert_array_equal(again.chi2, r2.chi2)


def test_elementwise_matches_reference_run():
    """test/python/elementwise_maxent.py:101-188 on the reference's own fixture (G_tau_noise of
    elementwise_g_tau.npz): elementwise, diagonal and hermiticity relations + the reference's numbers."""
    g = gc.load_golden("g6_elementwise_2x2.npz")

    def make(cls, **kw):
        ew = cls(use_hermiticity=True, **kw)
        ew.set_verbosity(mb.VerbosityFlags.Quiet)
        ew.set_G_tau_data(g["tau"], g["G"])
        ew.omega = mb.DataOmegaMesh(g["omega"])
        ew.alpha_mesh = mb.DataAlphaMesh(g["alpha_mesh"])
        ew.set_error(float(g["err"]))
        return ew
    ew = make(mb.ElementwiseMaxEnt)
    res = ew.run()
    assert res.matrix_structure == (2, 2) and res.chi2.shape == (2, 2, 8) and res.A.shape == (2, 2, 8, 80)
    np.testing.assert_allclose(res.alpha, g["ref_alpha"], rtol=1e-14)
    np.testing.assert_array_equal(res.A[1, 0], res.A[0, 1])           # hermiticity: exact copy
    assert np.all(np.isnan(res.chi2[1, 0]))                           # (1,0) was not computed
    # BASELINE config 2 under the tiered contract: 1e-8 where the reference reproduces itself, 10x its noise elsewhere
    gc.check_matrix_result(res, g)
    # diagonal-only run gives the same diagonal elements (bitwise: same kernel, same launch shape per element)
    dg = make(mb.DiagonalMaxEnt)
    rd = dg.run()
    for i in range(2):
        np.testing.assert_array_equal(rd.A[i, i], res.A[i, i])
    assert np.all(np.isnan(rd.A_out[0, 1]))
    # element-at-a-time equals the batched pass
    one = make(mb.ElementwiseMaxEnt)
    r1 = one.run_element((1, 1))
    np.testing.assert_array_equal(r1.A[1, 1], res.A[1, 1])
    assert np.all(np.isnan(r1.A[0, 0]))
    # norms of the diagonal spectra (test/python/elementwise_maxent.py:185-188: 2 decimals)
    om = np.asarray(ew.omega)
    for i in range(2):
        assert abs(np.trapezoid(res.A_out[i, i], om) - 1.0) < 1e-2
    assert abs(np.trapezoid(res.A_out[0, 1], om)) < 1e-2


def _matrix_front_end(cls, g, **kw):
    ew = cls(**kw)
    ew.set_verbosity(mb.VerbosityFlags.Quiet)
    ew.set_G_tau_data(g["tau"], g["G"])
    ew.omega = mb.DataOmegaMesh(g["omega"])
    ew.alpha_mesh = mb.DataAlphaMesh(g["alpha_mesh"])
    ew.set_error(float(g["err"]))
    return ew


def test_poorman_matches_reference_run():
    """PoormanMaxEnt (python/elementwise_maxent.py:562-653; test/python/elementwise_maxent.py:131-147) against the run
    of the real reference on its own 2x2 fixture (g10): every element under the tiered contract."""
    g = gc.load_golden("g10_poorman_2x2.npz")
    res = _matrix_front_end(mb.PoormanMaxEnt, g, use_hermiticity=False).run()
    gc.check_matrix_result(res, g)
    herm = _matrix_front_end(mb.PoormanMaxEnt, g, use_hermiticity=True).run()
    np.testing.assert_array_equal(herm.A_out[0, 1], herm.A_out[1, 0])            # test/python/elementwise_maxent.py:151-152
    np.testing.assert_allclose(herm.A[0, 1], res.A[0, 1], rtol=0, atol=1e-9 * np.max(np.abs(res.A[0, 1])))


def test_complex_elements_match_reference_runs():
    """use_complex=True (python/elementwise_maxent.py:244-268, 633-652; test/python/complex_elementwise_maxent.py:76-137)
    for ElementwiseMaxEnt and PoormanMaxEnt against runs of the real reference on a complex Hermitian G(tau) (g11, g12):
    real and imaginary parts of every element under the tiered contract, same NaN pattern, Hermitian A_out."""
    g = gc.load_golden("g11_complex_elementwise_2x2.npz")
    res = _matrix_front_end(mb.ElementwiseMaxEnt, g, use_hermiticity=False, use_complex=True).run()
    assert res.A.shape == (2, 2, 2, 8, 80)
    gc.check_matrix_result(res, g)
    g = gc.load_golden("g12_complex_poorman_2x2.npz")
    res = _matrix_front_end(mb.PoormanMaxEnt, g, use_hermiticity=True, use_complex=True).run()
    gc.check_matrix_result(res, g)
    A_out = res.A_out
    for iw in range(A_out.shape[-1]):                                             # complex_elementwise_maxent.py:139-143
        assert np.all(A_out[..., iw] == A_out[..., iw].conjugate().transpose())


def test_poorman_runs_and_uses_diagonal_default_model():
    """PoormanMaxEnt (python/elementwise_maxent.py:562-653): off-diagonal default model from the diagonal A_out."""
    g = gc.load_golden("g6_elementwise_2x2.npz")
    pm = mb.PoormanMaxEnt(use_hermiticity=True)
    pm.set_verbosity(mb.VerbosityFlags.Quiet)
    pm.set_G_tau_data(g["tau"], g["G"])
    pm.omega = mb.DataOmegaMesh(g["omega"])
    pm.alpha_mesh = mb.DataAlphaMesh(g["alpha_mesh"])
    pm.set_error(float(g["err"]))
    res = pm.run()
    A00 = res.analyzer_results[0][0]['LineFitAnalyzer']['A_out']
    A11 = res.analyzer_results[1][1]['LineFitAnalyzer']['A_out']
    D = pm.maxent_offdiagonal.D.D
    np.testing.assert_allclose(D, (np.sqrt(A00 * A11) + 1e-6) * pm.omega.delta, rtol=1e-13)
    np.testing.assert_allclose(res.A[0, 0], g["ref_A"][0, 0], rtol=0, atol=1e-4 * np.max(g["ref_A"][0, 0]))
    assert np.all(np.isfinite(res.A_out))
    # large alpha: the off-diagonal H = D(e^x - e^-x) stays near zero
    assert np.max(np.abs(res.A[0, 1, 0])) < np.max(np.abs(res.A[0, 1, -1])) + 1e-12
    # the off-diagonal element went through the per-spectrum-model launch; a plain TauMaxEnt with that model agrees
    tm = mb.TauMaxEnt(cost_function='plusminus')
    tm.set_verbosity(mb.VerbosityFlags.Quiet)
    tm.set_G_tau_data(g["tau"], g["G"][0, 1])
    tm.omega = mb.DataOmegaMesh(g["omega"])
    tm.alpha_mesh = mb.DataAlphaMesh(g["alpha_mesh"])
    tm.D = mb.DataDefaultModel(np.sqrt(A00 * A11) + 1e-6, tm.omega)
    tm.set_error(float(g["err"]))
    r1 = tm.run()
    k = res.analyzer_results[0][1]['LineFitAnalyzer']['alpha_index']
    assert r1.analyzer_results['LineFitAnalyzer']['alpha_index'] == k
    assert np.max(np.abs(r1.A[:k + 1] - res.A[0, 1, :k + 1])) <= 1e-7 * np.max(np.abs(r1.A[:k + 1]))


def test_covariance_vs_error_vector():
    """test/python/cov.py:53-90: a diagonal covariance matrix and the matching error vector describe the same
    problem (the rotation only permutes / reorders the data space)."""
    rng = np.random.RandomState(3)
    g = gc.load_golden("g2_synth_200x100.npz")
    err = 1e-4 * (1.0 + rng.rand(len(g["tau"])))

    def base():
        tm = mb.TauMaxEnt(reduce_singular_space=1e-11)
        tm.set_verbosity(mb.VerbosityFlags.Quiet)
        tm.set_G_tau_data(g["tau"], g["G"])
        tm.omega = mb.DataOmegaMesh(g["omega"])
        tm.alpha_mesh = mb.DataAlphaMesh(g["alpha_mesh"][:12])
        return tm
    t1 = base()
    t1.set_error(err)
    r1 = t1.run()
    t2 = base()
    t2.set_cov(np.diag(err ** 2))
    assert t2.K._T is not None and t2.K.K.shape == (len(err), len(g["omega"]))
    np.testing.assert_allclose(np.sort(t2.err), np.sort(err), rtol=1e-12)
    r2 = t2.run()
    np.testing.assert_allclose(r2.chi2, r1.chi2, rtol=1e-7)
    assert np.max(np.abs(r2.A - r1.A)) < 1e-6 * np.max(np.abs(r1.A))
    np.testing.assert_array_equal(r2.G_orig, g["G"])                   # the original data are kept
    assert np.max(np.abs(r2.G_rec[-1] - g["G"])) < 5e-3               # G_rec lives in the original basis
    # a dense covariance: whitening must reduce to the same chi2 definition r^T C^-1 r
    C = np.diag(err ** 2) + 1e-9 * np.exp(-np.abs(np.subtract.outer(g["tau"], g["tau"])))
    t3 = base()
    t3.set_cov(C)
    r3 = t3.run()
    H = r3.H[3]
    r = t3.K.K_delta @ r3.A[3] - g["G"]
    np.testing.assert_allclose(r3.chi2[3], r @ np.linalg.solve(C, r), rtol=1e-6)
    assert np.all(np.isfinite(H))
    # switching back to a plain error undoes the rotation
    t3.set_error(1e-4)
    assert t3.K._T is None and t3.K.K.shape == (len(g["tau"]), len(g["omega"]))


def test_rank_deficient_covariance():
    """A covariance matrix estimated from fewer samples than singular values: set_cov drops its null space
    (python/tau_maxent.py:253-288), the rotated kernel keeps k < n_sv rows, v stays n_sv-dimensional
    (python/kernels.py:160-180).  Oracle: the reference algorithm on the rotated problem with the rotated U."""
    pr = mo.synthetic_problem(120, 60, beta=20.0, mu=0.5, sigma=2e-3, seed=11)
    rng = np.random.RandomState(5)
    n_tau, k = 120, 30
    B = rng.randn(n_tau, k)
    cov = 4e-6 * (B @ B.T) / k                                 # rank 30 < n_sv
    e, v = np.linalg.eigh(cov)
    keep = e >= 1e-14
    assert keep.sum() == k
    T = v[:, keep].T
    G = pr["G"][0] if pr["G"].ndim == 2 else pr["G"]
    tm = mb.TauMaxEnt(alpha_mesh=mb.LogAlphaMesh(0.5, 500, 6), reduce_singular_space=1e-9)
    tm.set_verbosity(mb.VerbosityFlags.Quiet)
    tm.omega = mb.DataOmegaMesh(pr["omega"])
    tm.set_G_tau_data(pr["tau"], G)
    tm.set_cov(cov)
    res = tm.run()
    n_sv = len(tm.K.S)
    assert n_sv > k
    U0, S0, V0 = mo.kernel_svd(pr["K"], 1e-9)
    o = mo.maxent_loop(T @ pr["K"], T @ G, np.sqrt(e[keep]), pr["omega"], np.asarray(tm.alpha_mesh),
                       svd=(T @ U0, S0, V0))
    assert o["n_sv"] == n_sv
    assert np.all(gc.rel_A(res.A, o["A"]) <= 1e-7), gc.rel_A(res.A, o["A"])
    np.testing.assert_allclose(res.chi2, o["chi2"], rtol=1e-7)
    for name in ('LineFitAnalyzer', 'Chi2CurvatureAnalyzer'):
        assert res.analyzer_results[name]['alpha_index'] == o["analyzers"][name]["alpha_index"]


def test_rank_above_the_fused_limit_raises():
    """A kernel whose numerical rank exceeds MX_MAX_NSV (256) is an error (the reference keeps every S >= threshold,
    python/kernels.py:101-122; a silent truncation would change A), unless the caller asks for the truncation."""
    from maxent_b200 import engine, _lib
    rng = np.random.RandomState(2)
    K = rng.randn(400, 300)
    om = np.linspace(-5, 5, 300)
    with pytest.raises(_lib.MaxEntLibraryError):
        engine.SharedProblem(K, 1e-2, mo.flat_default_model(om), mo.omega_delta(om), reduce_singular_space=1e-14)
    prob = engine.SharedProblem(K, 1e-2, mo.flat_default_model(om), mo.omega_delta(om), reduce_singular_space=1e-14,
                                max_nsv=64)
    assert prob.n_sv == 64 and prob.n_sv_uncapped == 300


def test_per_spectrum_error_models_in_one_launch():
    """BatchedTauMaxEnt with an error model PER SPECTRUM (TauMaxEnt.set_error / set_cov are per data set in the
    reference, python/tau_maxent.py:227-288; test/python/cov.py:53-90): one scalar sigma per spectrum, one error vector
    per spectrum (three distinct vectors -> three whitening groups in ONE launch of the sweep), one covariance matrix
    per spectrum.  Every spectrum against the oracle run with its own error model, tiered contract."""
    from maxent_b200 import batched
    pr = mo.synthetic_problem(120, 60, beta=20.0, mu=[0.5, -0.5, 1.0, 0.0, 0.8, -0.2], sigma=1e-3, seed=3)
    n_tau, B = 120, 6
    mesh = mo.log_alpha_mesh(0.5, 500.0, 8)
    rng = np.random.RandomState(8)

    def job():
        j = batched.BatchedTauMaxEnt(reduce_singular_space=1e-10)
        j.set_kernel_tau(pr["tau"], batched.DataOmegaMesh(pr["omega"]), beta=20.0)
        j.alpha_mesh = mb.DataAlphaMesh(mesh)
        return j

    def check(out, b, o, what):
        tol = np.maximum(1e-8, 10 * o["noise_A"])
        dA = gc.rel_A(out.A(b), o["A"])
        assert np.all(dA <= tol), (what, b, dA / tol)
        np.testing.assert_allclose(out.chi2[b], o["chi2"], rtol=1e-7, err_msg="%s %d" % (what, b))
        assert int(out.alpha_index[b, 0]) == o["analyzers"]["LineFitAnalyzer"]["alpha_index"], (what, b)
        assert int(out.alpha_index[b, 1]) == o["analyzers"]["Chi2CurvatureAnalyzer"]["alpha_index"], (what, b)

    def oracle(K, G, err, svd=None):
        o = mo.maxent_loop(K, G, err, pr["omega"], mesh, reduce_singular_space=1e-10, svd=svd)
        o["noise_A"], _ = gc.oracle_floor(o, lambda f: mo.maxent_loop(K, G * f, err, pr["omega"], mesh, reduce_singular_space=1e-10,
                                                                     svd=svd, analyzers=False))
        return o

    # (a) one scalar error bar per spectrum
    sig = np.array([1e-3, 2e-3, 5e-4, 1e-3, 3e-3, 5e-4])
    j = job()
    j.set_error(sig)
    out = j.run(pr["G"])
    for b in range(B):
        check(out, b, oracle(pr["K"], pr["G"][b], sig[b]), "sigma")
    # (b) one error vector per spectrum, three distinct ones
    vecs = 1e-3 * (1.0 + 0.5 * rng.rand(3, n_tau))
    which = np.array([0, 1, 2, 1, 0, 2])
    j = job()
    j.set_error(vecs[which])
    out = j.run(pr["G"])
    mode, index, specs = j._error_plan(B)
    assert mode == "groups" and len(specs) == 3
    for b in range(B):
        check(out, b, oracle(pr["K"], pr["G"][b], vecs[which[b]]), "vector")
    # (c) one covariance matrix per spectrum (two distinct ones, the second rank deficient -> fewer rotated rows)
    i = np.arange(n_tau)
    C0 = 1e-6 * (np.eye(n_tau) + 0.4 * np.exp(-np.abs(i[:, None] - i[None, :]) / 2.5))
    Bm = rng.randn(n_tau, 90)
    C1 = 4e-6 * (Bm @ Bm.T) / 90
    covs = np.stack([C0, C1, C0, C1, C1, C0])
    j = job()
    j.set_cov(covs)
    out = j.run(pr["G"])
    U0, S0, V0 = mo.kernel_svd(pr["K"], 1e-10)
    for b in range(B):
        e, v = np.linalg.eigh(covs[b])
        keep = e >= 1e-14
        T = v[:, keep].T
        check(out, b, oracle(T @ pr["K"], T @ pr["G"][b], np.sqrt(e[keep]), svd=(T @ U0, S0, V0)), "cov")


def test_config4_matches_reference_run():
    """BASELINE config 4 (n_tau = 10000, n_omega = 2000, 100 alphas, probability, cut 1e-11) against the run of the REAL
    reference stored in tests/golden/g13 (218 s on 8 cores + its reproducibility run): chi2, S, Q, probability for all 100
    alphas and A at the analyzer picks and every tenth alpha under the tiered contract, identical picks."""
    g = gc.load_golden("g13_config4_10000x2000.npz")
    pr = mo.synthetic_problem(10000, 2000, mu=1.0, seed=1234)
    G = pr["G"][0]
    assert abs(float(np.sum(G)) - float(g["G_sha_check"])) < 1e-9
    tm = mb.TauMaxEnt(probability='normal', reduce_singular_space=1e-11)
    tm.set_verbosity(mb.VerbosityFlags.Quiet)
    tm.set_G_tau_data(pr["tau"], G)
    tm.omega = mb.HyperbolicOmegaMesh(-10, 10, 2000)
    tm.alpha_mesh = mb.LogAlphaMesh(0.01, 2000, 100)
    tm.set_error(1e-4)
    res = tm.run()
    assert len(tm.K.S) == int(g["ref_n_sv"]) == 54
    tolA, tolc, tolS = gc.tolerances(g, "A"), gc.tolerances(g, "chi2"), gc.tolerances(g, "S")
    rows = g["A_rows"]
    dA = gc.rel_A(res.A[rows], g["ref_A"])
    assert np.all(dA <= tolA[rows]), (dA / tolA[rows])
    dc = np.abs(res.chi2 / g["ref_chi2"] - 1)
    assert np.all(dc <= np.maximum(tolc, 2e-7)), (dc / tolc).max()     # device TauKernel: see gc.check_against_reference
    dQ = np.abs(res.Q / g["ref_Q"] - 1)
    assert np.all(dQ <= np.maximum(tolc, tolS)), dQ.max()
    p, pref = res.probability, g["ref_probability"]
    assert np.all(np.abs(p - pref) <= np.maximum(gc.PROB_RTOL, gc.NOISE_FACTOR * gc.running_max(g["noise_chi2"])) * np.abs(pref))
    for name in ('LineFitAnalyzer', 'Chi2CurvatureAnalyzer', 'EntropyAnalyzer', 'ClassicAnalyzer'):
        assert res.analyzer_results[name]['alpha_index'] == int(g["ref_idx_" + name]), name
        k = int(g["ref_idx_" + name])
        ref = g["ref_Aout_" + name]
        assert np.max(np.abs(res.analyzer_results[name]['A_out'] - ref)) <= tolA[k] * np.max(np.abs(ref)), name


def test_batched_threshold_skip_on_the_device():
    """max|G| < G_threshold (python/maxent_loop.py:174-179) in a batch: such spectra are left out of the launch, come back
    with MX_STATUS_SKIPPED, NaN chi2, alpha_index -1 and A_out = 0 (device tensors as well as host arrays), and do not
    change the other spectra by a bit; replacing a public attribute rebuilds the cached problem."""
    import torch
    from maxent_b200 import batched, _lib
    pr = mo.synthetic_problem(120, 60, beta=20.0, mu=[0.5, -0.5, 1.0, 0.2], sigma=1e-3, seed=3)
    j = batched.BatchedTauMaxEnt(reduce_singular_space=1e-10)
    j.set_kernel_tau(pr["tau"], batched.DataOmegaMesh(pr["omega"]), beta=20.0)
    j.set_alpha_mesh_log(0.5, 500.0, 6)
    j.set_error(np.array([1e-3, 2e-3, 1e-3, 5e-4]))                # per-spectrum rows must follow the kept spectra
    G = pr["G"].copy()
    G[1] = 0.0
    G[3] *= 1e-12
    out = j.run(G)
    assert out.zero_elements == [1, 3]
    assert np.all(np.isnan(out.chi2[[1, 3]])) and np.all(out.alpha_index[[1, 3]] == -1) and np.all(out.A_out[[1, 3]] == 0.0)
    assert bool((out.device.status[[1, 3]] == _lib.STATUS_SKIPPED).all()) and bool((out.device.A_out[[1, 3]] == 0).all())
    assert not out.converged[1].any() and out.converged[0].all()
    j2 = batched.BatchedTauMaxEnt(reduce_singular_space=1e-10)
    j2.set_kernel_tau(pr["tau"], batched.DataOmegaMesh(pr["omega"]), beta=20.0)
    j2.set_alpha_mesh_log(0.5, 500.0, 6)
    j2.set_error(np.array([1e-3, 1e-3]))
    ref = j2.run(G[[0, 2]])
    np.testing.assert_array_equal(out.chi2[[0, 2]], ref.chi2)
    np.testing.assert_array_equal(out.A_out[[0, 2]], ref.A_out)
    p1 = j2.prepare()
    j2.reduce_singular_space = 1e-8
    assert j2.prepare() is not p1 and j2.prepare().n_sv < p1.n_sv


def test_iomega_kernel_matches_reference_run():
    """G(i omega_n) -> real A(omega): IOmegaKernel + ComplexChi2 (python/kernels.py:283-346, python/functions.py:380-437)
    through MaxEntLoop with complex data, against the run of the real reference on the stacked [Re; Im] system (g14):
    200 Matsubara frequencies, 100 omega points, 20 alphas, probability.  Tiered contract, identical picks."""
    g = gc.load_golden("g14_iomega_200x100.npz")
    om = mb.DataOmegaMesh(g["omega"])
    K = mb.IOmegaKernel(g["iomega"], om, beta=float(g["beta"]))
    np.testing.assert_allclose(K.K[0], g["ref_K_row0"], rtol=1e-15)
    np.testing.assert_allclose(K.K_delta[0], g["ref_K_delta_row0"], rtol=1e-15)
    D = mb.FlatDefaultModel(omega=om)
    G = g["G"]
    assert np.iscomplexobj(G)
    Q = mb.MaxEntCostFunction(chi2=mb.ComplexChi2(K=K, G=G, err=float(g["err"]) * np.ones(len(G))), S=mb.NormalEntropy(D=D),
                              H_of_v=mb.NormalH_of_v(D=D, K=K))
    ml = mb.MaxEntLoop(cost_function=Q, alpha_mesh=mb.DataAlphaMesh(g["alpha_mesh"]), reduce_singular_space=1e-11,
                       scale_alpha=float(g["scale_alpha"]), probability='normal')
    ml.set_verbosity(mb.VerbosityFlags.Quiet)
    res = ml.run()
    assert len(K.S) == int(g["ref_n_sv"]) and K.U.shape[0] == 2 * len(G) and not np.iscomplexobj(K.V)
    gc.check_against_reference(g, _ResView(res))
    # the reconstructed data are complex again: G_rec = K_delta A (python/maxent_result.py:905-908)
    assert np.iscomplexobj(res.G_rec) and res.G_rec.shape == (20, len(G))
    k = res.analyzer_results['LineFitAnalyzer']['alpha_index']
    assert np.max(np.abs(res.G_rec[k] - G)) < 6 * float(g["err"])
    # scale_alpha = 'Ndata' counts the complex data points (len(G), python/maxent_loop.py:216-220)
    ml2 = mb.MaxEntLoop(cost_function=Q, alpha_mesh=mb.DataAlphaMesh(g["alpha_mesh"]), reduce_singular_space=1e-11)
    ml2.set_verbosity(mb.VerbosityFlags.Quiet)
    np.testing.assert_allclose(ml2.run().alpha, res.alpha, rtol=1e-15)


def test_preblur_b_scan_in_one_launch():
    """The b scan of the preblur workflow (doc/guide/preblur_example.py:46-75) as one batch: every b is a whitening group
    of ONE launch of the sweep (its own kernel SVD, its own n_sv, V', xi).  Each b must reproduce the run of that b alone
    (same bits: padding adds exact zeros), b = 0.3 the run of the real reference (fixture g7)."""
    g = gc.load_golden("g7_preblur_200x100.npz")
    tm = _tau_maxent_from_fixture(g)
    K_tau = tm.K
    bs = [0.1, 0.2, 0.3, 0.4]
    scan = mb.preblur_scan(tm, bs)
    assert sorted(scan) == bs and tm.K is K_tau
    n_svs = []
    for b in bs:
        t1 = _tau_maxent_from_fixture(g)
        K1 = t1.K
        t1.A_of_H = mb.PreblurA_of_H(b=b, omega=t1.omega)
        t1.K = mb.PreblurKernel(K=K1, b=b)
        r1 = t1.run()
        n_svs.append(len(t1.K.S))
        np.testing.assert_array_equal(scan[b].chi2, r1.chi2)
        np.testing.assert_array_equal(scan[b].A, r1.A)
        assert scan[b].analyzer_results['LineFitAnalyzer']['alpha_index'] == r1.analyzer_results['LineFitAnalyzer']['alpha_index']
    assert len(set(n_svs)) > 1, n_svs                      # the groups really had different singular-space dimensions
    gc.check_against_reference(g, _ResView(scan[0.3]), rtol_chi2_S=2e-7)


def test_threshold_skip_and_unsupported_combinations():
    g = gc.load_golden("g2_synth_200x100.npz")
    tm = _tau_maxent_from_fixture(g)
    tm.set_G_tau_data(g["tau"], 1e-12 * g["G"])
    tm.set_error(1e-4)
    assert tm.run() is None and 'G below threshold' in tm.logtaker.get_error_messages()[0]
    res = mb.MaxEntResult(matrix_structure=(1, 1))
    assert tm.run(result=res, matrix_element=(0, 0)) is None and res.zero_elements == [(0, 0)]
    tm2 = _tau_maxent_from_fixture(g)
    tm2.cost_function.d_dv = True
    with pytest.raises(NotImplementedError):
        tm2.run()
    tm3 = _tau_maxent_from_fixture(g)
    tm3.scale_alpha = 'nonsense'
    with pytest.raises(Exception):
        tm3.run()
    tm4 = _tau_maxent_from_fixture(g)
    tm4.scale_alpha = 1.0
    tm4.alpha_mesh = mb.DataAlphaMesh(g["ref_alpha"])                # already-scaled alphas with scale 1
    r4 = tm4.run()
    np.testing.assert_allclose(r4.chi2, g["ref_chi2"], rtol=1e-7)


def test_alpha_loop_log_lines(capsys):
    """The per-alpha report of python/maxent_loop.py:248-255, printed from the device counters."""
    g = gc.load_golden("g2_synth_200x100.npz")
    tm = _tau_maxent_from_fixture(g)
    tm.set_verbosity(mb.VerbosityFlags.AlphaLoop | mb.VerbosityFlags.Timing)
    tm.minimizer.maxiter = 5
    res = tm.run()
    out = capsys.readouterr().out.splitlines()
    lines = [l for l in out if l.startswith("alpha[")]
    assert len(lines) == len(res.alpha)
    width = int(np.ceil(np.log10(len(res.alpha))))
    assert lines[0].startswith("alpha[%*d] = %16.8e, chi2 = %16.8e, n_iter=" % (width, 0, res.alpha[0], res.chi2[0]))
    assert any(l.endswith("!") for l in lines) and not np.all(res.converged)
    assert any("did not converge" in l for l in out) and any(l.startswith("MaxEnt loop finished in") for l in out)
    assert tm.minimizer.n_iter_last == int(res.n_iter[-1]) <= 5


def test_config4_large_kernel_full_probability_analysis():
    """BASELINE config 4: n_tau = 10000, n_omega = 2000, 100 alphas, probability + Bryan/Classic analysis, through
    TauMaxEnt.  The reference's own run of this case (205 s + 298 s setup on 8 cores, BASELINE.md / SURVEY.md 6)
    gave n_sv = 54 and the analyzer picks LineFit 38, Chi2Curvature 48, Entropy 92, Classic 99."""
    from oracle import maxent_oracle as mo
    pr = mo.synthetic_problem(10000, 2000, mu=1.0, seed=1234)
    tm = mb.TauMaxEnt(probability='normal', reduce_singular_space=1e-11)
    tm.set_verbosity(mb.VerbosityFlags.Quiet)
    tm.set_G_tau_data(pr["tau"], pr["G"][0])
    tm.omega = mb.HyperbolicOmegaMesh(-10, 10, 2000)
    tm.alpha_mesh = mb.LogAlphaMesh(0.01, 2000, 100)
    tm.set_error(1e-4)
    res = tm.run()
    assert len(tm.K.S) == 54
    picks = {k: res.analyzer_results[k]['alpha_index'] for k in
             ('LineFitAnalyzer', 'Chi2CurvatureAnalyzer', 'EntropyAnalyzer', 'ClassicAnalyzer')}
    assert picks == {'LineFitAnalyzer': 38, 'Chi2CurvatureAnalyzer': 48, 'EntropyAnalyzer': 92, 'ClassicAnalyzer': 99}
    assert np.all(res.converged) and np.all(np.isfinite(res.probability))
    assert np.all(np.diff(res.probability) > 0)                     # p still rising at the smallest alpha
    # chi2 and S reported by the device equal the values recomputed from A with the full kernel matrix
    H = res.H[[0, 38, 48, 99]]
    r = (H @ pr["K"].T - pr["G"][0][None, :]) / 1e-4
    np.testing.assert_allclose((r * r).sum(1), res.chi2[[0, 38, 48, 99]], rtol=1e-8)
    D = tm.D.D
    np.testing.assert_allclose((H - D - H * np.log(H / D)).sum(1), res.S[[0, 38, 48, 99]], rtol=1e-9)
    # Bryan's average is a convex combination of the A_alpha
    br = res.analyzer_results['BryanAnalyzer']['A_out']
    assert np.all(br >= res.A.min(0) - 1e-12) and np.all(br <= res.A.max(0) + 1e-12)
    assert abs(np.trapezoid(res.A_out, np.asarray(tm.omega)) - 1.0) < 1e-2


def test_complex_matrix_elements():
    """use_complex=True (python/elementwise_maxent.py:244-268): real and imaginary parts of the off-diagonal
    elements are continued separately (plus-minus entropy), the diagonal is real, A_out is Hermitian."""
    g = gc.load_golden("g6_elementwise_2x2.npz")
    tau, Gr = g["tau"], g["G"]
    Gr = 0.5 * (Gr + np.transpose(Gr, (1, 0, 2)))               # the fixture's noise is independent per element
    # a Hermitian G(tau): rotate the real symmetric matrix with a complex unitary
    th = 0.3
    U = np.array([[np.cos(th), 1j * np.sin(th)], [1j * np.sin(th), np.cos(th)]])
    Gc = np.einsum('ab,bct,dc->adt', U, Gr.astype(complex), U.conj())
    assert np.max(np.abs(Gc - np.conj(np.transpose(Gc, (1, 0, 2))))) < 1e-12
    ew = mb.ElementwiseMaxEnt(use_hermiticity=True, use_complex=True)
    ew.set_verbosity(mb.VerbosityFlags.Quiet)
    ew.set_G_tau_data(tau, Gc)
    ew.omega = mb.DataOmegaMesh(g["omega"])
    ew.alpha_mesh = mb.DataAlphaMesh(g["alpha_mesh"])
    ew.set_error(float(g["err"]))
    res = ew.run()
    assert res.effective_matrix_structure == (2, 2, 2) and res.A.shape == (2, 2, 2, 8, 80)
    assert (0, 0, 1) in res.zero_elements and (1, 1, 1) in res.zero_elements
    A_out = res.A_out
    assert A_out.dtype == complex and A_out.shape == (2, 2, 80)
    np.testing.assert_array_equal(A_out[1, 0], np.conj(A_out[0, 1]))
    assert np.max(np.abs(A_out[0, 0].imag)) == 0 and np.all(np.isfinite(A_out.real))
    # trace is invariant under the rotation: compare with the real run's diagonal sum at the LineFit alphas
    ref = mb.ElementwiseMaxEnt(use_hermiticity=True)
    ref.set_verbosity(mb.VerbosityFlags.Quiet)
    ref.set_G_tau_data(tau, Gr)
    ref.omega = mb.DataOmegaMesh(g["omega"])
    ref.alpha_mesh = mb.DataAlphaMesh(g["alpha_mesh"])
    ref.set_error(float(g["err"]))
    r0 = ref.run()
    om = np.asarray(ew.omega)
    t_c = np.trapezoid((A_out[0, 0] + A_out[1, 1]).real, om)
    t_r = np.trapezoid(r0.A_out[0, 0] + r0.A_out[1, 1], om)
    assert abs(t_c - t_r) < 2e-2 and abs(t_c - 2.0) < 5e-2


def test_maxent_loop_with_data_kernel():
    """A user-supplied kernel matrix (DataKernel, python/kernels.py:183-207) through a hand-assembled MaxEntLoop,
    against the oracle; also Bryan's cost function on the same problem (test/python/bryan_cost_function.py)."""
    from oracle import maxent_oracle as mo
    rng = np.random.RandomState(21)
    n_tau, n_om = 90, 64
    om = mb.LinearOmegaMesh(-3, 3, n_om)
    U, _ = np.linalg.qr(rng.randn(n_tau, n_om))
    V, _ = np.linalg.qr(rng.randn(n_om, n_om))
    S = np.concatenate([np.logspace(0, -5, 40), 1e-15 * np.ones(n_om - 40)])
    Kmat = (U * S) @ V.T
    A_true = np.exp(-(np.asarray(om) - 0.3) ** 2)
    G = (Kmat * om.delta[None, :]) @ A_true + 1e-3 * rng.randn(n_tau)
    mesh = mb.LogAlphaMesh(0.5, 200, 6)
    for kind, variant in (('normal', 'normal'), ('bryan', 'bryan')):
        K = mb.DataKernel(np.arange(n_tau, dtype=float), om, Kmat)
        D = mb.FlatDefaultModel(om)
        ml = mb.MaxEntLoop(cost_function=kind, alpha_mesh=mesh, reduce_singular_space=1e-9)
        ml.set_verbosity(mb.VerbosityFlags.Quiet)
        ml.K, ml.D, ml.G, ml.err = K, D, G, 1e-3 * np.ones(n_tau)
        res = ml.run()
        assert len(K.S) == 40 and np.all(res.converged)
        o = mo.maxent_loop(Kmat, G, 1e-3, np.asarray(om), np.asarray(mesh), variant=variant, reduce_singular_space=1e-9)
        noise, _ = gc.oracle_floor(o, lambda f: mo.maxent_loop(Kmat, G * f, 1e-3, np.asarray(om), np.asarray(mesh), variant=variant,
                                                              reduce_singular_space=1e-9, analyzers=False))
        tol = np.maximum(1e-8, 10 * noise)
        assert np.all(gc.rel_A(res.A, o["A"]) <= tol), (kind, gc.rel_A(res.A, o["A"]) / tol)
        np.testing.assert_allclose(res.chi2, o["chi2"], rtol=1e-7)
        for name in ('LineFitAnalyzer', 'Chi2CurvatureAnalyzer', 'EntropyAnalyzer'):
            assert res.analyzer_results[name]['alpha_index'] == o["analyzers"][name]["alpha_index"]
        np.testing.assert_allclose(res.G_rec[-1], (Kmat * om.delta[None, :]) @ res.A[-1], rtol=1e-12)


def test_preblur_matches_reference_run():
    """Preblur formalism (PreblurKernel + PreblurA_of_H, doc/guide/preblur_example.py:46-52):
    the same user script as oracle/make_golden.py ran on the reference (fixture g7)."""
    g = gc.load_golden("g7_preblur_200x100.npz")
    b = float(g["preblur_b"])
    tm = _tau_maxent_from_fixture(g)
    K_tau = tm.K
    tm.A_of_H = mb.PreblurA_of_H(b=b, omega=tm.omega)
    tm.K = mb.PreblurKernel(K=K_tau, b=b)
    res = tm.run()
    assert len(tm.K.S) == int(g["ref_n_sv"])
    np.testing.assert_allclose(tm.K.K, K_tau.K @ (tm.omega.delta[:, None] * mb.get_preblur(tm.omega, b)), rtol=1e-12, atol=1e-300)
    gc.check_against_reference(g, _ResView(res), rtol_chi2_S=2e-7)
    # the hidden image itself is only determined up to the near-null space of the blur: compare it where the
    # problem is well determined (large alpha), A = B H everywhere (done above)
    dH = gc.rel_A(res.H, g["ref_H"])
    assert np.all(dH[:9] <= 1e-8), dH
    np.testing.assert_allclose(res.A, res.H @ mb.get_preblur(tm.omega, b).T, rtol=1e-12, atol=1e-300)    # A = B H
    np.testing.assert_allclose(res.G_rec[-1], g["ref_G_rec_last"], rtol=0, atol=1e-8)
    # switching the blur off again gives the plain result
    tm.A_of_H = mb.IdentityA_of_H(tm.omega)
    tm.K = K_tau
    r0 = tm.run()
    np.testing.assert_allclose(r0.H, r0.A * tm.omega.delta[None, :], rtol=1e-15)
    assert not np.allclose(r0.A[5], res.A[5], rtol=1e-3)
