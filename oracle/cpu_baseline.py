#!/usr/bin/env python
"""Host-CPU baseline of the alpha-sweep hot path.  TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Times ``oracle/maxent_oracle.py`` (the numpy port that is pinned bit-identically to TRIQS/maxent's
``TauMaxEnt.run``; the Python reference itself cannot travel to the GPU box) on the host cores:
one spectrum per process, one BLAS thread per process -- the best-throughput configuration found
for the reference.  Called by ``bench.py`` (``cpu_baseline`` leg and
``--impl reference``) in a subprocess so that no CUDA context is forked.

    python -m oracle.cpu_baseline --n-tau 2000 --n-omega 1000 --n-alpha 60 --spectra 8 --procs 8

prints one JSON line: {"spectra": n, "wall_s": t, "spectra_per_s": n/t, "cores": procs, ...}.
"""
import argparse
import json
import os
import sys
import time

for _v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = os.environ.get("MAXENT_CPU_BLAS_THREADS", "1")

import numpy as np  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import maxent_oracle as mo  # noqa: E402

_P = {}
PERTURBATIONS = (1, 2, 3, 4, 5)      # see _perturbed


def kresolved_inputs(n_tau, n_omega, kpoints, rows, seed=3):
    """The k-resolved batch (SURVEY.md 8(d) C3 recipe = BASELINE config 3): spectrum k of `kpoints` is a Gaussian at
    mu_k = 2 cos(2 pi k / kpoints) with noise row k of default_rng(seed).standard_normal((kpoints, n_tau));
    `rows` selects the k values."""
    rows = np.asarray(rows, dtype=int)
    noise = np.random.default_rng(seed).standard_normal((kpoints, n_tau))[rows]
    return mo.synthetic_problem(n_tau, n_omega, mu=2.0 * np.cos(2.0 * np.pi * rows / kpoints), noise=noise)


def bench_inputs(n_tau, n_omega, n_spectra, seed=5, first=0):
    """The benchmark's synthetic batch (SURVEY.md 8(d) C5 recipe): rows of
    default_rng(seed).standard_normal((B, n_tau)) as noise on a Gaussian A(omega), mu = 1."""
    if _P.get("krows") is not None:
        return kresolved_inputs(n_tau, n_omega, _P["kpoints"], _P["krows"])
    rng = np.random.default_rng(seed)
    noise = rng.standard_normal((first + n_spectra, n_tau))[first:]
    return mo.synthetic_problem(n_tau, n_omega, mu=np.ones(n_spectra), noise=noise)


def _init(n_tau, n_omega, n_alpha, n_spectra, thr, seed):
    _P["pr"] = bench_inputs(n_tau, n_omega, n_spectra, seed, first=_P.get("first", 0))
    _P["mesh"] = mo.log_alpha_mesh(0.01, 2000, n_alpha)
    _P["thr"] = thr


def _perturbed(task):
    """One rounding-level perturbation of the oracle run of spectrum b (SURVEY.md 8(c) tier T4: the oracle's own
    reproducibility): v = 1, 2: G * (1 +- 1e-15); v = 3: the other LAPACK SVD driver (gesvd instead of numpy's gesdd --
    the singular vectors next to the cut are only determined to ~eps * S[0] / S[k]); v = 4: every entry of
    the kernel matrix moved by +-1e-15 relative (the rounding of the kernel itself); v = 5: G * (1 + 2e-15)."""
    b, v = task
    pr = _P["pr"]
    G = pr["G"][b]
    K = pr["K"]
    svd = None
    if v == 1:
        G = G * (1.0 + 1.e-15)
    elif v == 2:
        G = G * (1.0 - 1.e-15)
    elif v == 4:
        K = K * (1.0 + 1.e-15 * np.random.default_rng(1000 + b).choice([-1.0, 1.0], size=K.shape))
    elif v == 5:
        G = G * (1.0 + 2.e-15)
    else:
        import scipy.linalg
        U, S, Vh = scipy.linalg.svd(pr["K"], full_matrices=False, lapack_driver="gesvd")
        keep = S >= _P["thr"]
        svd = (U[:, keep], S[keep], Vh.T[:, keep])
    o2 = mo.maxent_loop(K, G, pr["err"], pr["omega"], _P["mesh"], reduce_singular_space=_P["thr"], fast_d2=True,
                        analyzers=False, svd=svd)
    return dict(b=b, v=v, A=o2["A"], chi2=o2["chi2"])


def _one(b):
    pr = _P["pr"]
    t0 = time.perf_counter()
    # fast_d2: the K^T W K setup product through BLAS instead of the reference's einsum
    # (python/functions.py:372-377) -- identical numbers to rounding, and it only favours the CPU side.
    out = mo.maxent_loop(pr["K"], pr["G"][b], pr["err"], pr["omega"], _P["mesh"],
                         reduce_singular_space=_P["thr"], fast_d2=True)
    an = out["analyzers"]
    if _P.get("keep"):
        _P["keep_rows"] = dict(A=out["A"], chi2=out["chi2"], S=out["S"], Q=out["Q"], n_iter=out["n_iter"])
    return dict(b=b, arrays=_P.pop("keep_rows", None), wall=time.perf_counter() - t0, n_iter=int(out["n_iter"].sum()), n_qeval=int(out["n_qeval"]),
                n_solve=int(out["n_solve"]), n_sv=int(out["n_sv"]),
                linefit=int(an["LineFitAnalyzer"]["alpha_index"]), chi2curv=int(an["Chi2CurvatureAnalyzer"]["alpha_index"]))


def run(n_tau, n_omega, n_alpha, spectra, procs, thr=1e-11, seed=5, dump=None, kpoints=0, krows=None, first=0, noise=False):
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    _P["keep"] = dump is not None            # inherited by the forked workers
    _P["first"], _P["noise"] = first, bool(noise)
    _P["kpoints"], _P["krows"] = kpoints, (None if not kpoints else list(krows))
    if kpoints:
        spectra = len(krows)
    if procs <= 1:
        _init(n_tau, n_omega, n_alpha, spectra, thr, seed)
        rows = [_one(b) for b in range(spectra)]
    else:
        with ctx.Pool(procs, initializer=_init, initargs=(n_tau, n_omega, n_alpha, spectra, thr, seed)) as pool:
            pert = pool.map_async(_perturbed, [(b, v) for v in PERTURBATIONS for b in range(spectra)], chunksize=1) if noise else None
            rows = pool.map(_one, range(spectra), chunksize=1)
            pert = pert.get() if pert is not None else []
        for r in pert:                       # noise floor = largest movement over the perturbations
            a = rows[r["b"]]["arrays"]
            nA = np.max(np.abs(r["A"] - a["A"]), axis=1) / np.max(np.abs(a["A"]), axis=1)
            nc = np.abs(r["chi2"] / a["chi2"] - 1.0)
            a["noise_A"] = np.maximum(a.get("noise_A", 0.0), nA)
            a["noise_chi2"] = np.maximum(a.get("noise_chi2", 0.0), nc)
    wall = time.perf_counter() - t0
    if dump is not None:                     # the oracle's own outputs, for the full-size parity test
        pr = bench_inputs(n_tau, n_omega, spectra, seed, first=first)
        extra = {}
        if noise:
            extra = dict(noise_A=np.stack([r["arrays"]["noise_A"] for r in rows]),
                         noise_chi2=np.stack([r["arrays"]["noise_chi2"] for r in rows]))
        np.savez(dump, G=pr["G"], **extra, A=np.stack([r["arrays"]["A"] for r in rows]), chi2=np.stack([r["arrays"]["chi2"] for r in rows]),
                 S=np.stack([r["arrays"]["S"] for r in rows]), Q=np.stack([r["arrays"]["Q"] for r in rows]),
                 n_iter=np.stack([r["arrays"]["n_iter"] for r in rows]), linefit=np.array([r["linefit"] for r in rows]),
                 chi2curv=np.array([r["chi2curv"] for r in rows]))
    return dict(spectra=spectra, wall_s=wall, spectra_per_s=spectra / wall, cores=procs,
                per_spectrum_s=[round(r["wall"], 3) for r in rows], n_iter=[r["n_iter"] for r in rows],
                n_qeval=[r["n_qeval"] for r in rows], n_solve=[r["n_solve"] for r in rows], n_sv=rows[0]["n_sv"],
                linefit=[r["linefit"] for r in rows], chi2curv=[r["chi2curv"] for r in rows],
                numpy=np.__version__, blas_threads=int(os.environ["OPENBLAS_NUM_THREADS"]),
                reduce_singular_space=thr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-tau", type=int, default=2000)
    ap.add_argument("--n-omega", type=int, default=1000)
    ap.add_argument("--n-alpha", type=int, default=60)
    ap.add_argument("--spectra", type=int, default=0)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--thr", type=float, default=1e-11)
    ap.add_argument("--dump", default=None, help="write the oracle's A/chi2/S/Q/picks of the sample to this .npz")
    ap.add_argument("--kpoints", type=int, default=0, help="k-resolved recipe (C3): size of the k mesh")
    ap.add_argument("--krows", default="", help="comma-separated k values to run (with --kpoints)")
    ap.add_argument("--first", type=int, default=0, help="first row of the benchmark batch to run (rows of another rank's shard)")
    ap.add_argument("--noise-floor", action="store_true",
                    help="with --dump and --procs > 1: also run the rounding-level perturbations of every spectrum")
    a = ap.parse_args()
    procs = a.procs or (os.cpu_count() or 1)
    spectra = a.spectra or procs
    krows = [int(x) for x in a.krows.split(",") if x]
    print(json.dumps(run(a.n_tau, a.n_omega, a.n_alpha, spectra, procs, a.thr, dump=a.dump, kpoints=a.kpoints,
                         krows=krows, first=a.first, noise=a.noise_floor)))


if __name__ == "__main__":
    main()
