"""A/B timing of sweep-kernel builds on the benchmark shape (GPU box).

    python tools/ab_bench.py [--spectra N] [--phases] libA.so libB.so ...

Every library (built with `python -m maxent_b200.build --variant <tag> -D...`) runs in its own process on the same
synthetic batch: device-resident sweep timed with CUDA events (best of 3), optional device phase timers, and the
results (chi2, analyzer picks, LM counters) are compared with those of the FIRST library: a pure scheduling change
must reproduce them bit for bit, a change of summation order within rounding.
One JSON line per library; everything also lands in gpurun_out/ab_<tag>.json.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["planner", "solver", "T-pass", "H-pass", "gradient", "J assembly", "accept/convergence/output", "replay"]


def child(lib, spectra, phases, out_npz, cost_function):
    os.environ["MAXENT_B200_LIB"] = lib
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from maxent_b200 import batched, engine
    job = batched.BatchedTauMaxEnt(reduce_singular_space=1e-11, cost_function=cost_function,
                                   svd=os.environ.get("MX_AB_SVD", "jacobi"))
    G = batched.synthetic_bootstrap_batch(2000, 1000, spectra, seed=5)
    job.set_kernel_tau(np.linspace(0.0, 40.0, 2000), batched.hyperbolic_omega(-10.0, 10.0, 1000), beta=40.0)
    job.set_alpha_mesh_log(0.01, 2000.0, 60)
    job.set_error(1.e-4)
    prob = job.prepare()
    Gd = G.cuda()
    lm = engine.LMParams()
    best = 1e30
    res = None
    for i in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = engine.run_sweep(prob, Gd, job.alpha_effective(), want_v=False, analyze_results=True, lm=lm)
        e1.record()
        torch.cuda.synchronize()
        if i:
            best = min(best, e0.elapsed_time(e1))
    it = float(res.n_iter.sum())
    out = dict(lib=os.path.basename(lib), spectra=spectra, ms=round(best, 3), spectra_per_s=round(spectra / best * 1e3, 1),
               lm_iterations_per_spectrum=round(it / spectra, 4), trials_per_iteration=round(float(res.n_trial.sum()) / it, 3),
               batches_per_iteration=round(float(res.n_batch.sum()) / it, 4), n_sv=prob.n_sv)
    np.savez(out_npz, chi2=res.chi2.cpu().numpy(), S=res.S.cpu().numpy(), alpha_index=res.alpha_index.cpu().numpy(),
             n_iter=res.n_iter.cpu().numpy(), n_solve=res.n_solve.cpu().numpy(),
             A_pick=res.A_out[:, 0].cpu().numpy())
    if phases:
        for per_sm in ("full", "1"):
            os.environ["MX_MAX_CTAS_PER_SM"] = "99" if per_sm == "full" else per_sm
            n = spectra if per_sm == "full" else max(148, spectra // 4)
            for _ in range(2):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = engine.run_sweep(prob, Gd[:n], job.alpha_effective(), want_v=False, analyze_results=False,
                                     phase_timers=True, lm=lm)
                e1.record()
                torch.cuda.synchronize()
            cyc = r.phase_cycles.double().cpu().numpy()
            itn = float(r.n_iter.sum())
            nb = float(r.n_batch.sum())
            us = cyc.sum(0) / 1965.0 / itn
            out["phases_%s_per_sm" % per_sm] = dict(zip(NAMES, np.round(us, 2).tolist()))
            out["phases_%s_per_sm" % per_sm]["total"] = round(float(us.sum()), 2)
            out["phases_%s_per_sm" % per_sm]["tpass_us_per_batch"] = round(float(cyc[:, 2].sum() / 1965.0 / nb), 2)
            out["phases_%s_per_sm" % per_sm]["solver_us_per_batch"] = round(float(cyc[:, 1].sum() / 1965.0 / nb), 2)
            out["phases_%s_per_sm" % per_sm]["spectra_per_s_timing_build"] = round(n / e0.elapsed_time(e1) * 1e3, 1)
        os.environ.pop("MX_MAX_CTAS_PER_SM")
    print("ABRESULT " + json.dumps(out), flush=True)


def main():
    args = sys.argv[1:]
    if args and args[0] == "--child":
        child(args[1], int(args[2]), args[3] == "1", args[4], args[5])
        return
    import numpy as np
    spectra, phases, cf = 1184, False, "normal"
    libs = []
    i = 0
    while i < len(args):
        if args[i] == "--spectra":
            spectra = int(args[i + 1]); i += 2
        elif args[i] == "--cost-function":
            cf = args[i + 1]; i += 2
        elif args[i] == "--phases":
            phases = True; i += 1
        else:
            libs.append(os.path.abspath(args[i])); i += 1
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    ref = None
    for lib in libs:
        tag = os.path.basename(lib).replace("libmaxent_b200", "").replace(".so", "").strip("_") or "base"
        npz = os.path.join(ROOT, "gpurun_out", "ab_%s.npz" % tag)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", lib, str(spectra), "1" if phases else "0",
                            npz, cf], capture_output=True, text=True, timeout=900)
        line = [l for l in r.stdout.splitlines() if l.startswith("ABRESULT ")]
        if not line:
            print(json.dumps(dict(lib=tag, error=(r.stderr or r.stdout)[-2000:])), flush=True)
            continue
        out = json.loads(line[0][9:])
        d = np.load(npz)
        if ref is None:
            ref = {k: d[k] for k in d.files}
        else:
            c0, c1 = ref["chi2"], d["chi2"]
            rel = np.abs(c1 / c0 - 1.0)
            out["vs_first"] = dict(bitwise_chi2=bool(np.array_equal(c0, c1)),
                                   chi2_rel_max_alpha_lt_37=float(rel[:, :37].max()), chi2_rel_max=float(rel.max()),
                                   picks_equal=bool(np.array_equal(ref["alpha_index"], d["alpha_index"])),
                                   picks_differ=int((ref["alpha_index"] != d["alpha_index"]).any(axis=1).sum()),
                                   A_pick_rel_max=float(np.max(np.abs(d["A_pick"] - ref["A_pick"])) / np.max(np.abs(ref["A_pick"]))),
                                   n_iter_equal=bool(np.array_equal(ref["n_iter"], d["n_iter"])))
        print(json.dumps(out), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "ab_%s.json" % tag), "w") as f:
            json.dump(out, f)


if __name__ == "__main__":
    main()
