#!/bin/bash
# The measurement batch behind profiles/r02_*: run on a B200 box from the repo root.
#   gpurun --timeout 2400 -- 'bash tools/round_batch.sh'
# Everything lands in gpurun_out/; numbers printed under ncu or the sanitizer are never bench values.
set -u
O=gpurun_out; mkdir -p $O
python -c 'import __graft_entry__ as g; g.smoke()' > $O/r02_smoke_final.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_gpu_tests_final.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_gpu_tests_final.log
python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err; echo "bench rc=$?"; cat $O/r02_bench_default.json
python bench.py --impl reference --steps 1 --warmup 1 > $O/r02_bench_reference_arm.json 2>/dev/null; echo "ref arm rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sweep2_kernel|project_kernel|analyze_kernel|jacobi_kernel|cgs_kernel|gemm_kernel|layout_V_kernel|tau_kernel_kernel|finish_kernel|col_norm_kernel|randn_kernel|transpose_kernel" -c 400 --csv --log-file $O/r02_launches_step_8192.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r02_launches_bench.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sweep2_kernel -c 1 -f -o $O/r02c_sweep2_full \
    python bench.py --spectra 296 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02c_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 compute-sanitizer --tool memcheck python -c 'import __graft_entry__ as g; g.smoke()' > $O/r02_sanitizer_memcheck_smoke.txt 2>&1; echo "memcheck smoke rc=$?"; tail -2 $O/r02_sanitizer_memcheck_smoke.txt
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/r02_sanitizer_memcheck_fullshape.txt 2>&1; echo "memcheck full-shape rc=$?"; tail -2 $O/r02_sanitizer_memcheck_fullshape.txt
python tools/config_latency.py > $O/r02_config_latency.json 2> $O/r02_config_latency.err; echo "latency rc=$?"
python tools/phase_times.py > $O/r02_phase_times.txt 2>&1; echo "phases rc=$?"
python tools/wide_bench.py > $O/r02_wide_bench.json 2> $O/r02_wide_bench.err; echo "wide rc=$?"
ls -la $O/r02c_sweep2_full.ncu-rep
