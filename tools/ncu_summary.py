#!/usr/bin/env python
"""Headline metrics of an `ncu --set full` report as one JSON object (the profiles/*_ncu_full_summary.json files).

    python tools/ncu_summary.py gpurun_out/r02b_sweep2_full.ncu-rep "what was captured" > profiles/r02b_sweep2_ncu_full_summary.json
"""
import csv
import io
import json
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active")


def main():
    rep, what = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for h, u, v in zip(head, units, vals):
        if h in KEEP or h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                out[h] = float(v.replace(",", ""))
            except ValueError:
                out[h] = v
            if u and h in KEEP:
                out[h + " [unit]"] = u
    out["_kernel"] = vals[head.index("Kernel Name")] if "Kernel Name" in head else ""
    out["_what"] = what
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
