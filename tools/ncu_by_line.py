#!/usr/bin/env python
"""Aggregate the warp-stall samples of an ncu report by CUDA source line.

    python tools/ncu_by_line.py gpurun_out/prof.ncu-rep maxent_b200/build/mx_sweep2_nt7.o [--top 40]

ncu's CSV source page lists SASS only; the line table comes from nvdisasm --print-line-info on the
cubin of the same object file (instruction order is identical, opcodes are cross-checked).
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(obj):
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
    out = []
    for f in sorted(os.listdir(d)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(d, f)], capture_output=True, text=True).stdout
        cur = (None, 0)
        chain = []
        fresh = True
        func = None
        for line in txt.splitlines():
            m = re.match(r"\s*//## File \"([^\"]*)\", line (\d+)", line)
            if m:
                if fresh:
                    chain = []
                    fresh = False
                chain.append((os.path.basename(m.group(1)), int(m.group(2))))
                # innermost location inside the kernel source file, outermost = phase
                inner = next((c for c in chain if c[0] == MAINFILE), chain[0])
                cur = (inner, chain[-1])
                continue
            m = re.match(r"\s*\.text\.(\S+):", line)
            if m:
                func = m.group(1)
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                out.append((func, int(m.group(1), 16), m.group(2).strip(), cur))
                fresh = True
    return out


MAINFILE = "mx_sweep2.cuh"


def main():
    global MAINFILE
    rep, obj = sys.argv[1], sys.argv[2]
    if "--file" in sys.argv:
        MAINFILE = sys.argv[sys.argv.index("--file") + 1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = rows[2:]
    sass = sass_lines(obj)
    # pick the function whose instruction count matches
    byf = defaultdict(list)
    for f, off, op, cur in sass:
        byf[f].append((off, op, cur))
    cand = [f for f, v in byf.items() if len(v) == len(body)]
    if not cand:
        print("no function with %d instructions; have %s" % (len(body), {f: len(v) for f, v in byf.items()}))
        return
    ins = byf[cand[0]]
    for which, title in ((1, "by outermost line (phase)"), (0, "by innermost line in " + MAINFILE)):
        per_line = defaultdict(lambda: defaultdict(float))
        total = 0.0
        for r, (off, op, cur) in zip(body, ins):
            n = float(r[isamp] or 0)
            total += n
            key = cur[which] if isinstance(cur[0], tuple) else cur
            per_line[key]["samples"] += n
            per_line[key]["exec"] += float(r[iexec] or 0)
            if "DMMA" in op:
                per_line[key]["dmma"] += float(r[iexec] or 0)
            for i, h in stall_cols:
                per_line[key][h] += float(r[i] or 0)
        print("== %s; total samples %d" % (title, total))
        for cur, d in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
            st = sorted(((v, k) for k, v in d.items() if k.startswith("stall_")), reverse=True)[:4]
            print("%-18s:%4d  %6.2f%%  exec %12d dmma %11d  %s" % (cur[0], cur[1], 100 * d["samples"] / total, d["exec"], d["dmma"],
                                                       " ".join("%s=%.0f%%" % (k[6:], 100 * v / max(d["samples"], 1)) for v, k in st)))


if __name__ == "__main__":
    main()
