#!/usr/bin/env python
"""Summarise an .ncu-rep of the tile kernel: headline counters + where the warp-stall samples fall.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--top N]"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
m = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__lsu_writeback_active_mem_lg.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
for k in keys:
    if k in m:
        print("%-80s %s" % (k, m[k]))
for k, v in m.items():
    if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", k):
        try:
            if float(v) > 0.08:
                print("%-80s %s" % (k.replace("smsp__average_warps_issue_stalled_", "stall/issue: ").replace("_per_issue_active.ratio", ""), v))
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        break  # next kernel's section
    data.append(r)
S = lambda r, k: int(r[ix[k]] or 0)
tot = sum(S(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
# classify regions: loader = between first LDGSTS-related and ..., handlers = FFMA blocks
cls = Counter()
exe = Counter()
for r in data:
    t = r[ix["Source"]]
    op = t.split()[1] if t.startswith("@") else t.split()[0]
    op = op.split(".")[0]
    cls[op] += S(r, "# Samples")
    exe[op] += S(r, "Instructions Executed")
for op, c in cls.most_common(14):
    print("  %-10s samples %6d (%4.1f%%)  execs %d" % (op, c, 100.0 * c / tot, exe[op]))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("top instructions:")
for r in sorted(data, key=lambda r: -S(r, "# Samples"))[:top]:
    st = sorted(((S(r, h), h[6:]) for h in stall_cols), reverse=True)[:3]
    print("  %s %-48s %6d (%4.1f%%) exec %-9s %s" % (r[ix["Address"]][-5:], r[ix["Source"]][:48], S(r, "# Samples"),
                                                   100.0 * S(r, "# Samples") / tot, r[ix["Instructions Executed"]],
                                                   " ".join("%s=%d" % (n, v) for v, n in st if v)))
