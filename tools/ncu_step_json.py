#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of one bench step (one launch per layer, in layer order) into JSON and merge the
per-launch DRAM traffic into profiles/traffic.json (what bench.py reports as roofline.traffic).
python tools/ncu_step_json.py <rep> <bench line .json> <out .json>"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, bench_json, out = sys.argv[1:4]
bench = json.loads(open(bench_json).read().strip().splitlines()[-1])
layers = [(L["layer"], L["kernel"]) for L in bench["layers"] if L.get("op", "fwd") == "fwd"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
ix = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
launches = []
for (layer, kernel), r in zip(layers, data):
    d = {"layer": layer, "kernel": kernel, "ncu_kernel": r[ix["Kernel Name"]][:60]}
    for k in KEYS:
        if k in ix:
            d[k] = "%s %s" % (r[ix[k]], units[ix[k]])
    tr = sum(float(r[ix[k]].replace(",", "")) * SCALE.get(units[ix[k]], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    d["dram_traffic_bytes"] = tr
    launches.append(d)
json.dump({"source": "ncu --set full --clock-control none, one launch per layer of bench.py's step (tools/reproduce_profiles.sh r02)",
           "launches": launches}, open(out, "w"), indent=1)
tp = os.path.join(ROOT, "profiles", "traffic.json")
traffic = json.load(open(tp)) if os.path.exists(tp) else {}
for d in launches:
    traffic.setdefault(d["layer"], {})[d["kernel"]] = d["dram_traffic_bytes"]
json.dump(traffic, open(tp, "w"), indent=1)
for d in launches:
    print(d["layer"], d["kernel"], d["gpu__time_duration.sum"], "traffic %.1f MB" % (d["dram_traffic_bytes"] / 1e6),
          "fma", d["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed"], "issue", d["smsp__issue_active.avg.pct_of_peak_sustained_active"])
