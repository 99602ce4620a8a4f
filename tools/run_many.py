#!/usr/bin/env python
"""Time several (net, layer index, variant) triples in one process: python tools/run_many.py net:idx:variant[,variant...] ..."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402

for arg in sys.argv[1:]:
    net, idx, vs = arg.split(":")
    idx = int(idx)
    spec = wl.sweep_specs(64)[idx] if net == "sweep" else wl.NETWORKS[net][idx]
    d = wl.make_layer_data(spec, idx)
    geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
    csr = capi.weight_align(torch.from_numpy(d["w"]).cuda(), geom)
    x = torch.from_numpy(d["x"]).cuda()
    b = torch.from_numpy(d["bias"]).cuda() if d["bias"] is not None else None
    plan = capi.Plan(geom, csr)
    y = torch.empty((spec.N, spec.Cout, plan.Ho, plan.Wo), device="cuda")
    flops, _ = wl.alg_work(spec, plan.nnz)
    yref = None
    for vt in vs.split(","):
        v, rank = (int(t) for t in (vt.split("r") + ["0"])[:2])   # "4r2" = variant 4, tiling candidate 2
        try:
            plan.set_config(v, rank)
        except capi.EscortError:
            print("%s v%d unsupported" % (spec.name, v))
            continue
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.forward(x, b, top=y)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        if yref is None:
            yref = y.clone()
        err = float((y - yref).norm() / yref.norm())
        print("%-28s v%-2d %.3f ms %5.2f TF err %.1e | %s" % (spec.name, v, best, flops / best / 1e9, err, plan.describe()),
              flush=True)
    del plan
