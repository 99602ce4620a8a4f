#!/usr/bin/env python
"""Run one layer's forward a few times (for ncu / quick timing): python tools/run_layer.py <net> <idx> <variant> [iters] [N]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402

net, idx, variant = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
spec = wl.NETWORKS[net][idx]
if len(sys.argv) > 5:
    spec = spec._replace(N=int(sys.argv[5]))
d = wl.make_layer_data(spec, idx)
geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
csr = capi.weight_align(torch.from_numpy(d["w"]).cuda(), geom)
x = torch.from_numpy(d["x"]).cuda()
b = torch.from_numpy(d["bias"]).cuda() if d["bias"] is not None else None
plan = capi.Plan(geom, csr)
if variant != -1:
    plan.set_variant(variant)
print(spec.name, plan.describe(), flush=True)
y = torch.empty((spec.N, spec.Cout, plan.Ho, plan.Wo), device="cuda")
flops, _ = wl.alg_work(spec, plan.nnz)
best = 1e9
for _ in range(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.forward(x, b, top=y)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1)
    best = min(best, ms)
print("RESULT %s v%d %s: %.3f ms  %.2f TFLOP/s  %.0f img/s" % (spec.name, variant, plan.kernel_name, best, flops / best / 1e9, spec.N / best * 1e3), flush=True)
del plan
