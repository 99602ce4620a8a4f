#!/usr/bin/env python
"""For every layer of the given networks: autotune the plan and print the chosen kernel / layout / time.
python tools/tune_report.py alexnet googlenet resnet50"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402

seen = {}
for net in sys.argv[1:]:
    tot_ms, tot_fl = 0.0, 0.0
    for idx, spec in enumerate(wl.NETWORKS[net]):
        key = (spec.Cin, spec.Cout, spec.H, spec.k, spec.group, spec.sparsity, spec.N)
        d = wl.make_layer_data(spec, idx)
        geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
        csr = capi.weight_align(torch.from_numpy(d["w"]).cuda(), geom)
        x = torch.from_numpy(d["x"]).cuda()
        b = torch.from_numpy(d["bias"]).cuda() if d["bias"] is not None else None
        plan = capi.Plan(geom, csr)
        plan.autotune(spec.N)
        y = torch.empty((spec.N, spec.Cout, plan.Ho, plan.Wo), device="cuda")
        flops, _ = wl.alg_work(spec, plan.nnz)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.forward(x, b, top=y)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        tot_ms += best
        tot_fl += flops
        print("%-30s %.3f ms %5.2f TF | %s" % (spec.name, best, flops / best / 1e9, plan.describe()[:150]), flush=True)
        del plan
    print("== %s: %.3f ms total, %.2f TF average" % (net, tot_ms, tot_fl / tot_ms / 1e9), flush=True)
