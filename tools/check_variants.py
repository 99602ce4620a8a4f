#!/usr/bin/env python
"""Every forward variant that applies to a layer, at a given batch, against the generic kernel (one thread per output, the
always-correct path): python tools/check_variants.py <net>:<idx> <N> [N ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402

net, idx = sys.argv[1].split(":")
for N in [int(a) for a in sys.argv[2:]]:
    spec = wl.NETWORKS[net][int(idx)]._replace(N=N)
    d = wl.make_layer_data(spec, int(idx))
    geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
    w = torch.from_numpy(d["w"]).cuda()
    x = torch.from_numpy(d["x"]).cuda()
    b = torch.from_numpy(d["bias"]).cuda() if d["bias"] is not None else None
    plan = capi.Plan(geom, capi.weight_align(w, geom))
    plan.set_variant(0)
    ref = plan.forward(x, b, relu=True).clone()
    bad = 0
    for v in range(1, 80):
        for rank in range(4):
            try:
                plan.set_config(v, rank)
            except capi.EscortError:
                break
            y = plan.forward(x, b, relu=True)
            torch.cuda.synchronize()
            err = float((y - ref).norm() / ref.norm())
            flag = "" if err < 1e-4 else "   <-- MISMATCH"
            bad += err >= 1e-4
            if flag or rank == 0:
                print("N %4d v%-3d rank %d %-44s rel_l2 %.2e%s" % (N, v, rank, plan.kernel_name, err, flag), flush=True)
    print("N %d: %d mismatching configurations" % (N, bad), flush=True)
    del plan
