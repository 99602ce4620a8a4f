#!/usr/bin/env python
"""Print the tilings of the forward / backward-data / backward-weight plans of layers: python tools/desc.py net:idx ..."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402

for arg in sys.argv[1:]:
    net, idx = arg.split(":")
    spec = wl.NETWORKS[net][int(idx)]._replace(N=8)
    d = wl.make_layer_data(spec, int(idx))
    geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
    w = torch.from_numpy(d["w"]).cuda()
    plan = capi.Plan(geom, capi.weight_align(w, geom))
    x = torch.from_numpy(d["x"]).cuda()
    y = plan.forward(x, None)
    plan.backward_data(y)
    plan.backward_weight(x, y, wd_dense=torch.zeros_like(w))
    torch.cuda.synchronize()
    print(spec.name)
    for part in plan.describe().split(" | "):
        print("   ", part)
