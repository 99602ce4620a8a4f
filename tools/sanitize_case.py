#!/usr/bin/env python
"""One thin layer through a given forward variant + the default backward kernels, checked against the oracle: small enough
to run under `compute-sanitizer --tool memcheck|racecheck|synccheck`.  python tools/sanitize_case.py <variant> [variant ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

H = int(os.environ.get("SANITIZE_H", "13"))   # 13: rows not 16-byte multiples (cp.async loaders); 28: TMA-eligible
spec = wl.ConvSpec("thin_conv3", 3, 32, 48, H, 3, 1, 1, 1, 0.88, True, True)
d = wl.make_layer_data(spec, 1)
g = po.Geom(spec.N, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
ocsr = po.weight_align(d["w"], g)
y_ref = po.conv_forward(d["x"], ocsr, g, d["bias"], relu=True)
geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
w = torch.from_numpy(d["w"]).cuda()
csr = capi.weight_align(w, geom)
x = torch.from_numpy(d["x"]).cuda()
b = torch.from_numpy(d["bias"]).cuda()
plan = capi.Plan(geom, csr)
bad = 0
for v in [int(a) for a in sys.argv[1:]] or [-1]:
    if v != -1:
        try:
            plan.set_variant(v)
        except capi.EscortError:
            print("forward  v%-3d does not apply to this geometry" % v, flush=True)
            continue
    y = plan.forward(x, b, relu=True)
    torch.cuda.synchronize()
    err = po.rel_l2(y.cpu().numpy(), y_ref)
    print("forward  v%-3d %-44s rel_l2 %.2e  %s" % (v, plan.kernel_name, err, plan.describe()[len(plan.kernel_name):60]), flush=True)
    bad += err > 1e-4
dy = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, y_ref.shape).astype(np.float32)).cuda()
wd = torch.zeros_like(w)
plan.backward_weight(x, dy, wd_dense=wd)
dx = plan.backward_data(dy)
torch.cuda.synchronize()
wd_o, _, dx_o = po.conv_backward(d["x"], dy.cpu().numpy(), d["w"], g, mask_only=True, want_b=False)
ew, ed = po.rel_l2(wd.cpu().numpy(), wd_o), po.rel_l2(dx.cpu().numpy(), dx_o)
print("backward %s  weight rel_l2 %.2e  data rel_l2 %.2e" % (plan.kernel_names(), ew, ed), flush=True)
bad += (ew > 1e-4) + (ed > 1e-4)
del plan
sys.exit(1 if bad else 0)
