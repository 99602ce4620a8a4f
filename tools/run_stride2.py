#!/usr/bin/env python
"""Stride-2 layers of the config-5 sweep (N = 64, 3x3, pad 1, 70 % sparse): forward and backward data through the
space-to-depth sub-plans against the kernels on the strided input / the generic backward kernel.
python tools/run_stride2.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402


def best_ms(fn, n=4):
    fn()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


specs = [s for s in wl.sweep_specs(64) if s.stride == 2 and abs(s.sparsity - 0.7) < 1e-6 and s.Cin >= 128 and s.H >= 14]
for li, spec in enumerate(specs):
    d = wl.make_layer_data(spec, li)
    geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
    w = torch.from_numpy(d["w"]).cuda()
    x = torch.from_numpy(d["x"]).cuda()
    res = {}
    for mode in ("s2d", "direct"):
        if mode == "direct":
            os.environ["ESCORT_NO_S2D"] = "1"
        else:
            os.environ.pop("ESCORT_NO_S2D", None)
        plan = capi.Plan(geom, capi.weight_align(w, geom))
        if mode == "s2d":
            plan.set_config(-2, 0)
        else:
            plan.autotune(spec.N)
        y = plan.forward(x, None)
        dy = torch.rand(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(li)) * 2 - 1   # the same in both modes
        dx = torch.empty_like(x)
        res[mode] = (best_ms(lambda: plan.forward(x, None, top=y)), best_ms(lambda: plan.backward_data(dy, dx)), dx.clone(),
                     plan.kernel_name)
        del plan
    flops, _ = wl.alg_work(spec, int((d["w"] != 0).sum()))
    err = float((res["s2d"][2] - res["direct"][2]).norm() / res["direct"][2].norm())
    print("%-22s fwd s2d %.3f ms %5.1f TF | strided %.3f ms %5.1f TF (%s) || bwd-data s2d %.3f ms %5.1f TF | generic %.3f ms %5.1f TF | rel_l2 %.1e"
          % (spec.name, res["s2d"][0], flops / res["s2d"][0] / 1e9, res["direct"][0], flops / res["direct"][0] / 1e9, res["direct"][3][:28],
             res["s2d"][1], flops / res["s2d"][1] / 1e9, res["direct"][1], flops / res["direct"][1] / 1e9, err), flush=True)
