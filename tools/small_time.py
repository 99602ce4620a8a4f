#!/usr/bin/env python
"""LeNet-sized layers (BASELINE configs[0]): GPU time per launch of the variant-0 kernels (sconv_fwd_small / sconv_fwd_generic)
and of the autotuned plan, measured inside a CUDA graph of 50 launches so that the host's ~10 us per ctypes call is not
what is timed.  python tools/small_time.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402

REP = 50
for idx in (0, 1):
    spec = wl.LENET[idx]
    d = wl.make_layer_data(spec, idx)
    geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
    w = torch.from_numpy(d["w"]).cuda()
    x = torch.from_numpy(d["x"]).cuda()
    b = torch.from_numpy(d["bias"]).cuda()
    for mode in ("small", "generic", "autotuned"):
        if mode == "generic":
            os.environ["ESCORT_NO_SMALL_MAPS"] = "1"
        else:
            os.environ.pop("ESCORT_NO_SMALL_MAPS", None)
        plan = capi.Plan(geom, capi.weight_align(w, geom))
        if mode == "autotuned":
            plan.autotune(spec.N)
        else:
            plan.set_variant(0)
        y = plan.forward(x, b, relu=True)
        st = torch.cuda.Stream()
        plan.forward(x, b, relu=True, top=y, stream=st)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for _ in range(REP):
                plan.forward(x, b, relu=True, top=y, stream=torch.cuda.current_stream())
        graph.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) / REP)
        flops, _ = wl.alg_work(spec, plan.nnz)
        print("%-12s %-10s %-28s %.4f ms per launch  %.2f TFLOP/s" % (spec.name, mode, plan.kernel_name, best, flops / best / 1e9), flush=True)
