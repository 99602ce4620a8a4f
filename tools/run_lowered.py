#!/usr/bin/env python
"""The paper's second baseline on B200: LOWERED_SPARSE (per-image im2col + cusparseSpMM, escort_lowered_sparse_forward)
timed next to the direct sparse convolution on the BASELINE layers.  python tools/run_lowered.py [net ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402


def best_ms(fn, n=3):
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for net in sys.argv[1:] or ["alexnet"]:
    seen = set()
    for idx, spec in enumerate(wl.NETWORKS[net]):
        shape = (spec.Cin, spec.Cout, spec.H, spec.k, spec.stride, spec.pad, spec.group)
        if shape in seen:
            continue
        seen.add(shape)
        d = wl.make_layer_data(spec, idx)
        geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
        w = torch.from_numpy(d["w"]).cuda()
        x = torch.from_numpy(d["x"]).cuda()
        b = torch.from_numpy(d["bias"]).cuda() if d["bias"] is not None else None
        csr = capi.weight_align(w, geom)
        csr_raw = capi.weight_align(w, geom, stretch=False)
        plan = capi.Plan(geom, csr)
        plan.autotune(spec.N)
        y = torch.empty((spec.N, spec.Cout, plan.Ho, plan.Wo), device="cuda")
        flops, _ = wl.alg_work(spec, plan.nnz)
        t_direct = best_ms(lambda: plan.forward(x, b, relu=spec.relu, top=y))
        y2 = capi.lowered_sparse_forward(geom, x, csr_raw, b, relu=spec.relu)   # warm-up (handle, workspace)
        t_low = best_ms(lambda: capi.lowered_sparse_forward(geom, x, csr_raw, b, relu=spec.relu), n=2)
        err = float(torch.linalg.vector_norm((y2 - y).double()) / torch.linalg.vector_norm(y.double()))
        print("%-30s direct %-34s %8.3f ms %6.2f TFLOP/s | lowered-sparse (im2col + cusparseSpMM, per image) %9.3f ms %6.2f TFLOP/s | x%.1f  rel_l2 %.1e"
              % (spec.name, plan.kernel_name, t_direct, flops / t_direct / 1e9, t_low, flops / t_low / 1e9, t_low / t_direct, err), flush=True)
        del plan
