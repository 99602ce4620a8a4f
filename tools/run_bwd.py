#!/usr/bin/env python
"""Time forward / backward-data / backward-weight of layers: python tools/run_bwd.py net:idx ...  (ESCORT_GENERIC_BACKWARD=1 for the generic kernels)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402


def timeit(fn, n=4):
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for arg in sys.argv[1:]:
    net, idx = arg.split(":")
    idx = int(idx)
    spec = wl.NETWORKS[net][idx]
    d = wl.make_layer_data(spec, idx)
    geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
    w = torch.from_numpy(d["w"]).cuda()
    csr = capi.weight_align(w, geom)
    x = torch.from_numpy(d["x"]).cuda()
    plan = capi.Plan(geom, csr)
    y = plan.forward(x, None)
    dy = torch.rand_like(y) * 2 - 1
    dx = torch.empty_like(x)
    wd = torch.zeros_like(w)
    flops, _ = wl.alg_work(spec, plan.nnz)
    t_f = timeit(lambda: plan.forward(x, None, top=y))
    t_d = timeit(lambda: plan.backward_data(dy, dx))
    t_w = timeit(lambda: plan.backward_weight(x, dy, wd_dense=wd, accumulate=False))
    wd.zero_()   # the dense diff accumulates (Caffe's contract): one clean pass for the cross-check
    plan.backward_weight(x, dy, wd_dense=wd, accumulate=False)
    # cross-check against the generic kernels
    os.environ["ESCORT_GENERIC_BACKWARD"] = "1"
    dx2 = torch.empty_like(x)
    wd2 = torch.zeros_like(w)
    plan.backward_data(dy, dx2)
    plan.backward_weight(x, dy, wd_dense=wd2, accumulate=False)
    torch.cuda.synchronize()
    del os.environ["ESCORT_GENERIC_BACKWARD"]
    e_d = float((dx - dx2).norm() / dx2.norm())
    e_w = float((wd - wd2).norm() / wd2.norm())
    print("%-28s fwd %.3f ms %5.2f TF | bwd-data %.3f ms %5.2f TF err %.1e | bwd-weight %.3f ms %5.2f TF err %.1e | %s" % (
        spec.name, t_f, flops / t_f / 1e9, t_d, flops / t_d / 1e9, e_d, t_w, flops / t_w / 1e9, e_w, plan.kernel_name), flush=True)
    del plan
