#!/usr/bin/env python
"""Level 1 of INTEGRATION.md timed once, so that nobody mistakes it for the product: escort_copy_input + escort_sconv_padded
driven per image and per group exactly like the reference's forward_gpu_sconv loop (base_conv_layer.cpp:749-798; the
compat kernel is the reference-class algorithm behind the reference's own argument list), next to Level 2 (one native
launch for the whole batch).  The per-image loop is replayed from a CUDA graph: GPU time only, no Python in it.
python tools/run_level1.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402

N = 64
print("AlexNet layers, %d images, forward with bias + ReLU (ms; images/s)" % N)
for idx in range(4):
    spec = wl.ALEXNET[idx]._replace(N=N)
    d = wl.make_layer_data(spec, idx)
    geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
    w = torch.from_numpy(d["w"]).cuda()
    x = torch.from_numpy(d["x"]).cuda()
    b = torch.from_numpy(d["bias"]).cuda()
    csr = capi.weight_align(w, geom)
    plan = capi.Plan(geom, csr)
    plan.autotune(N)
    y = plan.forward(x, b, relu=True)
    H, P, K, G = spec.H, spec.pad, spec.k, spec.group
    M, Cg = spec.Cout // G, spec.Cin // G
    Ho = plan.Ho
    plen = spec.Cin * (H + P) * (H + P) + P * (H + 2 * P)
    padded = torch.zeros(plen, device="cuda")
    top = torch.zeros_like(y)
    ifmap = Cg * (H + P) * (H + P)
    woff = M * Cg * K * K

    def level1():
        for n in range(N):
            capi.copy_input(padded, x[n], spec.Cin, H, H, P, P)
            for gi in range(G):
                capi.sconv_padded(True, 1, padded.data_ptr() + 4 * gi * ifmap, ifmap, csr["rowptr"].data_ptr() + 4 * (M + 1) * gi,
                                  csr["colidx"].data_ptr() + 4 * woff * gi, csr["values"].data_ptr() + 4 * woff * gi,
                                  b.data_ptr() + 4 * M * gi, H, H, P, P, 1, 1, 1, 1, K, K, top[n].data_ptr() + 4 * gi * M * Ho * Ho, M, G)

    st = torch.cuda.Stream()
    torch.cuda.synchronize()   # (the zero fills above ran on the default stream)
    with torch.cuda.stream(st):
        level1()
    torch.cuda.synchronize()
    err = float((top - y).norm() / y.norm())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=st):
        level1()

    def best(fn):
        fn()
        t = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            t = min(t, e0.elapsed_time(e1))
        return t

    t1 = best(graph.replay)
    t2 = best(lambda: plan.forward(x, b, relu=True, top=y))
    print("  %-14s Level 1 (per image: copy_input + %d x sconv_padded) %8.3f ms %9.0f | Level 2 (%s) %7.3f ms %9.0f | %5.1fx | rel_l2 between them %.1e"
          % (spec.name, G, t1, N / t1 * 1e3, plan.kernel_name, t2, N / t2 * 1e3, t1 / t2, err), flush=True)
