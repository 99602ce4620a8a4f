#!/usr/bin/env python
"""Experiment: is the tile interpreter limited by instruction fetch?  Same nnz, but nonzeros restricted to few taps."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl

def run(w, spec, variant, tag):
    geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
    csr = capi.weight_align(torch.from_numpy(w).cuda(), geom)
    x = torch.rand((spec.N, spec.Cin, spec.H, spec.H), device="cuda")
    plan = capi.Plan(geom, csr); plan.set_variant(variant)
    y = torch.empty((spec.N, spec.Cout, plan.Ho, plan.Wo), device="cuda")
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.forward(x, None, top=y); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flops = 2.0 * plan.nnz * plan.Ho * plan.Wo * spec.N
    print("%-28s v%d %s nnz=%d: %.3f ms %.2f TF" % (tag, variant, plan.kernel_name, plan.nnz, best, flops / best / 1e9), flush=True)
    del plan

spec = wl.ALEXNET[1]
rng = np.random.default_rng(0)
shape = (spec.Cout, spec.Cin, 3, 3)
nnz_target = int(0.12 * np.prod(shape))
for variant in (1, 2):
    # (a) uniformly random positions
    w = np.zeros(shape, np.float32); idx = rng.choice(w.size, nnz_target, replace=False); w.reshape(-1)[idx] = 0.01
    run(w, spec, variant, "random taps")
    # (b) only centre tap, same nnz count (density 1.0 on that tap would be 98304 < nnz_target; use all centre + some (0,0))
    w = np.zeros(shape, np.float32); w[:, :, 1, 1] = 0.01
    rest = nnz_target - spec.Cout * spec.Cin
    if rest > 0:
        sub = rng.choice(spec.Cout * spec.Cin, rest, replace=False); w[:, :, 0, 0].reshape(-1)[sub] = 0.01
        ww = w[:, :, 0, 0].copy().reshape(-1); ww[sub] = 0.01; w[:, :, 0, 0] = ww.reshape(spec.Cout, spec.Cin)
    run(w, spec, variant, "centre(+corner) taps only")
    # (c) one output channel per block dense-ish: rows 0 mod OT only (few handlers: 9), same nnz
    w = np.zeros(shape, np.float32)
    rows = np.arange(0, spec.Cout, 8)
    per_row = nnz_target // len(rows)
    for r in rows:
        idx = rng.choice(spec.Cin * 9, min(per_row, spec.Cin * 9), replace=False); w[r].reshape(-1)[idx] = 0.01
    run(w, spec, variant, "1 row per 8 (9 handlers)")
