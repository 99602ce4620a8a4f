#!/usr/bin/env python
"""The round-2 kernels outside the sparse tile / TMEM families, on thin cases small enough for `compute-sanitizer
--tool memcheck|racecheck|synccheck`: the tcgen05 inner product, the dense convolution through both A sources (implicit
GEMM from NCHW with ragged pixel / channel / output-channel tails + residual; column buffer), and the small-map sparse
forward with and without padding.  python tools/sanitize_dense.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi, workloads as wl  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

rng = np.random.default_rng(7)
bad = 0
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

x = rng.uniform(-1, 1, (37, 132)).astype(np.float32)
w = (rng.standard_normal((130, 132)) / 11).astype(np.float32)
b = rng.standard_normal(130).astype(np.float32)
y = capi.inner_product_forward(cu(x), cu(w), cu(b), relu=True)
torch.cuda.synchronize()
ref = np.maximum(x.astype(np.float64) @ w.astype(np.float64).T + b, 0)
err = np.linalg.norm(y.cpu().numpy() - ref) / np.linalg.norm(ref)
print("inner product            dense_gemm_tf32_kernel          rel_l2 %.2e" % err, flush=True)
bad += err > 2e-3

for name, N, Cin, Cout, H, k, s, p in [("implicit 1x1", 3, 36, 300, 14, 1, 1, 0), ("column buffer 3x3", 2, 20, 130, 13, 3, 1, 1),
                                       ("column buffer 1x1 7px", 2, 64, 48, 7, 1, 1, 0)]:
    xx = rng.uniform(-1, 1, (N, Cin, H, H)).astype(np.float32)
    ww = (rng.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float32)
    bb = rng.standard_normal(Cout).astype(np.float32)
    g = po.Geom(N, Cin, H, H, Cout, k, s, p, 1, 1)
    Ho = (H + 2 * p - k) // s + 1
    res = rng.uniform(-1, 1, (N, Cout, Ho, Ho)).astype(np.float32)
    ref = np.maximum(po.dense_conv(xx, ww, g, bb, relu=False) + res, 0)
    geom = capi.make_geom(Cin, Cout, H, H, k, s, p, 1, 1)
    yy = capi.dense_conv_forward(geom, cu(xx), cu(ww), cu(bb), relu=True, residual=cu(res))
    torch.cuda.synchronize()
    err = po.rel_l2(yy.cpu().numpy(), ref)
    print("dense conv %-22s dense_conv_tf32_kernel  rel_l2 %.2e" % (name, err), flush=True)
    bad += err > 2e-3

for name, spec in [("no padding", wl.ConvSpec("lenet2", 4, 20, 50, 12, 5, 1, 0, 1, 0.80, True, False)),
                   ("padding", wl.ConvSpec("thin3", 3, 32, 48, 13, 3, 1, 1, 1, 0.88, True, True))]:
    d = wl.make_layer_data(spec, 2)
    g = po.Geom(spec.N, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
    ref = po.conv_forward(d["x"], po.weight_align(d["w"], g), g, d["bias"], relu=True)
    geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
    plan = capi.Plan(geom, capi.weight_align(cu(d["w"]), geom))
    plan.set_variant(0)
    yy = plan.forward(cu(d["x"]), cu(d["bias"]), relu=True)
    torch.cuda.synchronize()
    err = po.rel_l2(yy.cpu().numpy(), ref)
    print("small map %-23s %-23s rel_l2 %.2e" % (name, plan.kernel_name, err), flush=True)
    bad += err > 1e-4
    del plan
# stride 2 through the space-to-depth sub-plan: s2d_pad_kernel + stride-1 kernel, and backward data = the sub-plan's backward
# plan + d2s_unpad_kernel
spec = wl.ConvSpec("stride2", 3, 16, 24, 15, 3, 2, 1, 1, 0.7, True, False)
d = wl.make_layer_data(spec, 3)
g = po.Geom(spec.N, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
ref = po.conv_forward(d["x"], po.weight_align(d["w"], g), g, d["bias"], relu=False)
geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
plan = capi.Plan(geom, capi.weight_align(cu(d["w"]), geom))
plan.set_config(-2, 0)
yy = plan.forward(cu(d["x"]), cu(d["bias"]), relu=False)
dy = rng.uniform(-1, 1, ref.shape).astype(np.float32)
dx = plan.backward_data(cu(dy))
torch.cuda.synchronize()
_, _, dx_o = po.conv_backward(d["x"], dy, d["w"], g, mask_only=True, want_w=False, want_b=False)
e1, e2 = po.rel_l2(yy.cpu().numpy(), ref), po.rel_l2(dx.cpu().numpy(), dx_o)
print("stride 2 space-to-depth           %-23s fwd rel_l2 %.2e  bwd-data rel_l2 %.2e" % (plan.describe().split(" ")[0] + " (s2d)", e1, e2), flush=True)
bad += (e1 > 1e-4) + (e2 > 1e-4)
del plan
sys.exit(1 if bad else 0)
