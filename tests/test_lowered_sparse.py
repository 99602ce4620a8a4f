"""The LOWERED_SPARSE comparator (im2col + CSR x dense, src/caffe/layers/base_conv_layer.cpp:715-745,
src/caffe/util/math_functions.cu:48-62) against the same oracle as the direct path: 1e-4 relative L2 in fp32."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [  # name, N, Cin, Cout, H, k, stride, pad, group, sparsity, bias, relu
    ("thin_conv3", 5, 32, 48, 13, 3, 1, 1, 1, 0.88, True, True),
    ("thin_conv2_g2", 3, 16, 32, 27, 5, 1, 2, 2, 0.85, True, True),
    ("stride2_nopad", 2, 12, 20, 15, 3, 2, 0, 1, 0.7, False, False),
    ("pointwise", 3, 40, 24, 14, 1, 1, 0, 1, 0.6, True, False),
    ("empty_group", 2, 8, 8, 6, 3, 1, 1, 2, 0.5, True, False),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_lowered_sparse_matches_the_oracle(case):
    import torch
    from caffe_escoin_b200 import capi, workloads as wl
    from oracle import pyoracle as po
    name, N, Cin, Cout, H, k, stride, pad, group, sparsity, has_bias, relu = case
    spec = wl.ConvSpec(name, N, Cin, Cout, H, k, stride, pad, group, sparsity, has_bias, relu)
    d = wl.make_layer_data(spec, 7)
    w = d["w"].copy()
    if name == "empty_group":
        w[Cout // 2:] = 0.0          # the second group has no nonzero at all
    g = po.Geom(N, Cin, H, H, Cout, k, stride, pad, 1, group)
    ocsr = po.weight_align(w, g)
    y_ref = po.conv_forward(d["x"], ocsr, g, d["bias"], relu=relu)
    geom = capi.make_geom(Cin, Cout, H, H, k, stride, pad, 1, group)
    csr_raw = capi.weight_align(torch.from_numpy(w).cuda(), geom, stretch=False)
    bias = torch.from_numpy(d["bias"]).cuda() if d["bias"] is not None else None
    y = capi.lowered_sparse_forward(geom, torch.from_numpy(d["x"]).cuda(), csr_raw, bias, relu=relu)
    torch.cuda.synchronize()
    assert y.shape == y_ref.shape
    assert po.rel_l2(y.cpu().numpy(), y_ref) < 1e-4
