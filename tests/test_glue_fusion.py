"""f2: conv -> BatchNorm(use_global_stats) -> Scale -> ReLU as one forward launch (escort_plan_fold_affine) against the
four layers evaluated one after the other in fp32 (oracle conv, then the arithmetic of
src/caffe/layers/batch_norm_layer.cpp:98-152 and the Scale layer)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [-1, 0, 51, 59])
@pytest.mark.parametrize("with_bias", [False, True])
def test_conv_bn_scale_relu_in_one_launch(variant, with_bias):
    import torch
    from caffe_escoin_b200 import capi, workloads as wl
    from oracle import pyoracle as po
    spec = wl.ConvSpec("res_like", 4, 32, 40, 14, 3, 1, 1, 1, 0.7, with_bias, False)
    d = wl.make_layer_data(spec, 11)
    rng = np.random.default_rng(5)
    M = spec.Cout
    sf_blob = np.float32(3.0)                                  # BatchNorm's third blob (moving-average normaliser)
    mean = (rng.standard_normal(M) * 0.5).astype(np.float32) * sf_blob
    var = rng.uniform(0.2, 2.0, M).astype(np.float32) * sf_blob
    gamma = rng.uniform(-1.5, 1.5, M).astype(np.float32)
    beta = rng.standard_normal(M).astype(np.float32)
    eps = np.float32(1e-5)
    g = po.Geom(spec.N, spec.Cin, spec.H, spec.H, M, spec.k, spec.stride, spec.pad, 1, spec.group)
    ocsr = po.weight_align(d["w"], g)
    conv = po.conv_forward(d["x"], ocsr, g, d["bias"], relu=False)
    # the reference's layers, one after the other, in fp32
    m_, v_ = mean * np.float32(1.0 / sf_blob), var * np.float32(1.0 / sf_blob)
    bn = (conv - m_[None, :, None, None]) / np.sqrt(v_ + eps)[None, :, None, None]
    ref = np.maximum(bn * gamma[None, :, None, None] + beta[None, :, None, None], 0).astype(np.float32)

    geom = capi.make_geom(spec.Cin, M, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
    w = torch.from_numpy(d["w"]).cuda()
    plan = capi.Plan(geom, capi.weight_align(w, geom))
    if variant >= 0:
        plan.set_variant(variant)
    cu = lambda a: torch.from_numpy(a).cuda()
    a, b = capi.bn_scale_to_affine(cu(mean), cu(var), float(sf_blob), float(eps), cu(gamma), cu(beta))
    bias = capi.fold_affine(plan, w, a, b, cu(d["bias"]) if with_bias else None)
    y = plan.forward(cu(d["x"]), bias, relu=True)
    torch.cuda.synchronize()
    assert po.rel_l2(y.cpu().numpy(), ref) < 1e-4
    # the fold is a refresh, not a re-pack: the mask is unchanged
    assert plan.nnz == int(np.count_nonzero(d["w"]))
