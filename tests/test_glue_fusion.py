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


def _bn_params(rng, M):
    sf_blob = np.float32(2.0)
    mean = (rng.standard_normal(M) * 0.3).astype(np.float32) * sf_blob
    var = rng.uniform(0.5, 2.0, M).astype(np.float32) * sf_blob
    gamma = rng.uniform(0.5, 1.5, M).astype(np.float32)
    beta = (rng.standard_normal(M) * 0.2).astype(np.float32)
    return mean, var, sf_blob, gamma, beta


@pytest.mark.gpu
@pytest.mark.parametrize("H", [12, 7])   # 12: 1x1 layers run the implicit GEMM; 7: the column-buffer path (49 pixels)
def test_resnet_bottleneck_block_in_three_launches(H):
    """A ResNet-50 bottleneck (models/resnet/test_sconv.prototxt: branch2a 1x1 -> BN -> Scale -> ReLU -> branch2b pruned 3x3 ->
    BN -> Scale -> ReLU -> branch2c 1x1 -> BN -> Scale -> Eltwise SUM with the block input -> ReLU): 13 layers, 3 launches
    (dense tcgen05, sparse direct, dense tcgen05 with the residual in its epilogue) against torch fp64 layer by layer."""
    import torch
    from caffe_escoin_b200 import capi, workloads as wl
    N, Cio, Cmid, eps = 3, 96, 32, 1e-5
    rng = np.random.default_rng(H)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    x = cu(rng.uniform(-1, 1, (N, Cio, H, H)).astype(np.float32))
    w_a = cu((rng.standard_normal((Cmid, Cio, 1, 1)) / np.sqrt(Cio)).astype(np.float32))
    w_b_np = wl.prune_magnitude((rng.standard_normal((Cmid, Cmid, 3, 3)) / np.sqrt(Cmid * 9)).astype(np.float32), 0.7)
    w_b = cu(w_b_np)
    w_c = cu((rng.standard_normal((Cio, Cmid, 1, 1)) / np.sqrt(Cmid)).astype(np.float32))
    bns = [_bn_params(rng, M) for M in (Cmid, Cmid, Cio)]

    def bn_ref(t, prm):
        mean, var, sf, gamma, beta = (torch.from_numpy(np.asarray(v, dtype=np.float64)).cuda() for v in prm)
        m, v = mean / sf, var / sf
        return (t - m[None, :, None, None]) / torch.sqrt(v + eps)[None, :, None, None] * gamma[None, :, None, None] + beta[None, :, None, None]

    F = torch.nn.functional
    r = torch.relu(bn_ref(F.conv2d(x.double(), w_a.double()), bns[0]))
    r = torch.relu(bn_ref(F.conv2d(r, w_b.double(), padding=1), bns[1]))
    ref = torch.relu(bn_ref(F.conv2d(r, w_c.double()), bns[2]) + x.double())

    affine = [capi.bn_scale_to_affine(cu(p[0]), cu(p[1]), float(p[2]), eps, cu(p[3]), cu(p[4])) for p in bns]
    g_a = capi.make_geom(Cio, Cmid, H, H, 1, 1, 0, 1, 1)
    g_b = capi.make_geom(Cmid, Cmid, H, H, 3, 1, 1, 1, 1)
    g_c = capi.make_geom(Cmid, Cio, H, H, 1, 1, 0, 1, 1)
    wa_f, ba_f = capi.dense_fold_affine(w_a, *affine[0])
    wc_f, bc_f = capi.dense_fold_affine(w_c, *affine[2])
    plan = capi.Plan(g_b, capi.weight_align(w_b, g_b))
    bb_f = capi.fold_affine(plan, w_b, *affine[1])
    t1 = capi.dense_conv_forward(g_a, x, wa_f, ba_f, relu=True)
    t2 = plan.forward(t1, bb_f, relu=True)
    y = capi.dense_conv_forward(g_c, t2, wc_f, bc_f, relu=True, residual=x)
    torch.cuda.synchronize()
    err = float((y.double() - ref).norm() / ref.norm())
    assert err < 2e-3, err   # TF32 in the two dense layers
    # in place on the residual (Caffe's Eltwise is often in place on the shortcut blob)
    xin = x.clone()
    y2 = capi.dense_conv_forward(g_c, t2, wc_f, bc_f, relu=True, residual=xin, top=xin)
    torch.cuda.synchronize()
    assert torch.equal(y2, y)
