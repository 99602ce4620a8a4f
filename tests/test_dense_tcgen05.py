"""f1: the layers the reference keeps dense, on tcgen05 (TF32 products, fp32 accumulate).  Inner product against
numpy fp32 / fp64 (src/caffe/layers/inner_product_layer.cu:9-31), dense convolution against the oracle's dense caffe_conv
(src/caffe/test/test_convolution_layer.cpp:20-150).  Tolerance 2e-3 relative L2: TF32 keeps 10 mantissa bits (the
sparse path's 1e-4 bar is for the fp32 FMA sparse path)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TOL = 2e-3

IP_CASES = [(1, 32, 8), (64, 256, 96), (256, 1024, 1000), (37, 132, 130), (130, 36, 257), (256, 9216, 512)]


@pytest.mark.gpu
@pytest.mark.parametrize("num,K,M", IP_CASES)
@pytest.mark.parametrize("relu", [False, True])
def test_inner_product(num, K, M, relu):
    import torch
    from caffe_escoin_b200 import capi
    rng = np.random.default_rng(num * 7 + K)
    x = rng.uniform(-1, 1, (num, K)).astype(np.float32)
    w = (rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(M).astype(np.float32)
    ref = x.astype(np.float64) @ w.astype(np.float64).T + b
    if relu:
        ref = np.maximum(ref, 0)
    y = capi.inner_product_forward(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda(), relu=relu)
    torch.cuda.synchronize()
    err = np.linalg.norm(y.cpu().numpy() - ref) / np.linalg.norm(ref)
    assert err < TOL, err
    y2 = capi.inner_product_forward(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), None, relu=False)   # no bias
    ref2 = x.astype(np.float64) @ w.astype(np.float64).T
    assert np.linalg.norm(y2.cpu().numpy() - ref2) / np.linalg.norm(ref2) < TOL


def test_inner_product_rejects_unaligned_k():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a device pointer")
    from caffe_escoin_b200 import capi
    with pytest.raises(capi.EscortError):
        capi.inner_product_forward(torch.zeros((4, 30), device="cuda"), torch.zeros((8, 30), device="cuda"))


CONV_CASES = [  # name, N, Cin, Cout, H, k, stride, pad
    ("alexnet_conv1_thin", 3, 3, 24, 67, 11, 4, 0),
    ("googlenet_conv1_thin", 2, 3, 16, 56, 7, 2, 3),
    ("pointwise_reduce", 5, 96, 40, 14, 1, 1, 0),
    ("pointwise_wide", 2, 64, 256, 28, 1, 1, 0),
    ("dense_3x3", 4, 20, 130, 13, 3, 1, 1),
    # 1x1 / stride 1: the implicit GEMM straight from NCHW (MN-major A operand); ragged K, N and pixel tails
    ("pointwise_k36_m300", 3, 36, 300, 14, 1, 1, 0),
    ("pointwise_small_map", 2, 128, 64, 8, 1, 1, 0),
    ("pointwise_deep_k", 2, 544, 96, 12, 1, 1, 0),
    # 1x1 shapes the implicit GEMM does not take (odd pixel count, stride 2): column-buffer path
    ("pointwise_7px", 2, 64, 48, 7, 1, 1, 0),
    ("pointwise_stride2", 2, 32, 48, 14, 1, 2, 0),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_dense_conv(case):
    import torch
    from caffe_escoin_b200 import capi
    from oracle import pyoracle as po
    name, N, Cin, Cout, H, k, stride, pad = case
    rng = np.random.default_rng(len(name))
    x = rng.uniform(-1, 1, (N, Cin, H, H)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    g = po.Geom(N, Cin, H, H, Cout, k, stride, pad, 1, 1)
    ref = po.dense_conv(x, w, g, b, relu=True)
    geom = capi.make_geom(Cin, Cout, H, H, k, stride, pad, 1, 1)
    y = capi.dense_conv_forward(geom, torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda(), relu=True)
    torch.cuda.synchronize()
    assert y.shape == ref.shape
    assert po.rel_l2(y.cpu().numpy(), ref) < TOL


@pytest.mark.gpu
def test_pointwise_implicit_gemm_matches_column_buffer_path(monkeypatch):
    """The same layer through both dense-conv paths (ESCORT_DENSE_NO_IMPLICIT forces the column buffer), without bias / ReLU."""
    import torch
    from caffe_escoin_b200 import capi
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.uniform(-1, 1, (3, 72, 10, 10)).astype(np.float32)).cuda()
    w = torch.from_numpy((rng.standard_normal((136, 72, 1, 1)) / np.sqrt(72)).astype(np.float32)).cuda()
    geom = capi.make_geom(72, 136, 10, 10, 1, 1, 0, 1, 1)
    y_imp = capi.dense_conv_forward(geom, x, w, None, relu=False)
    monkeypatch.setenv("ESCORT_DENSE_NO_IMPLICIT", "1")
    y_col = capi.dense_conv_forward(geom, x, w, None, relu=False)
    torch.cuda.synchronize()
    ref = torch.einsum("nchw,mc->nmhw", x.double(), w.double()[:, :, 0, 0])
    assert float((y_imp.double() - ref).norm() / ref.norm()) < TOL
    assert float((y_col.double() - ref).norm() / ref.norm()) < TOL
    assert float((y_imp - y_col).norm() / y_col.norm()) < TOL
