"""Host logic of the stride-2 path (caffe_escoin_b200/csrc/core.cu build_s2d_plan / s2d_pad_kernel / d2s_unpad_kernel), checked on
the CPU with the oracle alone: a stride-2 K x K convolution with padding equals a VALID stride-1 convolution with a
ceil(K / 2) kernel over the four parity planes of the padded input, with the nonzeros mapped one to one
(ic, kh, kw) -> (4 ic + 2 (kh & 1) + (kw & 1), kh >> 1, kw >> 1).  The GPU tests
(tests/test_gpu_parity.py::test_stride2_through_space_to_depth) check the kernels that implement it."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def s2d_pad(x, pad, H2, W2):
    """x [N, C, H, W] -> parity planes [N, 4C, H2, W2] of the zero-padded input (s2d_pad_kernel)."""
    N, C, H, W = x.shape
    P = np.zeros((N, C, 2 * H2 + 2, 2 * W2 + 2), np.float32)
    P[:, :, pad:pad + H, pad:pad + W] = x
    out = np.zeros((N, 4 * C, H2, W2), np.float32)
    for py in range(2):
        for px in range(2):
            out[:, 2 * py + px::4] = P[:, :, py:py + 2 * H2:2, px:px + 2 * W2:2]
    return out


def map_weights(w, group):
    """w [M, C/g, K, K] -> [M, 4C/g, K2, K2] by the nonzero map of build_s2d_plan."""
    M, Cg, K, _ = w.shape
    K2 = (K - 1) // 2 + 1
    w2 = np.zeros((M, 4 * Cg, K2, K2), np.float32)
    for kh in range(K):
        for kw in range(K):
            w2[:, 2 * (kh & 1) + (kw & 1)::4, kh >> 1, kw >> 1] = w[:, :, kh, kw]
    return w2


@pytest.mark.parametrize("C,M,H,W,K,pad,group", [(6, 8, 14, 14, 3, 1, 1), (8, 12, 15, 13, 3, 1, 2), (4, 6, 28, 28, 5, 2, 1),
                                                (4, 4, 9, 9, 1, 0, 1), (3, 5, 11, 11, 3, 0, 1), (4, 8, 12, 10, 7, 3, 2)])
def test_stride2_equals_valid_stride1_over_parity_planes(C, M, H, W, K, pad, group):
    from oracle import pyoracle as po
    po.build()
    rng = np.random.default_rng(C * 100 + K)
    N = 2
    x = rng.uniform(-1, 1, (N, C, H, W)).astype(np.float32)
    w = rng.standard_normal((M, C // group, K, K)).astype(np.float32)
    w[rng.uniform(size=w.shape) < 0.6] = 0
    bias = rng.standard_normal(M).astype(np.float32)
    g = po.Geom(N, C, H, W, M, K, 2, pad, 1, group)
    ref = po.dense_conv(x, w, g, bias, relu=True)
    Ho, Wo = ref.shape[2:]
    K2 = (K - 1) // 2 + 1
    H2, W2 = Ho + K2 - 1, Wo + K2 - 1
    x2, w2 = s2d_pad(x, pad, H2, W2), map_weights(w, group)
    assert np.count_nonzero(w2) == np.count_nonzero(w)            # one to one
    g2 = po.Geom(N, 4 * C, H2, W2, M, K2, 1, 0, 1, group)
    got = po.dense_conv(x2, w2, g2, bias, relu=True)
    assert got.shape == ref.shape                                  # the sub-plan writes `top` directly
    assert po.rel_l2(got, ref) < 1e-6
    # and through the sparse CSR walk the kernels restate
    got_sparse = po.conv_forward(x2, po.weight_align(w2, g2), g2, bias, relu=True)
    assert po.rel_l2(got_sparse, ref) < 1e-6
    # backward data: gradient of the parity planes (masked backward of the stride-1 sub-problem), then depth-to-space
    dy = rng.uniform(-1, 1, ref.shape).astype(np.float32)
    _, _, dx_ref = po.conv_backward(x, dy, w, g, mask_only=True, want_w=False, want_b=False)
    _, _, dx2 = po.conv_backward(x2, dy, w2, g2, mask_only=True, want_w=False, want_b=False)
    dx = np.zeros_like(x)
    for y in range(H):
        for xx in range(W):
            yp, xp = y + pad, xx + pad
            if (yp >> 1) < H2 and (xp >> 1) < W2:
                dx[:, :, y, xx] = dx2[:, 2 * (yp & 1) + (xp & 1)::4, yp >> 1, xp >> 1]
    assert po.rel_l2(dx, dx_ref) < 1e-6
