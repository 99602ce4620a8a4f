#!/usr/bin/env python
"""f1 timing: the tcgen05 (TF32) inner-product and dense-convolution entries on the dense layers of the BASELINE networks,
with cuBLAS / cuDNN through PyTorch beside them for context (fp32 and TF32).  python tools/run_dense.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caffe_escoin_b200 import capi  # noqa: E402


def best_ms(fn, n=5):
    fn()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def lib_ms(fn, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    return best_ms(fn)


N = 256
print("inner product (InnerProductLayer::Forward_gpu, batch %d)" % N)
for name, K, M in [("alexnet/fc6", 9216, 4096), ("alexnet/fc7", 4096, 4096), ("alexnet/fc8", 4096, 1000), ("googlenet/loss3", 1024, 1000),
                   ("resnet50/fc1000", 2048, 1000)]:
    x = torch.rand(N, K, device="cuda") - 0.5
    w = (torch.rand(M, K, device="cuda") - 0.5) / K ** 0.5
    b = torch.rand(M, device="cuda")
    fl = 2.0 * N * K * M
    t = best_ms(lambda: capi.inner_product_forward(x, w, b, relu=True))
    t32 = lib_ms(lambda: torch.relu(torch.addmm(b, x, w.t())), False)
    ttf = lib_ms(lambda: torch.relu(torch.addmm(b, x, w.t())), True)
    ref = torch.relu(torch.addmm(b.double(), x.double(), w.double().t()))
    err = float((capi.inner_product_forward(x, w, b, relu=True).double() - ref).norm() / ref.norm())
    print("  %-20s K %5d M %5d | tcgen05 tf32 %7.3f ms %7.1f TFLOP/s rel_l2 %.1e | cuBLAS fp32 %7.3f ms %6.1f | cuBLAS tf32 %7.3f ms %6.1f"
          % (name, K, M, t, fl / t / 1e9, err, t32, fl / t32 / 1e9, ttf, fl / ttf / 1e9), flush=True)

print("dense convolution (EscConvolutionLayer::Forward_gpu, batch %d; ours = implicit GEMM from NCHW for 1x1 / stride 1, else "
      "transposed im2col + GEMM, both kernels timed; GB/s = (bottom + top bytes) / time)" % N)
for name, Cin, Cout, H, k, s, p in [("alexnet/conv1", 3, 96, 227, 11, 4, 0), ("googlenet/conv1", 3, 64, 224, 7, 2, 3),
                                   ("googlenet/conv2_reduce", 64, 64, 56, 1, 1, 0), ("resnet50/res2a_branch2a", 64, 64, 56, 1, 1, 0),
                                   ("resnet50/res2a_branch2c", 64, 256, 56, 1, 1, 0), ("resnet50/res2b_branch2a", 256, 64, 56, 1, 1, 0),
                                   ("resnet50/res3a_branch2c", 128, 512, 28, 1, 1, 0), ("resnet50/res4a_branch2a", 512, 256, 14, 1, 1, 0),
                                   ("resnet50/res4a_branch2c", 256, 1024, 14, 1, 1, 0), ("resnet50/res4b_branch2a", 1024, 256, 14, 1, 1, 0),
                                   ("resnet50/res5a_branch2c", 512, 2048, 7, 1, 1, 0)]:
    x = torch.rand(N, Cin, H, H, device="cuda") - 0.5
    w = (torch.rand(Cout, Cin, k, k, device="cuda") - 0.5) / (Cin * k * k) ** 0.5
    b = torch.rand(Cout, device="cuda")
    geom = capi.make_geom(Cin, Cout, H, H, k, s, p, 1, 1)
    Ho = (H + 2 * p - k) // s + 1
    fl = 2.0 * N * Cout * Ho * Ho * Cin * k * k
    t = best_ms(lambda: capi.dense_conv_forward(geom, x, w, b, relu=True), n=3)
    conv = lambda: torch.relu(torch.nn.functional.conv2d(x, w, b, stride=s, padding=p))
    t32 = lib_ms(conv, False)
    ttf = lib_ms(conv, True)
    torch.backends.cudnn.allow_tf32 = False
    ref = conv()
    err = float((capi.dense_conv_forward(geom, x, w, b, relu=True) - ref).norm() / ref.norm())
    gbs = 4.0 * N * (Cin * H * H + Cout * Ho * Ho) / t / 1e6
    print("  %-26s | tcgen05 tf32 %7.3f ms %7.1f TFLOP/s %5.0f GB/s rel_l2 %.1e | cuDNN fp32 %7.3f ms %6.1f | cuDNN tf32 %7.3f ms %6.1f"
          % (name, t, fl / t / 1e9, gbs, err, t32, fl / t32 / 1e9, ttf, fl / ttf / 1e9), flush=True)
