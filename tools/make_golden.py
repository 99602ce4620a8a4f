#!/usr/bin/env python
"""Generate tests/golden/*.npz by running THE REFERENCE'S OWN CODE compiled in place (oracle/_ref, built by
oracle/Makefile from /root/reference/include/caffe/util/sconv.hpp) on seeded inputs.

The reference has no golden vectors / tests for its sparse path, so these fixtures --
outputs of the reference's caffe_cpu_sconv_default<> and sconv_unit_stride<> kernels -- are what pins the
oracle (tests/test_oracle.py) and, through it, the CUDA path.  Run in the build container only
(/root/reference must exist); the .npz files are committed.  CSR inputs stored in the fixtures are produced
by the oracle's restatement of caffe_cpu_sparse_dense2csr (math_functions.cpp:92-105) because the file that
holds the original (boost/glog includes) cannot be compiled here; the fixtures also carry the dense weights
so the pack itself is re-checked against an independent numpy scan in the tests.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from caffe_escoin_b200 import workloads as wl  # noqa: E402

CASES = [
    # name, N, Cin, Cout, H, k, stride, pad, dilation, group, sparsity
    ("lenet_like_k5", 2, 4, 6, 12, 5, 1, 0, 1, 1, 0.80),
    ("alexnet_conv3_like_13x13_k3", 2, 8, 16, 13, 3, 1, 1, 1, 1, 0.88),
    ("alexnet_conv2_like_27x27_k5_g2", 1, 8, 8, 27, 5, 1, 2, 1, 2, 0.85),
    ("googlenet_like_14x14_k3", 2, 6, 20, 14, 3, 1, 1, 1, 1, 0.75),
    ("resnet_like_7x7_k3", 3, 16, 16, 7, 3, 1, 1, 1, 1, 0.70),
    ("stride2_14x14_k3", 2, 6, 10, 14, 3, 2, 1, 1, 1, 0.70),
    ("dilation2_10x10_k3", 1, 4, 6, 10, 3, 1, 2, 2, 1, 0.60),
    ("pointwise_7x7_k1", 2, 16, 8, 7, 1, 1, 0, 1, 1, 0.50),
]


def main():
    assert po.have_ref() or os.path.exists("/root/reference"), "needs the reference to generate fixtures"
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for idx, (name, N, Cin, Cout, H, k, s, p, d, grp, sp) in enumerate(CASES):
        rng = np.random.default_rng(1701 + idx)
        w = (rng.standard_normal((Cout, Cin // grp, k, k)) * 0.01).astype(np.float32)
        w = wl.prune_magnitude(w, sp)
        if idx == 1:  # exercise the != 0 test: a negative zero and an empty row
            w.reshape(-1)[3] = np.float32(-0.0)
            w[5] = 0.0
        bias = (rng.standard_normal(Cout) * 0.1).astype(np.float32)
        x = rng.uniform(-1, 1, (N, Cin, H, H)).astype(np.float32)
        g = po.Geom(N, Cin, H, H, Cout, k, s, p, d, grp)
        csr = po.weight_align(w, g, stretch=True)
        csr_raw = po.weight_align(w, g, stretch=False)
        y_def, _ = po.ref_conv_forward(x, csr, g, bias, relu=False, threads=1, blocked=False)
        y_def_relu, _ = po.ref_conv_forward(x, csr, g, bias, relu=True, threads=1, blocked=False)
        y_blk, used = po.ref_conv_forward(x, csr, g, bias, relu=False, threads=1, blocked=True)
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"), geom=np.array([N, Cin, Cout, H, k, s, p, d, grp], np.int32),
            w=w, bias=bias, x=x, values=csr["values"], colidx=csr["colidx"], colidx_raw=csr_raw["colidx"],
            rowptr=csr["rowptr"], nz_num=csr["nz_num"], nnz_per_row=csr["nnz_per_row"],
            y_ref_default=y_def, y_ref_default_relu=y_def_relu, y_ref_blocked=y_blk,
            blocked_used=np.array([int(used)], np.int32))
        print("%-36s nnz=%s blocked=%s default-vs-blocked rel_l2=%.2e" % (name, list(csr["nz_num"]), used,
                                                                          po.rel_l2(y_blk, y_def)))


if __name__ == "__main__":
    main()
