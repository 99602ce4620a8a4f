"""CPU tests: the oracle against the committed golden vectors (outputs of the reference's own kernels, made by
tools/make_golden.py), against the reference compiled in place when oracle/_ref exists, and against the dense
ground truth (reference src/caffe/test/test_convolution_layer.cpp:20-150 semantics)."""
import os

import numpy as np
import pytest

from conftest import golden_files, load_golden, geom_from_golden

TOL = 1e-4  # north_star: relative L2, fp32


def numpy_dense2csr(A):
    """Independent restatement of the scan (math_functions.cpp:92-105) used to double-check the oracle's pack."""
    M, N = A.shape
    vals, cols, rowptr = [], [], [0]
    for i in range(M):
        nz = np.nonzero(A[i] != 0)[0]
        vals.extend(A[i, nz].tolist())
        cols.extend(nz.tolist())
        rowptr.append(rowptr[-1] + len(nz))
    return np.array(vals, np.float32), np.array(cols, np.int32), np.array(rowptr, np.int32)


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_golden(po, path):
    d = load_golden(path)
    g = geom_from_golden(po, d)
    csr = po.weight_align(d["w"], g, stretch=True)
    # CSR: bit-exact against the fixture and against an independent numpy scan
    for key in ("values", "colidx", "rowptr", "nz_num", "nnz_per_row"):
        assert np.array_equal(csr[key].view(np.int32), d[key].view(np.int32)), key
    raw = po.weight_align(d["w"], g, stretch=False)
    assert np.array_equal(raw["colidx"], d["colidx_raw"])
    M, N = g.M, g.N
    for gi in range(g.group):
        v, c, r = numpy_dense2csr(d["w"].reshape(g.Cout, N)[gi * M:(gi + 1) * M])
        n = int(csr["nz_num"][gi])
        assert n == len(v)
        assert np.array_equal(raw["values"][gi * M * N: gi * M * N + n].view(np.int32), v.view(np.int32))
        assert np.array_equal(raw["colidx"][gi * M * N: gi * M * N + n], c)
        assert np.array_equal(raw["rowptr"][gi * (M + 1):(gi + 1) * (M + 1)], r)
    # forward vs the reference's caffe_cpu_sconv_default / sconv_unit_stride outputs
    y = po.conv_forward(d["x"], csr, g, d["bias"], relu=False, threads=1)
    yr = po.conv_forward(d["x"], csr, g, d["bias"], relu=True, threads=1)
    assert po.rel_l2(y, d["y_ref_default"]) < 1e-6
    assert po.rel_l2(yr, d["y_ref_default_relu"]) < 1e-6
    assert po.rel_l2(y, d["y_ref_blocked"]) < TOL
    # and vs the dense ground truth on the zero-filled weights
    yd = po.dense_conv(d["x"], d["w"], g, d["bias"])
    assert po.rel_l2(y, yd) < TOL


def test_negative_zero_and_nan_semantics(po):
    """`!= 0` keeps NaN, drops -0.0 (math_functions.cpp:96)."""
    A = np.zeros((3, 5), np.float32)
    A[0, 1] = -0.0
    A[0, 3] = np.nan
    A[2, 4] = 1e-38  # denormal-ish value is kept
    g = po.Geom(1, 5, 1, 1, 3, 1)
    csr = po.weight_align(A.reshape(3, 5, 1, 1), g, stretch=False)
    assert list(csr["rowptr"][:4]) == [0, 1, 1, 2]
    assert list(csr["colidx"][:2]) == [3, 4]
    assert np.isnan(csr["values"][0])


@pytest.mark.skipif(not os.path.exists("/root/reference"), reason="reference tree not present on this box")
@pytest.mark.parametrize("case", [
    dict(N=3, Cin=12, Cout=24, H=13, k=3, s=1, p=1, d=1, grp=1, sp=0.88),
    dict(N=2, Cin=8, Cout=12, H=27, k=5, s=1, p=2, d=1, grp=2, sp=0.85),
    dict(N=2, Cin=16, Cout=16, H=28, k=3, s=1, p=1, d=1, grp=1, sp=0.7),
    dict(N=2, Cin=8, Cout=8, H=15, k=3, s=2, p=1, d=1, grp=1, sp=0.6),
    dict(N=1, Cin=4, Cout=8, H=12, k=3, s=1, p=2, d=2, grp=1, sp=0.5),
])
def test_oracle_vs_reference_compiled_in_place(po, case):
    from caffe_escoin_b200 import workloads as wl
    rng = np.random.default_rng(7)
    w = wl.prune_magnitude((rng.standard_normal((case["Cout"], case["Cin"] // case["grp"], case["k"], case["k"]))
                            * 0.01).astype(np.float32), case["sp"])
    b = (rng.standard_normal(case["Cout"]) * 0.1).astype(np.float32)
    x = rng.uniform(-1, 1, (case["N"], case["Cin"], case["H"], case["H"])).astype(np.float32)
    g = po.Geom(case["N"], case["Cin"], case["H"], case["H"], case["Cout"], case["k"], case["s"], case["p"],
                case["d"], case["grp"])
    csr = po.weight_align(w, g)
    y = po.conv_forward(x, csr, g, b)
    for blocked in (False, True):
        yr, _ = po.ref_conv_forward(x, csr, g, b, blocked=blocked)
        assert po.rel_l2(y, yr) < 1e-6 if not blocked else po.rel_l2(y, yr) < TOL
    assert po.rel_l2(y, po.dense_conv(x, w, g, b)) < TOL


def test_backward_oracle_is_adjoint_of_forward(po):
    """<dY, conv(X)> == <dX, X> and == <dW, W> (bilinearity) -- pins the backward restatement to the forward."""
    from caffe_escoin_b200 import workloads as wl
    rng = np.random.default_rng(11)
    for (Cin, Cout, H, k, s, p, d, grp) in [(6, 8, 9, 3, 1, 1, 1, 1), (4, 6, 10, 3, 2, 1, 1, 2), (4, 4, 9, 3, 1, 2, 2, 1)]:
        g = po.Geom(2, Cin, H, H, Cout, k, s, p, d, grp)
        w = wl.prune_magnitude(rng.standard_normal(g.wshape()).astype(np.float32), 0.6)
        x = rng.uniform(-1, 1, (2, Cin, H, H)).astype(np.float32)
        dy = rng.uniform(-1, 1, (2, Cout, g.Ho, g.Wo)).astype(np.float32)
        y = po.dense_conv(x, w, g)
        wd, bd, xd = po.conv_backward(x, dy, w, g, mask_only=True)
        lhs = float(np.sum(dy.astype(np.float64) * y))
        assert abs(lhs - float(np.sum(xd.astype(np.float64) * x))) < 1e-3 * max(1.0, abs(lhs))
        assert abs(lhs - float(np.sum(wd.astype(np.float64) * w))) < 1e-3 * max(1.0, abs(lhs))
        assert np.all(wd[w == 0] == 0)  # gradient restricted to the mask
        assert np.allclose(bd, dy.sum(axis=(0, 2, 3)), rtol=1e-5, atol=1e-5)
        wd_full, _, _ = po.conv_backward(x, dy, w, g, mask_only=False)
        assert np.allclose(wd_full[w != 0], wd[w != 0])
