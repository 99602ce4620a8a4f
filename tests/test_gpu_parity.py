"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs, against the committed golden fixtures (outputs of the reference's own CPU
kernels), and -- when oracle/_ref/libescort_ref_gpu.so travelled with the snapshot -- against the reference's own
GPU direct-sconv kernels run on this box.  Bars: CSR pack bit-exact; fp32 outputs <= 1e-4 relative L2."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import golden_files, load_golden, geom_from_golden

pytestmark = pytest.mark.gpu
TOL = 1e-4  # BASELINE.json north_star: "within 1e-4 relative L2 in fp32"


def _torch():
    import torch
    return torch


def to_dev(a):
    torch = _torch()
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def capi_geom(capi, g):
    return capi.make_geom(g.Cin, g.Cout, g.H, g.W, g.kh, g.stride_h, g.pad_h, g.dil_h, g.group, kw=g.kw,
                          pad_w=g.pad_w, stride_w=g.stride_w, dilation_w=g.dil_w)


def assert_csr_equal(dev_csr, ocsr):
    for k in ("values", "colidx", "rowptr", "nnz_per_row"):
        a = dev_csr[k].cpu().numpy().view(np.int32)
        b = ocsr[k].view(np.int32)
        assert np.array_equal(a, b), "CSR field %s differs" % k
    assert [int(v) for v in dev_csr["nz_num"]] == [int(v) for v in ocsr["nz_num"]]


# ---------------------------------------------------------------- pack (a1-a3): bit-exact
@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_pack_bit_exact_vs_golden(capi, po, path):
    d = load_golden(path)
    g = geom_from_golden(po, d)
    geom = capi_geom(capi, g)
    for stretch in (True, False):
        csr = capi.weight_align(to_dev(d["w"]), geom, stretch=stretch)
        ocsr = po.weight_align(d["w"], g, stretch=stretch)
        assert_csr_equal(csr, ocsr)
        ref_col = d["colidx"] if stretch else d["colidx_raw"]
        assert np.array_equal(csr["colidx"].cpu().numpy(), ref_col)
        assert np.array_equal(csr["values"].cpu().numpy().view(np.int32), d["values"].view(np.int32))
        assert np.array_equal(csr["rowptr"].cpu().numpy(), d["rowptr"])


@pytest.mark.parametrize("M,N,density", [(1, 1, 1.0), (3, 5, 0.5), (7, 31, 0.3), (33, 32, 0.2), (20, 25, 0.2),
                                          (50, 500, 0.2), (384, 2304, 0.12), (512, 4608, 0.3), (5, 1000, 0.0),
                                          (1500, 33, 0.9), (4, 2400, 1.0)])
def test_pack_edge_cases(capi, po, M, N, density):
    rng = np.random.default_rng(M * 1000 + N)
    A = rng.standard_normal((M, N)).astype(np.float32)
    A[rng.uniform(size=A.shape) >= density] = 0.0
    if A.size > 10:
        A.reshape(-1)[1] = -0.0           # dropped
        A.reshape(-1)[A.size // 2] = np.nan  # kept
        A.reshape(-1)[A.size - 1] = 1e-42    # denormal kept
    if M > 2:
        A[M // 2] = 0.0                   # empty row
    values, colidx, rowptr, npr, nnz = capi.pack_csr(to_dev(A))
    g = po.Geom(1, N, 1, 1, M, 1)
    o = po.weight_align(A.reshape(M, N, 1, 1), g, stretch=False)
    assert nnz == int(o["nz_num"][0])
    assert np.array_equal(rowptr.cpu().numpy(), o["rowptr"][:M + 1])
    assert np.array_equal(npr.cpu().numpy(), o["nnz_per_row"])
    assert np.array_equal(colidx.cpu().numpy()[:nnz], o["colidx"][:nnz])
    assert np.array_equal(values.cpu().numpy()[:nnz].view(np.int32), o["values"][:nnz].view(np.int32))


def test_pack_full_size_layers_bit_exact(capi, po):
    """BASELINE full sizes: AlexNet conv2 (groups), ResNet res5 (2.36 M weights)."""
    from caffe_escoin_b200 import workloads as wl
    for spec in (wl.ALEXNET[0], wl.ALEXNET[1], wl.RESNET50[-1]):
        d = wl.make_layer_data(spec, 5, with_input=False)
        g = po.Geom(1, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
        csr = capi.weight_align(to_dev(d["w"]), capi_geom(capi, g))
        assert_csr_equal(csr, po.weight_align(d["w"], g))


# ---------------------------------------------------------------- forward (a4-a9)
def run_forward_all_paths(capi, po, g, w, bias, x, relu, variants=None):
    """Every forward kernel that accepts the geometry: auto choice (-1), generic (0), each tile-interpreter variant."""
    geom = capi_geom(capi, g)
    csr = capi.weight_align(to_dev(w), geom)
    outs = {}
    if variants is None:
        variants = list(range(-1, 90))
    for v in variants:
        plan = capi.Plan(geom, csr)
        if v != -1:
            try:
                plan.set_variant(v)
            except capi.EscortError:
                continue   # tile variant does not apply to this kernel size / stride
        y = plan.forward(to_dev(x), to_dev(bias), relu=relu)
        _torch().cuda.synchronize()
        outs["%d:%s" % (v, plan.kernel_name)] = y.cpu().numpy()
    return outs, csr


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_forward_vs_golden(capi, po, path):
    d = load_golden(path)
    g = geom_from_golden(po, d)
    for relu, key in ((False, "y_ref_default"), (True, "y_ref_default_relu")):
        outs, _ = run_forward_all_paths(capi, po, g, d["w"], d["bias"], d["x"], relu)
        for name, y in outs.items():
            assert po.rel_l2(y, d[key]) < TOL, (name, key)
    outs, _ = run_forward_all_paths(capi, po, g, d["w"], d["bias"], d["x"], False)
    for name, y in outs.items():
        assert po.rel_l2(y, d["y_ref_blocked"]) < TOL, name


FWD_CASES = [
    # N, Cin, Cout, H, W, kh, kw, stride, pad, dil, group, sparsity, bias, relu
    (4, 20, 50, 12, 12, 5, 5, 1, 0, 1, 1, 0.80, True, False),    # LeNet conv2
    (3, 1, 20, 28, 28, 5, 5, 1, 0, 1, 1, 0.80, True, False),     # LeNet conv1
    (5, 32, 48, 13, 13, 3, 3, 1, 1, 1, 1, 0.88, True, True),     # AlexNet conv3, thin, odd batch
    (2, 16, 32, 27, 27, 5, 5, 1, 2, 1, 2, 0.85, True, False),    # AlexNet conv2, thin, groups
    (2, 48, 32, 13, 13, 3, 3, 1, 1, 1, 2, 0.88, True, True),     # conv4/5 style groups
    (3, 24, 40, 14, 14, 3, 3, 1, 1, 1, 1, 0.75, True, False),    # GoogLeNet 4x
    (2, 16, 16, 7, 7, 3, 3, 1, 1, 1, 1, 0.70, False, False),     # ResNet res5 style, no bias
    (1, 8, 8, 56, 56, 3, 3, 1, 1, 1, 1, 0.70, False, True),      # ResNet res2 style
    (2, 16, 24, 28, 28, 5, 5, 1, 2, 1, 1, 0.75, True, False),    # GoogLeNet 5x5
    (2, 12, 12, 14, 14, 3, 3, 2, 1, 1, 1, 0.50, True, False),    # sweep stride 2
    (2, 12, 12, 15, 15, 3, 3, 2, 1, 1, 1, 0.90, True, True),     # stride 2, odd size
    (1, 6, 10, 11, 9, 3, 3, 1, 2, 2, 1, 0.60, True, False),      # dilation 2, non-square
    (2, 8, 6, 10, 12, 3, 5, 1, 1, 1, 1, 0.60, True, False),      # non-square kernel / image
    (2, 16, 8, 7, 7, 1, 1, 1, 0, 1, 1, 0.50, True, True),        # 1x1
    (2, 8, 8, 9, 9, 3, 3, 1, 0, 1, 1, 0.95, True, False),        # very sparse, empty rows, pad 0
    (1, 4, 4, 5, 5, 3, 3, 1, 1, 1, 1, 1.00, True, True),         # all weights pruned: output = relu(bias)
]


@pytest.mark.parametrize("case", FWD_CASES, ids=lambda c: "N%d_C%d_M%d_H%dx%d_k%dx%d_s%d_p%d_d%d_g%d_sp%g" % c[:12])
def test_forward_vs_oracle(capi, po, case):
    from caffe_escoin_b200 import workloads as wl
    N, Cin, Cout, H, W, kh, kw, s, p, dil, grp, sp, has_bias, relu = case
    rng = np.random.default_rng(hash(case[:12]) % (2 ** 31))
    w = (rng.standard_normal((Cout, Cin // grp, kh, kw)) * 0.01).astype(np.float32)
    w = wl.prune_magnitude(w, sp) if sp < 1.0 else np.zeros_like(w)
    bias = (rng.standard_normal(Cout) * 0.1).astype(np.float32) if has_bias else None
    x = rng.uniform(-1, 1, (N, Cin, H, W)).astype(np.float32)
    g = po.Geom(N, Cin, H, W, Cout, kh, s, p, dil, grp, kw=kw)
    ocsr = po.weight_align(w, g)
    y_or = po.conv_forward(x, ocsr, g, bias, relu=relu)
    outs, csr = run_forward_all_paths(capi, po, g, w, bias, x, relu)
    assert_csr_equal(csr, ocsr)
    y_dense = po.dense_conv(x, w, g, bias, relu=relu)
    for name, y in outs.items():
        assert y.shape == y_or.shape
        assert po.rel_l2(y, y_or) < TOL, name
        assert po.rel_l2(y, y_dense) < TOL, name


@pytest.mark.parametrize("case", FWD_CASES, ids=lambda c: "N%d_C%d_M%d_H%dx%d_k%dx%d_s%d_p%d_d%d_g%d_sp%g" % c[:12])
def test_small_map_kernel_is_bit_identical_to_the_generic_kernel(capi, po, case, monkeypatch):
    """Variant 0 runs sconv_fwd_small where the input image fits in shared memory (LeNet-sized layers) and
    sconv_fwd_generic otherwise (or with ESCORT_NO_SMALL_MAPS, read at plan creation): same records, same accumulation
    order -- the two must agree bit for bit, and both with the oracle."""
    from caffe_escoin_b200 import workloads as wl
    torch = _torch()
    N, Cin, Cout, H, W, kh, kw, s, p, dil, grp, sp, has_bias, relu = case
    rng = np.random.default_rng(hash(case[:12]) % (2 ** 31))
    w = (rng.standard_normal((Cout, Cin // grp, kh, kw)) * 0.01).astype(np.float32)
    w = wl.prune_magnitude(w, sp) if sp < 1.0 else np.zeros_like(w)
    bias = (rng.standard_normal(Cout) * 0.1).astype(np.float32) if has_bias else None
    x = rng.uniform(-1, 1, (N, Cin, H, W)).astype(np.float32)
    g = po.Geom(N, Cin, H, W, Cout, kh, s, p, dil, grp, kw=kw)
    y_or = po.conv_forward(x, po.weight_align(w, g), g, bias, relu=relu)
    geom = capi.make_geom(Cin, Cout, H, W, kh, s, p, dil, grp, kw=kw)
    wd = torch.from_numpy(w).cuda()
    xd = torch.from_numpy(x).cuda()
    bd = torch.from_numpy(bias).cuda() if has_bias else None
    outs = {}
    for no_small in (False, True):
        if no_small:
            monkeypatch.setenv("ESCORT_NO_SMALL_MAPS", "1")
        plan = capi.Plan(geom, capi.weight_align(wd, geom))
        plan.set_variant(0)
        fits = Cin * H * W * 4 <= 96 * 1024   # (the 8 x 56 x 56 case is 2 KB over: generic both times)
        assert plan.kernel_name == ("sconv_fwd_small" if fits and not no_small else "sconv_fwd_generic")
        outs[no_small] = plan.forward(xd, bd, relu=relu).cpu().numpy()
    assert np.array_equal(outs[False], outs[True])
    assert po.rel_l2(outs[False], y_or) < TOL


S2D_CASES = [
    # N, Cin, Cout, H, W, k, pad, group, sparsity, bias, relu   (stride 2)
    (3, 12, 16, 14, 14, 3, 1, 1, 0.5, True, False),
    (2, 16, 24, 15, 13, 3, 1, 2, 0.7, True, True),     # odd, non-square, groups
    (2, 8, 12, 28, 28, 5, 2, 1, 0.8, False, False),    # 5x5: a 3x3 kernel over the parity planes
    (2, 8, 8, 9, 9, 1, 0, 1, 0.4, True, False),        # 1x1 stride 2 (ResNet's down-sampling shape)
    (2, 6, 10, 11, 11, 3, 0, 1, 0.6, True, True),      # no padding
    (1, 128, 32, 56, 56, 3, 1, 1, 0.7, False, False),  # large enough for the parity-plane path to be the default
]


@pytest.mark.parametrize("case", S2D_CASES, ids=lambda c: "N%d_C%d_M%d_H%dx%d_k%d_p%d_g%d_sp%g" % c[:9])
def test_stride2_through_space_to_depth(capi, po, case):
    """A stride-2 layer runs by default as a stride-1 convolution over the parity planes of the padded input (core.cu
    build_s2d_plan): against the oracle, after autotune (which may keep either path), after a refresh, and against the
    kernels that read the strided input directly (explicit variant)."""
    from caffe_escoin_b200 import workloads as wl
    torch = _torch()
    N, Cin, Cout, H, W, k, pad, grp, sp, has_bias, relu = case
    rng = np.random.default_rng(hash(case[:9]) % (2 ** 31))
    w = wl.prune_magnitude((rng.standard_normal((Cout, Cin // grp, k, k)) * 0.05).astype(np.float32), sp)
    bias = (rng.standard_normal(Cout) * 0.1).astype(np.float32) if has_bias else None
    x = rng.uniform(-1, 1, (N, Cin, H, W)).astype(np.float32)
    g = po.Geom(N, Cin, H, W, Cout, k, 2, pad, 1, grp)
    y_or = po.conv_forward(x, po.weight_align(w, g), g, bias, relu=relu)
    geom = capi.make_geom(Cin, Cout, H, W, k, 2, pad, 1, grp)
    wd, xd = torch.from_numpy(w).cuda(), torch.from_numpy(x).cuda()
    bd = torch.from_numpy(bias).cuda() if has_bias else None
    plan = capi.Plan(geom, capi.weight_align(wd, geom))
    # default without a measurement: the parity-plane path only where the sweep says it wins (channels x height >= 128 x 56)
    assert ("(s2d)" in plan.describe()) == ((Cin // grp) * H >= 128 * 56), plan.describe()
    plan.set_config(-2, 0)                   # the space-to-depth path by its variant id
    assert "(s2d)" in plan.describe(), plan.describe()
    y = plan.forward(xd, bd, relu=relu)
    assert po.rel_l2(y.cpu().numpy(), y_or) < TOL, plan.describe()
    plan.set_variant(0)                      # the generic / small-map kernel on the strided input
    assert "(s2d)" not in plan.describe()
    assert po.rel_l2(plan.forward(xd, bd, relu=relu).cpu().numpy(), y_or) < TOL
    plan.autotune(N)                         # measures both paths, keeps the faster
    assert po.rel_l2(plan.forward(xd, bd, relu=relu).cpu().numpy(), y_or) < TOL, plan.describe()
    v, rank = plan.get_config()
    plan.set_config(-2, 0)                   # the space-to-depth path by its variant id (what a tune cache stores)
    assert "(s2d)" in plan.describe()
    w2 = (w * 1.5).astype(np.float32)        # a solver update: same mask, new values
    plan.refresh_values(torch.from_numpy(w2).cuda())
    y2_or = po.conv_forward(x, po.weight_align(w2, g), g, bias, relu=relu)
    assert po.rel_l2(plan.forward(xd, bd, relu=relu).cpu().numpy(), y2_or) < TOL
    y_small = plan.forward(xd[:1].contiguous(), bd, relu=relu)   # another batch size through the same buffer
    assert po.rel_l2(y_small.cpu().numpy(), y2_or[:1]) < TOL
    # backward data of the stride-2 layer: the parity-plane sub-plan's own backward-data plan + depth-to-space, with the
    # refreshed weights (the sub-plan's backward plan is built after the refresh: it must gather the current values)
    dy = rng.uniform(-1, 1, y_or.shape).astype(np.float32)
    _, _, dx_or = po.conv_backward(x, dy, w2, g, mask_only=True, want_b=False)
    dx = plan.backward_data(torch.from_numpy(dy).cuda())
    assert po.rel_l2(dx.cpu().numpy(), dx_or) < TOL


def test_forward_raw_and_stretched_plans_agree(capi, po):
    from caffe_escoin_b200 import workloads as wl
    spec = wl.ALEXNET[3]._replace(N=2, Cin=32, Cout=32)
    d = wl.make_layer_data(spec, 2)
    g = po.Geom(spec.N, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
    geom = capi_geom(capi, g)
    ys = []
    for stretch in (True, False):
        csr = capi.weight_align(to_dev(d["w"]), geom, stretch=stretch)
        ys.append(capi.Plan(geom, csr).forward(to_dev(d["x"]), to_dev(d["bias"])).cpu().numpy())
    assert np.array_equal(ys[0], ys[1])


def test_compat_entry_matches_reference_calling_sequence(capi, po):
    """escort_copy_input + escort_sconv_padded driven exactly like forward_gpu_sconv (base_conv_layer.cpp:749-798)."""
    torch = _torch()
    from caffe_escoin_b200 import workloads as wl
    for spec, dil in ((wl.ALEXNET[0]._replace(N=2, Cin=16, Cout=16), 1), (wl.ALEXNET[1]._replace(N=3, Cin=16, Cout=24), 1),
                      (wl._c("d2", 2, 6, 8, 11, 3, 1, 2, 1, 0.6), 2), (wl._c("s2", 2, 6, 8, 14, 3, 2, 1, 1, 0.6), 1)):
        d = wl.make_layer_data(spec, 4)
        g = po.Geom(spec.N, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, dil, spec.group)
        geom = capi_geom(capi, g)
        csr = capi.weight_align(to_dev(d["w"]), geom)
        ocsr = po.weight_align(d["w"], g)
        for relu in (False, True):
            y_or = po.conv_forward(d["x"], ocsr, g, d["bias"] if relu else None, relu=relu)
            x = to_dev(d["x"])
            bias = to_dev(d["bias"])
            top = torch.zeros((g.num, g.Cout, g.Ho, g.Wo), device="cuda")
            plen = g.Cin * (g.H + g.pad_h) * (g.W + g.pad_w) + g.pad_h * (g.W + 2 * g.pad_w)
            padded = torch.zeros(plen, device="cuda")
            M, Cg = g.Cout // g.group, g.Cin // g.group
            ifmap = Cg * (g.H + g.pad_h) * (g.W + g.pad_w)
            woff = M * Cg * g.kh * g.kw
            for n in range(g.num):
                src = x[n]
                if g.pad_h or g.pad_w:
                    capi.copy_input(padded, x[n].contiguous(), g.Cin, g.H, g.W, g.pad_h, g.pad_w)
                    src = padded
                for gi in range(g.group):
                    capi.sconv_padded(relu, 1, src.data_ptr() + 4 * gi * ifmap, ifmap,
                                      csr["rowptr"].data_ptr() + 4 * (M + 1) * gi, csr["colidx"].data_ptr() + 4 * woff * gi,
                                      csr["values"].data_ptr() + 4 * woff * gi, bias.data_ptr() + 4 * M * gi, g.H, g.W,
                                      g.pad_h, g.pad_w, g.stride_h, g.stride_w, g.dil_h, g.dil_w, g.kh, g.kw,
                                      top[n].data_ptr() + 4 * gi * M * g.Ho * g.Wo, M, g.group)
            torch.cuda.synchronize()
            assert po.rel_l2(top.cpu().numpy(), y_or) < TOL


def test_forward_vs_reference_gpu_kernels(capi, po):
    """Our forward vs the reference's own caffe_gpu_sconv (oracle/_ref/libescort_ref_gpu.so), same CSR, same box."""
    if not po.have_ref_gpu():
        pytest.skip("oracle/_ref/libescort_ref_gpu.so not built (reference tree absent at build time)")
    torch = _torch()
    from caffe_escoin_b200 import workloads as wl
    R = C.CDLL(po.ref_gpu_path())
    for spec in (wl.ALEXNET[1]._replace(N=4, Cin=64, Cout=96), wl.ALEXNET[0]._replace(N=2, Cin=32, Cout=32),
                 wl.GOOGLENET[1]._replace(N=2), wl.RESNET50[-1]._replace(N=2, Cin=64, Cout=64),
                 wl._c("s2", 2, 32, 32, 28, 3, 2, 1, 1, 0.8)):
        d = wl.make_layer_data(spec, 6)
        g = po.Geom(spec.N, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
        geom = capi_geom(capi, g)
        csr = capi.weight_align(to_dev(d["w"]), geom)
        x, bias = to_dev(d["x"]), to_dev(d["bias"])
        y = capi.Plan(geom, csr).forward(x, bias)
        top = torch.zeros_like(y)
        plen = g.Cin * (g.H + g.pad_h) * (g.W + g.pad_w) + g.pad_h * (g.W + 2 * g.pad_w)
        padded = torch.zeros(plen, device="cuda")
        torch.cuda.synchronize()
        p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        R.refgpu_conv_forward(p(x), g.num, g.Cin, g.H, g.W, g.Cout, g.group, g.kh, g.kw, g.pad_h, g.pad_w, g.stride_h,
                              g.stride_w, g.dil_h, g.dil_w, p(csr["values"]), p(csr["colidx"]), p(csr["rowptr"]),
                              p(bias), 0, p(top), p(padded))
        torch.cuda.synchronize()
        assert po.rel_l2(y.cpu().numpy(), top.cpu().numpy()) < TOL, spec.name


def test_forward_full_size_properties(capi, po):
    """BASELINE full size (AlexNet conv3, N=256): size-independent checks -- linearity in the input, batch
    independence (image n of the batch == the same image run alone), and spot parity of 2 images vs the oracle."""
    torch = _torch()
    from caffe_escoin_b200 import workloads as wl
    spec = wl.ALEXNET[1]
    d = wl.make_layer_data(spec, 1)
    g = po.Geom(spec.N, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
    geom = capi_geom(capi, g)
    csr = capi.weight_align(to_dev(d["w"]), geom)
    plan = capi.Plan(geom, csr)
    x = to_dev(d["x"])
    y = plan.forward(x, None)
    x2 = torch.flip(x, dims=[0]).contiguous()
    y2 = plan.forward(x2, None)
    ysum = plan.forward((2.0 * x + x2).contiguous(), None)
    torch.cuda.synchronize()
    assert po.rel_l2(ysum.cpu().numpy(), (2.0 * y + y2).cpu().numpy()) < TOL
    assert torch.equal(torch.flip(y2, dims=[0]), y)          # batch independence / determinism
    y_one = plan.forward(x[7:8].contiguous(), None)
    assert po.rel_l2(y_one.cpu().numpy(), y[7:8].cpu().numpy()) < 1e-6
    g2 = po.Geom(2, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
    y_or = po.conv_forward(d["x"][:2], po.weight_align(d["w"], g2), g2, None)
    assert po.rel_l2(y[:2].cpu().numpy(), y_or) < TOL


# ---------------------------------------------------------------- backward (a9), masked
BWD_CASES = [
    (3, 16, 24, 13, 13, 3, 3, 1, 1, 1, 1, 0.88),
    (2, 16, 16, 27, 27, 5, 5, 1, 2, 1, 2, 0.85),
    (2, 16, 16, 7, 7, 3, 3, 1, 1, 1, 1, 0.70),
    (2, 8, 8, 28, 28, 3, 3, 1, 1, 1, 1, 0.70),
    (2, 12, 12, 14, 14, 3, 3, 2, 1, 1, 1, 0.50),
    (1, 6, 10, 11, 9, 3, 3, 1, 2, 2, 1, 0.60),
    (2, 16, 8, 7, 7, 1, 1, 1, 0, 1, 1, 0.50),
    (2, 4, 6, 12, 12, 5, 5, 1, 0, 1, 1, 0.80),      # LeNet-like: no padding -> backward data pads by k-1
    (5, 32, 32, 56, 56, 3, 3, 1, 1, 1, 1, 0.70),    # odd batch, TMA-staged backward data
]


@pytest.mark.parametrize("generic_bwd", [False, True], ids=["tile", "generic"])
@pytest.mark.parametrize("case", BWD_CASES, ids=lambda c: "N%d_C%d_M%d_H%dx%d_k%dx%d_s%d_p%d_d%d_g%d_sp%g" % c)
def test_backward_vs_oracle(capi, po, case, generic_bwd, monkeypatch):
    torch = _torch()
    if generic_bwd:
        monkeypatch.setenv("ESCORT_GENERIC_BACKWARD", "1")   # the one-thread-per-element kernels (any stride / dilation)
    else:
        monkeypatch.delenv("ESCORT_GENERIC_BACKWARD", raising=False)
    from caffe_escoin_b200 import workloads as wl
    N, Cin, Cout, H, W, kh, kw, s, p, dil, grp, sp = case
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    w = wl.prune_magnitude((rng.standard_normal((Cout, Cin // grp, kh, kw)) * 0.01).astype(np.float32), sp)
    x = rng.uniform(-1, 1, (N, Cin, H, W)).astype(np.float32)
    g = po.Geom(N, Cin, H, W, Cout, kh, s, p, dil, grp, kw=kw)
    dy = rng.uniform(-1, 1, (N, Cout, g.Ho, g.Wo)).astype(np.float32)
    wd0 = rng.standard_normal(w.shape).astype(np.float32)   # param diffs ACCUMULATE (conv_layer.cu:59-62)
    bd0 = rng.standard_normal(Cout).astype(np.float32)
    wd_o, bd_o, dx_o = po.conv_backward(x, dy, w, g, mask_only=True, w_diff=wd0, b_diff=bd0)
    geom = capi_geom(capi, g)
    csr = capi.weight_align(to_dev(w), geom)
    plan = capi.Plan(geom, csr)
    wd = to_dev(wd0.copy())
    bd = to_dev(bd0.copy())
    wd_csr = torch.zeros_like(csr["values"])
    dx = torch.full((N, Cin, H, W), 123.0, device="cuda")   # bottom diff is OVERWRITTEN (backward_gpu_gemm beta=0)
    plan.backward_weight(to_dev(x), to_dev(dy), wd_dense=wd, wd_csr=wd_csr, accumulate=False)
    capi.bias_backward(to_dev(dy), bd)
    plan.backward_data(to_dev(dy), dx)
    torch.cuda.synchronize()
    assert po.rel_l2(wd.cpu().numpy(), wd_o) < TOL
    assert np.array_equal(wd.cpu().numpy()[w == 0], wd0[w == 0])   # untouched outside the mask
    assert po.rel_l2(bd.cpu().numpy(), bd_o) < TOL
    assert po.rel_l2(dx.cpu().numpy(), dx_o) < TOL
    # CSR-ordered gradient == dense gradient gathered at the nonzero positions, in the reference blob layout
    M, Ng = Cout // grp, (Cin // grp) * kh * kw
    wd_pure, _, _ = po.conv_backward(x, dy, w, g, mask_only=True, want_b=False, want_x=False)
    raw = po.weight_align(w, g, stretch=False)
    got = wd_csr.cpu().numpy()
    for gi in range(grp):
        rp = raw["rowptr"][gi * (M + 1):(gi + 1) * (M + 1)]
        cols = raw["colidx"][gi * M * Ng: gi * M * Ng + rp[-1]]
        rows = np.repeat(np.arange(M), np.diff(rp))
        expect = wd_pure.reshape(Cout, Ng)[gi * M + rows, cols]
        assert po.rel_l2(got[gi * M * Ng: gi * M * Ng + rp[-1]], expect) < TOL
    # accumulate=True adds on top
    plan.backward_weight(to_dev(x), to_dev(dy), wd_csr=wd_csr, accumulate=True)
    torch.cuda.synchronize()
    assert po.rel_l2(wd_csr.cpu().numpy(), 2.0 * got) < TOL


def test_autotune_forward_and_backward(capi, po):
    """Plan-time autotune (cuDNN-find idiom) of the forward and of the backward-data sub-plan keeps results correct, and a
    training loop (refresh -> forward -> backward) on the tuned plan follows the updated weights in both directions."""
    torch = _torch()
    from caffe_escoin_b200 import workloads as wl
    spec = wl.RESNET50[3]._replace(N=6, Cin=32, Cout=48)        # 28x28: TMA-staged kernels are candidates
    d = wl.make_layer_data(spec, 5)
    g = po.Geom(spec.N, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
    geom = capi_geom(capi, g)
    w = to_dev(d["w"])
    csr = capi.weight_align(w, geom)
    plan = capi.Plan(geom, csr)
    plan.autotune(spec.N)
    plan.autotune_backward(spec.N)
    v, r = plan.get_config()
    assert v >= 0 and r >= 0
    mask = (w != 0).float()
    for it in range(2):
        w = (w + 0.03 * torch.randn_like(w) * mask).contiguous()
        plan.refresh_values(w, csr["values"])
        y = plan.forward(to_dev(d["x"]), None)
        dy = torch.rand_like(y) * 2 - 1
        dx = plan.backward_data(dy)
        wd = torch.zeros_like(w)
        plan.backward_weight(to_dev(d["x"]), dy, wd_dense=wd)
        torch.cuda.synchronize()
        wn = w.cpu().numpy()
        y_or = po.conv_forward(d["x"], po.weight_align(wn, g), g, None)
        assert po.rel_l2(y.cpu().numpy(), y_or) < TOL
        wd_o, _, dx_o = po.conv_backward(d["x"], dy.cpu().numpy(), wn, g, mask_only=True, want_b=False)
        assert po.rel_l2(dx.cpu().numpy(), dx_o) < TOL
        assert po.rel_l2(wd.cpu().numpy(), wd_o) < TOL


def test_refresh_values_after_update(capi, po):
    """Masked SGD step: update dense weights at mask positions, re-gather CSR values, forward must follow."""
    torch = _torch()
    from caffe_escoin_b200 import workloads as wl
    spec = wl.GOOGLENET[1]._replace(N=2, Cin=32, Cout=32)
    d = wl.make_layer_data(spec, 9)
    g = po.Geom(spec.N, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
    geom = capi_geom(capi, g)
    w = to_dev(d["w"])
    csr = capi.weight_align(w, geom)
    plan = capi.Plan(geom, csr)
    mask = (w != 0).float()
    w2 = (w + 0.05 * torch.randn_like(w) * mask).contiguous()
    plan.refresh_values(w2, csr["values"])
    y = plan.forward(to_dev(d["x"]), to_dev(d["bias"]))
    dx = plan.backward_data(y)
    torch.cuda.synchronize()
    w2n = w2.cpu().numpy()
    o2 = po.weight_align(w2n, g)
    assert np.array_equal(csr["values"].cpu().numpy().view(np.int32), o2["values"].view(np.int32))
    y_or = po.conv_forward(d["x"], o2, g, d["bias"])
    assert po.rel_l2(y.cpu().numpy(), y_or) < TOL
    _, _, dx_o = po.conv_backward(d["x"], y.cpu().numpy(), w2n, g, want_w=False, want_b=False)
    assert po.rel_l2(dx.cpu().numpy(), dx_o) < TOL


def test_streams_and_errors(capi, po):
    torch = _torch()
    from caffe_escoin_b200 import workloads as wl
    spec = wl.ALEXNET[1]._replace(N=2, Cin=16, Cout=16)
    d = wl.make_layer_data(spec, 1)
    g = po.Geom(spec.N, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
    geom = capi_geom(capi, g)
    csr = capi.weight_align(to_dev(d["w"]), geom)
    plan = capi.Plan(geom, csr)
    s = torch.cuda.Stream()
    x, b = to_dev(d["x"]), to_dev(d["bias"])
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        y = plan.forward(x, b, stream=s)
    s.synchronize()
    assert po.rel_l2(y.cpu().numpy(), po.conv_forward(d["x"], po.weight_align(d["w"], g), g, d["bias"])) < TOL
    bad = dict(csr)
    bad["stretched"] = False  # stretched indices declared raw -> out-of-range column must be rejected, not UB
    with pytest.raises(capi.EscortError):
        capi.Plan(geom, bad)
    assert plan.forward(x[:0].contiguous(), b).shape[0] == 0   # empty batch is a no-op
