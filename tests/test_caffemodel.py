"""The weight on-disk path: `.caffemodel` bytes -> blobs through the C ABI, the way
Net::CopyTrainedLayersFrom / Blob::FromProto read them (src/caffe/net.cpp:785-821, src/caffe/blob.cpp:466-520), the
writer round trip, and magnitude pruning.  Host code only; the GPU leg (load -> WeightAlign -> forward) is marked gpu."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import protowire as pw  # noqa: E402


def _capi():
    from caffe_escoin_b200 import capi
    return capi


def _model_bytes(rng):
    w1 = rng.standard_normal((8, 3, 3, 3)).astype(np.float32)
    b1 = rng.standard_normal(8).astype(np.float32)
    w2 = rng.standard_normal((6, 4, 5, 5)).astype(np.float32)        # group 2: 8 / 2 input channels
    w3 = rng.standard_normal((10, 24)).astype(np.float64)            # inner product stored as double_data
    b3 = rng.standard_normal(10).astype(np.float32)
    w4 = rng.standard_normal((4, 6, 1, 1)).astype(np.float32)
    layers = [
        pw.layer_v2("data", "Data", []),
        pw.layer_v2("conv1", "Convolution", [pw.blob_shape(w1.ravel(), w1.shape), pw.blob_shape(b1, b1.shape)],
                    conv=pw.conv_param(8, 3, stride=1, pad=1)),
        pw.layer_v2("conv2", "Convolution", [pw.blob_legacy(w2.ravel(), w2.shape)],
                    conv=pw.conv_param(6, 5, group=2, bias_term=False,
                                       hw=dict(kernel_h=5, kernel_w=5, pad_h=2, pad_w=2, stride_h=2, stride_w=2), dilation=1)),
        pw.layer_v2("fc", "InnerProduct", [pw.blob_double(w3.ravel(), w3.shape), pw.blob_legacy(b3, (1, 1, 1, 10), unpacked=True)],
                    ip=pw.f_uint(1, 10)),
        pw.layer_v1("old_conv", 4, [pw.blob_legacy(w4.ravel(), w4.shape)], conv=pw.conv_param(4, 1)),
    ]
    raw = pw.net("tiny", layers, extra=pw.f_bytes(3, "data") + pw.f_uint(5, 1))   # input = 3, force_backward = 5: untouched fields
    return raw, dict(w1=w1, b1=b1, w2=w2, w3=w3, b3=b3, w4=w4)


def test_reads_what_blob_from_proto_reads(tmp_path):
    capi = _capi()
    raw, ref = _model_bytes(np.random.default_rng(0))
    path = str(tmp_path / "tiny.caffemodel")
    open(path, "wb").write(raw)
    m = capi.CaffeModel(path)
    assert len(m) == 5
    assert [m.layer(i)["name"] for i in range(5)] == ["data", "conv1", "conv2", "fc", "old_conv"]
    assert m.find("conv2") == 2 and m.find("nope") < 0
    c1 = m.layer(1)
    assert (c1["type"], c1["is_conv"], c1["num_blobs"], c1["num_output"], c1["kernel_h"], c1["kernel_w"], c1["pad_h"], c1["stride_w"],
            c1["group"], c1["bias_term"]) == ("Convolution", 1, 2, 8, 3, 3, 1, 1, 1, 1)
    c2 = m.layer(2)
    assert (c2["kernel_h"], c2["pad_w"], c2["stride_h"], c2["group"], c2["bias_term"], c2["num_blobs"]) == (5, 2, 2, 2, 0, 1)
    assert np.array_equal(m.blob(1, 0), ref["w1"]) and m.blob(1, 0).shape == (8, 3, 3, 3)       # shape.dim
    assert np.array_equal(m.blob(1, 1), ref["b1"])
    assert np.array_equal(m.blob(2, 0), ref["w2"]) and m.blob(2, 0).shape == (6, 4, 5, 5)       # legacy num/channels/height/width
    fc = m.layer(3)
    assert fc["is_inner_product"] == 1 and fc["num_output"] == 10 and fc["bias_term"] == 1
    assert np.array_equal(m.blob(3, 0), ref["w3"].astype(np.float32))                          # double_data narrowed (blob.cpp:490-494)
    assert np.array_equal(m.blob(3, 1).ravel(), ref["b3"]) and m.blob(3, 1).shape == (1, 1, 1, 10)   # unpacked repeated float
    v1 = m.layer(4)
    assert v1["type"] == "V1:4" and v1["is_conv"] == 1 and v1["kernel_h"] == 1
    assert np.array_equal(m.blob(4, 0), ref["w4"])
    m.close()


def test_rejects_truncated_and_missing_files(tmp_path):
    capi = _capi()
    raw, _ = _model_bytes(np.random.default_rng(1))
    path = str(tmp_path / "cut.caffemodel")
    open(path, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(capi.EscortError):
        capi.CaffeModel(path)
    with pytest.raises(capi.EscortError):
        capi.CaffeModel(str(tmp_path / "absent.caffemodel"))
    # a blob whose data count disagrees with its shape is the CHECK_EQ of blob.cpp:496
    bad = pw.net("bad", [pw.layer_v2("c", "Convolution", [pw.blob_shape([1.0, 2.0, 3.0], (2, 2))])])
    open(path, "wb").write(bad)
    with pytest.raises(capi.EscortError):
        capi.CaffeModel(path)


def test_prune_then_save_round_trip(tmp_path):
    capi = _capi()
    raw, ref = _model_bytes(np.random.default_rng(2))
    path, out = str(tmp_path / "a.caffemodel"), str(tmp_path / "b.caffemodel")
    open(path, "wb").write(raw)
    m = capi.CaffeModel(path)
    w = m.blob(1, 0)
    thr, nnz = capi.prune_magnitude(w.reshape(-1), 0.75)
    k = int(np.floor(0.75 * w.size))
    assert nnz == w.size - k == np.count_nonzero(w)
    assert np.all(np.abs(w[w != 0]) >= thr) and np.all(np.abs(ref["w1"][w == 0]) <= thr)
    assert np.array_equal(w[w != 0], ref["w1"][w != 0])              # survivors untouched
    m.save(out)
    m2 = capi.CaffeModel(out)
    assert len(m2) == len(m)
    for i in range(len(m)):
        a, b = m.layer(i), m2.layer(i)
        assert a == b                                                # names, types, conv parameters survive the writer
        for j in range(a["num_blobs"]):
            assert np.array_equal(m.blob(i, j), m2.blob(i, j)) and m.blob(i, j).shape == m2.blob(i, j).shape
    assert np.count_nonzero(m2.blob(1, 0)) == nnz
    # the untouched NetParameter fields (input, force_backward) and the layers' bottoms are still in the bytes
    saved = open(out, "rb").read()
    assert pw.f_bytes(3, "data") in saved and pw.f_bytes(3, "bottom_of_conv1") in saved


def test_prune_edge_cases():
    capi = _capi()
    w = np.array([0.5, -0.5, 0.5, 0.25, 0.0, -1.0], dtype=np.float32)
    a = w.copy()
    assert capi.prune_magnitude(a, 0.0)[1] == 5 and np.array_equal(a, w)      # nothing to remove
    a = w.copy()
    thr, nnz = capi.prune_magnitude(a, 0.5)                                   # 3 go: 0.0, 0.25 and ONE of the tied 0.5s
    assert nnz == 3 and thr == 0.5 and a[5] == -1.0 and np.count_nonzero(np.abs(a) == 0.5) == 2
    a = w.copy()
    assert capi.prune_magnitude(a, 1.0)[1] == 0 and not a.any()
    e = np.zeros(0, dtype=np.float32)
    assert capi.prune_magnitude(e, 0.9) == (0.0, 0)


@pytest.mark.gpu
def test_caffemodel_to_weight_align_to_forward(tmp_path):
    """load -> prune -> WeightAlign (C ABI) -> forward == the oracle on the same pruned weights"""
    import torch
    capi = _capi()
    from oracle import pyoracle as po
    rng = np.random.default_rng(3)
    w = rng.standard_normal((48, 32, 3, 3)).astype(np.float32)
    b = rng.standard_normal(48).astype(np.float32)
    raw = pw.net("one", [pw.layer_v2("conv3", "Convolution", [pw.blob_shape(w.ravel(), w.shape), pw.blob_shape(b, b.shape)],
                                     conv=pw.conv_param(48, 3, pad=1))])
    path = str(tmp_path / "one.caffemodel")
    open(path, "wb").write(raw)
    m = capi.CaffeModel(path)
    i = m.find("conv3")
    info = m.layer(i)
    wd = m.blob(i, 0)
    capi.prune_magnitude(wd.reshape(-1), 0.88)
    x = rng.uniform(-1, 1, (3, 32, 13, 13)).astype(np.float32)
    g = po.Geom(3, 32, 13, 13, info["num_output"], info["kernel_h"], info["stride_h"], info["pad_h"], 1, info["group"])
    ocsr = po.weight_align(wd, g)
    y_ref = po.conv_forward(x, ocsr, g, m.blob(i, 1), relu=True)
    geom = capi.make_geom(32, info["num_output"], 13, 13, info["kernel_h"], info["stride_h"], info["pad_h"], 1, info["group"])
    csr = capi.weight_align(torch.from_numpy(wd.copy()).cuda(), geom)
    plan = capi.Plan(geom, csr)
    y = plan.forward(torch.from_numpy(x).cuda(), torch.from_numpy(m.blob(i, 1).copy()).cuda(), relu=True)
    torch.cuda.synchronize()
    assert plan.nnz == np.count_nonzero(wd)
    assert po.rel_l2(y.cpu().numpy(), y_ref) < 1e-4
