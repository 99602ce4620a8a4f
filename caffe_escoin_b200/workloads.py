"""Layer tables and synthetic data for BASELINE.json's configs (shapes from the reference's prototxts:
models/lenet5/train_test.prototxt, models/bvlc_reference_caffenet/test_sconv.prototxt,
models/bvlc_googlenet/test_sconv.prototxt, models/resnet/test_sconv.prototxt).

Synthetic data recipe: weights N(0, 0.01^2) (Caffe `gaussian` filler std), global magnitude
pruning per layer to the target sparsity (ties -> zero, zeros written +0.0f), bias N(0, 0.1^2), inputs
U(-1, 1), fp32, seed = 1701 + layer_index.
"""
from collections import namedtuple

import numpy as np

ConvSpec = namedtuple("ConvSpec", "name N Cin Cout H k stride pad group sparsity bias relu")


def _c(name, N, Cin, Cout, H, k, s, p, g, sp, bias=True, relu=False):
    return ConvSpec(name, N, Cin, Cout, H, k, s, p, g, sp, bias, relu)


LENET = [
    _c("lenet/conv1", 64, 1, 20, 28, 5, 1, 0, 1, 0.80),
    _c("lenet/conv2", 64, 20, 50, 12, 5, 1, 0, 1, 0.80),
]

ALEXNET = [
    _c("alexnet/conv2", 256, 96, 256, 27, 5, 1, 2, 2, 0.85),
    _c("alexnet/conv3", 256, 256, 384, 13, 3, 1, 1, 1, 0.88),
    _c("alexnet/conv4", 256, 384, 384, 13, 3, 1, 1, 2, 0.88),
    _c("alexnet/conv5", 256, 384, 256, 13, 3, 1, 1, 2, 0.88),
]

# (3x3 reduce -> 3x3, 5x5 reduce -> 5x5, spatial) per inception module, bvlc_googlenet/test_sconv.prototxt
_GOOGLENET_INC = [
    ("3a", 96, 128, 16, 32, 28), ("3b", 128, 192, 32, 96, 28),
    ("4a", 96, 208, 16, 48, 14), ("4b", 112, 224, 24, 64, 14), ("4c", 128, 256, 24, 64, 14),
    ("4d", 144, 288, 32, 64, 14), ("4e", 160, 320, 32, 128, 14),
    ("5a", 160, 320, 32, 128, 7), ("5b", 192, 384, 48, 128, 7),
]
GOOGLENET = [_c("googlenet/conv2_3x3", 128, 64, 192, 56, 3, 1, 1, 1, 0.75)]
for _n, _r3, _o3, _r5, _o5, _hw in _GOOGLENET_INC:
    GOOGLENET.append(_c("googlenet/inception_%s_3x3" % _n, 128, _r3, _o3, _hw, 3, 1, 1, 1, 0.75))
    GOOGLENET.append(_c("googlenet/inception_%s_5x5" % _n, 128, _r5, _o5, _hw, 5, 1, 2, 1, 0.75))

# 16 branch2b 3x3 convs, bias_term: false (models/resnet/test_sconv.prototxt:170-180)
RESNET50 = []
for _stage, _cnt, _ch, _hw in ((2, 3, 64, 56), (3, 4, 128, 28), (4, 6, 256, 14), (5, 3, 512, 7)):
    for _b in range(_cnt):
        RESNET50.append(_c("resnet50/res%d%s_branch2b" % (_stage, "abcdef"[_b]), 256, _ch, _ch, _hw, 3, 1, 1, 1,
                           0.70, bias=False))

NETWORKS = {"lenet": LENET, "alexnet": ALEXNET, "googlenet": GOOGLENET, "resnet50": RESNET50}


def sweep_specs(N=64):
    """Config 5: sparsity x C=M x H=W x stride, 3x3 p1."""
    out = []
    for sp in (0.50, 0.60, 0.70, 0.80, 0.90, 0.95):
        for ch in (64, 128, 256, 512):
            for hw in (7, 14, 28, 56):
                for s in (1, 2):
                    out.append(_c("sweep/s%02d_c%d_h%d_st%d" % (round(sp * 100), ch, hw, s), N, ch, ch, hw, 3, s, 1,
                                  1, sp))
    return out


def out_dim(i, pad, k, s, d=1):
    return (i + 2 * pad - (d * (k - 1) + 1)) // s + 1


def prune_magnitude(w, sparsity):
    """Global magnitude pruning of one layer: zero the k = round(sparsity*count) smallest |w| (ties -> zero)."""
    w = np.array(w, dtype=np.float32, copy=True)
    k = int(round(sparsity * w.size))
    if k <= 0:
        return w
    flat = np.abs(w).ravel()
    thr = np.partition(flat, k - 1)[k - 1]
    w[np.abs(w) <= thr] = np.float32(0.0)
    return w


def make_layer_data(spec, layer_index=0, N=None, with_input=True):
    """Seeded synthetic (weights, bias, input) for a ConvSpec. Returns dict of numpy arrays."""
    rng = np.random.default_rng(1701 + layer_index)
    n = spec.N if N is None else N
    w = (rng.standard_normal((spec.Cout, spec.Cin // spec.group, spec.k, spec.k)) * 0.01).astype(np.float32)
    w = prune_magnitude(w, spec.sparsity)
    b = (rng.standard_normal(spec.Cout) * 0.1).astype(np.float32) if spec.bias else None
    x = None
    if with_input:
        x = rng.uniform(-1.0, 1.0, (n, spec.Cin, spec.H, spec.H)).astype(np.float32)
    return dict(w=w, bias=b, x=x)


def alg_work(spec, nnz, N=None):
    """Algorithmic FLOPs / bytes of one forward launch."""
    n = spec.N if N is None else N
    Ho = out_dim(spec.H, spec.pad, spec.k, spec.stride)
    flops = 2.0 * nnz * Ho * Ho * n
    byts = 4.0 * n * spec.Cin * spec.H * spec.H + 4.0 * n * spec.Cout * Ho * Ho + 8.0 * nnz + \
        4.0 * (spec.Cout + spec.group) + (4.0 * spec.Cout if spec.bias else 0.0)
    return flops, byts
