"""ctypes bindings for the CPU oracle (oracle/libescort_oracle.so) and, when built, the reference's own
code compiled in place (oracle/_ref/libescort_ref.so, oracle/_ref/libescort_ref_gpu.so).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (caffe_escoin_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "libescort_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libescort_ref.so")
_REF_GPU_SO = os.path.join(_HERE, "_ref", "libescort_ref_gpu.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists). Building the checker is not using it."""
    if force or not os.path.exists(_ORACLE_SO) or (
            os.path.exists("/root/reference") and not (os.path.exists(_REF_SO) and os.path.exists(_REF_GPU_SO))):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_ORACLE_SO)
        L.oracle_out_dim.restype = C.c_int
        L.oracle_padded_len.restype = C.c_long
        L.oracle_dense2csr.restype = C.c_int
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def have_ref():
    return os.path.exists(_REF_SO)


def have_ref_gpu():
    return os.path.exists(_REF_GPU_SO)


def ref():
    global _ref
    if _ref is None:
        if not have_ref():
            build()
        R = C.CDLL(_REF_SO)
        R.ref_conv_forward.restype = C.c_int
        R.ref_has_unit_stride.restype = C.c_int
        R.ref_vlen.restype = C.c_int
        R.ref_max_threads.restype = C.c_int
        _ref = R
    return _ref


def ref_gpu_path():
    return _REF_GPU_SO


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def out_dim(i, pad, k, s, d):
    return (i + 2 * pad - (d * (k - 1) + 1)) // s + 1


class Geom:
    """Convolution geometry, named like the reference's layer members (base_conv_layer.hpp)."""

    def __init__(self, num, Cin, H, W, Cout, k, stride=1, pad=0, dilation=1, group=1, kw=None, pad_w=None,
                 stride_w=None, dilation_w=None):
        self.num, self.Cin, self.H, self.W, self.Cout, self.group = num, Cin, H, W, Cout, group
        self.kh, self.kw = k, (k if kw is None else kw)
        self.pad_h, self.pad_w = pad, (pad if pad_w is None else pad_w)
        self.stride_h, self.stride_w = stride, (stride if stride_w is None else stride_w)
        self.dil_h, self.dil_w = dilation, (dilation if dilation_w is None else dilation_w)
        self.Ho = out_dim(H, self.pad_h, self.kh, self.stride_h, self.dil_h)
        self.Wo = out_dim(W, self.pad_w, self.kw, self.stride_w, self.dil_w)
        self.M = Cout // group
        self.N = (Cin // group) * self.kh * self.kw

    def wshape(self):
        return (self.Cout, self.Cin // self.group, self.kh, self.kw)

    def conv_args(self):
        return (self.kh, self.kw, self.pad_h, self.pad_w, self.stride_h, self.stride_w, self.dil_h, self.dil_w)


def weight_align(w, g, stretch=True):
    """oracle_weight_align: returns dict(values, colidx, rowptr, nnz_per_row, nz_num) with the reference's
    worst-case buffer sizes (base_conv_layer.cpp:509-513)."""
    w = np.ascontiguousarray(w, dtype=np.float32)
    cnt = w.size
    values = np.zeros(cnt, np.float32)
    colidx = np.zeros(cnt, np.int32)
    rowptr = np.zeros(g.Cout + g.group, np.int32)
    nnz_per_row = np.zeros(g.Cout, np.int32)
    nz_num = np.zeros(g.group, np.int32)
    lib().oracle_weight_align(_p(w, C.c_float), g.Cout, g.Cin, g.group, g.kh, g.kw, g.H, g.W, g.pad_h, g.pad_w,
                              int(stretch), _p(values, C.c_float), _p(colidx, C.c_int), _p(rowptr, C.c_int),
                              _p(nnz_per_row, C.c_int), _p(nz_num, C.c_int))
    return dict(values=values, colidx=colidx, rowptr=rowptr, nnz_per_row=nnz_per_row, nz_num=nz_num)


def conv_forward(x, csr, g, bias=None, relu=False, threads=0):
    """oracle_conv_forward (layer-level, stretched CSR)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    top = np.zeros((g.num, g.Cout, g.Ho, g.Wo), np.float32)
    if threads <= 0:
        threads = lib().oracle_max_threads()
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    lib().oracle_conv_forward(_p(x, C.c_float), g.num, g.Cin, g.H, g.W, g.Cout, g.group, *g.conv_args(),
                              _p(csr["values"], C.c_float), _p(csr["colidx"], C.c_int),
                              _p(csr["rowptr"], C.c_int), _p(b, C.c_float), int(relu), _p(top, C.c_float),
                              int(threads))
    return top


def ref_conv_forward(x, csr, g, bias=None, relu=False, threads=0, blocked=True):
    """The reference's own CPU kernels (oracle/_ref). Returns (top, used_blocked_kernel)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    top = np.zeros((g.num, g.Cout, g.Ho, g.Wo), np.float32)
    if threads <= 0:
        threads = ref().ref_max_threads()
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    used = ref().ref_conv_forward(_p(x, C.c_float), g.num, g.Cin, g.H, g.W, g.Cout, g.group, *g.conv_args(),
                                  _p(csr["values"], C.c_float), _p(csr["colidx"], C.c_int),
                                  _p(csr["rowptr"], C.c_int), _p(b, C.c_float), int(relu), _p(top, C.c_float),
                                  int(threads), int(blocked))
    return top, bool(used)


def dense_conv(x, w, g, bias=None, relu=False):
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    top = np.zeros((g.num, g.Cout, g.Ho, g.Wo), np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    lib().oracle_dense_conv(_p(x, C.c_float), g.num, g.Cin, g.H, g.W, _p(w, C.c_float), g.Cout, g.group,
                            *g.conv_args(), _p(b, C.c_float), int(relu), _p(top, C.c_float))
    return top


def conv_backward(x, top_diff, w, g, mask_only=True, want_w=True, want_b=True, want_x=True, w_diff=None,
                  b_diff=None):
    """oracle_conv_backward. w_diff / b_diff are accumulated into (copies of) the given arrays."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    top_diff = np.ascontiguousarray(top_diff, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    wd = (np.zeros_like(w) if w_diff is None else np.array(w_diff, np.float32, copy=True)) if want_w else None
    bd = (np.zeros(g.Cout, np.float32) if b_diff is None else np.array(b_diff, np.float32, copy=True)) \
        if want_b else None
    xd = np.zeros_like(x) if want_x else None
    lib().oracle_conv_backward(_p(x, C.c_float), _p(top_diff, C.c_float), g.num, g.Cin, g.H, g.W, _p(w, C.c_float),
                               g.Cout, g.group, *g.conv_args(), int(mask_only), _p(wd, C.c_float),
                               _p(bd, C.c_float), _p(xd, C.c_float))
    return wd, bd, xd


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d > 0 else float(np.linalg.norm(a - b))
