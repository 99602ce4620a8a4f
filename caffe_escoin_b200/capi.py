"""ctypes binding of include/escort_b200.h.  Device buffers are torch CUDA tensors; every call goes through
the C ABI exactly as the Caffe overlay would.  No fallback: a missing library raises."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libescort_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError("caffe_escoin_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU or PyTorch fallback for the sparse-conv path)" % LIB_PATH)

lib = C.CDLL(LIB_PATH)


class Geom(C.Structure):
    """escort_geom (BaseConvolutionLayer member names)."""
    _fields_ = [(n, C.c_int) for n in ("channels", "num_output", "group", "height", "width", "kernel_h", "kernel_w",
                                       "pad_h", "pad_w", "stride_h", "stride_w", "dilation_h", "dilation_w")]


EXPORTS = ["escort_pack_csr", "escort_stretch", "escort_copy_input", "escort_sconv_padded", "escort_plan_create",
           "escort_plan_destroy", "escort_plan_nnz", "escort_plan_kernel_name", "escort_plan_describe",
           "escort_plan_set_variant", "escort_plan_set_config", "escort_plan_get_config", "escort_plan_autotune",
           "escort_plan_autotune_backward", "escort_plan_copy_tuning",
           "escort_sconv_forward", "escort_sconv_backward_data", "escort_sconv_backward_weight",
           "escort_bias_backward", "escort_refresh_values", "escort_allreduce_grads", "escort_broadcast",
           "escort_comm_unique_id", "escort_comm_init_rank", "escort_comm_destroy", "escort_tmem_debug", "escort_measure_fp32_peak",
           "escort_inner_product_forward", "escort_dense_conv_workspace_bytes", "escort_dense_conv_forward", "escort_dense_conv_forward_residual", "escort_dense_fold_affine",
           "escort_bn_scale_to_affine", "escort_plan_fold_affine", "escort_lowered_sparse_forward", "escort_caffemodel_open", "escort_caffemodel_close", "escort_caffemodel_save", "escort_caffemodel_num_layers",
           "escort_caffemodel_find", "escort_caffemodel_layer", "escort_caffemodel_blob", "escort_prune_magnitude",
           "escort_last_error", "escort_version"]

lib.escort_last_error.restype = C.c_char_p
lib.escort_version.restype = C.c_char_p
lib.escort_plan_kernel_name.restype = C.c_char_p
lib.escort_plan_kernel_name.argtypes = [C.c_void_p]
lib.escort_plan_nnz.restype = C.c_long
lib.escort_plan_nnz.argtypes = [C.c_void_p]
lib.escort_plan_destroy.argtypes = [C.c_void_p]
lib.escort_plan_set_variant.argtypes = [C.c_void_p, C.c_int]
lib.escort_plan_set_config.argtypes = [C.c_void_p, C.c_int, C.c_int]
lib.escort_plan_get_config.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
lib.escort_plan_autotune.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
lib.escort_plan_autotune_backward.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
lib.escort_plan_copy_tuning.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
lib.escort_plan_describe.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
lib.escort_caffemodel_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
lib.escort_caffemodel_close.argtypes = [C.c_void_p]
lib.escort_caffemodel_save.argtypes = [C.c_void_p, C.c_char_p]
lib.escort_caffemodel_num_layers.argtypes = [C.c_void_p]
lib.escort_caffemodel_find.argtypes = [C.c_void_p, C.c_char_p]
lib.escort_caffemodel_layer.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
lib.escort_caffemodel_blob.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_long),
                                       C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_long)]


class EscortError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise EscortError("%s failed (rc=%d): %s" % (what, rc, lib.escort_last_error().decode()))


def _ptr(t):
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "device-contiguous tensor required"
    return C.c_void_p(t.data_ptr())


def _stream(stream):
    if stream is None:
        stream = torch.cuda.current_stream()
    return C.c_void_p(stream.cuda_stream)


def out_dim(i, pad, k, s, d=1):
    return (i + 2 * pad - (d * (k - 1) + 1)) // s + 1


def make_geom(Cin, Cout, H, W, k, stride=1, pad=0, dilation=1, group=1, kw=None, pad_w=None, stride_w=None,
              dilation_w=None):
    return Geom(Cin, Cout, group, H, W, k, k if kw is None else kw, pad, pad if pad_w is None else pad_w, stride,
                stride if stride_w is None else stride_w, dilation, dilation if dilation_w is None else dilation_w)


def pack_csr(A2d, stream=None, want_nnz_per_row=True):
    """escort_pack_csr on a row-major M x N device matrix. Returns (values, colidx, rowptr, nnz_per_row, nnz)."""
    M, N = A2d.shape
    dev = A2d.device
    values = torch.zeros(M * N, dtype=torch.float32, device=dev)
    colidx = torch.zeros(M * N, dtype=torch.int32, device=dev)
    rowptr = torch.zeros(M + 1, dtype=torch.int32, device=dev)
    npr = torch.zeros(M, dtype=torch.int32, device=dev) if want_nnz_per_row else None
    nnz = C.c_int(-1)
    _check(lib.escort_pack_csr(M, N, _ptr(A2d), _ptr(npr), _ptr(values), _ptr(rowptr), _ptr(colidx), C.byref(nnz),
                               _stream(stream)), "escort_pack_csr")
    return values, colidx, rowptr, npr, nnz.value


def weight_align(weights, geom, stretch=True, stream=None):
    """BaseConvolutionLayer::WeightAlign, GPU branch (reference base_conv_layer.cpp:236-264): per group
    dense->CSR into the layer's worst-case-sized blobs, then stretch.  Returns dict like the oracle's."""
    g = geom
    M = g.num_output // g.group
    N = (g.channels // g.group) * g.kernel_h * g.kernel_w
    dev = weights.device
    w = weights.contiguous().view(g.num_output, N)
    values = torch.zeros(g.num_output * N, dtype=torch.float32, device=dev)
    colidx = torch.zeros(g.num_output * N, dtype=torch.int32, device=dev)
    rowptr = torch.zeros(g.num_output + g.group, dtype=torch.int32, device=dev)
    npr = torch.zeros(g.num_output, dtype=torch.int32, device=dev)
    nz_num = []
    woff, roff = M * N, M + 1
    for gi in range(g.group):
        nnz = C.c_int(-1)
        _check(lib.escort_pack_csr(M, N, C.c_void_p(w.data_ptr() + 4 * woff * gi),
                                   C.c_void_p(npr.data_ptr() + 4 * M * gi),
                                   C.c_void_p(values.data_ptr() + 4 * woff * gi),
                                   C.c_void_p(rowptr.data_ptr() + 4 * roff * gi),
                                   C.c_void_p(colidx.data_ptr() + 4 * woff * gi), C.byref(nnz), _stream(stream)),
               "escort_pack_csr")
        nz_num.append(nnz.value)
        if stretch:
            _check(lib.escort_stretch(C.c_void_p(rowptr.data_ptr() + 4 * roff * gi),
                                      C.c_void_p(colidx.data_ptr() + 4 * woff * gi), M, g.height, g.width, g.pad_h,
                                      g.pad_w, g.kernel_h, g.kernel_w, _stream(stream)), "escort_stretch")
    return dict(values=values, colidx=colidx, rowptr=rowptr, nnz_per_row=npr, nz_num=nz_num, stretched=stretch)


class Plan:
    """Owner of an escort_plan*."""

    def __init__(self, geom, csr, stream=None):
        self.geom = geom
        self.Ho = out_dim(geom.height, geom.pad_h, geom.kernel_h, geom.stride_h, geom.dilation_h)
        self.Wo = out_dim(geom.width, geom.pad_w, geom.kernel_w, geom.stride_w, geom.dilation_w)
        h = C.c_void_p(0)
        _check(lib.escort_plan_create(C.byref(geom), _ptr(csr["rowptr"]), _ptr(csr["colidx"]), _ptr(csr["values"]),
                                      int(bool(csr.get("stretched", True))), C.byref(h), _stream(stream)),
               "escort_plan_create")
        self.h = h

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:   # (module globals are gone at interpreter shutdown)
            lib.escort_plan_destroy(self.h)
            self.h = None

    @property
    def nnz(self):
        return lib.escort_plan_nnz(self.h)

    @property
    def kernel_name(self):
        return lib.escort_plan_kernel_name(self.h).decode()

    def describe(self):
        buf = C.create_string_buffer(2048)
        lib.escort_plan_describe(self.h, buf, 2048)
        return buf.value.decode()

    def kernel_names(self):
        """{'fwd': ..., 'bwd_data': ..., 'bwd_weight': ...}: kernels in use (backward entries once those plans exist)."""
        parts = self.describe().split(" | ")
        out = {"fwd": parts[0].split(" ")[0] if parts[0] != "generic" else self.kernel_name,   # sconv_fwd_generic / sconv_fwd_small
               "bwd_data": "sconv_bwd_data_generic", "bwd_weight": "sconv_bwd_weight_generic"}
        for p in parts[1:]:
            k, v = p.split(": ", 1)
            out[k] = v.split(" ")[0] + ("(transposed plan)" if k == "bwd_data" else "")
        return out

    def set_variant(self, v):
        _check(lib.escort_plan_set_variant(self.h, int(v)), "escort_plan_set_variant")

    def set_config(self, v, rank):
        _check(lib.escort_plan_set_config(self.h, int(v), int(rank)), "escort_plan_set_config")

    def get_config(self):
        v, r = C.c_int(0), C.c_int(0)
        _check(lib.escort_plan_get_config(self.h, C.byref(v), C.byref(r)), "escort_plan_get_config")
        return v.value, r.value

    def autotune(self, num, stream=None):
        _check(lib.escort_plan_autotune(self.h, int(num), _stream(stream)), "escort_plan_autotune")

    def copy_tuning(self, other, stream=None):
        _check(lib.escort_plan_copy_tuning(self.h, other.h, _stream(stream)), "escort_plan_copy_tuning")

    def autotune_backward(self, num, stream=None):
        _check(lib.escort_plan_autotune_backward(self.h, int(num), _stream(stream)), "escort_plan_autotune_backward")

    def forward(self, bottom, bias=None, relu=False, top=None, stream=None):
        g = self.geom
        num = bottom.shape[0]
        if top is None:
            top = torch.empty((num, g.num_output, self.Ho, self.Wo), dtype=torch.float32, device=bottom.device)
        _check(lib.escort_sconv_forward(self.h, num, _ptr(bottom), _ptr(bias), int(relu), _ptr(top), _stream(stream)),
               "escort_sconv_forward")
        return top

    def backward_data(self, top_diff, bottom_diff=None, stream=None):
        g = self.geom
        num = top_diff.shape[0]
        if bottom_diff is None:
            bottom_diff = torch.empty((num, g.channels, g.height, g.width), dtype=torch.float32,
                                      device=top_diff.device)
        _check(lib.escort_sconv_backward_data(self.h, num, _ptr(top_diff), _ptr(bottom_diff), _stream(stream)),
               "escort_sconv_backward_data")
        return bottom_diff

    def backward_weight(self, bottom, top_diff, wd_dense=None, wd_csr=None, accumulate=True, stream=None):
        _check(lib.escort_sconv_backward_weight(self.h, bottom.shape[0], _ptr(bottom), _ptr(top_diff), _ptr(wd_dense),
                                                _ptr(wd_csr), int(accumulate), _stream(stream)),
               "escort_sconv_backward_weight")

    def refresh_values(self, weights_dense, values_csr=None, stream=None):
        _check(lib.escort_refresh_values(self.h, _ptr(weights_dense), _ptr(values_csr), _stream(stream)),
               "escort_refresh_values")


def bias_backward(top_diff, bias_diff, stream=None):
    num, M = top_diff.shape[0], top_diff.shape[1]
    spatial = top_diff.shape[2] * top_diff.shape[3]
    _check(lib.escort_bias_backward(num, M, spatial, _ptr(top_diff), _ptr(bias_diff), _stream(stream)),
           "escort_bias_backward")


def copy_input(dst, src, Cn, H, W, pad_h, pad_w, stream=None):
    _check(lib.escort_copy_input(_ptr(dst), _ptr(src), Cn, H, W, pad_h, pad_w, _stream(stream)), "escort_copy_input")


def sconv_padded(fuse_relu, num, inp, ifmap_size, rowptr, colidx, values, bias, H, W, pad_h, pad_w, stride_h,
                 stride_w, dil_h, dil_w, kh, kw, out, num_oc, num_groups, stream=None):
    """caffe_gpu_sconv's exact argument list (pointers may be ints = device addresses)."""
    def p(x):
        return C.c_void_p(x) if isinstance(x, int) else _ptr(x)
    _check(lib.escort_sconv_padded(int(fuse_relu), num, p(inp), ifmap_size, p(rowptr), p(colidx), p(values), p(bias),
                                   H, W, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w, kh, kw, p(out), num_oc,
                                   num_groups, _stream(stream)), "escort_sconv_padded")


def allreduce_grads(flat, scale, comm=None, stream=None):
    """ncclAllReduce(sum) over `flat` on `comm` (an NcclComm, a raw ncclComm_t address, or None = scale only), then * scale."""
    handle = comm.handle if isinstance(comm, NcclComm) else (comm or 0)
    _check(lib.escort_allreduce_grads(C.c_void_p(handle), _ptr(flat), C.c_size_t(flat.numel()), C.c_float(scale),
                                      _stream(stream)), "escort_allreduce_grads")


def broadcast(buf, root, comm, stream=None):
    handle = comm.handle if isinstance(comm, NcclComm) else comm
    _check(lib.escort_broadcast(C.c_void_p(handle), _ptr(buf), C.c_size_t(buf.numel()), int(root), _stream(stream)),
           "escort_broadcast")


class NcclComm:
    """An ncclComm_t created through the library's own bring-up entries (escort_comm_unique_id / _init_rank).
    `exchange_id(id_bytes_or_None) -> id_bytes` ships rank 0's 128-byte id to every rank (e.g. a torch.distributed or
    MPI broadcast); with nranks == 1 no exchange is needed."""

    def __init__(self, nranks=1, rank=0, exchange_id=None):
        buf = (C.c_char * 128)()
        if rank == 0:
            _check(lib.escort_comm_unique_id(buf), "escort_comm_unique_id")
        raw = bytes(buf)
        if nranks > 1:
            raw = exchange_id(raw if rank == 0 else None)
        buf = (C.c_char * 128).from_buffer_copy(raw)
        h = C.c_void_p(0)
        _check(lib.escort_comm_init_rank(C.byref(h), nranks, buf, rank), "escort_comm_init_rank")
        self.handle, self.nranks, self.rank = h.value, nranks, rank

    def destroy(self):
        if getattr(self, "handle", None):
            lib.escort_comm_destroy(C.c_void_p(self.handle))
            self.handle = None


def measure_fp32_peak(variant=0, iters=4096):
    tf, sms, khz = C.c_double(0), C.c_int(0), C.c_int(0)
    _check(lib.escort_measure_fp32_peak(variant, iters, C.byref(tf), C.byref(sms), C.byref(khz)),
           "escort_measure_fp32_peak")
    return tf.value, sms.value, khz.value


lib.escort_dense_conv_workspace_bytes.restype = C.c_size_t


def inner_product_forward(bottom, weight, bias=None, relu=False, stream=None):
    """InnerProduct forward on tcgen05 (TF32): bottom [num x K], weight [num_output x K]."""
    num, K = bottom.shape
    top = torch.empty((num, weight.shape[0]), device=bottom.device)
    _check(lib.escort_inner_product_forward(num, K, weight.shape[0], _ptr(bottom), _ptr(weight), _ptr(bias), int(relu), _ptr(top),
                                            _stream(stream)), "escort_inner_product_forward")
    return top


def dense_conv_forward(geom, bottom, weight, bias=None, relu=False, stream=None, residual=None, top=None, workspace=None):
    """Dense convolution (conv1 / 1x1 / unpruned) on tcgen05 (TF32): implicit GEMM from NCHW for 1x1 / stride 1, else
    transposed im2col + GEMM.  residual: the other input of the Eltwise SUM behind the layer (fused into the epilogue)."""
    num = bottom.shape[0]
    Ho = out_dim(geom.height, geom.pad_h, geom.kernel_h, geom.stride_h, geom.dilation_h)
    Wo = out_dim(geom.width, geom.pad_w, geom.kernel_w, geom.stride_w, geom.dilation_w)
    nbytes = lib.escort_dense_conv_workspace_bytes(C.byref(geom), num)
    ws = workspace if workspace is not None else torch.empty(nbytes, dtype=torch.uint8, device=bottom.device)
    assert ws.numel() * ws.element_size() >= nbytes
    if top is None:
        top = torch.empty((num, geom.num_output, Ho, Wo), device=bottom.device)
    _check(lib.escort_dense_conv_forward_residual(C.byref(geom), num, _ptr(bottom), _ptr(weight), _ptr(bias), _ptr(residual), int(relu),
                                                  _ptr(ws), C.c_size_t(nbytes), _ptr(top), _stream(stream)),
           "escort_dense_conv_forward_residual")
    return top


def dense_conv_workspace_bytes(geom, num):
    return int(lib.escort_dense_conv_workspace_bytes(C.byref(geom), num))


def dense_fold_affine(weights, a, b, bias=None, stream=None):
    """W'[oc] = W[oc] * a[oc], bias' = bias * a + b for a layer that stays dense; returns (W', bias')."""
    wf = torch.empty_like(weights)
    bias_out = torch.empty_like(a)
    M = weights.shape[0]
    _check(lib.escort_dense_fold_affine(M, C.c_long(weights.numel() // M), _ptr(weights), _ptr(a), _ptr(b), _ptr(bias), _ptr(wf),
                                        _ptr(bias_out), _stream(stream)), "escort_dense_fold_affine")
    return wf, bias_out


def bn_scale_to_affine(mean, var, scale_factor_blob, eps, gamma=None, beta=None, stream=None):
    """(a, b) of y = x * a + b for BatchNorm(use_global_stats) followed by Scale (device tensors of num_output floats)."""
    a, b = torch.empty_like(mean), torch.empty_like(mean)
    _check(lib.escort_bn_scale_to_affine(mean.numel(), _ptr(mean), _ptr(var), C.c_float(scale_factor_blob), C.c_float(eps), _ptr(gamma),
                                         _ptr(beta), _ptr(a), _ptr(b), _stream(stream)), "escort_bn_scale_to_affine")
    return a, b


def fold_affine(plan, weights_dense, a, b, bias=None, stream=None):
    """Fold y = conv * a[oc] + b[oc] into the plan; returns the bias to pass to plan.forward(..., relu=True)."""
    wf = torch.empty_like(weights_dense)
    bias_out = torch.empty_like(a)
    _check(lib.escort_plan_fold_affine(plan.h, _ptr(weights_dense), _ptr(a), _ptr(b), _ptr(bias), _ptr(wf), _ptr(bias_out),
                                       _stream(stream)), "escort_plan_fold_affine")
    return bias_out


def lowered_sparse_forward(geom, bottom, csr_raw, bias=None, relu=False, stream=None):
    """The LOWERED_SPARSE comparator (im2col + cusparseSpMM); csr_raw = weight_align(..., stretch=False)."""
    num = bottom.shape[0]
    Ho = out_dim(geom.height, geom.pad_h, geom.kernel_h, geom.stride_h, geom.dilation_h)
    Wo = out_dim(geom.width, geom.pad_w, geom.kernel_w, geom.stride_w, geom.dilation_w)
    col = torch.empty(geom.channels * geom.kernel_h * geom.kernel_w * Ho * Wo, device=bottom.device)
    top = torch.empty((num, geom.num_output, Ho, Wo), device=bottom.device)
    _check(lib.escort_lowered_sparse_forward(C.byref(geom), num, _ptr(bottom), _ptr(csr_raw["rowptr"]), _ptr(csr_raw["colidx"]),
                                             _ptr(csr_raw["values"]), _ptr(bias), int(relu), _ptr(col), _ptr(top), _stream(stream)),
           "escort_lowered_sparse_forward")
    return top


class LayerInfo(C.Structure):
    """escort_layer_info"""
    _fields_ = [("name", C.c_char_p), ("type", C.c_char_p)] + [(n, C.c_int) for n in (
        "num_blobs", "is_conv", "is_inner_product", "num_output", "bias_term", "group", "kernel_h", "kernel_w", "stride_h",
        "stride_w", "pad_h", "pad_w", "dilation")]


class CaffeModel:
    """A `.caffemodel` opened through the C ABI (escort_caffemodel_*): what Net::CopyTrainedLayersFrom reads
    (src/caffe/net.cpp:785-821).  Blob arrays are numpy views of the model's own host memory (writable: pruning)."""

    def __init__(self, path):
        h = C.c_void_p(0)
        _check(lib.escort_caffemodel_open(path.encode(), C.byref(h)), "escort_caffemodel_open")
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            lib.escort_caffemodel_close(self.handle)
            self.handle = None

    __del__ = close

    def __len__(self):
        return lib.escort_caffemodel_num_layers(self.handle)

    def find(self, name):
        return lib.escort_caffemodel_find(self.handle, name.encode())

    def layer(self, i):
        info = LayerInfo()
        _check(lib.escort_caffemodel_layer(self.handle, i, C.byref(info)), "escort_caffemodel_layer")
        d = {n: getattr(info, n) for n, _ in LayerInfo._fields_}
        d["name"], d["type"] = info.name.decode(), info.type.decode()
        return d

    def blob(self, i, j):
        import numpy as np
        ndim, shape, data, count = C.c_int(0), (C.c_long * 8)(), C.POINTER(C.c_float)(), C.c_long(0)
        _check(lib.escort_caffemodel_blob(self.handle, i, j, C.byref(ndim), shape, C.byref(data), C.byref(count)),
               "escort_caffemodel_blob")
        if count.value == 0:
            return np.zeros(tuple(shape[:ndim.value]), dtype=np.float32)
        return np.ctypeslib.as_array(data, shape=(count.value,)).reshape(tuple(shape[:ndim.value]))

    def save(self, path):
        _check(lib.escort_caffemodel_save(self.handle, path.encode()), "escort_caffemodel_save")


def prune_magnitude(w, sparsity):
    """In place on a C-contiguous float32 numpy array; returns (threshold, nnz)."""
    import numpy as np
    assert w.dtype == np.float32 and w.flags["C_CONTIGUOUS"]
    thr, nnz = C.c_float(0), C.c_long(0)
    _check(lib.escort_prune_magnitude(w.ctypes.data_as(C.POINTER(C.c_float)), C.c_long(w.size), C.c_double(sparsity),
                                      C.byref(thr), C.byref(nnz)), "escort_prune_magnitude")
    return thr.value, nnz.value
