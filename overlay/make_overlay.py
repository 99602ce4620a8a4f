#!/usr/bin/env python
"""Generate overlay/*.patch: the edits that put libescort_b200.so behind the reference's own call sites
(Level 1 + Level 2), as unified diffs against chenxuhao/caffe-escoin.

    python overlay/make_overlay.py [/root/reference]

Each edit is anchored on a line of the reference that must occur exactly once; the script copies nothing from the
reference into this repository except the two context lines `diff -U2` keeps around every hunk.  tests/test_overlay.py
applies the committed patches with `patch --dry-run` to a scratch copy of the reference files (CPU suite; skipped where
/root/reference is absent)."""
import difflib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def once(text, anchor):
    assert text.count(anchor) == 1, "anchor not unique (%d): %r" % (text.count(anchor), anchor[:60])
    return text.index(anchor)


def insert_after(text, anchor, new):
    i = once(text, anchor) + len(anchor)
    return text[:i] + new + text[i:]


def insert_before(text, anchor, new):
    i = once(text, anchor)
    return text[:i] + new + text[i:]


def replace_between(text, start, end, new):
    """Replace from the line containing `start` up to (not including) the line containing `end`."""
    i = once(text, start)
    i = text.rfind("\n", 0, i) + 1
    j = text.index(end, i)
    j = text.rfind("\n", 0, j) + 1
    return text[:i] + new + text[j:]


EDITS = {}


def edit(path):
    def deco(fn):
        EDITS[path] = fn
        return fn
    return deco


@edit("include/caffe/util/device_alternate.hpp")
def _(t):
    return insert_before(t, "#define CUBLAS_CHECK(condition) \\", '''// escort-b200: the sparse-conv hot path lives in libescort_b200.so (C ABI, include/escort_b200.h)
#include "escort_b200.h"
#define ESCORT_CHECK(call) \\
  do { \\
    int escort_rc__ = (call); \\
    CHECK_EQ(escort_rc__, 0) << " " << escort_last_error(); \\
  } while (0)

''')


@edit("include/caffe/layers/base_conv_layer.hpp")
def _(t):
    t = insert_before(t, "template <typename Dtype>\nclass BaseConvolutionLayer : public Layer<Dtype> {",
                      "struct escort_plan;  // escort-b200 (include/escort_b200.h; forward declaration keeps CPU_ONLY builds intact)\n\n")
    t = t.replace(": Layer<Dtype>(param), input_padded_(NULL), output_scratch_(NULL) {}",
                  ": Layer<Dtype>(param), input_padded_(NULL), output_scratch_(NULL), escort_plan_(NULL) {}")
    return insert_after(t, "  Dtype *d_input_padded_; // GPU pointer\n",
                        "  escort_plan *escort_plan_;  // escort-b200: what WeightAlign leaves for Forward_gpu / Backward_gpu (float nets)\n")


@edit("src/caffe/layers/base_conv_layer.cpp")
def _(t):
    # destructor: the plan (the padded scratch stays: the reference's double-precision path still uses it)
    t = insert_before(t, "\t\t\t\tCUDA_CHECK(cudaFree(d_input_padded_));\n", "\t\t\t\tescort_plan_destroy(escort_plan_);  // escort-b200\n")
    # WeightAlign, GPU branch: after the last group's CSR is packed and stretched, build the plan
    t = insert_after(t, "\t\t\t\t\tcaffe_gpu_stretch(rowptr, colidx, M, height, width, pad_h, pad_w, kernel_h, kernel_w);\n", '''					if (g == group_ - 1 && sizeof(Dtype) == sizeof(float)) {
						// escort-b200: the plan (byte-code of the tile kernels, tuned once at load) replaces the per-image loop
						escort_geom eg = {conv_in_channels_, conv_out_channels_, group_, height, width, kernel_h, kernel_w,
								pad_h, pad_w, stride_.cpu_data()[0], stride_.cpu_data()[1],
								dilation_.cpu_data()[0], dilation_.cpu_data()[1]};
						escort_plan_destroy(escort_plan_);
						escort_plan_ = NULL;
						ESCORT_CHECK(escort_plan_create(&eg, nz_weight_index_pointers_.gpu_data(), nz_weight_indices_.gpu_data(),
								reinterpret_cast<const float*>(nz_weight_values_.gpu_data()), /*colidx_is_stretched=*/1,
								&escort_plan_, NULL));
						ESCORT_CHECK(escort_plan_autotune(escort_plan_, num_, NULL));
						if (this->phase_ == TRAIN) ESCORT_CHECK(escort_plan_autotune_backward(escort_plan_, num_, NULL));
					}
''')
    return t


@edit("src/caffe/layers/conv_layer.cu")
def _(t):
    t = insert_after(t, "    Dtype* top_data = top[i]->mutable_gpu_data();\n", '''    if ((Caffe::conv_mode() == Caffe::SCONV_PAR || Caffe::conv_mode() == Caffe::SCONV) && this->escort_plan_) {
      // escort-b200: the whole batch, all groups, bias fused: one launch instead of num_ x (pad copy + group launches + bias GEMM)
      ESCORT_CHECK(escort_sconv_forward(this->escort_plan_, this->num_, reinterpret_cast<const float*>(bottom_data),
          this->bias_term_ ? reinterpret_cast<const float*>(this->blobs_[1]->gpu_data()) : NULL, /*fuse_relu=*/0,
          reinterpret_cast<float*>(top_data), NULL));
      continue;
    }
''')
    t = insert_after(t, "    const Dtype* top_diff = top[i]->gpu_diff();\n", '''    if ((Caffe::conv_mode() == Caffe::SCONV_PAR || Caffe::conv_mode() == Caffe::SCONV) && this->escort_plan_) {
      // escort-b200: backward restricted to the sparsity mask; same contract (parameter diffs accumulate, bottom diff is overwritten)
      const float* td = reinterpret_cast<const float*>(top_diff);
      if (this->bias_term_ && this->param_propagate_down_[1])
        ESCORT_CHECK(escort_bias_backward(this->num_, this->num_output_, this->out_spatial_dim_, td,
            reinterpret_cast<float*>(this->blobs_[1]->mutable_gpu_diff()), NULL));
      if (this->param_propagate_down_[0])
        ESCORT_CHECK(escort_sconv_backward_weight(this->escort_plan_, this->num_,
            reinterpret_cast<const float*>(bottom[i]->gpu_data()), td, reinterpret_cast<float*>(weight_diff), NULL,
            /*accumulate=*/1, NULL));
      if (propagate_down[i])
        ESCORT_CHECK(escort_sconv_backward_data(this->escort_plan_, this->num_, td,
            reinterpret_cast<float*>(bottom[i]->mutable_gpu_diff()), NULL));
      continue;
    }
''')
    return t


@edit("src/caffe/layers/conv_relu_layer.cu")
def _(t):
    return insert_after(t, "    Dtype* top_data = top[i]->mutable_gpu_data();\n", '''    if ((Caffe::conv_mode() == Caffe::SCONV_PAR || Caffe::conv_mode() == Caffe::SCONV) && this->escort_plan_) {
      // escort-b200: bias + ReLU fused into the epilogue of the one launch
      ESCORT_CHECK(escort_sconv_forward(this->escort_plan_, this->num_, reinterpret_cast<const float*>(bottom_data),
          this->bias_term_ ? reinterpret_cast<const float*>(this->blobs_[1]->gpu_data()) : NULL, /*fuse_relu=*/1,
          reinterpret_cast<float*>(top_data), NULL));
      continue;
    }
''')


@edit("src/caffe/parallel.cpp")
def _(t):
    return replace_between(t, "    NCCL_CHECK(ncclAllReduce(diff_, diff_, static_cast<int>(size_),", "  }\n}\n\ntemplate<typename Dtype>\nclass Worker : public InternalThread {", '''    if (sizeof(Dtype) == sizeof(float)) {
      // escort-b200: all-reduce + 1/N in one entry (NCCL 2 communicator)
      ESCORT_CHECK(escort_allreduce_grads(comm_, reinterpret_cast<float*>(diff_), size_,
                                          1.f / Caffe::solver_count(), cudaStreamDefault));
    } else {
      NCCL_CHECK(ncclAllReduce(diff_, diff_, static_cast<int>(size_),
                               nccl::dataType<Dtype>::type, ncclSum, comm_,
                               cudaStreamDefault));
      caffe_gpu_scal(static_cast<int>(size_),
                     (Dtype) 1.0 / Caffe::solver_count(), diff_);
    }
''')


@edit("src/caffe/util/math_functions.cu")
def _(t):
    """Level 1: the four free functions of the hot path forward to the library (insertions only: the reference bodies
    stay behind `#ifdef ESCORT_KEEP_LEGACY_CUSPARSE` / an early return).  cusparseSnnz / cusparse?dense2csc no longer
    exist in CUDA 12, so this is also what makes the fork link at all."""
    t = insert_after(t, "    float* A_nonzero_buf, int* A_idx_pointer_buf, int* A_nonzero_idx_buf, int *nnz_total) {\n",
                     "  // escort-b200: count -> scan -> ordered scatter, bit-exact with the sequential scan of math_functions.cpp:92-105\n"
                     "  ESCORT_CHECK(escort_pack_csr(M, N, A, nnzPerRow, A_nonzero_buf, A_idx_pointer_buf, A_nonzero_idx_buf, nnz_total, NULL));\n"
                     "#ifdef ESCORT_KEEP_LEGACY_CUSPARSE\n")
    t = insert_before(t, "}\n\ntemplate <>\nvoid caffe_gpu_sparse_dense2csr<double>(", "#endif  // ESCORT_KEEP_LEGACY_CUSPARSE\n")
    t = insert_after(t, "    double* A_nonzero_buf, int* A_idx_pointer_buf, int* A_nonzero_idx_buf,int *nnz_total) {\n",
                     "#ifndef ESCORT_KEEP_LEGACY_CUSPARSE\n  NOT_IMPLEMENTED;  // escort-b200 is fp32 (the fork's sparse path is only exercised with float nets)\n#else\n")
    t = insert_before(t, "}\n\ntemplate <typename Dtype>\n__global__ void sconv_dilation(", "#endif  // ESCORT_KEEP_LEGACY_CUSPARSE\n")
    t = insert_after(t, "\tint stride_h, int stride_w, int dilation_h, int dilation_w, int kernel_h, int kernel_w, Dtype *output, int num_oc, int num_groups)\n{\n",
                     "\tif (sizeof(Dtype) == sizeof(float)) {  // escort-b200: identical argument list, whole batch in one launch\n"
                     "\t\tESCORT_CHECK(escort_sconv_padded(FUSE_RELU, num, reinterpret_cast<const float*>(input), ifmap_size, rowptr, colidx,\n"
                     "\t\t\t\treinterpret_cast<const float*>(values), reinterpret_cast<const float*>(bias), height, width, pad_h, pad_w,\n"
                     "\t\t\t\tstride_h, stride_w, dilation_h, dilation_w, kernel_h, kernel_w, reinterpret_cast<float*>(output), num_oc,\n"
                     "\t\t\t\tnum_groups, NULL));\n\t\treturn;\n\t}\n")
    t = insert_after(t, "void caffe_gpu_stretch(const int *rowptr, int *colidx, int M, \n\t\tint height, int width, int pad_h, int pad_w, int kernel_h, int kernel_w) {\n",
                     "\tESCORT_CHECK(escort_stretch(rowptr, colidx, M, height, width, pad_h, pad_w, kernel_h, kernel_w, NULL));  // escort-b200\n\treturn;\n")
    t = insert_after(t, "void copy_input_data(Dtype *dst, const Dtype *src, int num_channels, int height, int width, int pad_h, int pad_w) {\n",
                     "\tif (sizeof(Dtype) == sizeof(float)) {  // escort-b200\n"
                     "\t\tESCORT_CHECK(escort_copy_input(reinterpret_cast<float*>(dst), reinterpret_cast<const float*>(src), num_channels, height,\n"
                     "\t\t\t\twidth, pad_h, pad_w, NULL));\n\t\treturn;\n\t}\n")
    return t


# ---- the layers the reference keeps dense, on tcgen05 (TF32 products, fp32 accumulate) ----
@edit("src/caffe/layers/inner_product_layer.cu")
def _(t):
    return insert_before(t, "  if (M_ == 1) {\n    caffe_gpu_gemv<Dtype>(CblasNoTrans, N_, K_, (Dtype)1.,\n                         weight, bottom_data, (Dtype)0., top_data);",
                         """  if (sizeof(Dtype) == sizeof(float) && !transpose_ && M_ > 1 && K_ % 4 == 0) {
    // escort-b200: one tcgen05 GEMM (TMA-staged operands, TMEM accumulator) with the bias fused, instead of sgemm + rank-1 bias GEMM
    ESCORT_CHECK(escort_inner_product_forward(M_, K_, N_, reinterpret_cast<const float*>(bottom_data),
        reinterpret_cast<const float*>(weight),
        bias_term_ ? reinterpret_cast<const float*>(this->blobs_[1]->gpu_data()) : NULL, /*fuse_relu=*/0,
        reinterpret_cast<float*>(top_data), NULL));
    return;
  }
""")


@edit("include/caffe/layers/esc_conv_layer.hpp")
def _(t):
    t = t.replace(": BaseConvolutionLayer<Dtype>(param), handles_setup_(false) {}",
                  ": BaseConvolutionLayer<Dtype>(param), handles_setup_(false), escort_ws_(NULL), escort_ws_bytes_(0) {}")
    return insert_after(t, "  void **workspace;  // aliases into workspaceData\n",
                        "  void *escort_ws_;  // escort-b200: column buffer of the tcgen05 path (unused by 1x1 layers)\n  size_t escort_ws_bytes_;\n")


@edit("src/caffe/layers/esc_conv_layer.cpp")
def _(t):
    return insert_after(t, "EscConvolutionLayer<Dtype>::~EscConvolutionLayer() {\n", "  cudaFree(escort_ws_);  // escort-b200\n")


@edit("src/caffe/layers/esc_conv_layer.cu")
def _(t):
    return insert_after(t, "    Dtype* top_data = top[i]->mutable_gpu_data();\n", """    if (sizeof(Dtype) == sizeof(float) && this->group_ == 1) {
      // escort-b200: tcgen05 convolution, bias fused -- an implicit GEMM straight from NCHW for 1x1 / stride 1 layers,
      // a transposed column buffer for conv1-type layers -- instead of cuDNN IMPLICIT_GEMM + cudnnAddTensor
      escort_geom eg = {this->channels_, this->num_output_, 1, bottom[i]->height(), bottom[i]->width(),
          this->kernel_shape_.cpu_data()[0], this->kernel_shape_.cpu_data()[1],
          this->pad_.cpu_data()[0], this->pad_.cpu_data()[1], this->stride_.cpu_data()[0], this->stride_.cpu_data()[1],
          this->dilation_.cpu_data()[0], this->dilation_.cpu_data()[1]};
      const size_t need = escort_dense_conv_workspace_bytes(&eg, this->num_);
      if (need > escort_ws_bytes_) {
        cudaFree(escort_ws_);
        CUDA_CHECK(cudaMalloc(&escort_ws_, need));
        escort_ws_bytes_ = need;
      }
      ESCORT_CHECK(escort_dense_conv_forward(&eg, this->num_, reinterpret_cast<const float*>(bottom_data),
          reinterpret_cast<const float*>(weight),
          this->bias_term_ ? reinterpret_cast<const float*>(this->blobs_[1]->gpu_data()) : NULL, /*fuse_relu=*/0,
          escort_ws_, escort_ws_bytes_, reinterpret_cast<float*>(top_data), NULL));
      continue;
    }
""")


@edit("Makefile")
def _(t):
    return insert_after(t, "LIBRARIES += glog gflags protobuf boost_system boost_filesystem m hdf5_hl hdf5 spmp\n",
                        "# escort-b200: the sparse-conv hot path (set ESCORT_B200_DIR to the checkout; its include/ holds escort_b200.h)\n"
                        "LIBRARIES += escort_b200\nINCLUDE_DIRS += $(ESCORT_B200_DIR)/include\nLIBRARY_DIRS += $(ESCORT_B200_DIR)/caffe_escoin_b200\n")


def main():
    names = []
    for rel, fn in EDITS.items():
        old = open(os.path.join(REF, rel)).read()
        new = fn(old)
        assert new != old, rel
        diff = difflib.unified_diff(old.splitlines(True), new.splitlines(True), "a/" + rel, "b/" + rel, n=2)
        name = rel.replace("/", "__") + ".patch"
        open(os.path.join(HERE, name), "w").write("".join(diff))
        names.append(name)
    print("wrote", len(names), "patches:", " ".join(names))


if __name__ == "__main__":
    main()
