// escort_conv_layer.hpp -- C++ host mirror of the reference's operator surface for the Escort hot path, written
// above the C ABI (include/escort_b200.h).  Same names, argument meaning and call order as the reference:
//
//   Caffe::set_conv_mode / conv_mode            include/caffe/common.hpp:112,154,161 ; tools/caffe.cpp:58-60,293-301
//   ConvolutionLayer::LayerSetUp / Reshape      src/caffe/layers/base_conv_layer.cpp:276-446, 449-530
//   ConvolutionLayer::WeightAlign               src/caffe/layers/base_conv_layer.cpp:46-273 (GPU branch :236-264)
//   ConvolutionLayer::Forward_gpu               src/caffe/layers/conv_layer.cu:8-40  (+ conv_relu_layer.cu:8-30)
//   ConvolutionLayer::Backward_gpu              src/caffe/layers/conv_layer.cu:43-73
//
// Caffe itself cannot be built in this image (no glog / gflags / boost / protobuf / BLAS), so this header stands in for
// the layer in tests and examples; INTEGRATION.md shows the overlay for the real sources.  Blobs are caller-visible
// device buffers (cudaMalloc), like Blob::gpu_data() / mutable_gpu_diff().  Errors: the reference's CHECK_* abort the
// process (include/caffe/util/device_alternate.hpp:60-77); the mirror throws std::runtime_error with the shim's message.
// There is no CPU fallback and no dense (LOWERED_GEMM) path here: those modes stay Caffe's own cuBLAS code.
#pragma once
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "escort_b200.h"

namespace escort_caffe {

class Caffe {
 public:
  enum ConvMode { LOWERED_GEMM = 0, LOWERED_SPARSE = 1, SCONV = 2, SCONV_PAR = 3 };  // common.hpp:112
  static ConvMode conv_mode() { return mode_(); }
  static void set_conv_mode(ConvMode m) { mode_() = m; }

 private:
  static ConvMode &mode_() {
    static thread_local ConvMode m = LOWERED_GEMM;  // the reference leaves it uninitialised (common.cpp:107-110)
    return m;
  }
};

inline void check(int rc, const char *what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + " failed (rc=" + std::to_string(rc) + "): " + escort_last_error());
}
inline void cuda_check(cudaError_t e, const char *what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// Minimal Blob: shape + device data/diff (SyncedMemory's device half, src/caffe/syncedmem.cpp:66-80).
struct Blob {
  std::vector<int> shape;
  float *data = nullptr, *diff = nullptr;
  size_t count() const {
    size_t c = 1;
    for (int d : shape) c *= (size_t)d;
    return shape.empty() ? 0 : c;
  }
  void Reshape(const std::vector<int> &s) {
    Free();
    shape = s;
    if (count()) {
      cuda_check(cudaMalloc((void **)&data, count() * sizeof(float)), "Blob::Reshape(data)");
      cuda_check(cudaMalloc((void **)&diff, count() * sizeof(float)), "Blob::Reshape(diff)");
      cuda_check(cudaMemset(data, 0, count() * sizeof(float)), "Blob memset");
      cuda_check(cudaMemset(diff, 0, count() * sizeof(float)), "Blob memset");
    }
  }
  void Free() {
    cudaFree(data);
    cudaFree(diff);
    data = diff = nullptr;
  }
  const float *gpu_data() const { return data; }
  float *mutable_gpu_data() { return data; }
  const float *gpu_diff() const { return diff; }
  float *mutable_gpu_diff() { return diff; }
  ~Blob() { Free(); }
  Blob() = default;
  Blob(const Blob &) = delete;
  Blob &operator=(const Blob &) = delete;
  Blob(Blob &&o) noexcept : shape(std::move(o.shape)), data(o.data), diff(o.diff) { o.data = o.diff = nullptr; }
};

struct ConvolutionParameter {  // src/caffe/proto/caffe.proto ConvolutionParameter, the fields the path reads
  int num_output = 0;
  bool bias_term = true;
  int kernel_h = 0, kernel_w = 0, stride_h = 1, stride_w = 1, pad_h = 0, pad_w = 0, dilation_h = 1, dilation_w = 1;
  int group = 1;
};

class ConvolutionLayer {
 public:
  // type "Convolution" (fuse_relu = false) or "ConvolutionReLU" (conv_relu_layer.cu:66)
  explicit ConvolutionLayer(const ConvolutionParameter &p, bool fuse_relu = false) : param_(p), fuse_relu_(fuse_relu) {}
  ~ConvolutionLayer() { escort_plan_destroy(plan_); }
  const char *type() const { return fuse_relu_ ? "ConvolutionReLU" : "Convolution"; }

  // bottom shape {num, channels, height, width}; allocates blobs_[0] (weights) and blobs_[1] (bias) and the CSR blobs,
  // sized to the dense count like base_conv_layer.cpp:509-513
  void LayerSetUp(const std::vector<int> &bottom_shape) {
    if (bottom_shape.size() != 4) throw std::runtime_error("ConvolutionLayer: 4-D bottom expected");
    const ConvolutionParameter &p = param_;
    channels_ = bottom_shape[1];
    if (p.group < 1 || channels_ % p.group || p.num_output % p.group)
      throw std::runtime_error("ConvolutionLayer: channels and num_output must be multiples of group");
    geom_ = escort_geom{channels_, p.num_output, p.group, bottom_shape[2], bottom_shape[3], p.kernel_h, p.kernel_w,
                        p.pad_h, p.pad_w, p.stride_h, p.stride_w, p.dilation_h, p.dilation_w};
    blobs_.resize(p.bias_term ? 2 : 1);
    blobs_[0].Reshape({p.num_output, channels_ / p.group, p.kernel_h, p.kernel_w});
    if (p.bias_term) blobs_[1].Reshape({p.num_output});
    param_propagate_down_.assign(blobs_.size(), true);
    const int dense = (int)blobs_[0].count();
    nz_weight_values_.Reshape({dense});
    cuda_check(cudaMalloc((void **)&nz_weight_indices_, sizeof(int) * dense), "cudaMalloc(indices)");
    cuda_check(cudaMalloc((void **)&nz_weight_index_pointers_, sizeof(int) * (p.num_output + p.group)), "cudaMalloc(rowptr)");
    cuda_check(cudaMalloc((void **)&nz_per_row_, sizeof(int) * p.num_output), "cudaMalloc(nz_per_row)");
    nz_num_.assign(p.group, 0);
    Reshape(bottom_shape);
  }
  void Reshape(const std::vector<int> &bottom_shape) {  // compute_output_shape, conv_layer.cpp:8-22
    num_ = bottom_shape[0];
    const ConvolutionParameter &p = param_;
    height_out_ = (geom_.height + 2 * p.pad_h - (p.dilation_h * (p.kernel_h - 1) + 1)) / p.stride_h + 1;
    width_out_ = (geom_.width + 2 * p.pad_w - (p.dilation_w * (p.kernel_w - 1) + 1)) / p.stride_w + 1;
  }
  std::vector<int> top_shape() const { return {num_, param_.num_output, height_out_, width_out_}; }

  // Dense -> CSR per group (bit-exact with caffe_cpu_sparse_dense2csr), stretch, then the plan the forward executes.
  // Called where Net::CopyTrainedLayersFrom calls it (src/caffe/net.cpp:819), i.e. after the weights are in blobs_[0].
  bool tune_backward_ = false;  // training hosts (phase TRAIN) also tune the backward-data kernel
  void WeightAlign(int tune_batch = 0) {
    const ConvolutionParameter &p = param_;
    const int M = p.num_output / p.group, N = (channels_ / p.group) * p.kernel_h * p.kernel_w;
    const int weight_offset = M * N, row_offset = M + 1;
    for (int g = 0; g < p.group; ++g) {
      check(escort_pack_csr(M, N, blobs_[0].gpu_data() + weight_offset * g, nz_per_row_ + M * g,
                            nz_weight_values_.mutable_gpu_data() + weight_offset * g, nz_weight_index_pointers_ + row_offset * g,
                            nz_weight_indices_ + weight_offset * g, &nz_num_[g], nullptr),
            "escort_pack_csr");
      check(escort_stretch(nz_weight_index_pointers_ + row_offset * g, nz_weight_indices_ + weight_offset * g, M, geom_.height,
                           geom_.width, p.pad_h, p.pad_w, p.kernel_h, p.kernel_w, nullptr),
            "escort_stretch");
    }
    escort_plan_destroy(plan_);
    plan_ = nullptr;
    check(escort_plan_create(&geom_, nz_weight_index_pointers_, nz_weight_indices_, nz_weight_values_.gpu_data(), 1, &plan_,
                             nullptr),
          "escort_plan_create");
    if (tune_batch > 0) check(escort_plan_autotune(plan_, tune_batch, nullptr), "escort_plan_autotune");
    if (tune_batch > 0 && tune_backward_)
      check(escort_plan_autotune_backward(plan_, tune_batch, nullptr), "escort_plan_autotune_backward");
  }

  // conv_layer.cu:8-40: in SCONV / SCONV_PAR mode the whole batch, all groups, bias (and ReLU for ConvolutionReLU).
  void Forward_gpu(const Blob &bottom, Blob &top) {
    const Caffe::ConvMode m = Caffe::conv_mode();
    if (m != Caffe::SCONV && m != Caffe::SCONV_PAR)
      throw std::runtime_error("ConvolutionLayer::Forward_gpu: conv_mode must be SCONV or SCONV_PAR on this path "
                               "(LOWERED_GEMM / LOWERED_SPARSE are Caffe's own cuBLAS / cuSPARSE code)");
    if (!plan_) throw std::runtime_error("ConvolutionLayer::Forward_gpu: WeightAlign() has not run");
    check(escort_sconv_forward(plan_, bottom.shape[0], bottom.gpu_data(), param_.bias_term ? blobs_[1].gpu_data() : nullptr,
                               fuse_relu_ ? 1 : 0, top.mutable_gpu_data(), nullptr),
          "escort_sconv_forward");
  }

  // conv_layer.cu:43-73 with the gradient restricted to the sparsity mask: parameter diffs ACCUMULATE into
  // blobs_[i].diff, bottom diff is OVERWRITTEN; honours param_propagate_down_ and propagate_down.
  void Backward_gpu(const Blob &top, bool propagate_down, Blob &bottom) {
    if (!plan_) throw std::runtime_error("ConvolutionLayer::Backward_gpu: WeightAlign() has not run");
    const int num = top.shape[0];
    if (param_.bias_term && param_propagate_down_[1])
      check(escort_bias_backward(num, param_.num_output, height_out_ * width_out_, top.gpu_diff(), blobs_[1].mutable_gpu_diff(),
                                 nullptr),
            "escort_bias_backward");
    if (param_propagate_down_[0])
      check(escort_sconv_backward_weight(plan_, num, bottom.gpu_data(), top.gpu_diff(), blobs_[0].mutable_gpu_diff(), nullptr, 1,
                                         nullptr),
            "escort_sconv_backward_weight");
    if (propagate_down)
      check(escort_sconv_backward_data(plan_, num, top.gpu_diff(), bottom.mutable_gpu_diff(), nullptr),
            "escort_sconv_backward_data");
  }

  // after a solver update of blobs_[0]: keep the CSR snapshot coherent (positions fixed by the mask)
  void RefreshValues() { check(escort_refresh_values(plan_, blobs_[0].gpu_data(), nz_weight_values_.mutable_gpu_data(), nullptr), "escort_refresh_values"); }

  std::vector<Blob> &blobs() { return blobs_; }
  std::vector<bool> &param_propagate_down() { return param_propagate_down_; }
  const std::vector<int> &nz_num() const { return nz_num_; }
  const int *nz_weight_indices() const { return nz_weight_indices_; }
  const int *nz_weight_index_pointers() const { return nz_weight_index_pointers_; }
  const Blob &nz_weight_values() const { return nz_weight_values_; }
  const char *kernel_name() const { return escort_plan_kernel_name(plan_); }

 private:
  ConvolutionParameter param_;
  bool fuse_relu_;
  escort_geom geom_{};
  int channels_ = 0, num_ = 0, height_out_ = 0, width_out_ = 0;
  std::vector<Blob> blobs_;
  std::vector<bool> param_propagate_down_;
  // base_conv_layer.hpp:184-192
  Blob nz_weight_values_;
  int *nz_weight_indices_ = nullptr, *nz_weight_index_pointers_ = nullptr, *nz_per_row_ = nullptr;
  std::vector<int> nz_num_;
  escort_plan *plan_ = nullptr;
};

}  // namespace escort_caffe
