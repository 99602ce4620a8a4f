// escort_dense_layers.hpp -- C++ host mirror of the two layers the reference keeps DENSE, written
// above the C ABI like escort_conv_layer.hpp:
//
//   InnerProductLayer::LayerSetUp / Reshape / Forward_gpu   src/caffe/layers/inner_product_layer.cpp:10-78,
//                                                           src/caffe/layers/inner_product_layer.cu:9-31
//   EscConvolutionLayer::Forward_gpu                        src/caffe/layers/esc_conv_layer.cu:11-46 (cuDNN IMPLICIT_GEMM)
//
// Same blob layout as the reference (InnerProduct: blobs_[0] = N_ x K_, transpose: false; convolution: blobs_[0] =
// num_output x channels x kh x kw), bias fused.  TF32 products, fp32 accumulation: a host that needs bit-level fp32 keeps
// the reference's cuBLAS / cuDNN call.  Backward of these layers stays Caffe's own code.
#pragma once
#include "escort_conv_layer.hpp"

namespace escort_caffe {

struct InnerProductParameter {  // src/caffe/proto/caffe.proto InnerProductParameter
  int num_output = 0;
  bool bias_term = true;
  int axis = 1;
  bool transpose = false;  // only false is supported here (the reference's default)
};

class InnerProductLayer {
 public:
  explicit InnerProductLayer(const InnerProductParameter &p, bool fuse_relu = false) : param_(p), fuse_relu_(fuse_relu) {}
  const char *type() const { return "InnerProduct"; }
  // inner_product_layer.cpp:10-58: K_ = count from `axis`, N_ = num_output, weight shape {N_, K_}
  void LayerSetUp(const std::vector<int> &bottom_shape) {
    if (param_.transpose) throw std::runtime_error("InnerProductLayer: transpose is not supported on the tcgen05 path");
    if (param_.axis < 0 || param_.axis >= (int)bottom_shape.size()) throw std::runtime_error("InnerProductLayer: bad axis");
    N_ = param_.num_output;
    K_ = 1;
    for (size_t i = param_.axis; i < bottom_shape.size(); ++i) K_ *= bottom_shape[i];
    blobs_.resize(param_.bias_term ? 2 : 1);
    blobs_[0].Reshape({N_, K_});
    if (param_.bias_term) blobs_[1].Reshape({N_});
    Reshape(bottom_shape);
  }
  void Reshape(const std::vector<int> &bottom_shape) {  // inner_product_layer.cpp:60-78
    int k = 1;
    M_ = 1;
    for (int i = 0; i < param_.axis; ++i) M_ *= bottom_shape[i];
    for (size_t i = param_.axis; i < bottom_shape.size(); ++i) k *= bottom_shape[i];
    if (k != K_) throw std::runtime_error("InnerProductLayer: input size incompatible with inner product parameters");
  }
  std::vector<int> top_shape() const { return {M_, N_}; }
  void Forward_gpu(const Blob &bottom, Blob &top) {
    check(escort_inner_product_forward(M_, K_, N_, bottom.gpu_data(), blobs_[0].gpu_data(),
                                       param_.bias_term ? blobs_[1].gpu_data() : nullptr, fuse_relu_ ? 1 : 0, top.mutable_gpu_data(),
                                       nullptr),
          "escort_inner_product_forward");
  }
  std::vector<Blob> &blobs() { return blobs_; }

 private:
  InnerProductParameter param_;
  bool fuse_relu_;
  int M_ = 0, K_ = 0, N_ = 0;
  std::vector<Blob> blobs_;
};

// type "EscConvolution": the dense convolution of conv1 / 1x1 / unpruned layers (group 1)
class EscConvolutionLayer {
 public:
  explicit EscConvolutionLayer(const ConvolutionParameter &p, bool fuse_relu = false) : param_(p), fuse_relu_(fuse_relu) {}
  ~EscConvolutionLayer() { cudaFree(workspace_); }
  const char *type() const { return "EscConvolution"; }
  void LayerSetUp(const std::vector<int> &bottom_shape) {
    if (bottom_shape.size() != 4) throw std::runtime_error("EscConvolutionLayer: 4-D bottom expected");
    if (param_.group != 1) throw std::runtime_error("EscConvolutionLayer: group > 1 is not supported on the tcgen05 path");
    const ConvolutionParameter &p = param_;
    geom_ = escort_geom{bottom_shape[1], p.num_output, 1, bottom_shape[2], bottom_shape[3], p.kernel_h, p.kernel_w,
                        p.pad_h, p.pad_w, p.stride_h, p.stride_w, p.dilation_h, p.dilation_w};
    blobs_.resize(p.bias_term ? 2 : 1);
    blobs_[0].Reshape({p.num_output, bottom_shape[1], p.kernel_h, p.kernel_w});
    if (p.bias_term) blobs_[1].Reshape({p.num_output});
    Reshape(bottom_shape);
  }
  void Reshape(const std::vector<int> &bottom_shape) {
    num_ = bottom_shape[0];
    const ConvolutionParameter &p = param_;
    height_out_ = (geom_.height + 2 * p.pad_h - (p.dilation_h * (p.kernel_h - 1) + 1)) / p.stride_h + 1;
    width_out_ = (geom_.width + 2 * p.pad_w - (p.dilation_w * (p.kernel_w - 1) + 1)) / p.stride_w + 1;
    const size_t need = escort_dense_conv_workspace_bytes(&geom_, num_);  // the cuDNN workspace's role (esc_conv_layer.cpp Reshape)
    if (need > workspace_bytes_) {
      cudaFree(workspace_);
      workspace_ = nullptr;
      cuda_check(cudaMalloc(&workspace_, need), "cudaMalloc(workspace)");
      workspace_bytes_ = need;
    }
  }
  std::vector<int> top_shape() const { return {num_, param_.num_output, height_out_, width_out_}; }
  // `residual`: the other bottom of an Eltwise SUM that follows the layer (NULL = the reference's plain forward)
  void Forward_gpu(const Blob &bottom, Blob &top, const Blob *residual = nullptr) {
    check(escort_dense_conv_forward_residual(&geom_, bottom.shape[0], bottom.gpu_data(), blobs_[0].gpu_data(),
                                             param_.bias_term ? blobs_[1].gpu_data() : nullptr,
                                             residual ? residual->gpu_data() : nullptr, fuse_relu_ ? 1 : 0, workspace_,
                                             workspace_bytes_, top.mutable_gpu_data(), nullptr),
          "escort_dense_conv_forward");
  }
  std::vector<Blob> &blobs() { return blobs_; }

 private:
  ConvolutionParameter param_;
  bool fuse_relu_;
  escort_geom geom_{};
  int num_ = 0, height_out_ = 0, width_out_ = 0;
  std::vector<Blob> blobs_;
  void *workspace_ = nullptr;
  size_t workspace_bytes_ = 0;
};

}  // namespace escort_caffe
