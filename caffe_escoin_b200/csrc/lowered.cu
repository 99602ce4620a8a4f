// lowered.cu -- the reference's second baseline, conv_mode LOWERED_SPARSE: im2col + CSR x dense,
// BaseConvolutionLayer::forward_gpu_gemm (src/caffe/layers/base_conv_layer.cpp:715-745) with
// caffe_gpu_sparse_csrmm = cusparseScsrmm2 + a cublasSgeam transpose (src/caffe/util/math_functions.cu:48-62).  csrmm2 was
// removed from CUDA 12; this is the same computation on cusparseSpMM (CSR x row-major dense -> row-major dense, so the
// reference's transposed output buffer and geam disappear).  A COMPARATOR: it exists so that the direct sparse
// convolution can be timed against "cuSPARSE" on B200 the way the paper does, not as a product path.
#include <cusparse.h>
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace escort {
namespace {

struct CusparseApi {
  void *lib = nullptr;
  cusparseStatus_t (*Create)(cusparseHandle_t *) = nullptr;
  cusparseStatus_t (*Destroy)(cusparseHandle_t) = nullptr;
  cusparseStatus_t (*SetStream)(cusparseHandle_t, cudaStream_t) = nullptr;
  cusparseStatus_t (*CreateCsr)(cusparseSpMatDescr_t *, int64_t, int64_t, int64_t, void *, void *, void *, cusparseIndexType_t,
                                cusparseIndexType_t, cusparseIndexBase_t, cudaDataType) = nullptr;
  cusparseStatus_t (*CreateDnMat)(cusparseDnMatDescr_t *, int64_t, int64_t, int64_t, void *, cudaDataType, cusparseOrder_t) = nullptr;
  cusparseStatus_t (*DestroySpMat)(cusparseConstSpMatDescr_t) = nullptr;
  cusparseStatus_t (*DestroyDnMat)(cusparseConstDnMatDescr_t) = nullptr;
  cusparseStatus_t (*SpMM_bufferSize)(cusparseHandle_t, cusparseOperation_t, cusparseOperation_t, const void *, cusparseConstSpMatDescr_t,
                                      cusparseConstDnMatDescr_t, const void *, cusparseDnMatDescr_t, cudaDataType, cusparseSpMMAlg_t,
                                      size_t *) = nullptr;
  cusparseStatus_t (*SpMM)(cusparseHandle_t, cusparseOperation_t, cusparseOperation_t, const void *, cusparseConstSpMatDescr_t,
                           cusparseConstDnMatDescr_t, const void *, cusparseDnMatDescr_t, cudaDataType, cusparseSpMMAlg_t, void *) = nullptr;
  bool ok = false;
};

const CusparseApi &cusparse_api() {
  static const CusparseApi api = [] {
    CusparseApi a;
    for (const char *name : {"libcusparse.so.12", "libcusparse.so", "/usr/local/cuda/lib64/libcusparse.so.12"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) return a;
#define ESCORT_SYM(field, sym) *(void **)(&a.field) = dlsym(a.lib, sym)
    ESCORT_SYM(Create, "cusparseCreate");
    ESCORT_SYM(Destroy, "cusparseDestroy");
    ESCORT_SYM(SetStream, "cusparseSetStream");
    ESCORT_SYM(CreateCsr, "cusparseCreateCsr");
    ESCORT_SYM(CreateDnMat, "cusparseCreateDnMat");
    ESCORT_SYM(DestroySpMat, "cusparseDestroySpMat");
    ESCORT_SYM(DestroyDnMat, "cusparseDestroyDnMat");
    ESCORT_SYM(SpMM_bufferSize, "cusparseSpMM_bufferSize");
    ESCORT_SYM(SpMM, "cusparseSpMM");
#undef ESCORT_SYM
    a.ok = a.Create && a.Destroy && a.SetStream && a.CreateCsr && a.CreateDnMat && a.DestroySpMat && a.DestroyDnMat && a.SpMM_bufferSize && a.SpMM;
    return a;
  }();
  return api;
}

// col[(c * kh + r) * kw + s][oy * Wo + ox] = in[c][oy * stride - pad + r * dil][ox * stride - pad + s * dil] (0 outside):
// the layout of caffe's im2col_gpu (src/caffe/util/im2col.cu), one thread per column-buffer element.
__global__ void im2col_kernel(long total, const float *__restrict__ in, int H, int W, int kh, int kw, int pad_h, int pad_w, int stride_h,
                              int stride_w, int dil_h, int dil_w, int Ho, int Wo, float *__restrict__ col) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int N = Ho * Wo;
  const int n = (int)(e % N);
  const int k = (int)(e / N);
  const int s = k % kw, r = (k / kw) % kh, c = k / (kw * kh);
  const int oy = n / Wo, ox = n - oy * Wo;
  const int y = oy * stride_h - pad_h + r * dil_h, x = ox * stride_w - pad_w + s * dil_w;
  col[e] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(in + ((size_t)c * H + y) * W + x) : 0.f;
}

__global__ void bias_relu_kernel(long total, int N, int M, const float *__restrict__ bias, int relu, float *__restrict__ y) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  float v = y[e];
  if (bias) v += __ldg(bias + (e / N) % M);
  if (relu) v = fmaxf(v, 0.f);
  y[e] = v;
}

}  // namespace
}  // namespace escort

using namespace escort;

extern "C" ESCORT_API int escort_lowered_sparse_forward(const escort_geom *g, int num, const float *bottom, const int *rowptr,
                                                        const int *colidx_raw, const float *values, const float *bias, int fuse_relu,
                                                        float *col_buffer, float *top, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(g && num >= 0 && bottom && rowptr && colidx_raw && values && col_buffer && top, "escort_lowered_sparse_forward: null argument");
  ESCORT_REQUIRE(g->group > 0 && g->channels % g->group == 0 && g->num_output % g->group == 0, "escort_lowered_sparse_forward: bad group");
  const CusparseApi &cs = cusparse_api();
  if (!cs.ok) {
    set_last_error("escort_lowered_sparse_forward: libcusparse.so.12 not found (the comparator needs cuSPARSE; the sparse-conv path does not)");
    return ESCORT_EINVAL;
  }
  const int Ho = (g->height + 2 * g->pad_h - (g->dilation_h * (g->kernel_h - 1) + 1)) / g->stride_h + 1;
  const int Wo = (g->width + 2 * g->pad_w - (g->dilation_w * (g->kernel_w - 1) + 1)) / g->stride_w + 1;
  const int N = Ho * Wo, Cg = g->channels / g->group, Mg = g->num_output / g->group, K = Cg * g->kernel_h * g->kernel_w;
  const long weight_offset = (long)Mg * K;          // base_conv_layer.cpp: weight_offset_ = conv_out_channels_ * kernel_dim_ / group_
  const long col_total = (long)g->group * K * N;    // col_offset_ * group_
  if (num == 0 || N <= 0) return 0;

  static thread_local cusparseHandle_t handle = nullptr;
  static thread_local void *workspace = nullptr;
  static thread_local size_t workspace_bytes = 0;
  if (!handle && cs.Create(&handle) != CUSPARSE_STATUS_SUCCESS) {
    set_last_error("escort_lowered_sparse_forward: cusparseCreate failed");
    return ESCORT_EINVAL;
  }
  cs.SetStream(handle, stream);
  // nnz per group = rowptr[Mg] of the group's block (host copy of G integers)
  std::vector<int> nnz(g->group);
  for (int gi = 0; gi < g->group; ++gi)
    ESCORT_CUDA(cudaMemcpyAsync(&nnz[gi], rowptr + (size_t)(Mg + 1) * gi + Mg, sizeof(int), cudaMemcpyDeviceToHost, stream));
  ESCORT_CUDA(cudaStreamSynchronize(stream));
  const float one = 1.f, zero = 0.f;
  const size_t in_image = (size_t)g->channels * g->height * g->width, out_image = (size_t)g->num_output * N;
  for (int n = 0; n < num; ++n) {  // the reference's per-image loop (conv_layer.cu:15-26)
    im2col_kernel<<<(unsigned)((col_total + 255) / 256), 256, 0, stream>>>(col_total, bottom + n * in_image, g->height, g->width, g->kernel_h,
                                                                           g->kernel_w, g->pad_h, g->pad_w, g->stride_h, g->stride_w,
                                                                           g->dilation_h, g->dilation_w, Ho, Wo, col_buffer);
    ESCORT_LAUNCH_CHECK();
    for (int gi = 0; gi < g->group; ++gi) {
      float *C = top + n * out_image + (size_t)gi * Mg * N;
      if (nnz[gi] == 0) {
        ESCORT_CUDA(cudaMemsetAsync(C, 0, (size_t)Mg * N * sizeof(float), stream));
        continue;
      }
      cusparseSpMatDescr_t A = nullptr;
      cusparseDnMatDescr_t B = nullptr, Cd = nullptr;
      bool ok = cs.CreateCsr(&A, Mg, K, nnz[gi], (void *)(rowptr + (size_t)(Mg + 1) * gi), (void *)(colidx_raw + weight_offset * gi),
                             (void *)(values + weight_offset * gi), CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO,
                             CUDA_R_32F) == CUSPARSE_STATUS_SUCCESS;
      ok = ok && cs.CreateDnMat(&B, K, N, N, (void *)(col_buffer + (size_t)gi * K * N), CUDA_R_32F, CUSPARSE_ORDER_ROW) == CUSPARSE_STATUS_SUCCESS;
      ok = ok && cs.CreateDnMat(&Cd, Mg, N, N, (void *)C, CUDA_R_32F, CUSPARSE_ORDER_ROW) == CUSPARSE_STATUS_SUCCESS;
      size_t need = 0;
      ok = ok && cs.SpMM_bufferSize(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, A, B, &zero, Cd, CUDA_R_32F,
                                    CUSPARSE_SPMM_CSR_ALG2, &need) == CUSPARSE_STATUS_SUCCESS;
      if (ok && need > workspace_bytes) {
        ESCORT_CUDA(cudaStreamSynchronize(stream));
        cudaFree(workspace);
        workspace = nullptr;
        workspace_bytes = 0;
        ok = cudaMalloc(&workspace, need) == cudaSuccess;
        if (ok) workspace_bytes = need;
      }
      ok = ok && cs.SpMM(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, A, B, &zero, Cd, CUDA_R_32F,
                         CUSPARSE_SPMM_CSR_ALG2, workspace) == CUSPARSE_STATUS_SUCCESS;
      if (A) cs.DestroySpMat(A);
      if (B) cs.DestroyDnMat(B);
      if (Cd) cs.DestroyDnMat(Cd);
      if (!ok) {
        set_last_error("escort_lowered_sparse_forward: cusparseSpMM failed");
        return ESCORT_EINVAL;
      }
    }
  }
  if (bias || fuse_relu) {  // forward_gpu_bias (base_conv_layer.cpp:747-753) + the ReLU layer that follows, one pass
    const long total = (long)num * out_image;
    bias_relu_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(total, N, g->num_output, bias, fuse_relu, top);
    ESCORT_LAUNCH_CHECK();
  }
  return 0;
}
