// pack.cu -- dense -> CSR weight packing, index stretch and the padded-input copy (hot-path rows a2-a4).
//
// Replaces the cuSPARSE pair cusparseSnnz + cusparseSdense2csc that caffe_gpu_sparse_dense2csr<float>
// calls (reference src/caffe/util/math_functions.cu:103-128; removed from CUDA 12) with a
// count -> scan -> ordered-scatter pipeline that is bit-exact with the reference's CPU conversion
// (src/caffe/util/math_functions.cpp:92-105).  HBM-bound integer/byte work: one warp per row, coalesced
// 128-byte row reads, ballot/popc compaction so the (row, ascending column) order is preserved.
#include "common.cuh"

namespace escort {

static constexpr int kPackWarps = 8;  // warps per CTA

// rowcnt[i] = #{j : A[i][j] != 0}.  The test is `!= 0` exactly as the reference: -0.0f dropped, NaN kept.
__global__ void __launch_bounds__(kPackWarps * 32) pack_count_kernel(int M, int N, const float *__restrict__ A,
                                                                     int *__restrict__ rowptr,
                                                                     int *__restrict__ nnz_per_row) {
  const int row = blockIdx.x * kPackWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float *a = A + (size_t)row * N;
  int cnt = 0;
  for (int j0 = 0; j0 < N; j0 += 32) {
    const int j = j0 + lane;
    const bool nz = (j < N) && (__ldg(a + j) != 0.0f);
    cnt += __popc(__ballot_sync(0xffffffffu, nz));
  }
  if (lane == 0) {
    rowptr[row + 1] = cnt;
    if (nnz_per_row) nnz_per_row[row] = cnt;
  }
}

// In-place inclusive scan of rowptr[1..M] (rowptr[0] = 0) by one CTA; M is a few thousand at most.
__global__ void __launch_bounds__(1024) pack_scan_kernel(int M, int *__restrict__ rowptr) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) {
    carry_s = 0;
    rowptr[0] = 0;
  }
  __syncthreads();
  for (int base = 0; base < M; base += 1024) {
    const int i = base + tid;
    int v = (i < M) ? rowptr[i + 1] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += t;
    }
    if (lane == 31) warp_sums[wid] = v;
    __syncthreads();
    if (wid == 0) {
      int s = warp_sums[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += t;
      }
      warp_sums[lane] = s;
    }
    __syncthreads();
    const int carry = carry_s;
    const int prefix = (wid > 0 ? warp_sums[wid - 1] : 0) + carry;
    if (i < M) rowptr[i + 1] = v + prefix;
    __syncthreads();
    if (tid == 1023) carry_s = v + prefix;
    __syncthreads();
  }
}

// Ordered scatter: values / colidx of row i go to [rowptr[i], rowptr[i+1]) in ascending column order.
__global__ void __launch_bounds__(kPackWarps * 32) pack_scatter_kernel(int M, int N, const float *__restrict__ A,
                                                                       const int *__restrict__ rowptr,
                                                                       float *__restrict__ values,
                                                                       int *__restrict__ colidx) {
  const int row = blockIdx.x * kPackWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float *a = A + (size_t)row * N;
  int off = rowptr[row];
  const unsigned lt = (1u << lane) - 1u;
  for (int j0 = 0; j0 < N; j0 += 32) {
    const int j = j0 + lane;
    const float v = (j < N) ? __ldg(a + j) : 0.0f;
    const bool nz = (j < N) && (v != 0.0f);
    const unsigned m = __ballot_sync(0xffffffffu, nz);
    if (nz) {
      const int p = off + __popc(m & lt);
      values[p] = v;
      colidx[p] = j;
    }
    off += __popc(m);
  }
}

// reference: stretch_kernel (math_functions.cu:706-719) runs one THREAD per row; here one thread per
// nonzero slot of the row range so the pass is a single coalesced read-modify-write of colidx.
__global__ void stretch_kernel(const int *__restrict__ rowptr, int *__restrict__ colidx, int M, int Hp, int Wp,
                               int kernel_h, int kernel_w) {
  const int begin = rowptr[0], end = rowptr[M];
  for (int j = begin + blockIdx.x * blockDim.x + threadIdx.x; j < end; j += gridDim.x * blockDim.x) {
    const int col = colidx[j];
    const int kc = col % kernel_w;
    const int kr = (col / kernel_w) % kernel_h;
    const int ic = col / (kernel_w * kernel_h);
    colidx[j] = (ic * Hp + kr) * Wp + kc;
  }
}

// reference: copy_input (math_functions.cu:729-749).  One thread per source element, x fastest.
__global__ void copy_input_kernel(float *__restrict__ dst, const float *__restrict__ src, int C, int H, int W,
                                  int pad_h, int pad_w) {
  const long total = (long)C * H * W;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long t = i / W;
    const int y = (int)(t % H);
    const int c = (int)(t / H);
    dst[((long)c * (H + pad_h) + y + pad_h) * (W + pad_w) + pad_w + x] = __ldg(src + i);
  }
}

}  // namespace escort

using namespace escort;

extern "C" int escort_pack_csr(int M, int N, const float *A, int *nnz_per_row, float *values, int *rowptr,
                               int *colidx, int *nnz_total_host, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(M > 0 && N > 0 && A && values && rowptr && colidx, "escort_pack_csr: bad arguments");
  const int blocks = ceil_div(M, kPackWarps);
  pack_count_kernel<<<blocks, kPackWarps * 32, 0, stream>>>(M, N, A, rowptr, nnz_per_row);
  ESCORT_LAUNCH_CHECK();
  pack_scan_kernel<<<1, 1024, 0, stream>>>(M, rowptr);
  ESCORT_LAUNCH_CHECK();
  pack_scatter_kernel<<<blocks, kPackWarps * 32, 0, stream>>>(M, N, A, rowptr, values, colidx);
  ESCORT_LAUNCH_CHECK();
  if (nnz_total_host) {
    ESCORT_CUDA(cudaMemcpyAsync(nnz_total_host, rowptr + M, sizeof(int), cudaMemcpyDeviceToHost, stream));
    ESCORT_CUDA(cudaStreamSynchronize(stream));
  }
  return 0;
}

extern "C" int escort_stretch(const int *rowptr, int *colidx, int M, int height, int width, int pad_h, int pad_w,
                              int kernel_h, int kernel_w, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(M > 0 && rowptr && colidx && kernel_h > 0 && kernel_w > 0, "escort_stretch: bad arguments");
  stretch_kernel<<<148, 256, 0, stream>>>(rowptr, colidx, M, height + pad_h, width + pad_w, kernel_h, kernel_w);
  ESCORT_LAUNCH_CHECK();
  return 0;
}

extern "C" int escort_copy_input(float *dst, const float *src, int num_channels, int height, int width, int pad_h,
                                 int pad_w, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(dst && src && num_channels > 0 && height > 0 && width > 0, "escort_copy_input: bad arguments");
  const long total = (long)num_channels * height * width;
  const int blocks = (int)std::min<long>((total + 255) / 256, 148L * 8);
  copy_input_kernel<<<blocks, 256, 0, stream>>>(dst, src, num_channels, height, width, pad_h, pad_w);
  ESCORT_LAUNCH_CHECK();
  return 0;
}
