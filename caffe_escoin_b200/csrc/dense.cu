// dense.cu -- SURVEY section 8 (f1): the layers the reference keeps DENSE, on the 5th-generation tensor cores.
//
// The reference runs conv1 / 1x1 / any unpruned convolution through EscConvolutionLayer (cuDNN IMPLICIT_GEMM,
// src/caffe/layers/esc_conv_layer.cu:21-29) and the fully connected layers through InnerProductLayer (cuBLAS sgemm,
// src/caffe/layers/inner_product_layer.cu:9-31).  Both are D[i][j] = sum_k A[i][k] * B[j][k] with K contiguous in both
// operands ("TN"):  FC: A = bottom [num x K], B = weight [num_output x K];  conv: A = weight [M x K], B = the transposed
// column buffer [(image, pixel) x K] written by im2colT_kernel.  One kernel does both:
//   * TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) stages 128 x 32-float tiles of A and B into a 6-deep mbarrier ring;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = 128, K = 8; fp32 bits read as TF32,
//     fp32 accumulate) from shared-memory descriptors into a 128-lane x 128-column TMEM accumulator and releases each
//     stage with tcgen05.commit;
//   * four epilogue warps (one per TMEM lane quadrant) read the accumulator with tcgen05.ld, add the bias, apply ReLU and
//     store.
// Precision: TF32 products (10-bit mantissa), fp32 accumulation -- about 5e-4 relative L2 against an fp32 GEMM; the
// sparse path's 1e-4 bar does not apply here ("bf16/TF32 questions live here"), the tests state 2e-3.
// First cut: the conv path materialises the column buffer (not an implicit GEMM yet), group = 1.
#include <cuda.h>

#include <algorithm>
#include <mutex>
#include <string>

#include "common.cuh"

namespace escort {
namespace {

constexpr int kBM = 128, kBN = 128, kBK = 32;  // tile: rows of A, rows of B, floats of K per stage (128 bytes = one swizzle row)
constexpr int kStages = 6;
constexpr int kStageBytes = (kBM + kBN) * kBK * 4;  // 32 KiB
constexpr int kDenseThreads = 192;                  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr int kDenseSmem = kStages * kStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;

typedef CUresult (*TmaEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);
TmaEncodeFn dense_tma_encoder() {
  static const TmaEncodeFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    TmaEncodeFn f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      f = (TmaEncodeFn)p;
    cudaGetLastError();
    return f;
  }();
  return fn;
}

__device__ __forceinline__ void dn_mbar_init(unsigned addr, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
// bounded wait (a protocol bug must not hang the GPU): ~2 s, then trap
__device__ __forceinline__ void dn_mbar_wait(unsigned addr, unsigned parity) {
  unsigned ok = 0;
  unsigned long long t0 = 0;
  for (;;) {
    for (int i = 0; i < 256 && !ok; ++i)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok)
                   : "r"(addr), "r"(parity)
                   : "memory");
    if (ok) return;
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t0 == 0) t0 = t1;
    if (t1 - t0 > 2000000000ull) __trap();
  }
}

// shared-memory matrix descriptor of a K-major tile in the 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes
// apart (SBO), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B; a K step of 8 floats advances the start by 32 B
__device__ __forceinline__ uint64_t dn_smem_desc(unsigned saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // leading byte offset: unused for swizzled K-major tiles
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;            // descriptor version
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}

struct DenseParams {
  int rows_a, rows_b, K;     // D is rows_a x rows_b
  int mode;                  // 0: out[i * ldo + j], bias[j] (inner product)   1: out[(j / HW) * rows_a * HW + i * HW + j % HW], bias[i] (conv)
  int ldo, HW;
  int fuse_relu;
  const float *bias;
  float *out;
};

__global__ void __launch_bounds__(kDenseThreads, 1)
    dense_gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const DenseParams p) {
  extern __shared__ unsigned char dsm_raw[];
  const unsigned raw_addr = (unsigned)__cvta_generic_to_shared(dsm_raw);
  const unsigned base = (raw_addr + 1023u) & ~1023u;  // the swizzle atom (8 rows x 128 B) needs 1024-byte alignment
  const unsigned bars = base + kStages * kStageBytes;  // full[kStages] | empty[kStages] | tmem_full | tmem base slot
  const unsigned full0 = bars, empty0 = bars + 8 * kStages, tmem_full = bars + 16 * kStages, tslot = tmem_full + 8;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN;
  const int nkb = (p.K + kBK - 1) / kBK;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      dn_mbar_init(full0 + 8 * s, 1);
      dn_mbar_init(empty0 + 8 * s, 1);
    }
    dn_mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(tslot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  unsigned tbase;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tbase) : "r"(tslot));

  if (wid == 0) {
    if (lane == 0) {  // ---- TMA producer ----
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        dn_mbar_wait(empty0 + 8 * s, (unsigned)(((kb / kStages) & 1) ^ 1));  // a fresh barrier passes the first round
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + 8 * s), "r"((unsigned)kStageBytes) : "memory");
        const unsigned sa = base + s * kStageBytes, sb = sa + kBM * kBK * 4;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(sa),
                     "l"(&tmap_a), "r"(kb * kBK), "r"(m0), "r"(full0 + 8 * s)
                     : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(sb),
                     "l"(&tmap_b), "r"(kb * kBK), "r"(n0), "r"(full0 + 8 * s)
                     : "memory");
      }
    }
  } else if (wid == 1) {
    if (lane == 0) {  // ---- MMA issuer ----
      // instruction descriptor (kind::tf32): D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major,
      // N >> 3 at bits 17-22, M >> 4 at bits 24-28
      const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(kBN >> 3) << 17) | ((unsigned)(kBM >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        dn_mbar_wait(full0 + 8 * s, (unsigned)((kb / kStages) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned sa = base + s * kStageBytes, sb = sa + kBM * kBK * 4;
        const uint64_t da = dn_smem_desc(sa), db = dn_smem_desc(sb);
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k) {
          const unsigned acc = (kb > 0 || k > 0) ? 1u : 0u;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase),
                       "l"(da + (uint64_t)(2 * k)), "l"(db + (uint64_t)(2 * k)), "r"(idesc), "r"(acc)
                       : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + 8 * s) : "memory");  // frees the stage when the MMAs retire
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tmem_full) : "memory");
    }
  } else {
    // ---- epilogue: warp w owns TMEM lanes 32 * (w % 4) .. + 31 = rows m0 + 32 * (w % 4) + lane ----
    const int q = wid & 3;
    dn_mbar_wait(tmem_full, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int i = m0 + 32 * q + lane;
    const float bias_i = (p.mode == 1 && p.bias && i < p.rows_a) ? __ldg(p.bias + i) : 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < kBN; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tbase + ((uint32_t)(32 * q) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
            "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
            "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
            "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (i < p.rows_a) {
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const int j = n0 + c0 + t;
          if (j >= p.rows_b) continue;
          float x = __uint_as_float(v[t]);
          size_t off;
          if (p.mode == 0) {
            if (p.bias) x += __ldg(p.bias + j);
            off = (size_t)i * p.ldo + j;
          } else {
            x += bias_i;
            const int img = j / p.HW, px = j - img * p.HW;
            off = ((size_t)img * p.rows_a + i) * p.HW + px;
          }
          if (p.fuse_relu) x = fmaxf(x, 0.f);
          p.out[off] = x;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wid == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tbase) : "memory");
}

// transposed column buffer: colT[(image, oy, ox)][k], k = (c * kh + r) * kw + s, rows of Kp floats (zero padded): the same
// elements as caffe's im2col (src/caffe/util/im2col.cu) with the K index contiguous, which is what a K-major MMA operand wants
__global__ void im2colT_kernel(long total, const float *__restrict__ in, int C, int H, int W, int kh, int kw, int pad_h, int pad_w, int stride_h,
                               int stride_w, int dil_h, int dil_w, int Ho, int Wo, int K, int Kp, float *__restrict__ colT) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int k = (int)(e % Kp);
  const long row = e / Kp;
  float v = 0.f;
  if (k < K) {
    const int HW = Ho * Wo;
    const int img = (int)(row / HW), px = (int)(row - (long)img * HW);
    const int oy = px / Wo, ox = px - oy * Wo;
    const int s = k % kw, r = (k / kw) % kh, c = k / (kw * kh);
    const int y = oy * stride_h - pad_h + r * dil_h, x = ox * stride_w - pad_w + s * dil_w;
    if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(in + (((size_t)img * C + c) * H + y) * W + x);
  }
  colT[e] = v;
}

__global__ void pad_rows_kernel(long total, const float *__restrict__ src, int K, int Kp, float *__restrict__ dst) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int k = (int)(e % Kp);
  dst[e] = k < K ? __ldg(src + (e / Kp) * K + k) : 0.f;
}

int encode_2d(CUtensorMap *out, const float *ptr, int rows, int K, const char *what) {
  TmaEncodeFn enc = dense_tma_encoder();
  if (!enc) {
    set_last_error(std::string(what) + ": cuTensorMapEncodeTiled is not available");
    return ESCORT_EINVAL;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kBM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error(std::string(what) + ": cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return ESCORT_EINVAL;
  }
  return 0;
}

int launch_gemm(const float *A, int rows_a, const float *B, int rows_b, int K, const DenseParams &prm, cudaStream_t stream, const char *what) {
  CUtensorMap ta, tb;
  int rc;
  if ((rc = encode_2d(&ta, A, rows_a, K, what)) || (rc = encode_2d(&tb, B, rows_b, K, what))) return rc;
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(dense_gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDenseSmem); });
  const dim3 grid((unsigned)ceil_div(rows_a, kBM), (unsigned)ceil_div(rows_b, kBN));
  dense_gemm_tf32_kernel<<<grid, kDenseThreads, kDenseSmem, stream>>>(ta, tb, prm);
  ESCORT_LAUNCH_CHECK();
  return 0;
}

}  // namespace
}  // namespace escort

using namespace escort;

extern "C" ESCORT_API int escort_inner_product_forward(int num, int K, int num_output, const float *bottom, const float *weight,
                                                       const float *bias, int fuse_relu, float *top, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(num >= 0 && K > 0 && num_output > 0 && bottom && weight && top, "escort_inner_product_forward: bad arguments");
  ESCORT_REQUIRE(K % 4 == 0 && ((uintptr_t)bottom & 15) == 0 && ((uintptr_t)weight & 15) == 0,
                 "escort_inner_product_forward: K must be a multiple of 4 and the operands 16-byte aligned (TMA row pitch)");
  if (num == 0) return 0;
  DenseParams prm = {num, num_output, K, 0, num_output, 1, fuse_relu, bias, top};
  return launch_gemm(bottom, num, weight, num_output, K, prm, stream, "escort_inner_product_forward");
}

static int dense_conv_dims(const escort_geom *g, int *Ho, int *Wo, int *K, int *Kp) {
  *Ho = (g->height + 2 * g->pad_h - (g->dilation_h * (g->kernel_h - 1) + 1)) / g->stride_h + 1;
  *Wo = (g->width + 2 * g->pad_w - (g->dilation_w * (g->kernel_w - 1) + 1)) / g->stride_w + 1;
  *K = g->channels * g->kernel_h * g->kernel_w;
  *Kp = (*K + 31) / 32 * 32;
  return 0;
}

extern "C" ESCORT_API size_t escort_dense_conv_workspace_bytes(const escort_geom *g, int num) {
  if (!g || num < 0) return 0;
  int Ho, Wo, K, Kp;
  dense_conv_dims(g, &Ho, &Wo, &K, &Kp);
  return ((size_t)num * Ho * Wo + (size_t)g->num_output) * Kp * sizeof(float) + 256;
}

extern "C" ESCORT_API int escort_dense_conv_forward(const escort_geom *g, int num, const float *bottom, const float *weight, const float *bias,
                                                    int fuse_relu, void *workspace, size_t workspace_bytes, float *top,
                                                    escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(g && num >= 0 && bottom && weight && workspace && top, "escort_dense_conv_forward: null argument");
  ESCORT_REQUIRE(g->group == 1, "escort_dense_conv_forward: group > 1 is not implemented (the reference's dense layers have group 1, AlexNet conv1 included)");
  ESCORT_REQUIRE(workspace_bytes >= escort_dense_conv_workspace_bytes(g, num), "escort_dense_conv_forward: workspace too small");
  int Ho, Wo, K, Kp;
  dense_conv_dims(g, &Ho, &Wo, &K, &Kp);
  if (num == 0 || Ho <= 0 || Wo <= 0) return 0;
  const long rows = (long)num * Ho * Wo;
  ESCORT_REQUIRE(rows < 2147483647L, "escort_dense_conv_forward: batch too large for 32-bit pixel indices");
  float *wpad = reinterpret_cast<float *>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  float *colT = wpad + (size_t)g->num_output * Kp;
  const long wtot = (long)g->num_output * Kp, ctot = rows * Kp;
  pad_rows_kernel<<<(unsigned)((wtot + 255) / 256), 256, 0, stream>>>(wtot, weight, K, Kp, wpad);
  ESCORT_LAUNCH_CHECK();
  im2colT_kernel<<<(unsigned)((ctot + 255) / 256), 256, 0, stream>>>(ctot, bottom, g->channels, g->height, g->width, g->kernel_h, g->kernel_w,
                                                                     g->pad_h, g->pad_w, g->stride_h, g->stride_w, g->dilation_h, g->dilation_w, Ho,
                                                                     Wo, K, Kp, colT);
  ESCORT_LAUNCH_CHECK();
  DenseParams prm = {g->num_output, (int)rows, Kp, 1, 0, Ho * Wo, fuse_relu, bias, top};
  return launch_gemm(wpad, g->num_output, colT, (int)rows, Kp, prm, stream, "escort_dense_conv_forward");
}
