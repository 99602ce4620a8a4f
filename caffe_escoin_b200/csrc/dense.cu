// dense.cu -- SURVEY section 8 (f1): the layers the reference keeps DENSE, on the 5th-generation tensor cores.
//
// The reference runs conv1 / 1x1 / any unpruned convolution through EscConvolutionLayer (cuDNN IMPLICIT_GEMM,
// src/caffe/layers/esc_conv_layer.cu:21-29) and the fully connected layers through InnerProductLayer (cuBLAS sgemm,
// src/caffe/layers/inner_product_layer.cu:9-31).  Both are D[i][j] = sum_k A[i][k] * B[j][k] on tcgen05:
//   * TMA (cp.async.bulk.tensor, 128-byte swizzles) stages 32-float K slices of A and B into an mbarrier ring;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, K = 8; fp32 bits read as TF32, fp32
//     accumulate) from shared-memory descriptors into a TMEM accumulator and releases each stage with tcgen05.commit;
//   * four epilogue warps (one per TMEM lane quadrant) read the accumulator with tcgen05.ld, add the bias, apply ReLU and
//     store.
// Two kernels:
//   dense_gemm_tf32_kernel       inner product: A = bottom [num x K], B = weight [num_output x K], both K-major;
//   dense_conv_tf32_kernel<I>    convolution with the pixels on the M side: I = true is the implicit GEMM of 1x1 / stride 1
//                                layers straight from NCHW (MN-major A operand, no column buffer, no weight copy);
//                                I = false reads the transposed column buffer im2colT_kernel writes (conv1-type layers).
// Precision: TF32 products (the operands' low 13 mantissa bits are dropped), fp32 accumulation -- up to 7e-4 relative L2
// against an fp32 GEMM on full-mantissa data; the sparse path's 1e-4 bar does not apply here (SURVEY section 8 f1:
// "bf16/TF32 questions live here"), the tests state 2e-3.  group = 1.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <string>

#include "common.cuh"

namespace escort {
namespace {

constexpr int kBM = 128, kBN = 128, kBK = 32;  // tile: rows of A, most rows of B, floats of K per stage (128 bytes = one swizzle row)
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
  int rows_a, rows_b, K;     // D is rows_a x rows_b: out[i * ldo + j], bias[j]
  int ldo;
  int BN;                    // rows of B per tile: 128, or 64 when 128 would leave SMs without a tile
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
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * p.BN;
  const int stage_bytes = (kBM + p.BN) * kBK * 4;
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
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + 8 * s), "r"((unsigned)stage_bytes) : "memory");
        const unsigned sa = base + s * stage_bytes, sb = sa + kBM * kBK * 4;
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
      const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(p.BN >> 3) << 17) | ((unsigned)(kBM >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        dn_mbar_wait(full0 + 8 * s, (unsigned)((kb / kStages) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned sa = base + s * stage_bytes, sb = sa + kBM * kBK * 4;
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
#pragma unroll 1
    for (int c0 = 0; c0 < p.BN; c0 += 32) {
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
          if (p.bias) x += __ldg(p.bias + j);
          if (p.fuse_relu) x = fmaxf(x, 0.f);
          p.out[(size_t)i * p.ldo + j] = x;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wid == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tbase) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// Convolution as D[pixel][m] = sum_k A[pixel][k] * W[m][k]: the pixels are the M side of the MMA (128 per tile), the
// output channels the N side (BN = 64 / 128 / 256 columns of TMEM), so a layer with 64 output channels wastes nothing and
// -- a TMEM lane being a pixel -- for every output channel the 32 lanes of an epilogue warp hold 32 consecutive pixels
// of one NCHW row: each store instruction writes one full 128-byte line (bias from shared memory, ReLU fused).
// Two sources of A:
//   kImplicit (1x1, stride 1, no padding): an IMPLICIT GEMM straight from NCHW, no column buffer.  K = input channel, and
//     pixels are contiguous in memory => an MN-major A operand.  A stage is four TMA boxes {32 pixels x 32 channels} of
//     the 3-D tensor (pixel, channel, image) in the "128-byte swizzle with 32-byte atoms"
//     (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B <-> UMMA layout type 1, SWIZZLE_128B_BASE32B: the only MN-major layout
//     tcgen05 takes for 32-bit operands; 32-byte chunks of a 128-byte row XOR the row index mod 4).  A box is eight such
//     atoms (4 channel rows of 128 bytes) stacked along K: canonical layout ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte
//     units with LBO = 4096 B (next 32 pixels = next box), SBO = 512 B (next 4 channels); a K step of 8 channels
//     advances the descriptor start by 1024 B.  Ragged pixel / channel tails are TMA zero fill.  B = the caller's
//     weight matrix [M x C] itself (K-major) -- no padded copy either.
//   !kImplicit (any other geometry): A = the transposed column buffer [(image, pixel) x Kp], K-major, one box per stage.
// The ring is two stages deep (48-96 KB of dynamic shared memory), so two to four CTAs share an SM and the loads of one
// tile overlap the MMAs and stores of another without a persistent loop (measured: 2 stages beat 4 on every layer).
constexpr int kPwBM = 128;
constexpr int kPwABytes = kPwBM * kBK * 4;  // 16 KiB (implicit: 4 boxes of 4 KiB)

struct PwParams {
  int M, HW;               // output channels, pixels per image
  long rows;               // !kImplicit: num * HW
  int tiles_per_img, ntn;  // kImplicit: pixel tiles per image; output-channel tiles
  int BN, stages, nkb;
  int fuse_relu;
  const float *bias;
  const float *residual;   // optional second input of an Eltwise SUM behind the convolution (same shape as out; may alias it)
  float *out;
};

__device__ __forceinline__ uint64_t pw_desc_a(unsigned saddr) {  // MN-major, SWIZZLE_128B_BASE32B
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(4096 >> 4) << 16;  // leading byte offset: the next 32-pixel atom column (one TMA box)
  d |= (uint64_t)(512 >> 4) << 32;   // stride byte offset: the next group of 4 channel rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

template <bool kImplicit, bool kResidual>  // kResidual keeps 32 more registers live in the epilogue: its own instantiation
__global__ void __launch_bounds__(kDenseThreads)
    dense_conv_tf32_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const PwParams p) {
  extern __shared__ unsigned char dsm_raw[];
  const unsigned raw_addr = (unsigned)__cvta_generic_to_shared(dsm_raw);
  const unsigned base = (raw_addr + 1023u) & ~1023u;
  const int stage_bytes = kPwABytes + p.BN * kBK * 4;
  const unsigned bars = base + p.stages * stage_bytes;  // full[4] | empty[4] | tmem_full | tmem slot | bias[256]
  const unsigned full0 = bars, empty0 = bars + 32, tmem_full = bars + 64, tslot = bars + 72, sbias = bars + 128;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int nt = (int)(blockIdx.x % (unsigned)p.ntn), tile = (int)(blockIdx.x / (unsigned)p.ntn);
  const int img = kImplicit ? tile / p.tiles_per_img : 0;
  const int p0 = kImplicit ? (tile - img * p.tiles_per_img) * kPwBM : 0;  // first pixel of the tile in its image
  const long row0 = (long)tile * kPwBM;                                    // !kImplicit: first row of the column buffer
  const int n0 = nt * p.BN;
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      dn_mbar_init(full0 + 8 * s, 1);
      dn_mbar_init(empty0 + 8 * s, 1);
    }
    dn_mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"((unsigned)p.BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (wid >= 2) {  // bias of this tile's output channels -> shared memory (zero beyond M or without a bias)
    for (int j = tid - 64; j < p.BN; j += 128) {
      const float b = (p.bias && n0 + j < p.M) ? __ldg(p.bias + n0 + j) : 0.f;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + 4 * j), "f"(b) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  unsigned tbase;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tbase) : "r"(tslot));

  if (wid == 0) {
    if (lane == 0) {  // ---- TMA producer ----
      const int nbox = kImplicit ? min(4, (p.HW - p0 + 31) / 32) : 1;  // implicit: boxes that hold at least one pixel of the image
      const unsigned tx = (unsigned)((kImplicit ? nbox * 4096 : kPwABytes) + p.BN * kBK * 4);
      for (int kb = 0; kb < p.nkb; ++kb) {
        const int s = kb % p.stages;
        dn_mbar_wait(empty0 + 8 * s, (unsigned)(((kb / p.stages) & 1) ^ 1));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + 8 * s), "r"(tx) : "memory");
        const unsigned sa = base + s * stage_bytes, sb = sa + kPwABytes;
        if (kImplicit) {
          for (int b = 0; b < nbox; ++b)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                             sa + 4096 * b),
                         "l"(&tmap_x), "r"(p0 + 32 * b), "r"(kb * kBK), "r"(img), "r"(full0 + 8 * s)
                         : "memory");
        } else {
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(sa),
                       "l"(&tmap_x), "r"(kb * kBK), "r"((int)row0), "r"(full0 + 8 * s)
                       : "memory");
        }
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(sb),
                     "l"(&tmap_w), "r"(kb * kBK), "r"(n0), "r"(full0 + 8 * s)
                     : "memory");
      }
    }
  } else if (wid == 1) {
    if (lane == 0) {  // ---- MMA issuer: A MN-major (bit 15) when implicit, B K-major ----
      const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | (kImplicit ? (1u << 15) : 0u) | ((unsigned)(p.BN >> 3) << 17) |
                             ((unsigned)(kPwBM >> 4) << 24);
      for (int kb = 0; kb < p.nkb; ++kb) {
        const int s = kb % p.stages;
        dn_mbar_wait(full0 + 8 * s, (unsigned)((kb / p.stages) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned sa = base + s * stage_bytes, sb = sa + kPwABytes;
        const uint64_t da = kImplicit ? pw_desc_a(sa) : dn_smem_desc(sa), db = dn_smem_desc(sb);
        constexpr int kstep_a = kImplicit ? (1024 >> 4) : (32 >> 4);
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k) {
          const unsigned acc = (kb > 0 || k > 0) ? 1u : 0u;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase),
                       "l"(da + (uint64_t)(kstep_a * k)), "l"(db + (uint64_t)(2 * k)), "r"(idesc), "r"(acc)
                       : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + 8 * s) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tmem_full) : "memory");
    }
  } else {
    // ---- epilogue: warp w owns TMEM lanes 32 * (w % 4) .. + 31 = pixels (rows) 32 * (w % 4) + lane of the tile ----
    const int q = wid & 3;
    dn_mbar_wait(tmem_full, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    bool live;
    size_t o0;  // index of out[image][n0][pixel]
    if (kImplicit) {
      const int px = p0 + 32 * q + lane;
      live = px < p.HW;
      o0 = ((size_t)img * p.M + n0) * p.HW + px;
    } else {
      const long row = row0 + 32 * q + lane;
      live = row < p.rows;
      const long im = row / p.HW;
      o0 = ((size_t)im * p.M + n0) * p.HW + (row - im * p.HW);
    }
    float *orow = p.out + o0;
    [[maybe_unused]] const float *rrow = kResidual ? p.residual + o0 : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < p.BN && n0 + c0 < p.M; c0 += 32) {
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
      const int mleft = p.M - n0 - c0;
      [[maybe_unused]] float r[kResidual ? 32 : 1];  // all residual loads of the chunk before its first store: they may alias
                                                     // (in-place Eltwise), and 32 load -> store pairs in program order would
                                                     // pay 32 dependent global latencies
      if constexpr (kResidual) {
#pragma unroll
        for (int t = 0; t < 32; ++t) r[t] = (live && t < mleft) ? rrow[(size_t)(c0 + t) * p.HW] : 0.f;
      }
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        float b;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(b) : "r"(sbias + 4 * (c0 + t)));
        float x = __uint_as_float(v[t]) + b;
        if constexpr (kResidual) x += r[t];
        if (p.fuse_relu) x = fmaxf(x, 0.f);
        if (live && t < mleft) orow[(size_t)(c0 + t) * p.HW] = x;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wid == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"((unsigned)p.BN) : "memory");
}

// transposed column buffer: colT[(image, oy, ox)][k], k = (c * kh + r) * kw + s, rows of Kp floats (zero padded): the same
// elements as caffe's im2col (src/caffe/util/im2col.cu) with the K index contiguous, which is what a K-major MMA operand wants
__global__ void im2colT_kernel(long total4, const float *__restrict__ in, int C, int H, int W, int kh, int kw, int pad_h, int pad_w,
                               int stride_h, int stride_w, int dil_h, int dil_w, int Ho, int Wo, int K, int Kp, float *__restrict__ colT) {
  // one thread = four consecutive k of one row (Kp is a multiple of 32): one 16-byte store, the (c, r, s) decode done once
  const long e4 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e4 >= total4) return;
  const int kq = Kp >> 2;
  const int k0 = (int)(e4 % kq) * 4;
  const long row = e4 / kq;
  const int HW = Ho * Wo;
  const int img = (int)(row / HW), px = (int)(row - (long)img * HW);
  const int oy = px / Wo, ox = px - oy * Wo;
  int s = k0 % kw, r = (k0 / kw) % kh, c = k0 / (kw * kh);
  const int y0 = oy * stride_h - pad_h, x0 = ox * stride_w - pad_w;
  const float *src = in + (size_t)img * C * H * W;
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t = 0.f;
    if (k0 + i < K) {
      const int y = y0 + r * dil_h, x = x0 + s * dil_w;
      if (y >= 0 && y < H && x >= 0 && x < W) t = __ldg(src + ((size_t)c * H + y) * W + x);
    }
    v[i] = t;
    if (++s == kw) {
      s = 0;
      if (++r == kh) {
        r = 0;
        ++c;
      }
    }
  }
  reinterpret_cast<float4 *>(colT)[e4] = make_float4(v[0], v[1], v[2], v[3]);
}

__global__ void pad_rows_kernel(long total, const float *__restrict__ src, int K, int Kp, float *__restrict__ dst) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int k = (int)(e % Kp);
  dst[e] = k < K ? __ldg(src + (e / Kp) * K + k) : 0.f;
}

int encode_2d(CUtensorMap *out, const float *ptr, int rows, int K, int box_rows, const char *what) {
  TmaEncodeFn enc = dense_tma_encoder();
  if (!enc) {
    set_last_error(std::string(what) + ": cuTensorMapEncodeTiled is not available");
    return ESCORT_EINVAL;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error(std::string(what) + ": cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return ESCORT_EINVAL;
  }
  return 0;
}

int encode_tiled(CUtensorMap *out, const float *ptr, int rank, const cuuint64_t *dims, const cuuint64_t *strides, const cuuint32_t *box,
                 CUtensorMapSwizzle swizzle, const char *what) {
  TmaEncodeFn enc = dense_tma_encoder();
  if (!enc) {
    set_last_error(std::string(what) + ": cuTensorMapEncodeTiled is not available");
    return ESCORT_EINVAL;
  }
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float *>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error(std::string(what) + ": cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return ESCORT_EINVAL;
  }
  return 0;
}

// the 1x1 / stride 1 / no padding convolution the implicit GEMM takes: TMA needs 16-byte pitches (HW and C multiples of 4)
bool pointwise_applies(const escort_geom *g, const float *bottom, const float *weight) {
  return g->kernel_h == 1 && g->kernel_w == 1 && g->stride_h == 1 && g->stride_w == 1 && g->pad_h == 0 && g->pad_w == 0 && g->group == 1 &&
         (g->height * g->width) % 4 == 0 && g->channels % 4 == 0 && ((uintptr_t)bottom & 15) == 0 && ((uintptr_t)weight & 15) == 0 &&
         !getenv("ESCORT_DENSE_NO_IMPLICIT");
}

// launch of dense_conv_tf32_kernel: implicit = straight from NCHW (1x1), else from the transposed column buffer
int launch_conv(bool implicit, const escort_geom *g, int num, int HW, const float *a_src, int K, const float *w_src, const float *bias,
                const float *residual, int fuse_relu, float *top, cudaStream_t stream) {
  const char *what = "escort_dense_conv_forward";
  const int M = g->num_output;
  PwParams prm;
  prm.M = M, prm.HW = HW, prm.rows = (long)num * HW;
  prm.tiles_per_img = ceil_div(HW, kPwBM);
  prm.nkb = ceil_div(K, kBK);
  // tile width and ring depth, measured (profiles/r02_dense_tcgen05.txt): CTAs per SM matter more than ring depth -- two
  // stages leave room for 2-4 co-resident CTAs whose loads, MMAs and stores overlap; 256 columns only pay when one tile
  // covers all output channels and the K loop is long
  prm.BN = M <= 64 ? 64 : (M > 128 && M <= 256 && prm.nkb >= 8) ? 256 : 128;
  if (const char *e = getenv("ESCORT_DENSE_BN")) prm.BN = atoi(e) == 64 ? 64 : atoi(e) == 128 ? 128 : 256;  // (measurement knob)
  prm.ntn = ceil_div(M, prm.BN);
  prm.stages = std::min(prm.nkb, 2);
  if (const char *e = getenv("ESCORT_DENSE_STAGES")) prm.stages = std::max(1, std::min(std::min(prm.nkb, 4), atoi(e)));  // (measurement knob)
  prm.fuse_relu = fuse_relu, prm.bias = bias, prm.residual = residual, prm.out = top;
  CUtensorMap tx, tw;
  int rc;
  if (implicit) {
    const cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)K, (cuuint64_t)num};
    const cuuint64_t strides[2] = {(cuuint64_t)HW * 4, (cuuint64_t)K * HW * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)kBK, 1};
    rc = encode_tiled(&tx, a_src, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, what);
  } else {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)prm.rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kPwBM};
    rc = encode_tiled(&tx, a_src, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, what);
  }
  if (rc) return rc;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)prm.BN};
    if ((rc = encode_tiled(&tw, w_src, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, what))) return rc;
  }
  const int smem = prm.stages * (kPwABytes + prm.BN * kBK * 4) + 1024 /* alignment slack */ + 128 /* barriers */ + 1024 /* bias */;
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const PwParams);
  static const Kern kerns[4] = {dense_conv_tf32_kernel<false, false>, dense_conv_tf32_kernel<false, true>, dense_conv_tf32_kernel<true, false>,
                                dense_conv_tf32_kernel<true, true>};
  static std::once_flag once;
  std::call_once(once, [] {
    const int most = 4 * (kPwABytes + 256 * kBK * 4) + 2176;
    for (Kern k : kerns) {
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
      // (the driver's default carve-out keeps 3 CTAs of the 48 KB shape per SM; forcing the maximum carve-out for a 4th
      // was measured 5-10 % slower on every layer)
    }
  });
  const long tiles = implicit ? (long)num * prm.tiles_per_img : (prm.rows + kPwBM - 1) / kPwBM;
  const long ctas = tiles * prm.ntn;
  ESCORT_REQUIRE(ctas < 2147483647L, "escort_dense_conv_forward: batch too large for one launch");
  kerns[(implicit ? 2 : 0) + (residual ? 1 : 0)]<<<(unsigned)ctas, kDenseThreads, smem, stream>>>(tx, tw, prm);
  ESCORT_LAUNCH_CHECK();
  return 0;
}

int launch_gemm(const float *A, int rows_a, const float *B, int rows_b, int K, DenseParams prm, cudaStream_t stream, const char *what) {
  // 128 x 128 tiles unless they would leave SMs idle: a small-batch inner product streams its weights once, and one SM
  // cannot pull more than its share of the HBM bandwidth (fc6 at batch 256: 64 tiles on 148 SMs)
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  prm.BN = (ceil_div(rows_a, kBM) * ceil_div(rows_b, kBN) < sms && rows_b > 64) ? 64 : kBN;
  if (const char *e = getenv("ESCORT_DENSE_FC_BN")) prm.BN = atoi(e) == 64 ? 64 : kBN;  // (measurement knob)
  CUtensorMap ta, tb;
  int rc;
  if ((rc = encode_2d(&ta, A, rows_a, K, kBM, what)) || (rc = encode_2d(&tb, B, rows_b, K, prm.BN, what))) return rc;
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(dense_gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDenseSmem); });
  const dim3 grid((unsigned)ceil_div(rows_a, kBM), (unsigned)ceil_div(rows_b, prm.BN));
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
  DenseParams prm = {num, num_output, K, num_output, kBN, fuse_relu, bias, top};
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

extern "C" ESCORT_API int escort_dense_conv_forward_residual(const escort_geom *g, int num, const float *bottom, const float *weight,
                                                             const float *bias, const float *residual, int fuse_relu, void *workspace,
                                                             size_t workspace_bytes, float *top, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(g && num >= 0 && bottom && weight && workspace && top, "escort_dense_conv_forward: null argument");
  ESCORT_REQUIRE(g->group == 1, "escort_dense_conv_forward: group > 1 is not implemented (the reference's dense layers have group 1, AlexNet conv1 included)");
  ESCORT_REQUIRE(workspace_bytes >= escort_dense_conv_workspace_bytes(g, num), "escort_dense_conv_forward: workspace too small");
  int Ho, Wo, K, Kp;
  dense_conv_dims(g, &Ho, &Wo, &K, &Kp);
  if (num == 0 || Ho <= 0 || Wo <= 0) return 0;
  if (pointwise_applies(g, bottom, weight))
    return launch_conv(true, g, num, Ho * Wo, bottom, g->channels, weight, bias, residual, fuse_relu, top, stream);
  const long rows = (long)num * Ho * Wo;
  ESCORT_REQUIRE(rows < 2147483647L, "escort_dense_conv_forward: batch too large for 32-bit pixel indices");
  float *wpad = reinterpret_cast<float *>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  float *colT = wpad + (size_t)g->num_output * Kp;
  const long wtot = (long)g->num_output * Kp, ctot = rows * Kp;
  pad_rows_kernel<<<(unsigned)((wtot + 255) / 256), 256, 0, stream>>>(wtot, weight, K, Kp, wpad);
  ESCORT_LAUNCH_CHECK();
  im2colT_kernel<<<(unsigned)((ctot / 4 + 255) / 256), 256, 0, stream>>>(ctot / 4, bottom, g->channels, g->height, g->width, g->kernel_h,
                                                                         g->kernel_w, g->pad_h, g->pad_w, g->stride_h, g->stride_w,
                                                                         g->dilation_h, g->dilation_w, Ho, Wo, K, Kp, colT);
  ESCORT_LAUNCH_CHECK();
  return launch_conv(false, g, num, Ho * Wo, colT, Kp, wpad, bias, residual, fuse_relu, top, stream);
}

extern "C" ESCORT_API int escort_dense_conv_forward(const escort_geom *g, int num, const float *bottom, const float *weight, const float *bias,
                                                    int fuse_relu, void *workspace, size_t workspace_bytes, float *top,
                                                    escort_stream_t stream_) {
  return escort_dense_conv_forward_residual(g, num, bottom, weight, bias, nullptr, fuse_relu, workspace, workspace_bytes, top, stream_);
}
