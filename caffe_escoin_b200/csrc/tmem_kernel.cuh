// tmem_kernel.cuh -- device side of the TMEM-window forward kernel (design notes: DESIGN.md section 4.4, host side:
// sconv_tmem.cu).
//
// The reference walks one CSR row per output channel and reads the input through the "stretched" column index
// (include/caffe/util/sconv.hpp:594-678, src/caffe/util/math_functions.cu:264-319): out[q] += w * in[q + off] over a
// padded, flattened image.  Here the same walk runs with the input window resident in TENSOR MEMORY: a lane owns T
// consecutive flattened output positions, its T + halo input floats of one channel sit in T + halo TMEM columns of its
// own TMEM lane, and the shift of a nonzero (kh * pitch + kw) is the COLUMN ADDRESS of one tcgen05.ld -- an address,
// not a register index, so there are no per-tap handlers, no masks and no patch registers.  Measured on B200
// (tools/ubench/tmem_bench.cu): tcgen05.ld delivers 960 B/clk/SM against 128 B/clk/SM for shared memory, and the loop
// "tcgen05.ld.x32 ; 16 x FFMA2" runs at 63 TFLOP/s (0.88 of the FP32 peak) at any weight density.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "common.cuh"

namespace escort {

static constexpr int kTmMaxStages = 8;
static constexpr int kTmMaxSlots = 16;
// barrier area: smem_full[8] | smem_empty[8] | tm_full[4][16] | tm_empty[4][16] | tmem base slot
static constexpr int kTmBarBytes = (2 * kTmMaxStages + 8 * kTmMaxSlots) * 8 + 128;

struct TmParams {
  // geometry
  int C, H, W, M, Ho, Wo, pad_h, pad_w;
  int Cg, Mg, ngroups;
  int PW, IMGR, IMG;      // padded row pitch (W + pad_w), rows per image block (H + pad_h), positions per image block
  int TILE, HALO, SW;     // flattened positions per work unit (32 * T), window halo, floats per staged channel row
  int SLOTW, CHS, NSLOT;  // TMEM columns per channel window, channels per slot group, slot groups in the TMEM ring
  int CI, nchunks, NS;    // channels per staged chunk, chunks per conv group, shared-memory stages
  int nsg;                // slot groups per full chunk (CI / CHS)
  int nblk, ogroups;      // channel blocks (OT channels each) per conv group, CTA passes (NCW blocks each) per conv group
  int ntiles;             // filled per launch: ceil(num * IMG / TILE)
  int stage0_off, stage_bytes, in_bytes, hdr_counts_off;
  int ostage_off;         // per compute warp: TILE floats of output staging (transposes lane-major tiles for coalesced stores)
  int lpr_shift, RO;      // loader: log2(lanes per padded row), rows per warp-wide copy instruction
  int ltab_off, ltab_n;   // loader table: smem offset, steps (entries = steps * 32 lanes)
  int SWP;                // floats per staged channel row incl. the skew padding (SW / 32 * 36)
  const int *oc_list;     // [ngroups * nblk * OT] global output channel or -1
  const uint4 *prog;      // record regions (16-byte units)
  const int2 *rtab;       // [ngroups * ogroups * nchunks] {offset, length} of a region in 16-byte units
  int *dbg;               // host-mapped debug words (bounded barrier waits report here before trapping), may be null
  int skip;               // measurement only (ESCORT_TM_SKIP, results are WRONG): 1 = no window fill, 2 = no taps, 4 = no input loads
};

#ifndef ESCORT_TMEM_HOST_ONLY
// ---- optional event trace (make TMTRACE=1; tools/tm_trace.py): lane 0 of every warp of CTA 0 logs {clock, code} ------
#ifdef ESCORT_TM_TRACE
static constexpr int kTmTraceWarps = 24, kTmTraceEvents = 4096;
__device__ unsigned long long g_tm_trace[kTmTraceWarps * kTmTraceEvents];
__device__ int g_tm_trace_n[kTmTraceWarps];
// per-warp event counters live in the spare bytes behind the TMEM base slot of the barrier area (shared memory)
__device__ __forceinline__ void tm_ev(unsigned code) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    int *cnt = reinterpret_cast<int *>(smem_raw + (2 * 8 + 8 * 16) * 8 + 16) + w;
    const int n = *cnt;
    if (n < kTmTraceEvents) {
      g_tm_trace[w * kTmTraceEvents + n] = ((unsigned long long)clock64() << 8) | code;
      *cnt = n + 1;
    }
  }
}
__device__ __forceinline__ void tm_ev_init() {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  if (threadIdx.x < kTmTraceWarps) reinterpret_cast<int *>(smem_raw + (2 * 8 + 8 * 16) * 8 + 16)[threadIdx.x] = 0;
}
__device__ __forceinline__ void tm_ev_fini() {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  if (blockIdx.x == 0 && threadIdx.x < kTmTraceWarps) g_tm_trace_n[threadIdx.x] = reinterpret_cast<int *>(smem_raw + (2 * 8 + 8 * 16) * 8 + 16)[threadIdx.x];
}
#define TM_EV(code) tm_ev(code)
#else
#define TM_EV(code) do { } while (0)
#define tm_ev_init() do { } while (0)
#define tm_ev_fini() do { } while (0)
#endif
// ---- small PTX helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tm_mbar_init(unsigned addr, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void tm_mbar_arrive(unsigned addr) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}
// Bounded wait: a protocol bug (or a lost arrival) must not hang the GPU.  After ~2 s without progress the first warp to
// give up records {code, block, warp, barrier offset, parity} in the host-mapped debug words and traps.
__device__ __forceinline__ void tm_mbar_wait(unsigned addr, unsigned parity, int code, int *dbg) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(addr), "r"(parity)
               : "memory");
  if (ok) return;
  unsigned long long t0 = 0;
  unsigned backoff = 32;
  for (;;) {
    // Polls with a growing plain sleep in between.  try_wait's own suspension ends at every barrier event of the CTA
    // (hundreds of cp.async arrivals per chunk), so four waiting warps polled ~70 times per wait and took about a
    // third of their scheduler's issue slots from the one warp everybody was waiting for (profiles/r02_ncu_*_polls.txt).
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
      asm volatile("nanosleep.u32 %0;" ::"r"(backoff));
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok)
                   : "r"(addr), "r"(parity)
                   : "memory");
      if (ok) return;
      if (backoff < 256) backoff += backoff >> 1;
    }
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t0 == 0) t0 = t1;
    if (t1 - t0 > 2000000000ull) {
      if (dbg && (threadIdx.x & 31) == 0 && atomicCAS(dbg, 0, code) == 0) {
        dbg[1] = (int)blockIdx.x;
        dbg[2] = (int)(threadIdx.x >> 5);
        dbg[3] = (int)(addr & 0xffffu);
        dbg[4] = (int)parity;
        __threadfence_system();
      }
      __trap();
    }
  }
}
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// cp.async with zero fill: src_bytes = 0 writes zeros (halo positions and rows outside the batch)
__device__ __forceinline__ void tm_cp_async4(unsigned dst_smem, const float *src, unsigned src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void tm_cp_async16(unsigned dst_smem, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}

#define TM_O8(r, b) "=r"(r[b + 0]), "=r"(r[b + 1]), "=r"(r[b + 2]), "=r"(r[b + 3]), "=r"(r[b + 4]), "=r"(r[b + 5]), "=r"(r[b + 6]), "=r"(r[b + 7])
#define TM_I8(r, b) "r"(r[b + 0]), "r"(r[b + 1]), "r"(r[b + 2]), "r"(r[b + 3]), "r"(r[b + 4]), "r"(r[b + 5]), "r"(r[b + 6]), "r"(r[b + 7])

template <int T>
__device__ __forceinline__ void tm_ld_window(uint32_t (&r)[T], uint32_t taddr);
template <>
__device__ __forceinline__ void tm_ld_window<32>(uint32_t (&r)[32], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : TM_O8(r, 0), TM_O8(r, 8), TM_O8(r, 16), TM_O8(r, 24)
      : "r"(taddr)
      : "memory");
}
template <>
__device__ __forceinline__ void tm_ld_window<16>(uint32_t (&r)[16], uint32_t taddr) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : TM_O8(r, 0), TM_O8(r, 8)
               : "r"(taddr)
               : "memory");
}
// One nonzero: tcgen05.ld of the lane's T-column window at `taddr`, then T/2 packed FMAs acc += {w, w} * x.  Kept as ONE
// asm block so that ptxas sees the accumulators as plain in/out operands of the FMAs (as separate statements it
// computed the FMAs into the window registers and copied all of them back: 16 extra MOVs per nonzero).
template <int T>
__device__ __forceinline__ void tm_tap(unsigned long long (&acc)[T / 2], uint32_t taddr, unsigned wbits);
template <>
__device__ __forceinline__ void tm_tap<16>(unsigned long long (&a)[8], uint32_t taddr, unsigned wbits) {
  asm volatile(
      "{\n\t.reg .b32 t<16>;\n\t.reg .b64 p<8>, w2;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15}, [%8];\n\t"
      "mov.b64 w2, {%9, %9};\n\t"
      "tcgen05.wait::ld.sync.aligned;\n\t"
      "mov.b64 p0, {t0, t1};\n\tmov.b64 p1, {t2, t3};\n\tmov.b64 p2, {t4, t5};\n\tmov.b64 p3, {t6, t7};\n\t"
      "mov.b64 p4, {t8, t9};\n\tmov.b64 p5, {t10, t11};\n\tmov.b64 p6, {t12, t13};\n\tmov.b64 p7, {t14, t15};\n\t"
      "fma.rn.f32x2 %0, w2, p0, %0;\n\tfma.rn.f32x2 %1, w2, p1, %1;\n\tfma.rn.f32x2 %2, w2, p2, %2;\n\tfma.rn.f32x2 %3, w2, p3, %3;\n\t"
      "fma.rn.f32x2 %4, w2, p4, %4;\n\tfma.rn.f32x2 %5, w2, p5, %5;\n\tfma.rn.f32x2 %6, w2, p6, %6;\n\tfma.rn.f32x2 %7, w2, p7, %7;\n\t}"
      : "+l"(a[0]), "+l"(a[1]), "+l"(a[2]), "+l"(a[3]), "+l"(a[4]), "+l"(a[5]), "+l"(a[6]), "+l"(a[7])
      : "r"(taddr), "r"(wbits)
      : "memory");
}
template <>
__device__ __forceinline__ void tm_tap<32>(unsigned long long (&a)[16], uint32_t taddr, unsigned wbits) {
  asm volatile(
      "{\n\t.reg .b32 t<32>;\n\t.reg .b64 p<16>, w2;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15,t16,t17,t18,t19,t20,t21,t22,t23,t24,t25,t26,t27,t28,t29,t30,t31}, [%16];\n\t"
      "mov.b64 w2, {%17, %17};\n\t"
      "tcgen05.wait::ld.sync.aligned;\n\t"
      "mov.b64 p0, {t0, t1};\n\tmov.b64 p1, {t2, t3};\n\tmov.b64 p2, {t4, t5};\n\tmov.b64 p3, {t6, t7};\n\t"
      "mov.b64 p4, {t8, t9};\n\tmov.b64 p5, {t10, t11};\n\tmov.b64 p6, {t12, t13};\n\tmov.b64 p7, {t14, t15};\n\t"
      "mov.b64 p8, {t16, t17};\n\tmov.b64 p9, {t18, t19};\n\tmov.b64 p10, {t20, t21};\n\tmov.b64 p11, {t22, t23};\n\t"
      "mov.b64 p12, {t24, t25};\n\tmov.b64 p13, {t26, t27};\n\tmov.b64 p14, {t28, t29};\n\tmov.b64 p15, {t30, t31};\n\t"
      "fma.rn.f32x2 %0, w2, p0, %0;\n\tfma.rn.f32x2 %1, w2, p1, %1;\n\tfma.rn.f32x2 %2, w2, p2, %2;\n\tfma.rn.f32x2 %3, w2, p3, %3;\n\t"
      "fma.rn.f32x2 %4, w2, p4, %4;\n\tfma.rn.f32x2 %5, w2, p5, %5;\n\tfma.rn.f32x2 %6, w2, p6, %6;\n\tfma.rn.f32x2 %7, w2, p7, %7;\n\t"
      "fma.rn.f32x2 %8, w2, p8, %8;\n\tfma.rn.f32x2 %9, w2, p9, %9;\n\tfma.rn.f32x2 %10, w2, p10, %10;\n\tfma.rn.f32x2 %11, w2, p11, %11;\n\t"
      "fma.rn.f32x2 %12, w2, p12, %12;\n\tfma.rn.f32x2 %13, w2, p13, %13;\n\tfma.rn.f32x2 %14, w2, p14, %14;\n\tfma.rn.f32x2 %15, w2, p15, %15;\n\t}"
      : "+l"(a[0]), "+l"(a[1]), "+l"(a[2]), "+l"(a[3]), "+l"(a[4]), "+l"(a[5]), "+l"(a[6]), "+l"(a[7]), "+l"(a[8]), "+l"(a[9]),
        "+l"(a[10]), "+l"(a[11]), "+l"(a[12]), "+l"(a[13]), "+l"(a[14]), "+l"(a[15])
      : "r"(taddr), "r"(wbits)
      : "memory");
}
// Two nonzeros of the same output-channel slot: both window loads are issued before the first FMA, so the second
// load's latency hides behind the first tap's FMAs (the accumulation order stays a-then-b: bit-identical to two taps).
template <int T>
__device__ __forceinline__ void tm_tap2(unsigned long long (&acc)[T / 2], uint32_t ta, unsigned wa, uint32_t tb, unsigned wb);
template <>
__device__ __forceinline__ void tm_tap2<16>(unsigned long long (&a)[8], uint32_t ta, unsigned wa, uint32_t tb, unsigned wb) {
  asm volatile(
      "{\n\t.reg .b32 t<16>, u<16>;\n\t.reg .b64 p<8>, q<8>, w2, v2;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15}, [%8];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {u0,u1,u2,u3,u4,u5,u6,u7,u8,u9,u10,u11,u12,u13,u14,u15}, [%10];\n\t"
      "mov.b64 w2, {%9, %9};\n\tmov.b64 v2, {%11, %11};\n\t"
      "tcgen05.wait::ld.sync.aligned;\n\t"
      "mov.b64 p0, {t0, t1};\n\tmov.b64 p1, {t2, t3};\n\tmov.b64 p2, {t4, t5};\n\tmov.b64 p3, {t6, t7};\n\t"
      "mov.b64 p4, {t8, t9};\n\tmov.b64 p5, {t10, t11};\n\tmov.b64 p6, {t12, t13};\n\tmov.b64 p7, {t14, t15};\n\t"
      "mov.b64 q0, {u0, u1};\n\tmov.b64 q1, {u2, u3};\n\tmov.b64 q2, {u4, u5};\n\tmov.b64 q3, {u6, u7};\n\t"
      "mov.b64 q4, {u8, u9};\n\tmov.b64 q5, {u10, u11};\n\tmov.b64 q6, {u12, u13};\n\tmov.b64 q7, {u14, u15};\n\t"
      "fma.rn.f32x2 %0, w2, p0, %0;\n\tfma.rn.f32x2 %1, w2, p1, %1;\n\tfma.rn.f32x2 %2, w2, p2, %2;\n\tfma.rn.f32x2 %3, w2, p3, %3;\n\t"
      "fma.rn.f32x2 %4, w2, p4, %4;\n\tfma.rn.f32x2 %5, w2, p5, %5;\n\tfma.rn.f32x2 %6, w2, p6, %6;\n\tfma.rn.f32x2 %7, w2, p7, %7;\n\t"
      "fma.rn.f32x2 %0, v2, q0, %0;\n\tfma.rn.f32x2 %1, v2, q1, %1;\n\tfma.rn.f32x2 %2, v2, q2, %2;\n\tfma.rn.f32x2 %3, v2, q3, %3;\n\t"
      "fma.rn.f32x2 %4, v2, q4, %4;\n\tfma.rn.f32x2 %5, v2, q5, %5;\n\tfma.rn.f32x2 %6, v2, q6, %6;\n\tfma.rn.f32x2 %7, v2, q7, %7;\n\t}"
      : "+l"(a[0]), "+l"(a[1]), "+l"(a[2]), "+l"(a[3]), "+l"(a[4]), "+l"(a[5]), "+l"(a[6]), "+l"(a[7])
      : "r"(ta), "r"(wa), "r"(tb), "r"(wb)
      : "memory");
}
template <>
__device__ __forceinline__ void tm_tap2<32>(unsigned long long (&a)[16], uint32_t ta, unsigned wa, uint32_t tb, unsigned wb) {
  tm_tap<32>(a, ta, wa);  // 32-column windows: 64 more registers for the pair are not available; one after the other
  tm_tap<32>(a, tb, wb);
}
__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               TM_I8(r, 0), TM_I8(r, 8)
               : "memory");
}

// Skew padding of a staged channel row / an output staging tile: one 16-byte pad chunk after every 8 chunks.  A lane's
// 128-bit accesses have a stride of T/4 chunks, which alone would hit 2 (T = 16) or 1 (T = 32) of the 8 bank groups;
// with the skew the 8 lanes of a quarter warp land on 8 different groups, and -- unlike an XOR swizzle -- the 4
// chunks of an aligned 16-column block stay consecutive, so a fill block is one address + immediate offsets.
__device__ __forceinline__ unsigned tm_skew(unsigned chunk) { return chunk + (chunk >> 3); }

struct TmUnit {
  int og, cg, tile;
};
__device__ __forceinline__ TmUnit tm_decode_unit(const TmParams &p, int u) {
  // og fastest: CTAs that share an input tile run at the same time and hit it in L2
  TmUnit c;
  c.og = u % p.ogroups; u /= p.ogroups;
  c.cg = u % p.ngroups; u /= p.ngroups;
  c.tile = u;
  return c;
}

static constexpr int kTm2BufCols = 256;  // TMEM columns per buffer

// One output channel's records of one slot group: n records of {absolute TMEM address, weight} at shared address rp
// (advanced past them).  Accumulation order = record order (the reference's CSR order).  Reads up to 32 bytes past the
// last record (prefetch; the host pads every region).
#define TM2_T16 "{t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15}"
#define TM2_U16 "{u0,u1,u2,u3,u4,u5,u6,u7,u8,u9,u10,u11,u12,u13,u14,u15}"
#define TM2_PACK(P, R)                                                                                                    \
  "mov.b64 " #P "0, {" #R "0, " #R "1};\n\tmov.b64 " #P "1, {" #R "2, " #R "3};\n\tmov.b64 " #P "2, {" #R "4, " #R "5};\n\t"         \
  "mov.b64 " #P "3, {" #R "6, " #R "7};\n\tmov.b64 " #P "4, {" #R "8, " #R "9};\n\tmov.b64 " #P "5, {" #R "10, " #R "11};\n\t"      \
  "mov.b64 " #P "6, {" #R "12, " #R "13};\n\tmov.b64 " #P "7, {" #R "14, " #R "15};\n\t"
#define TM2_FMA8(W, P)                                                                                                    \
  "fma.rn.f32x2 %0, " #W ", " #P "0, %0;\n\tfma.rn.f32x2 %1, " #W ", " #P "1, %1;\n\tfma.rn.f32x2 %2, " #W ", " #P "2, %2;\n\t"     \
  "fma.rn.f32x2 %3, " #W ", " #P "3, %3;\n\tfma.rn.f32x2 %4, " #W ", " #P "4, %4;\n\tfma.rn.f32x2 %5, " #W ", " #P "5, %5;\n\t"     \
  "fma.rn.f32x2 %6, " #W ", " #P "6, %6;\n\tfma.rn.f32x2 %7, " #W ", " #P "7, %7;\n\t"
// Invariant at entry and exit: (c0, w0) = the record at rp, (c1, w1) = the record at rp + 8 -- the next section's first
// records are already in registers when it starts (their load overlapped this section's last FMAs).
template <int BUF>
__device__ __forceinline__ void tm2_section16(unsigned long long (&a)[8], unsigned &rp, unsigned n, unsigned &c0, unsigned &w0, unsigned &c1,
                                              unsigned &w1) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 n2, ca, cb;\n\t"
      ".reg .b32 t<16>, u<16>;\n\t"
      ".reg .b64 x<8>, y<8>, ww, vv;\n\t"
      "and.b32 n2, %13, 1;\n\t"
      "setp.eq.u32 p, n2, 0;\n\t"
      "@p bra TM2_PAIRS;\n\t"
      "add.u32 ca, %9, %14;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 " TM2_T16 ", [ca];\n\t"
      "mov.b64 ww, {%10, %10};\n\t"
      "mov.b32 %9, %11;\n\t"
      "mov.b32 %10, %12;\n\t"
      "ld.shared.v2.u32 {%11, %12}, [%8+16];\n\t"
      "add.u32 %8, %8, 8;\n\t"
      "tcgen05.wait::ld.sync.aligned;\n\t"
      TM2_PACK(x, t)
      TM2_FMA8(ww, x)
      "TM2_PAIRS:\n\t"
      "shr.u32 n2, %13, 1;\n\t"
      "setp.eq.u32 p, n2, 0;\n\t"
      "@p bra TM2_DONE;\n\t"
      "TM2_LOOP:\n\t"
      "add.u32 ca, %9, %14;\n\t"
      "add.u32 cb, %11, %14;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 " TM2_T16 ", [ca];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 " TM2_U16 ", [cb];\n\t"
      "mov.b64 ww, {%10, %10};\n\t"
      "mov.b64 vv, {%12, %12};\n\t"
      "ld.shared.v2.u32 {%9, %10}, [%8+16];\n\t"
      "ld.shared.v2.u32 {%11, %12}, [%8+24];\n\t"
      "add.u32 %8, %8, 16;\n\t"
      "tcgen05.wait::ld.sync.aligned;\n\t"
      TM2_PACK(x, t)
      TM2_PACK(y, u)
      TM2_FMA8(ww, x)
      TM2_FMA8(vv, y)
      "sub.u32 n2, n2, 1;\n\t"
      "setp.ne.u32 p, n2, 0;\n\t"
      "@p bra TM2_LOOP;\n\t"
      "TM2_DONE:\n\t"
      "}"
      : "+l"(a[0]), "+l"(a[1]), "+l"(a[2]), "+l"(a[3]), "+l"(a[4]), "+l"(a[5]), "+l"(a[6]), "+l"(a[7]), "+r"(rp), "+r"(c0), "+r"(w0),
        "+r"(c1), "+r"(w1)
      : "r"(n), "n"(BUF * kTm2BufCols)
      : "memory");
}

// the same with the window base (slot group position in the TMEM ring) in a register
__device__ __forceinline__ void tm2_section16_r(unsigned long long (&a)[8], unsigned &rp, unsigned n, unsigned base) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 n2, c0, w0, c1, w1;\n\t"
      ".reg .b32 t<16>, u<16>;\n\t"
      ".reg .b64 x<8>, y<8>, ww, vv;\n\t"
      "and.b32 n2, %9, 1;\n\t"
      "setp.eq.u32 p, n2, 0;\n\t"
      "@p bra TM2_PAIRS;\n\t"
      "ld.shared.v2.u32 {c0, w0}, [%8];\n\t"
      "add.u32 %8, %8, 8;\n\t"
      "add.u32 c0, c0, %10;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 " TM2_T16 ", [c0];\n\t"
      "mov.b64 ww, {w0, w0};\n\t"
      "tcgen05.wait::ld.sync.aligned;\n\t"
      TM2_PACK(x, t)
      TM2_FMA8(ww, x)
      "TM2_PAIRS:\n\t"
      "shr.u32 n2, %9, 1;\n\t"
      "setp.eq.u32 p, n2, 0;\n\t"
      "@p bra TM2_DONE;\n\t"
      "ld.shared.v2.u32 {c0, w0}, [%8];\n\t"
      "ld.shared.v2.u32 {c1, w1}, [%8+8];\n\t"
      "TM2_LOOP:\n\t"
      "add.u32 c0, c0, %10;\n\t"
      "add.u32 c1, c1, %10;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 " TM2_T16 ", [c0];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 " TM2_U16 ", [c1];\n\t"
      "mov.b64 ww, {w0, w0};\n\t"
      "mov.b64 vv, {w1, w1};\n\t"
      "ld.shared.v2.u32 {c0, w0}, [%8+16];\n\t"
      "ld.shared.v2.u32 {c1, w1}, [%8+24];\n\t"
      "add.u32 %8, %8, 16;\n\t"
      "tcgen05.wait::ld.sync.aligned;\n\t"
      TM2_PACK(x, t)
      TM2_PACK(y, u)
      TM2_FMA8(ww, x)
      TM2_FMA8(vv, y)
      "sub.u32 n2, n2, 1;\n\t"
      "setp.ne.u32 p, n2, 0;\n\t"
      "@p bra TM2_LOOP;\n\t"
      "TM2_DONE:\n\t"
      "}"
      : "+l"(a[0]), "+l"(a[1]), "+l"(a[2]), "+l"(a[3]), "+l"(a[4]), "+l"(a[5]), "+l"(a[6]), "+l"(a[7]), "+r"(rp)
      : "r"(n), "r"(base)
      : "memory");
}

// ring position of a chunk: unit, chunk inside the unit, shared-memory stage and its phase parity
struct Tm2Pos {
  int u, c;
  unsigned st, ph;
};
__device__ __forceinline__ void tm2_advance(const TmParams &p, Tm2Pos &x) {
  if (++x.c == p.nchunks) {
    x.c = 0;
    x.u += (int)gridDim.x;
  }
  if (++x.st == (unsigned)p.NS) {
    x.st = 0;
    x.ph ^= 1u;
  }
}

// ---- loader: the chunk at ring position `ld` global -> shared memory (4-byte cp.async driven by the per-unit table) --
struct Tm2Load {
  Tm2Pos pos;
  int count;      // chunks issued so far
  int tab_unit;   // unit the loader table was built for
  int chan0;      // first input channel of the unit's conv group
  const int2 *rt; // the unit's row of the region table
  int2 r;         // region {offset, length} of `pos` (fetched one call ahead: a dependent global load otherwise)
};
template <int NL, int BAR>
__device__ __forceinline__ void tm2_unit_setup(const TmParams &p, int num, unsigned char *smem_raw, Tm2Load &ls, int lw, int lane) {
  const TmUnit uc = tm_decode_unit(p, ls.pos.u);
  ls.tab_unit = ls.pos.u;
  ls.chan0 = uc.cg * p.Cg;
  ls.rt = p.rtab + ((size_t)uc.cg * p.ogroups + uc.og) * p.nchunks;
  int2 *ltab = reinterpret_cast<int2 *>(smem_raw + p.ltab_off);
  asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NL * 32) : "memory");  // every loader warp is done with the previous unit's table
  const int tile_start = uc.tile * p.TILE;
  const int R0 = tile_start / p.PW;
  const int nrows = (tile_start + p.SW - 1) / p.PW - R0 + 1;
  const int nxb = (p.PW + 31) >> 5;
  const int lpr = 1 << p.lpr_shift;
  for (int e = lw * 32 + lane; e < p.ltab_n * 32; e += NL * 32) {
    const int j = e >> 5, l = e & 31;
    const int jr = j / nxb, xb = j - jr * nxb;
    const int row = jr * p.RO + (l >> p.lpr_shift), x = (l & (lpr - 1)) + 32 * xb;
    int2 ent = make_int2(-1, -1);
    if (row < nrows && x < p.PW) {
      const int R = R0 + row;
      const int d = R * p.PW + x - tile_start;
      if (d >= 0 && d < p.SW) {
        const int n = R / p.IMGR, yy = R - n * p.IMGR;
        ent.y = (int)((tm_skew((unsigned)d >> 2) << 4) + (((unsigned)d & 3u) << 2));
        if (n < num && yy >= p.pad_h && x >= p.pad_w) ent.x = ((n * p.C * p.H + (yy - p.pad_h)) * p.W + (x - p.pad_w)) * 4;  // byte offset
      }
    }
    ltab[e] = ent;
  }
  asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NL * 32) : "memory");
}
template <int NL, int BAR>
__device__ __forceinline__ void tm2_issue_load(const TmParams &p, int num, const float *__restrict__ bottom, unsigned char *smem_raw,
                                               unsigned smem_base, Tm2Load &ls, int lw, int lane) {
  TM_EV(7);
  if (ls.pos.u != ls.tab_unit) {
    tm2_unit_setup<NL, BAR>(p, num, smem_raw, ls, lw, lane);
    ls.r = ls.rt[ls.pos.c];
  }
  const unsigned smem_full = smem_base, smem_empty = smem_base + 8 * kTmMaxStages;
  if (ls.count >= p.NS) tm_mbar_wait(smem_empty + 8 * ls.pos.st, ls.pos.ph ^ 1u, 1, p.dbg);
  TM_EV(8);
  const int2 *ltab = reinterpret_cast<const int2 *>(smem_raw + p.ltab_off);
  const unsigned stage_addr = smem_base + p.stage0_off + ls.pos.st * (unsigned)p.stage_bytes;
  const int nch = min(p.CI, p.Cg - ls.pos.c * p.CI);
  const size_t HW4 = (size_t)(p.H * p.W) * 4;
  const char *src0 = reinterpret_cast<const char *>(bottom) + (size_t)(ls.chan0 + ls.pos.c * p.CI) * HW4;
  const unsigned row_bytes = (unsigned)p.SWP * 4u;
#define TM2_CP4(K) asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + (K) * row_bytes), "l"(src + (K) * sstep), "r"(nbytes))
#pragma unroll 1
  for (int j = lw; j < ((p.skip & 4) ? 0 : p.ltab_n); j += NL) {  // this warp's table steps, all channels of the chunk
    const int2 e = ltab[j * 32 + lane];
    if (e.y >= 0) {
      const char *src = e.x >= 0 ? src0 + (unsigned)e.x : reinterpret_cast<const char *>(bottom);
      const unsigned nbytes = e.x >= 0 ? 4u : 0u;
      const size_t sstep = e.x >= 0 ? HW4 : 0;
      unsigned dst = stage_addr + (unsigned)e.y;
      int ch = nch;
#pragma unroll 1
      for (; ch >= 4; ch -= 4) {  // four copies back to back (ptxas pads a lone LDGSTS with dummy LDS)
        TM2_CP4(0);
        TM2_CP4(1);
        TM2_CP4(2);
        TM2_CP4(3);
        src += 4 * sstep;
        dst += 4 * row_bytes;
      }
      if (ch & 2) {
        TM2_CP4(0);
        TM2_CP4(1);
        src += 2 * sstep;
        dst += 2 * row_bytes;
      }
      if (ch & 1) TM2_CP4(0);
    }
  }
#undef TM2_CP4
  {  // record region of this (pass, chunk): contiguous 16-byte async copies
    const uint4 *src = p.prog + ls.r.x;
    const unsigned dst = stage_addr + p.in_bytes;
#pragma unroll 1
    for (int i = lw * 32 + lane; i < ls.r.y; i += NL * 32) tm_cp_async16(dst + 16u * i, src + i);
  }
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_full + 8 * ls.pos.st) : "memory");
  ++ls.count;
  tm2_advance(p, ls.pos);
  if (ls.pos.u == ls.tab_unit) ls.r = ls.rt[ls.pos.c];  // next call's region (same unit: same table row)
  TM_EV(9);
}


// ---- producer warps (NPW / 4 per TMEM lane quadrant): shared memory -> TMEM windows ------------------------------------
// NPW producer warps = NPW / 4 per quadrant; pw = producer warp index, quadrant = pw % 4, pa = pw / 4 = which share of a
// slot group's 16-column blocks this warp fills.
template <int T, int NCW, int NPW, int FB, int LC>
__device__ __forceinline__ void tm_producer_loop(const TmParams &p, int num, const float *__restrict__ bottom, int nunits,
                                                 unsigned char *smem_raw, unsigned smem_base, uint32_t tbase, int pw, int lane) {
  constexpr int NPQ = NPW / 4;
  const int q = pw & 3, pa = pw >> 2;
  const unsigned smem_full = smem_base, smem_empty = smem_base + 8 * kTmMaxStages;
  const unsigned tm_full = smem_base + 16 * kTmMaxStages + (unsigned)q * 8 * kTmMaxSlots;
  const unsigned tm_empty = tm_full + 4 * 8 * kTmMaxSlots;
  const uint32_t tq = tbase + ((uint32_t)(q * 32) << 16);
  const int my_units = blockIdx.x < (unsigned)nunits ? (nunits - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total = my_units * p.nchunks;
  unsigned slot = 0, round = 0;  // TMEM ring position of the next slot group to fill
  unsigned gsel = 0;             // slot groups seen so far (which producer of the quadrant fills the next one)
  const int nblocks = p.SLOTW >> 4;  // 16-column blocks per channel window

  // lane's first chunk of a window (before the skew): lane * T/4; a 16-column block never straddles a pad chunk
  const unsigned lane_chunk = (unsigned)lane * (T / 4);
  Tm2Load ls;
  ls.pos = {(int)blockIdx.x, 0, 0u, 0u};
  ls.count = 0;
  ls.tab_unit = -1;
  const int LA = p.NS - 2;  // chunks the loads run ahead: the stage being refilled was released a whole chunk ago
  if (!LC)
    for (int it = 0; it < LA && it < total; ++it) tm2_issue_load<NPW, 2>(p, num, bottom, smem_raw, smem_base, ls, pw, lane);
  Tm2Pos cur = {(int)blockIdx.x, 0, 0u, 0u};  // ring position of chunk `it` (no divisions in the loop)
  for (int it = 0; it < total; ++it) {
    if (!LC && ls.count < total) tm2_issue_load<NPW, 2>(p, num, bottom, smem_raw, smem_base, ls, pw, lane);
    const int s = (int)cur.st;
    TM_EV(11);
    tm_mbar_wait(smem_full + 8 * s, cur.ph, 2, p.dbg);
    TM_EV(12);
    const int c = cur.c;
    tm2_advance(p, cur);
    const int nch = min(p.CI, p.Cg - c * p.CI);
    const unsigned stage_addr = smem_base + p.stage0_off + (unsigned)s * p.stage_bytes;
    for (int ch0 = 0; ch0 < nch; ch0 += p.CHS) {
      if (NPQ > 1 && (int)(gsel++ & (NPQ - 1)) != pa) {  // the quadrant's producers take slot groups in turn
        if (++slot == (unsigned)p.NSLOT) {
          slot = 0;
          ++round;
        }
        continue;
      }
      if (round > 0) {
        tm_mbar_wait(tm_empty + 8 * slot, (round - 1) & 1u, 3, p.dbg);
        tm_fence_after();
      }
      TM_EV(13);
      const int chn = min(p.CHS, nch - ch0);
      unsigned row_addr = stage_addr + (unsigned)(ch0 * p.SWP) * 4u;
      uint32_t tcol = tq + slot * (unsigned)(p.CHS * p.SLOTW);
      for (int k = 0; k < ((p.skip & 1) ? 0 : chn); ++k, row_addr += (unsigned)p.SWP * 4u, tcol += (unsigned)p.SLOTW) {
        // the channel window in 16-column blocks, FB per round: all the 128-bit loads first, then the TMEM stores
        int b0 = 0;
#define TM_LDS128(B, J) asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4+" #J "*16];" : "=r"(v[B][4 * J]), "=r"(v[B][4 * J + 1]), "=r"(v[B][4 * J + 2]), "=r"(v[B][4 * J + 3]) : "r"(a))
#define TM_LDBLOCK(B, BLK) { const unsigned a = row_addr + (tm_skew(lane_chunk + 4u * (unsigned)(BLK)) << 4); TM_LDS128(B, 0); TM_LDS128(B, 1); TM_LDS128(B, 2); TM_LDS128(B, 3); }
        for (; b0 + (FB - 1) < nblocks; b0 += FB) {
          uint32_t v[FB][16];
#pragma unroll
          for (int b = 0; b < FB; ++b) TM_LDBLOCK(b, b0 + b)
#pragma unroll
          for (int b = 0; b < FB; ++b) tm_st16(tcol + 16u * (unsigned)(b0 + b), v[b]);
        }
        if (FB > 1 && b0 < nblocks) {  // tail: 1 .. FB-1 blocks
          uint32_t v[FB][16];
#pragma unroll
          for (int b = 0; b < FB - 1; ++b)
            if (b0 + b < nblocks) TM_LDBLOCK(b, b0 + b)
#pragma unroll
          for (int b = 0; b < FB - 1; ++b)
            if (b0 + b < nblocks) tm_st16(tcol + 16u * (unsigned)(b0 + b), v[b]);
        }
#undef TM_LDBLOCK
#undef TM_LDS128
      }
      tm_wait_st();
      TM_EV(14);
      tm_fence_before();
      __syncwarp();
      if (lane == 0) tm_mbar_arrive(tm_full + 8 * slot);
      if (++slot == (unsigned)p.NSLOT) {
        slot = 0;
        ++round;
      }
    }
    __syncwarp();
    if (lane == 0) tm_mbar_arrive(smem_empty + 8 * s);  // the producer's share; the compute warps still read the records
  }
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
// warps 0 .. NCW-1: compute (TMEM lane quadrant = wid % 4, channel block = wid); warps NCW .. NCW+NPW-1: producers.
template <int T, int OT, int NCW, int NPW, int CREGS, int PREGS, int LC>
__global__ void __launch_bounds__((NCW + NPW) * 32, 1)
    sconv_tmem_kernel(const TmParams p, int num, const float *__restrict__ bottom, const float *__restrict__ bias, int fuse_relu,
                      float *__restrict__ top, int nunits) {
  // setmaxnreg moves registers inside the CTA's OWN pool (what the launch allocated: R0 per thread), not the whole
  // register file: the compute warps can only grow by what the producer warps give back.  (Measured the hard way: a
  // split that needed the SM's unallocated registers left the last compute warpgroup spinning in setmaxnreg.inc.)
  constexpr int R0 = 65536 / ((NCW + NPW) * 32) / 8 * 8;
  static_assert(NCW * (CREGS - R0) <= NPW * (R0 - PREGS), "setmaxnreg split exceeds the CTA register pool");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const unsigned smem_full = smem_base, smem_empty = smem_base + 8 * kTmMaxStages;
  const unsigned tm_full0 = smem_base + 16 * kTmMaxStages, tm_empty0 = tm_full0 + 4 * 8 * kTmMaxSlots;
  uint32_t *tbase_slot = reinterpret_cast<uint32_t *>(smem_raw + (2 * kTmMaxStages + 8 * kTmMaxSlots) * 8);

  tm_ev_init();
  if (tid == 0) {
    for (int s = 0; s < p.NS; ++s) {
      tm_mbar_init(smem_full + 8 * s, (LC ? NCW : NPW) * 32);  // every loader thread arrives through its cp.asyncs
      tm_mbar_init(smem_empty + 8 * s, NCW + NPW);   // every warp releases the stage
    }
    for (int qq = 0; qq < 4; ++qq)
      for (int s = 0; s < p.NSLOT; ++s) {
        tm_mbar_init(tm_full0 + (qq * kTmMaxSlots + s) * 8, 1);
        tm_mbar_init(tm_empty0 + (qq * kTmMaxSlots + s) * 8, NCW / 4);
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"l"((uint64_t)__cvta_generic_to_shared(tbase_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tm_fence_before();
  __syncthreads();
  tm_fence_after();
  const uint32_t tbase = *tbase_slot;
  if (tbase != 0u) {  // the records hold absolute addresses: the CTA owns all 512 columns, so its base is column 0
    if (p.dbg && tid == 0 && atomicCAS(p.dbg, 0, 9) == 0) p.dbg[1] = (int)tbase;
    __trap();
  }

  if (wid >= NCW) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PREGS) : "memory");
    tm_producer_loop<T, NCW, NPW, (PREGS >= 72 ? 3 : PREGS >= 64 ? 2 : 1), LC>(p, num, bottom, nunits, smem_raw, smem_base, tbase, wid - NCW, lane);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CREGS) : "memory");
    const int q = wid & 3;
    const unsigned tm_full = tm_full0 + (unsigned)q * 8 * kTmMaxSlots, tm_empty = tm_empty0 + (unsigned)q * 8 * kTmMaxSlots;
    const unsigned slot_cols = (unsigned)(p.CHS * p.SLOTW);
    unsigned long long acc[OT][T / 2];
    unsigned st = 0, ph = 0;          // shared-memory stage ring
    unsigned slot = 0, sph = 0;       // TMEM slot ring
    // LC: the compute warps are also the loader (at low weight density they idle on the producers, which never idle)
    const int my_units = blockIdx.x < (unsigned)nunits ? (nunits - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total = my_units * p.nchunks;
    const int LA = p.NS - 2;  // chunks the loads run ahead: the stage being refilled was released a whole chunk ago
    Tm2Load ls;
    ls.pos = {(int)blockIdx.x, 0, 0u, 0u};
    ls.count = 0;
    ls.tab_unit = -1;
    for (int k = 0; LC && k < LA && k < total; ++k) tm2_issue_load<NCW, 1>(p, num, bottom, smem_raw, smem_base, ls, wid, lane);
#pragma unroll 1
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
#pragma unroll
      for (int o = 0; o < OT; ++o)
#pragma unroll
        for (int k = 0; k < T / 2; ++k) acc[o][k] = 0ull;
#pragma unroll 1
      for (int c = 0; c < p.nchunks; ++c) {
        if (LC && ls.count < total) tm2_issue_load<NCW, 1>(p, num, bottom, smem_raw, smem_base, ls, wid, lane);
        TM_EV(1);
        tm_mbar_wait(smem_full + 8 * st, ph, 4, p.dbg);
        TM_EV(2);
        const unsigned region = smem_base + p.stage0_off + st * p.stage_bytes + p.in_bytes;
        unsigned rp;  // this warp's records (8 bytes each: {TMEM column inside the slot group, fp32 weight})
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rp) : "r"(region + 4u * (unsigned)wid));
        rp += region;
        unsigned col = 0, wbits = 0;  // the NEXT record, always one tap ahead (its load overlaps the current tap's FMAs)
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(col), "=r"(wbits) : "r"(rp));
        unsigned cp = region + p.hdr_counts_off + (unsigned)(wid * p.nsg) * 8u;  // per slot group: 8 tap counts (one byte per o)
        const int nch = min(p.CI, p.Cg - c * p.CI);
#pragma unroll 1
        for (int ch0 = 0; ch0 < nch; ch0 += p.CHS) {
          unsigned cnt_lo, cnt_hi;
          asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cnt_lo), "=r"(cnt_hi) : "r"(cp));
          cp += 8;
          tm_mbar_wait(tm_full + 8 * slot, sph, 5, p.dbg);
          tm_fence_after();
          TM_EV(3);
          const uint32_t tslot = slot * slot_cols;  // (the records carry the quadrant's lane bits: absolute addresses)
          if ((cnt_lo | cnt_hi) != 0u && !(p.skip & 2)) {
#pragma unroll
            for (int o = 0; o < OT; ++o) {
              unsigned n = ((o < 4 ? cnt_lo : cnt_hi) >> (8 * (o & 3))) & 0xffu;
#ifdef ESCORT_TM_SECTION_WALK  // measured: the C++ walk below (next record prefetched across sections) is 3-4 % faster here
              if constexpr (T == 16) {
                tm2_section16_r(acc[o], rp, n, tslot);
                continue;
              }
#endif
              // (col, wbits) always hold the NEXT record; an odd tap first, then pairs (half the loop overhead per tap)
              if (n & 1u) {
                const uint32_t taddr = tslot + col;
                const unsigned w = wbits;
                rp += 8;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(col), "=r"(wbits) : "r"(rp));
                tm_tap<T>(acc[o], taddr, w);
              }
#pragma unroll 1
              for (n >>= 1; n > 0; --n) {
                const uint32_t ta = tslot + col;
                const unsigned wa = wbits;
                unsigned colb, wb;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+8];" : "=r"(colb), "=r"(wb) : "r"(rp));
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+16];" : "=r"(col), "=r"(wbits) : "r"(rp));
                rp += 16;
                tm_tap2<T>(acc[o], ta, wa, tslot + colb, wb);
              }
            }
          }
          TM_EV(4);
          tm_fence_before();
          __syncwarp();
          if (lane == 0) tm_mbar_arrive(tm_empty + 8 * slot);
          if (++slot == (unsigned)p.NSLOT) {
            slot = 0;
            sph ^= 1u;
          }
        }
        __syncwarp();
        if (lane == 0) tm_mbar_arrive(smem_empty + 8 * st);
        if (++st == (unsigned)p.NS) {
          st = 0;
          ph ^= 1u;
        }
      }
      // ---- epilogue: bias + ReLU, lane-major tile -> (swizzled) shared memory -> coalesced predicated stores ----
      TM_EV(5);
      const TmUnit uc = tm_decode_unit(p, u);
      const int blk = uc.og * NCW + wid;
      if (blk < p.nblk) {
        const int HoWo = p.Ho * p.Wo;
        int ooff[T];  // where position i * 32 + lane of the tile goes (image offset + pixel), -1 = no such output
        {
          // (image, row, column) of the lane's first position by division, then 32 positions per step incrementally
          const int pos = uc.tile * p.TILE + lane;  // the host checks num * IMG < 2^31
          int n = pos / p.IMG;
          const int r = pos - n * p.IMG;
          int y = r / p.PW, x = r - y * p.PW;
          const int dy = 32 / p.PW, dx = 32 - dy * p.PW;
#pragma unroll
          for (int i = 0; i < T; ++i) {
            ooff[i] = (n < num && y < p.Ho && x < p.Wo) ? (n * p.M) * HoWo + y * p.Wo + x : -1;
            x += dx;
            y += dy;
            if (x >= p.PW) {
              x -= p.PW;
              ++y;
            }
            while (y >= p.IMGR) {
              y -= p.IMGR;
              ++n;
            }
          }
        }
        const unsigned ost = smem_base + p.ostage_off + (unsigned)wid * (unsigned)(36 * T * 4);  // TILE floats + skew padding
#pragma unroll
        for (int o = 0; o < OT; ++o) {
          const int oc = p.oc_list[((size_t)uc.cg * p.nblk + blk) * OT + o];
          if (oc < 0) continue;
          const float b = bias ? __ldg(bias + oc) : 0.f;
#pragma unroll
          for (int j = 0; j < T / 4; ++j) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              v[2 * e] = __uint_as_float((unsigned)(acc[o][2 * j + e] & 0xffffffffull)) + b;
              v[2 * e + 1] = __uint_as_float((unsigned)(acc[o][2 * j + e] >> 32)) + b;
            }
            if (fuse_relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            const unsigned ch = tm_skew((unsigned)lane * (T / 4) + (unsigned)j);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ost + (ch << 4)), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
          }
          __syncwarp();
          float *out = top + (size_t)oc * HoWo;
#pragma unroll
          for (int i = 0; i < T; ++i) {
            const unsigned pl = (unsigned)(i * 32 + lane);
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(ost + (tm_skew(pl >> 2) << 4) + ((pl & 3u) << 2)) : "memory");
            if (ooff[i] >= 0) out[ooff[i]] = v;
          }
          __syncwarp();
        }
      }
      TM_EV(6);
    }
  }
  tm_fence_before();
  __syncthreads();
  tm_ev_fini();
  if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

#endif  // !ESCORT_TMEM_HOST_ONLY

}  // namespace escort
