// tmem_bench.cu -- primitive-rate microbenchmarks that decided the round-2 forward inner loop.
// Measures, per SM, with W warps resident: tcgen05.ld / tcgen05.st throughput (TMEM <-> registers), LDS.128 throughput,
// and the rate of the candidate loops "X window from TMEM (or shared memory) + 16 FFMA2".
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bench tmem_bench.cu ; run: ./tmem_bench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

#define R8(r, b) "=r"(r[b+0]), "=r"(r[b+1]), "=r"(r[b+2]), "=r"(r[b+3]), "=r"(r[b+4]), "=r"(r[b+5]), "=r"(r[b+6]), "=r"(r[b+7])
#define I8(r, b) "r"(r[b+0]), "r"(r[b+1]), "r"(r[b+2]), "r"(r[b+3]), "r"(r[b+4]), "r"(r[b+5]), "r"(r[b+6]), "r"(r[b+7])

__device__ __forceinline__ void ldtm32(uint32_t (&r)[32], uint32_t ta) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24)
      : "r"(ta)
      : "memory");
}
__device__ __forceinline__ void ldtm16(uint32_t *r, uint32_t ta) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : R8(r, 0), R8(r, 8)
               : "r"(ta)
               : "memory");
}
__device__ __forceinline__ void ldtm8(uint32_t *r, uint32_t ta) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : R8(r, 0) : "r"(ta) : "memory");
}
__device__ __forceinline__ void sttm32(uint32_t ta, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(ta), I8(r, 0), I8(r, 8), I8(r, 16), I8(r, 24)
      : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void lds32f(uint32_t (&r)[32], uint32_t sa) {  // 8 x LDS.128, conflict free
#pragma unroll
  for (int k = 0; k < 8; ++k)
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[4 * k]), "=r"(r[4 * k + 1]), "=r"(r[4 * k + 2]), "=r"(r[4 * k + 3]) : "r"(sa + 512u * k) : "memory");
}
// 16 packed FMAs: acc pairs += {w,w} * x pairs
__device__ __forceinline__ void fma16(uint64_t (&acc)[16], uint64_t w2, const uint32_t (&r)[32]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint64_t x;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "r"(r[2 * i]), "r"(r[2 * i + 1]));
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(w2), "l"(x));
  }
}
__device__ __forceinline__ void fma32s(float (&acc)[32], float w, const uint32_t (&r)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(w), "f"(__uint_as_float(r[i])));
}

enum { M_LD32 = 0, M_LD32X2, M_LD16, M_LD8, M_ST32, M_LD32_FMA2, M_LDS_FMA2, M_FMA2, M_LD32_FMA, M_MIX, M_LDS, M_LD32_FMA2_DEP, M_COUNT };
static const char *kNames[] = {"ldtm.x32 + wait", "2 x ldtm.x32 + wait", "ldtm.x16 + wait", "ldtm.x8 + wait", "sttm.x32 + wait", "ldtm.x32 | 16 FFMA2 (dbl buf)",
                               "8 LDS.128 | 16 FFMA2 (dbl buf)", "16 FFMA2 only", "ldtm.x32 | 32 FFMA (dbl buf)", "mix: 3/4 warps ldtm|FFMA2, 1/4 warps LDS->sttm fill",
                               "8 LDS.128 only", "ldtm.x32 ; wait ; 16 FFMA2 (single buf)"};

template <int MODE>
__global__ void __launch_bounds__(512, 1) bench(int iters, float *out, unsigned long long *cyc) {
  __shared__ uint32_t tbase_s;
  extern __shared__ __align__(16) unsigned char dsm[];
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  for (int i = tid; i < 8192; i += blockDim.x) reinterpret_cast<float *>(dsm)[i] = 1.0f + i * 1e-6f;
  for (int i = tid; i < 1024; i += blockDim.x) {
    uint32_t *r = reinterpret_cast<uint32_t *>(dsm + 32768) + 4 * i;
    r[0] = (uint32_t)((i * 37 + 5) % 224); r[1] = 0; r[2] = r[3] = __float_as_uint(1e-3f + 1e-6f * i);
  }
  if (wid == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"l"((uint64_t)__cvta_generic_to_shared(&tbase_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase_s + ((uint32_t)((wid & 3) * 32) << 16);
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dsm) + lane * 16u;
  uint32_t rA[32], rB[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) rA[i] = __float_as_uint(1.0f + 0.001f * i + 0.01f * lane), rB[i] = rA[i];
  if (wid < 4) {  // initialise all 512 columns of this sub-partition's lanes
    for (int c = 0; c < 512; c += 32) sttm32(tb + c, rA);
    wait_st();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint64_t acc[16];
  float accs[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) accs[i] = 0.f;
  const float wf = 1.0e-3f;
  uint64_t w2;
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(w2) : "f"(wf));
  const unsigned long long t0 = clock64();
  if (MODE == M_LD32) {
#pragma unroll 1
    for (int it = 0; it < iters; ++it) { ldtm32(rA, tb + ((it * 32) & 448)); wait_ld(); }
  } else if (MODE == M_LD32X2) {
#pragma unroll 1
    for (int it = 0; it < iters; it += 2) { ldtm32(rA, tb + ((it * 32) & 448)); ldtm32(rB, tb + ((it * 32 + 32) & 448)); wait_ld(); }
  } else if (MODE == M_LD16) {
#pragma unroll 1
    for (int it = 0; it < iters; ++it) { ldtm16(rA, tb + ((it * 16) & 448)); wait_ld(); }
  } else if (MODE == M_LD8) {
#pragma unroll 1
    for (int it = 0; it < iters; ++it) { ldtm8(rA, tb + ((it * 8) & 448)); wait_ld(); }
  } else if (MODE == M_ST32) {
#pragma unroll 1
    for (int it = 0; it < iters; ++it) { sttm32(tb + ((it * 32) & 448), rA); wait_st(); }
  } else if (MODE == M_LD32_FMA2 || (MODE == M_MIX && (wid & 12) != 12)) {
    // the candidate loop: a 16-byte record {TMEM column, -, w, w} per nonzero, read as one broadcast LDS.128
    const uint32_t rec = (uint32_t)__cvta_generic_to_shared(dsm) + 32768u;
    uint32_t c0, p0, c1, p1;
    uint64_t wa, wb;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%3]; ld.shared.b64 %2, [%3+8];" : "=r"(c0), "=r"(p0), "=l"(wa) : "r"(rec) : "memory");
    ldtm32(rA, tb + c0);
#pragma unroll 1
    for (int it = 0; it < iters; it += 2) {
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%3]; ld.shared.b64 %2, [%3+8];" : "=r"(c1), "=r"(p1), "=l"(wb) : "r"(rec + (((it + 1) & 1023) << 4)) : "memory");
      wait_ld();
      ldtm32(rB, tb + c1);
      fma16(acc, wa, rA);
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%3]; ld.shared.b64 %2, [%3+8];" : "=r"(c0), "=r"(p0), "=l"(wa) : "r"(rec + (((it + 2) & 1023) << 4)) : "memory");
      wait_ld();
      ldtm32(rA, tb + c0);
      fma16(acc, wb, rB);
    }
    wait_ld();
  } else if (MODE == M_MIX) {  // fill warps: window from shared memory -> TMEM
#pragma unroll 1
    for (int it = 0; it < iters; ++it) { lds32f(rA, sa + ((it & 7) << 12)); sttm32(tb + 256 + ((it * 32) & 224), rA); wait_st(); }
  } else if (MODE == M_LD32_FMA2_DEP) {
#pragma unroll 1
    for (int it = 0; it < iters; ++it) { ldtm32(rA, tb + ((it * 37 + 5) & 255)); wait_ld(); fma16(acc, w2, rA); }
  } else if (MODE == M_LDS_FMA2) {
    lds32f(rA, sa);
#pragma unroll 1
    for (int it = 0; it < iters; it += 2) {
      lds32f(rB, sa + (((it + 1) & 7) << 12));
      fma16(acc, w2, rA);
      lds32f(rA, sa + (((it + 2) & 7) << 12));
      fma16(acc, w2, rB);
    }
  } else if (MODE == M_LDS) {
#pragma unroll 1
    for (int it = 0; it < iters; ++it) lds32f(rA, sa + ((it & 7) << 12));
  } else if (MODE == M_FMA2) {
#pragma unroll 1
    for (int it = 0; it < iters; ++it) fma16(acc, w2, rA);
  } else if (MODE == M_LD32_FMA) {
    ldtm32(rA, tb);
#pragma unroll 1
    for (int it = 0; it < iters; it += 2) {
      wait_ld();
      ldtm32(rB, tb + ((it * 37 + 5) & 255));
      fma32s(accs, wf, rA);
      wait_ld();
      ldtm32(rA, tb + ((it * 37 + 42) & 255));
      fma32s(accs, wf, rB);
    }
    wait_ld();
  }
  const unsigned long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += __uint_as_float((uint32_t)acc[i]) + __uint_as_float((uint32_t)(acc[i] >> 32));
#pragma unroll
  for (int i = 0; i < 32; ++i) s += accs[i] + __uint_as_float(rA[i]) + __uint_as_float(rB[i]);
  out[blockIdx.x * blockDim.x + tid] = s;
  if (lane == 0) cyc[blockIdx.x * 32 + wid] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  (void)nw;
  if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase_s) : "memory");
}

template <int MODE>
static void run(int warps, int iters, int sms, float *out, unsigned long long *cyc, double ghz_nominal) {
  CK(cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    bench<MODE><<<sms, warps * 32, 65536>>>(iters, out, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  static unsigned long long h[148 * 32];
  CK(cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms * 32, cudaMemcpyDeviceToHost));
  double cmax = 0, csum = 0;
  for (int b = 0; b < sms; ++b)
    for (int w = 0; w < warps; ++w) { csum += (double)h[b * 32 + w]; if ((double)h[b * 32 + w] > cmax) cmax = (double)h[b * 32 + w]; }
  const double cavg = csum / (sms * warps);
  // per-iteration payload: 4096 B for x32 transfers (scaled for x16 / x8), 32 lanes * 32 FMAs for the FMA loops
  double bytes_it = 4096.0;
  if (MODE == M_LD16) bytes_it = 2048.0;
  if (MODE == M_LD8) bytes_it = 1024.0;
  int work_warps = warps;
  if (MODE == M_MIX) work_warps = warps - warps / 4;
  const double b_per_clk_sm = bytes_it * iters * work_warps / cmax;
  const bool fma = (MODE == M_LD32_FMA2 || MODE == M_LDS_FMA2 || MODE == M_FMA2 || MODE == M_LD32_FMA || MODE == M_MIX || MODE == M_LD32_FMA2_DEP);
  const double flop = 2.0 * 32 * 32 * (double)iters * work_warps * sms;
  printf("%-52s warps/SM %2d  cycles/iter/warp avg %7.1f max %7.1f  %7.1f B/clk/SM", kNames[MODE], warps, cavg / iters, cmax / iters, b_per_clk_sm);
  if (fma) printf("  %6.2f TFLOP/s (events)  %5.1f FMA-lanes/clk/SM", flop / (best * 1e-3) / 1e12, 32.0 * 32 * iters * work_warps / cmax);
  printf("  %.3f ms\n", best);
  (void)ghz_nominal;
}

int main(int argc, char **argv) {
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float *out;
  unsigned long long *cyc;
  CK(cudaMalloc(&out, sizeof(float) * sms * 512));
  CK(cudaMalloc(&cyc, sizeof(unsigned long long) * sms * 32));
  const int iters = argc > 1 ? atoi(argv[1]) : 4096;
  printf("SMs %d, iters %d\n", sms, iters);
  const int ws[] = {1, 4, 8, 12, 16};
  for (int wi = 0; wi < 5; ++wi) {
    const int w = ws[wi];
    run<M_LD32>(w, iters, sms, out, cyc, 1.965);
    run<M_LD32X2>(w, iters, sms, out, cyc, 1.965);
    run<M_LD16>(w, iters, sms, out, cyc, 1.965);
    run<M_LD8>(w, iters, sms, out, cyc, 1.965);
    run<M_ST32>(w, iters, sms, out, cyc, 1.965);
    run<M_LDS>(w, iters, sms, out, cyc, 1.965);
    run<M_FMA2>(w, iters, sms, out, cyc, 1.965);
    run<M_LD32_FMA2>(w, iters, sms, out, cyc, 1.965);
    run<M_LD32_FMA2_DEP>(w, iters, sms, out, cyc, 1.965);
    run<M_LD32_FMA>(w, iters, sms, out, cyc, 1.965);
    run<M_LDS_FMA2>(w, iters, sms, out, cyc, 1.965);
    if (w >= 4) run<M_MIX>(w, iters, sms, out, cyc, 1.965);
    printf("\n");
  }
  return 0;
}
