// utccp_probe.cu -- what tcgen05.cp (smem -> TMEM copy engine) does with a given shared-memory matrix descriptor, and how
// fast.  Shared memory is filled with float(word index); after one copy every TMEM lane / column is read back, so the
// lane <- address mapping of each descriptor variant is observed, not assumed.  
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
#define R8(r, b) "=r"(r[b+0]), "=r"(r[b+1]), "=r"(r[b+2]), "=r"(r[b+3]), "=r"(r[b+4]), "=r"(r[b+5]), "=r"(r[b+6]), "=r"(r[b+7])

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off, uint32_t swz) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)(swz & 7) << 61;
  return d;
}

struct Cfg { int shape; uint32_t start_off, lbo, sbo, base_off, swz; };  // shape 0: 32x128b.warpx4, 1: 128x128b, 2: 128x256b

__global__ void __launch_bounds__(128, 1) probe(const Cfg *cfgs, int ncfg, float *out, int reps, unsigned long long *cyc) {
  extern __shared__ __align__(1024) unsigned char dsm[];
  __shared__ uint32_t tbase_s;
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x, wid = tid >> 5;
  float *sf = reinterpret_cast<float *>(dsm);
  for (int i = tid; i < 16384; i += blockDim.x) sf[i] = (float)i;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(dsm);
  const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"l"((uint64_t)__cvta_generic_to_shared(&tbase_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the async proxy (tcgen05.cp)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase_s, tl = tb + ((uint32_t)(wid * 32) << 16);
  uint32_t phase = 0;
  for (int k = 0; k < ncfg; ++k) {
    const Cfg c = cfgs[k];
    {  // sentinel
      uint32_t z[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) z[i] = __float_as_uint(-1.0f);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tl), "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned long long t0 = 0, t1 = 0;
    if (tid == 0) {
      const uint64_t d = make_desc(sbase + c.start_off, c.lbo, c.sbo, c.base_off, c.swz);
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (c.shape == 0) asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tb), "l"(d) : "memory");
        else if (c.shape == 1) asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(tb), "l"(d) : "memory");
        else asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tb), "l"(d) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb) : "memory");
    }
    {  // everyone waits for the copy
      asm volatile(
          "{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(mb), "r"(phase) : "memory");
      phase ^= 1;
    }
    if (tid == 0) { t1 = clock64(); cyc[k] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : R8(r, 0) : "r"(tl) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) out[(k * 128 + tid) * 8 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase_s) : "memory");
}

int main(int argc, char **argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 1;
  Cfg h[] = {
      {0, 0, 16, 128, 0, 0},      // A no swizzle, core matrices 128 B apart
      {0, 16, 16, 128, 0, 0},     // B start + 16
      {0, 0, 16, 1024, 0, 2},     // C 128B swizzle, atoms 1024 B apart
      {0, 16, 16, 1024, 0, 2},    // D chunk 1
      {0, 128, 16, 1024, 0, 2},   // E start one row down, base offset 0
      {0, 128, 16, 1024, 1, 2},   // F start one row down, base offset 1
      {0, 128 + 48, 16, 1024, 1, 2},  // G row 1, chunk 3, base offset 1
      {1, 0, 16, 128, 0, 0},      // H 128x128b no swizzle
      {1, 0, 16, 1024, 0, 2},     // I 128x128b swizzled
      {2, 0, 128 * 16, 128, 0, 0},  // J 128x256b no swizzle, LBO = 2048 (second 16-byte column 2 KB away)
      {2, 0, 16, 1024, 0, 2},     // K 128x256b swizzled
      {0, 0, 16, 256, 0, 0},      // L no swizzle, SBO 256
  };
  const int n = sizeof(h) / sizeof(h[0]);
  Cfg *d;
  float *out;
  unsigned long long *cyc;
  CK(cudaMalloc(&d, sizeof(h)));
  CK(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&out, sizeof(float) * n * 128 * 8));
  CK(cudaMalloc(&cyc, 8 * n));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024));
  probe<<<1, 128, 65536 + 1024>>>(d, n, out, reps, cyc);
  CK(cudaDeviceSynchronize());
  static float ho[16 * 128 * 8];
  unsigned long long hc[16];
  CK(cudaMemcpy(ho, out, sizeof(float) * n * 128 * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hc, cyc, 8 * n, cudaMemcpyDeviceToHost));
  for (int k = 0; k < n; ++k) {
    printf("cfg %c shape %d start+%u lbo %u sbo %u boff %u swz %u : %llu cycles for %d copies (%.1f / copy)\n", 'A' + k, h[k].shape, h[k].start_off, h[k].lbo,
           h[k].sbo, h[k].base_off, h[k].swz, hc[k], reps, (double)hc[k] / reps);
    if (reps > 1) continue;
    for (int l = 0; l < 128; ++l) {
      if (l >= 18 && l < 32) continue;
      if (l >= 36 && l < 64) continue;
      if (l >= 66 && l < 96) continue;
      if (l >= 98) continue;
      printf("  lane %3d:", l);
      for (int i = 0; i < 8; ++i) printf(" %6.0f", ho[(k * 128 + l) * 8 + i]);
      printf("\n");
    }
  }
  return 0;
}
