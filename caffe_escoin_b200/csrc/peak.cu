// peak.cu -- register-resident FP32 FMA microbenchmark: the measured P_fp32 roofline denominator
// ("the builder must measure it").  variant 0: scalar FFMA with one operand shared
// across the accumulators (the access pattern of the sparse-conv inner loop); variant 1: packed
// fma.rn.f32x2 (FFMA2, new in sm_100); variant 2: scalar FFMA, three distinct register operands.
#include "common.cuh"

namespace escort {

template <int VARIANT>
__global__ void __launch_bounds__(256) ffma_peak_kernel(float *out, int iters, float a0, float b0) {
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = (float)(threadIdx.x + i);
  float a = a0 + threadIdx.x * 1e-9f, b = b0;
  float xs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) xs[i] = b0 + i * 1e-7f;
  for (int it = 0; it < iters; ++it) {
    if (VARIANT == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = fmaf(a, xs[(i + r) & 7], acc[i]);
    } else if (VARIANT == 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          asm volatile(
              "{\n\t.reg .b64 ra, rb, rc;\n\t"
              "mov.b64 ra, {%2, %2};\n\t"
              "mov.b64 rb, {%3, %4};\n\t"
              "mov.b64 rc, {%0, %1};\n\t"
              "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
              "mov.b64 {%0, %1}, rc;\n\t}"
              : "+f"(acc[i]), "+f"(acc[i + 1])
              : "f"(a), "f"(xs[(i + r) & 7]), "f"(xs[(i + r + 1) & 7]));
        }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[(i + 1) & 31], xs[(i + r) & 7], acc[i]);
    }
    a += b;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace escort

using namespace escort;

extern "C" int escort_measure_fp32_peak(int variant, int iters, double *tflops_host, int *sm_count_host,
                                        int *clock_khz_host) {
  ESCORT_REQUIRE(tflops_host && iters > 0 && variant >= 0 && variant <= 2, "escort_measure_fp32_peak: bad arguments");
  int dev = 0, sms = 0, khz = 0;
  ESCORT_CUDA(cudaGetDevice(&dev));
  ESCORT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  ESCORT_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  const int blocks = sms * 8, threads = 256;
  float *out = nullptr;
  ESCORT_CUDA(cudaMalloc(&out, (size_t)blocks * threads * sizeof(float)));
  cudaEvent_t e0, e1;
  ESCORT_CUDA(cudaEventCreate(&e0));
  ESCORT_CUDA(cudaEventCreate(&e1));
  double best_ms = 1e30;
  for (int rep = 0; rep < 6; ++rep) {
    ESCORT_CUDA(cudaEventRecord(e0, 0));
    if (variant == 0) ffma_peak_kernel<0><<<blocks, threads>>>(out, iters, 1.0001f, 0.9999f);
    else if (variant == 1) ffma_peak_kernel<1><<<blocks, threads>>>(out, iters, 1.0001f, 0.9999f);
    else ffma_peak_kernel<2><<<blocks, threads>>>(out, iters, 1.0001f, 0.9999f);
    ESCORT_CUDA(cudaEventRecord(e1, 0));
    ESCORT_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    ESCORT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best_ms) best_ms = ms;
  }
  ESCORT_LAUNCH_CHECK();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  const double flops = 2.0 * 128.0 * (double)iters * (double)blocks * threads;
  *tflops_host = flops / (best_ms * 1e-3) / 1e12;
  if (sm_count_host) *sm_count_host = sms;
  if (clock_khz_host) *clock_khz_host = khz;
  return 0;
}
