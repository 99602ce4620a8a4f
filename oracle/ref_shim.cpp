// ref_shim.cpp -- thin extern "C" wrapper that compiles the REFERENCE'S OWN CPU direct sparse conv
// (/root/reference/include/caffe/util/sconv.hpp, included from where it lies; nothing is copied) into
// oracle/_ref/libescort_ref.so.  TEST INFRASTRUCTURE ONLY: used to pin oracle/escort_oracle.c, to
// generate tests/golden/, and as bench.py's CPU baseline (cpu_baseline.kind == "reference").
//
// The header needs four extern globals (sconv.hpp:29,42) and a NOT_IMPLEMENTED macro; defining USE_ICC
// additionally enables the AVX2 register-blocked kernel sconv_unit_stride<WIDTH,K> under g++
// (unknown-pragma warnings only).  The (WIDTH,K) pairs dispatched below are exactly the ones the
// reference itself instantiates in caffe_cpu_blocked_sconv (src/caffe/util/math_functions.cpp:225-383).
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NOT_IMPLEMENTED abort()
#define USE_ICC 1
#include "caffe/util/sconv.hpp"

unsigned long long conv_cycles_of_this_batch[1024 * 16], transpose_cycle, pool_cycle;
int flop_cnt;

#define REF_API extern "C" __attribute__((visibility("default")))

// caffe_cpu_sconv_default<FUSE_RELU> (sconv.hpp:594-678). bias must be non-NULL (the reference
// dereferences it unconditionally, sconv.hpp:627,646).
REF_API void ref_sconv_default(int fuse_relu, const float *input_padded, int in_channels, int height,
                               int width, int pad_h, int pad_w, int stride_h, int stride_w, int dilation_h,
                               int dilation_w, const int *rowptr, const int *colidx, const float *values,
                               int kernel_h, int kernel_w, const float *bias, float *output,
                               int out_channels) {
  if (fuse_relu)
    caffe_cpu_sconv_default<true>(input_padded, in_channels, height, width, pad_h, pad_w, stride_h, stride_w,
                                  dilation_h, dilation_w, rowptr, colidx, values, kernel_h, kernel_w, bias,
                                  output, out_channels);
  else
    caffe_cpu_sconv_default<false>(input_padded, in_channels, height, width, pad_h, pad_w, stride_h, stride_w,
                                   dilation_h, dilation_w, rowptr, colidx, values, kernel_h, kernel_w, bias,
                                   output, out_channels);
}

template <int WIDTH, int K>
static void unit_stride(int relu, const float *input, const int *rowptr, const int *colidx,
                        const float *values, const float *bias, float *output, int oc_begin, int oc_end,
                        float *scratch, int in_channels, int out_channels) {
  const int *rp[1] = {rowptr};
  const int *ci[1] = {colidx};
  const float *va[1] = {values};
  if (relu)
    sconv_unit_stride<WIDTH, K, true>(input, rp, ci, va, 1, bias, output, oc_begin, oc_end, scratch,
                                      in_channels, out_channels);
  else
    sconv_unit_stride<WIDTH, K, false>(input, rp, ci, va, 1, bias, output, oc_begin, oc_end, scratch,
                                       in_channels, out_channels);
}

// 1 if the reference has a sconv_unit_stride<WIDTH,K> specialisation for this square, unit-stride,
// matched-padding geometry (math_functions.cpp:225-383).
REF_API int ref_has_unit_stride(int width, int k) {
  if (k == 1) return width == 4 || width == 7 || width == 14 || width == 28 || width == 56;
  if (k == 3) return width == 12 || width == 13 || width == 7 || width == 14 || width == 28 || width == 56 ||
                     width == 3;
  if (k == 5) return width == 27 || width == 7 || width == 14 || width == 28;
  return 0;
}

// sconv_unit_stride<WIDTH,K,FUSE_RELU> with ncolblocks == 1 (sconv.hpp:428-585). input_padded must have
// VLEN-1 readable floats past oracle_padded_len (base_conv_layer.cpp:73,598). Returns 0 if dispatched.
REF_API int ref_sconv_unit_stride(int width, int k, int fuse_relu, const float *input_padded,
                                  const int *rowptr, const int *colidx, const float *values,
                                  const float *bias, float *output, int oc_begin, int oc_end, float *scratch,
                                  int in_channels, int out_channels) {
#define CASE(W_, K_)                                                                                       \
  if (width == W_ && k == K_) {                                                                            \
    unit_stride<W_, K_>(fuse_relu, input_padded, rowptr, colidx, values, bias, output, oc_begin, oc_end,   \
                        scratch, in_channels, out_channels);                                               \
    return 0;                                                                                              \
  }
  CASE(4, 1) CASE(7, 1) CASE(14, 1) CASE(28, 1) CASE(56, 1)
  CASE(12, 3) CASE(13, 3) CASE(7, 3) CASE(14, 3) CASE(28, 3) CASE(56, 3) CASE(3, 3)
  CASE(27, 5) CASE(7, 5) CASE(14, 5) CASE(28, 5)
#undef CASE
  return 1;
}

REF_API int ref_vlen(void) { return VLEN; }

// Whole-layer forward the way the reference's CPU path runs it (conv_layer.cpp:41-58 +
// base_conv_layer.cpp:570-661, ICC/BLOCKED_SCONV flavour when a specialisation exists, else the default
// kernel): omp-parallel over images, per image: pad copy (base_conv_layer.cpp:615-620), per group sconv
// with the bias fused ONCE (the reference's second forward_cpu_bias, conv_layer.cpp:55-58, is its
// double-bias bug under ICC and is not repeated).  use_blocked=0 forces caffe_cpu_sconv_default.
REF_API int ref_conv_forward(const float *bottom, int num, int Cin, int H, int W, int Cout, int group, int kh,
                             int kw, int pad_h, int pad_w, int stride_h, int stride_w, int dil_h, int dil_w,
                             const float *values, const int *colidx_stretched, const int *rowptr,
                             const float *bias /*nullable*/, int fuse_relu, float *top, int threads,
                             int use_blocked) {
  const int M = Cout / group;
  const long N = (long)(Cin / group) * kh * kw;
  const long weight_offset = (long)M * N;
  const int Ho = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
  const int Wo = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
  const long plen = (long)Cin * (H + pad_h) * (W + pad_w) + (long)pad_h * (W + 2 * pad_w) + (VLEN - 1);
  const long bottom_dim = (long)Cin * H * W, top_dim = (long)Cout * Ho * Wo;
  const bool blocked = use_blocked && dil_h == 1 && dil_w == 1 && stride_h == 1 && stride_w == 1 && H == W &&
                       kh == kw && pad_h == pad_w && kh == 2 * pad_h + 1 && ref_has_unit_stride(H, kh);
  float *zero_bias = NULL;
  if (!bias) zero_bias = (float *)calloc((size_t)Cout, sizeof(float));
  const float *b = bias ? bias : zero_bias;
  if (threads < 1) threads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
  {
    float *padded = NULL;
    if (posix_memalign((void **)&padded, 4096, sizeof(float) * (size_t)(plen + 64))) abort();
    memset(padded, 0, sizeof(float) * (size_t)(plen + 64));
    float *scratch = NULL;
    if (posix_memalign((void **)&scratch, 4096, sizeof(float) * (size_t)OC_BLOCK * Ho * ((Wo + 15) / 16 * 16)))
      abort();
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int n = 0; n < num; ++n) {
      const float *in = bottom + n * bottom_dim;
      // the blocked kernel reads VLEN-wide past each row, so it always runs from the padded scratch
      if (pad_h != 0 || pad_w != 0 || blocked) {
        for (int c = 0; c < Cin; ++c)
          for (int y = 0; y < H; ++y)
            memcpy(padded + ((long)c * (H + pad_h) + y + pad_h) * (W + pad_w) + pad_w, in + ((long)c * H + y) * W,
                   sizeof(float) * W);
        in = padded;
      }
      for (int g = 0; g < group; ++g) {
        const float *in_g = in + (long)(Cin / group) * g * (H + pad_h) * (W + pad_w);
        float *out_g = top + n * top_dim + (long)g * M * Ho * Wo;
        const int *rp = rowptr + (M + 1) * g;
        const int *ci = colidx_stretched + weight_offset * g;
        const float *va = values + weight_offset * g;
        if (blocked)
          ref_sconv_unit_stride(H, kh, fuse_relu, in_g, rp, ci, va, b + M * g, out_g, 0, M, scratch, Cin / group,
                                M);
        else
          ref_sconv_default(fuse_relu, in_g, Cin / group, H, W, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w,
                            rp, ci, va, kh, kw, b + M * g, out_g, M);
      }
    }
    free(padded);
    free(scratch);
  }
  free(zero_bias);
  return blocked ? 1 : 0;
}

REF_API int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
