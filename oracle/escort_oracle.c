/*
 * escort_oracle.c -- CPU restatement of the reference's Escort direct-sparse-convolution path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (caffe_escoin_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED against the reference's own code compiled in place
 * (oracle/_ref/libescort_ref.so = /root/reference/include/caffe/util/sconv.hpp built by
 * oracle/Makefile; see tests/test_oracle.py and the committed fixtures in tests/golden/
 * made by tools/make_golden.py).  The reference ships NO golden vectors / tests of its own for the
 * sparse path, so "the reference run here" is the pin.
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 * Plain C, fp32 arithmetic in the same order as the reference's scalar loops.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ---- geometry ---------------------------------------------------------------------------- */

/* output extent: src/caffe/layers/conv_layer.cpp:8-22 (compute_output_shape) and
 * include/caffe/util/sconv.hpp:612-615 */
ORACLE_API int oracle_out_dim(int in, int pad, int kernel, int stride, int dilation) {
  return (in + 2 * pad - (dilation * (kernel - 1) + 1)) / stride + 1;
}

/* padded scratch length for one image: src/caffe/layers/base_conv_layer.cpp:71 and :596
 * (C*(H+ph)*(W+pw) + ph*(W+2pw); only top/left padding is materialised, reads past the right/bottom
 * edge land in the next row's / next channel's pad zeros or in the trailing ph*(W+2pw) zeros). */
ORACLE_API long oracle_padded_len(int C, int H, int W, int pad_h, int pad_w) {
  return (long)C * (H + pad_h) * (W + pad_w) + (long)pad_h * (W + 2 * pad_w);
}

/* ---- a1/a2: dense -> CSR ------------------------------------------------------------------ */

/* src/caffe/util/math_functions.cpp:92-105 (caffe_cpu_sparse_dense2csr<float>, non-MKL branch).
 * Row-major M x N; keeps A[i][j] != 0 (so -0.0f is dropped, NaN kept) in (i, ascending j) order.
 * nnz_per_row (may be NULL) mirrors the extra output of the GPU twin
 * (src/caffe/util/math_functions.cu:103-128).  Returns total nnz. */
ORACLE_API int oracle_dense2csr(int M, int N, const float *A, float *values, int *colidx, int *rowptr,
                                int *nnz_per_row) {
  int nnz = 0;
  rowptr[0] = 0;
  for (int i = 0; i < M; ++i) {
    int cnt = 0;
    for (int j = 0; j < N; ++j) {
      if (A[(long)i * N + j] != 0) {
        values[nnz] = A[(long)i * N + j];
        colidx[nnz] = j;
        ++cnt;
        ++nnz;
      }
    }
    rowptr[i + 1] = rowptr[i] + cnt;
    if (nnz_per_row) nnz_per_row[i] = cnt;
  }
  return nnz;
}

/* a3: src/caffe/layers/base_conv_layer.cpp:96-107 (CPU) == src/caffe/util/math_functions.cu:706-719
 * (stretch_kernel): colidx (ic,kh,kw) -> offset into the top/left padded image. In place. */
ORACLE_API void oracle_stretch(const int *rowptr, int *colidx, int M, int H, int W, int pad_h, int pad_w,
                               int kernel_h, int kernel_w) {
  for (int oc = 0; oc < M; ++oc) {
    for (int j = rowptr[oc]; j < rowptr[oc + 1]; ++j) {
      int col = colidx[j];
      int kernel_col = col % kernel_w;
      int kernel_row = (col / kernel_w) % kernel_h;
      int in_channel = col / (kernel_w * kernel_h);
      colidx[j] = (in_channel * (H + pad_h) + kernel_row) * (W + pad_w) + kernel_col;
    }
  }
}

/* a4: src/caffe/layers/base_conv_layer.cpp:615-620 (CPU memcpy rows) ==
 * src/caffe/util/math_functions.cu:729-749 (copy_input). dst must be pre-zeroed once
 * (base_conv_layer.cpp:79, :259) and have oracle_padded_len() floats. */
ORACLE_API void oracle_pad_input(float *dst, const float *src, int C, int H, int W, int pad_h, int pad_w) {
  for (int c = 0; c < C; ++c)
    for (int y = 0; y < H; ++y)
      memcpy(dst + ((long)c * (H + pad_h) + y + pad_h) * (W + pad_w) + pad_w, src + ((long)c * H + y) * W,
             sizeof(float) * W);
}

/* ---- a6/a11: direct sparse convolution of one image, one group ---------------------------- */

/* include/caffe/util/sconv.hpp:594-678 (caffe_cpu_sconv_default<FUSE_RELU>), which is the bias-fused
 * twin of src/caffe/util/math_functions.cpp:128-176 (caffe_cpu_sconv) and of the GPU kernels
 * sconv_base / sconv_relu_base / sconv_dilation (src/caffe/util/math_functions.cu:154-223).
 * bias == NULL -> sum starts at 0 (caffe_cpu_sconv / sconv_base); otherwise at bias[oc].
 * colidx is the STRETCHED index. Sequential fp32 accumulation in CSR order. */
ORACLE_API void oracle_sconv(const float *input_padded, int H, int W, int pad_h, int pad_w, int stride_h,
                             int stride_w, int dilation_h, int dilation_w, const int *rowptr,
                             const int *colidx, const float *values, int kernel_h, int kernel_w,
                             const float *bias, int fuse_relu, float *output, int out_channels) {
  const int output_h = oracle_out_dim(H, pad_h, kernel_h, stride_h, dilation_h);
  const int output_w = oracle_out_dim(W, pad_w, kernel_w, stride_w, dilation_w);
  const int Wp = W + pad_w, Hp = H + pad_h;
  if (dilation_h != 1 || dilation_w != 1) {
    for (int orow = 0; orow < output_h; ++orow)
      for (int ocol = 0; ocol < output_w; ++ocol)
        for (int oc = 0; oc < out_channels; ++oc) {
          float sum = bias ? bias[oc] : 0.f;
          for (int j = rowptr[oc]; j < rowptr[oc + 1]; ++j) {
            int off = colidx[j];
            int kernel_col = off % Wp;
            int kernel_row = (off / Wp) % Hp;
            int in_channel = off / (Wp * Hp);
            int input_row = kernel_row * dilation_h + orow * stride_h;
            int input_col = kernel_col * dilation_w + ocol * stride_w;
            sum += values[j] * input_padded[((long)in_channel * Hp + input_row) * Wp + input_col];
          }
          output[((long)oc * output_h + orow) * output_w + ocol] = (fuse_relu && sum < 0.f) ? 0.f : sum;
        }
  } else {
    for (int orow = 0; orow < output_h; ++orow)
      for (int ocol = 0; ocol < output_w; ++ocol) {
        const float *in = input_padded + (long)orow * stride_h * Wp + ocol * stride_w;
        for (int oc = 0; oc < out_channels; ++oc) {
          float sum = bias ? bias[oc] : 0.f;
          for (int j = rowptr[oc]; j < rowptr[oc + 1]; ++j) sum += values[j] * in[colidx[j]];
          output[((long)oc * output_h + orow) * output_w + ocol] = (fuse_relu && sum < 0.f) ? 0.f : sum;
        }
      }
  }
}

/* ---- a1 whole layer: WeightAlign over groups --------------------------------------------- */

/* src/caffe/layers/base_conv_layer.cpp:60-63 (M, N, weight_offset, row_offset), :83-107 (CPU) and
 * :236-264 (GPU): per group g, CSR of the M x N slice written at offsets weight_offset*g (values,
 * colidx), row_offset*g (rowptr, each group restarts at 0), M*g (nnz_per_row); then stretch.
 * values/colidx must hold Cout*(Cin/g)*kh*kw entries (base_conv_layer.cpp:509-512), rowptr Cout+g,
 * nnz_per_row Cout, nz_num g. do_stretch=0 keeps the raw (ic,kh,kw) column. */
ORACLE_API void oracle_weight_align(const float *weights, int Cout, int Cin, int group, int kernel_h,
                                    int kernel_w, int H, int W, int pad_h, int pad_w, int do_stretch,
                                    float *values, int *colidx, int *rowptr, int *nnz_per_row,
                                    int *nz_num) {
  const int M = Cout / group;
  const int N = (Cin / group) * kernel_h * kernel_w;
  const long weight_offset = (long)M * N;
  const int row_offset = M + 1;
  for (int g = 0; g < group; ++g) {
    nz_num[g] = oracle_dense2csr(M, N, weights + weight_offset * g, values + weight_offset * g,
                                 colidx + weight_offset * g, rowptr + row_offset * g, nnz_per_row + M * g);
    if (do_stretch)
      oracle_stretch(rowptr + row_offset * g, colidx + weight_offset * g, M, H, W, pad_h, pad_w, kernel_h,
                     kernel_w);
  }
}

/* ---- a5/a9 whole layer forward: ConvolutionLayer::Forward_{cpu,gpu} in SCONV mode ---------- */

/* src/caffe/layers/conv_layer.cu:16-26 (loop over images, sconv then forward_gpu_bias) +
 * src/caffe/layers/base_conv_layer.cpp:749-798 / :570-661 (pad copy; per group input offset
 * (Cin/g)*g*(H+ph)*(W+pw), output offset g*M*Ho*Wo, rowptr offset (M+1)*g, values/colidx offset
 * weight_offset*g).  Bias semantics follow the NON-buggy configuration: out = sconv + bias[g*M+oc]
 * exactly once (the ICC double-bias and the un-offset group bias in the fused-ReLU path,
 * base_conv_layer.cpp:783-790, are reference quirks we do not reproduce).
 * Density gates (0.2 GPU / 0.5 CPU) are NOT applied: the sparse path is always taken.
 * threads>1 parallelises over images like conv_layer.cpp:41-44 does under ICC. */
ORACLE_API void oracle_conv_forward(const float *bottom, int num, int Cin, int H, int W, int Cout, int group,
                                    int kernel_h, int kernel_w, int pad_h, int pad_w, int stride_h,
                                    int stride_w, int dilation_h, int dilation_w, const float *values,
                                    const int *colidx_stretched, const int *rowptr, const float *bias,
                                    int fuse_relu, float *top, int threads) {
  const int M = Cout / group;
  const int N = (Cin / group) * kernel_h * kernel_w;
  const long weight_offset = (long)M * N;
  const int Ho = oracle_out_dim(H, pad_h, kernel_h, stride_h, dilation_h);
  const int Wo = oracle_out_dim(W, pad_w, kernel_w, stride_w, dilation_w);
  const long plen = oracle_padded_len(Cin, H, W, pad_h, pad_w);
  const long bottom_dim = (long)Cin * H * W, top_dim = (long)Cout * Ho * Wo;
  if (threads < 1) threads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
  {
    float *padded = (float *)calloc((size_t)plen, sizeof(float));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int n = 0; n < num; ++n) {
      const float *in = bottom + n * bottom_dim;
      const float *src = in;
      if (pad_h != 0 || pad_w != 0) {
        oracle_pad_input(padded, in, Cin, H, W, pad_h, pad_w);
        src = padded;
      }
      for (int g = 0; g < group; ++g) {
        oracle_sconv(src + (long)(Cin / group) * g * (H + pad_h) * (W + pad_w), H, W, pad_h, pad_w,
                     stride_h, stride_w, dilation_h, dilation_w, rowptr + (M + 1) * g,
                     colidx_stretched + weight_offset * g, values + weight_offset * g, kernel_h, kernel_w,
                     bias ? bias + M * g : NULL, fuse_relu, top + n * top_dim + (long)g * M * Ho * Wo, M);
      }
    }
    free(padded);
  }
}

/* ---- dense ground truth -------------------------------------------------------------------- */

/* src/caffe/test/test_convolution_layer.cpp:20-150 (caffe_conv): naive grouped conv with
 * stride/pad/dilation and bias, on the dense (zero-filled) weights.  2-D only.  Accumulates in
 * double so it is a ground truth rather than a summation-order twin. */
ORACLE_API void oracle_dense_conv(const float *bottom, int num, int Cin, int H, int W, const float *weights,
                                  int Cout, int group, int kernel_h, int kernel_w, int pad_h, int pad_w,
                                  int stride_h, int stride_w, int dilation_h, int dilation_w,
                                  const float *bias, int fuse_relu, float *top) {
  const int Ho = oracle_out_dim(H, pad_h, kernel_h, stride_h, dilation_h);
  const int Wo = oracle_out_dim(W, pad_w, kernel_w, stride_w, dilation_w);
  const int o_g = Cout / group, k_g = Cin / group;
  for (int n = 0; n < num; ++n)
    for (int g = 0; g < group; ++g)
      for (int o = 0; o < o_g; ++o) {
        const int oc = g * o_g + o;
        for (int y = 0; y < Ho; ++y)
          for (int x = 0; x < Wo; ++x) {
            double sum = bias ? (double)bias[oc] : 0.0;
            for (int k = 0; k < k_g; ++k) {
              const int ic = g * k_g + k;
              for (int p = 0; p < kernel_h; ++p)
                for (int q = 0; q < kernel_w; ++q) {
                  int in_y = y * stride_h - pad_h + p * dilation_h;
                  int in_x = x * stride_w - pad_w + q * dilation_w;
                  if (in_y >= 0 && in_y < H && in_x >= 0 && in_x < W)
                    sum += (double)bottom[(((long)n * Cin + ic) * H + in_y) * W + in_x] *
                           (double)weights[(((long)oc * k_g + k) * kernel_h + p) * kernel_w + q];
                }
            }
            float r = (float)sum;
            top[(((long)n * Cout + oc) * Ho + y) * Wo + x] = (fuse_relu && r < 0.f) ? 0.f : r;
          }
      }
}

/* ---- a9 backward ---------------------------------------------------------------------------- */

/* Semantics of src/caffe/layers/conv_layer.cu:43-73 + src/caffe/layers/base_conv_layer.cpp:859-897:
 *   bias_diff[oc]   += sum_{n,y,x} top_diff            (backward_gpu_bias: gemv, accumulates)
 *   weight_diff     += top_diff (x) im2col(bottom)     (weight_gpu_gemm: beta = 1, accumulates)
 *   bottom_diff      = W^T * top_diff -> col2im        (backward_gpu_gemm: beta = 0, overwrites)
 * The reference computes these densely; the rebuild restricts the weight gradient to the sparsity
 * mask (weights[i] != 0).  mask_only=1 leaves weight_diff untouched where weights == 0;
 * mask_only=0 is the reference's dense gradient.  Any output pointer may be NULL (param_propagate_down_
 * / propagate_down false).  Accumulation in double (ground truth, not an order twin of cuBLAS). */
ORACLE_API void oracle_conv_backward(const float *bottom, const float *top_diff, int num, int Cin, int H, int W,
                                     const float *weights, int Cout, int group, int kernel_h, int kernel_w,
                                     int pad_h, int pad_w, int stride_h, int stride_w, int dilation_h,
                                     int dilation_w, int mask_only, float *weight_diff, float *bias_diff,
                                     float *bottom_diff) {
  const int Ho = oracle_out_dim(H, pad_h, kernel_h, stride_h, dilation_h);
  const int Wo = oracle_out_dim(W, pad_w, kernel_w, stride_w, dilation_w);
  const int o_g = Cout / group, k_g = Cin / group;
  if (bias_diff) {
    for (int oc = 0; oc < Cout; ++oc) {
      double s = 0.0;
      for (int n = 0; n < num; ++n)
        for (int i = 0; i < Ho * Wo; ++i) s += (double)top_diff[((long)n * Cout + oc) * Ho * Wo + i];
      bias_diff[oc] += (float)s;
    }
  }
  if (weight_diff) {
    for (int oc = 0; oc < Cout; ++oc) {
      const int g = oc / o_g;
      for (int k = 0; k < k_g; ++k) {
        const int ic = g * k_g + k;
        for (int p = 0; p < kernel_h; ++p)
          for (int q = 0; q < kernel_w; ++q) {
            const long widx = (((long)oc * k_g + k) * kernel_h + p) * kernel_w + q;
            if (mask_only && !(weights[widx] != 0)) continue;
            double s = 0.0;
            for (int n = 0; n < num; ++n)
              for (int y = 0; y < Ho; ++y) {
                int in_y = y * stride_h - pad_h + p * dilation_h;
                if (in_y < 0 || in_y >= H) continue;
                for (int x = 0; x < Wo; ++x) {
                  int in_x = x * stride_w - pad_w + q * dilation_w;
                  if (in_x < 0 || in_x >= W) continue;
                  s += (double)top_diff[(((long)n * Cout + oc) * Ho + y) * Wo + x] *
                       (double)bottom[(((long)n * Cin + ic) * H + in_y) * W + in_x];
                }
              }
            weight_diff[widx] += (float)s;
          }
      }
    }
  }
  if (bottom_diff) {
    const long total = (long)num * Cin * H * W;
    double *acc = (double *)calloc((size_t)total, sizeof(double));
    for (int n = 0; n < num; ++n)
      for (int oc = 0; oc < Cout; ++oc) {
        const int g = oc / o_g;
        for (int k = 0; k < k_g; ++k) {
          const int ic = g * k_g + k;
          for (int p = 0; p < kernel_h; ++p)
            for (int q = 0; q < kernel_w; ++q) {
              const float w = weights[(((long)oc * k_g + k) * kernel_h + p) * kernel_w + q];
              if (!(w != 0)) continue;
              for (int y = 0; y < Ho; ++y) {
                int in_y = y * stride_h - pad_h + p * dilation_h;
                if (in_y < 0 || in_y >= H) continue;
                for (int x = 0; x < Wo; ++x) {
                  int in_x = x * stride_w - pad_w + q * dilation_w;
                  if (in_x < 0 || in_x >= W) continue;
                  acc[(((long)n * Cin + ic) * H + in_y) * W + in_x] +=
                      (double)w * (double)top_diff[(((long)n * Cout + oc) * Ho + y) * Wo + x];
                }
              }
            }
        }
      }
    for (long i = 0; i < total; ++i) bottom_diff[i] = (float)acc[i];
    free(acc);
  }
}

ORACLE_API int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
