/*
 * escort_b200.h -- C ABI of the B200-native (sm_100a) Escort direct-sparse-convolution hot path.
 *
 * This is the drop-in boundary: every entry point below stands behind one reference interface (cited
 * file:line, relative to chenxuhao/caffe-escoin).  Plain pointers and sizes only; all pointers are DEVICE
 * pointers unless a name ends in _host.  `stream` is a cudaStream_t passed as void* (NULL = legacy default
 * stream, which is what the reference uses everywhere).  Every function returns 0 on success, a positive
 * cudaError_t value on a CUDA failure, or a negative ESCORT_E* code; nothing aborts, nothing calls
 * cudaDeviceSynchronize (the reference's CudaTest() device sync after every launch,
 * include/caffe/util/cutil_subset.h:28-38, is deliberately not reproduced).  The library is re-entrant; plans
 * are bound to the device that was current when they were created.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point returns an error.
 */
#ifndef ESCORT_B200_H_
#define ESCORT_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESCORT_EINVAL (-1)   /* bad argument / unsupported geometry */
#define ESCORT_ENOMEM (-2)   /* host allocation failed */
#define ESCORT_ENCCL  (-3)   /* NCCL failure (see escort_last_error) */

#if defined(__GNUC__)
#define ESCORT_API __attribute__((visibility("default")))
#else
#define ESCORT_API
#endif

typedef void *escort_stream_t; /* cudaStream_t */

/* Convolution geometry of one layer; field names follow BaseConvolutionLayer's members
 * (include/caffe/layers/base_conv_layer.hpp:155-185; parsed in src/caffe/layers/base_conv_layer.cpp:276-446). */
typedef struct escort_geom {
  int channels;     /* conv_in_channels_  (C, all groups) */
  int num_output;   /* conv_out_channels_ (M, all groups) */
  int group;        /* group_ */
  int height, width;          /* conv_input_shape_[1], [2] */
  int kernel_h, kernel_w;     /* kernel_shape_ */
  int pad_h, pad_w;           /* pad_ */
  int stride_h, stride_w;     /* stride_ */
  int dilation_h, dilation_w; /* dilation_ */
} escort_geom;

/* ---- a2: dense -> CSR ---------------------------------------------------------------------------------
 * Replaces caffe_gpu_sparse_dense2csr<float> (include/caffe/util/math_functions.hpp:213-216,
 * src/caffe/util/math_functions.cu:103-128, called at src/caffe/layers/base_conv_layer.cpp:240-246), whose
 * cuSPARSE entry points no longer exist in CUDA 12.  Bit-exact with the CPU twin
 * caffe_cpu_sparse_dense2csr<float> (src/caffe/util/math_functions.cpp:92-105): row-major M x N matrix A,
 * keeps A[i][j] != 0 in (i, ascending j) order; rowptr has M+1 entries starting at 0; values / colidx must
 * hold M*N entries (worst case, as the reference sizes them, base_conv_layer.cpp:509-512); nnz_per_row (M
 * entries) may be NULL.  *nnz_total_host is written on the host after a stream synchronisation (the
 * reference's call is synchronous too). */
ESCORT_API int escort_pack_csr(int M, int N, const float *A, int *nnz_per_row, float *values, int *rowptr, int *colidx,
                    int *nnz_total_host, escort_stream_t stream);

/* ---- a3: stretch ---------------------------------------------------------------------------------------
 * Replaces caffe_gpu_stretch (math_functions.hpp:226-228, math_functions.cu:706-727, called at
 * base_conv_layer.cpp:263): in place, colidx (ic*kh*kw + r*kw + s) -> (ic*(H+pad_h) + r)*(W+pad_w) + s. */
ESCORT_API int escort_stretch(const int *rowptr, int *colidx, int M, int height, int width, int pad_h, int pad_w,
                   int kernel_h, int kernel_w, escort_stream_t stream);

/* ---- a4: padded input copy ------------------------------------------------------------------------------
 * Replaces copy_input_data<float> (math_functions.hpp:230-233, math_functions.cu:729-766, called at
 * base_conv_layer.cpp:771,828): dst[(c*(H+ph)+y+ph)*(W+pw)+pw+x] = src[(c*H+y)*W+x].  Only needed by the
 * compatibility entry escort_sconv_padded; the native forward reads the unpadded tensor. */
ESCORT_API int escort_copy_input(float *dst, const float *src, int num_channels, int height, int width, int pad_h,
                      int pad_w, escort_stream_t stream);

/* ---- a6 (compat): caffe_gpu_sconv with the reference's exact argument list ------------------------------
 * Replaces caffe_gpu_sconv<float> (math_functions.hpp:218-224, math_functions.cu:590-694, called at
 * base_conv_layer.cpp:788-795,841): `input` is the top/left padded image(s), colidx is STRETCHED,
 * out[oc,y,x] = (relu?)((fuse_relu ? bias[oc] : 0) + sum_j values[j]*input[y*sh*(W+pw) + x*sw + colidx[j]]).
 * num > 1 processes `num` images spaced ifmap_size*num_groups floats apart on input and
 * num_oc*num_groups*Ho*Wo on output (SCONV_PAR semantics, without the reference's odd-batch drop and
 * grid bug, math_functions.cu:664,675-689).  Dilation != 1 uses the decode of sconv_dilation (:154-179). */
ESCORT_API int escort_sconv_padded(int fuse_relu, int num, const float *input, int ifmap_size, const int *rowptr,
                        const int *colidx, const float *values, const float *bias, int height, int width,
                        int pad_h, int pad_w, int stride_h, int stride_w, int dilation_h, int dilation_w,
                        int kernel_h, int kernel_w, float *output, int num_oc, int num_groups,
                        escort_stream_t stream);

/* ---- plans -----------------------------------------------------------------------------------------------
 * State carried between WeightAlign and Forward/Backward: the derived, nnz-balanced blocked formats the
 * fast kernels execute.  Built from the layer's CSR blobs in the reference's layout
 * (base_conv_layer.cpp:240-246,509-513): values/colidx at offset (M/g)*(C/g)*kh*kw*g, rowptr at (M/g+1)*g
 * (each group restarting at 0).  colidx may be raw (ic,kh,kw) columns (colidx_is_stretched = 0) or the
 * stretched form (1).  The arrays are read during the call (device->host copy + stream sync) and not
 * retained.  One plan per (layer, device); owned by the library. */
typedef struct escort_plan escort_plan;

ESCORT_API int escort_plan_create(const escort_geom *geom, const int *rowptr, const int *colidx, const float *values,
                       int colidx_is_stretched, escort_plan **plan_out, escort_stream_t stream);
ESCORT_API int escort_plan_destroy(escort_plan *plan);
/* total nonzeros over all groups (sum of nz_num_, base_conv_layer.cpp:247) */
ESCORT_API long escort_plan_nnz(const escort_plan *plan);
/* name of the forward kernel variant the plan selected (static string) */
ESCORT_API const char *escort_plan_kernel_name(const escort_plan *plan);
/* human-readable tiling summary of the selected forward kernel, written to buf (NUL terminated) */
ESCORT_API int escort_plan_describe(const escort_plan *plan, char *buf, int buflen);
/* tuning knob for tests/bench: force a forward variant (-1 auto; 0 generic / small-map kernel; > 0 tile-interpreter and
 * TMEM-window variants; -2, stride-2 layers only: the space-to-depth path -- a stride-1 plan over the four parity planes
 * of the padded input, written by one extra pass into a buffer the plan owns (sized for the largest batch seen; a plan's
 * forward calls must therefore not overlap on different streams).  Auto uses it where the measured sweep says it wins,
 * escort_plan_autotune measures both paths) */
ESCORT_API int escort_plan_set_variant(escort_plan *plan, int variant);

/* tuning knob: variant as above plus which of the planner's tiling candidates to use (0 = its favourite) */
ESCORT_API int escort_plan_set_config(escort_plan *plan, int variant, int layout_rank);
/* the configuration in force (as chosen by escort_plan_autotune or set explicitly); lets a host cache tuning results */
ESCORT_API int escort_plan_get_config(const escort_plan *plan, int *variant_host, int *layout_rank_host);
/* measure every forward variant that supports the geometry on a scratch batch of `num` images and keep the
 * fastest (plan-time selection, like cuDNN's find); synchronises `stream`.  Optional: without it the plan
 * uses a static default. */
ESCORT_API int escort_plan_autotune(escort_plan *plan, int num, escort_stream_t stream);
/* the same for the backward kernels: the backward-weight variant, and the backward-data kernel (stride-1 layers, dilated
 * or not, run it as a forward plan over the transposed, rotated weights; stride-2 layers through the backward plan of
 * their space-to-depth sub-plan -- ConvolutionLayer::Backward_gpu's backward_gpu_gemm + col2im,
 * src/caffe/layers/conv_layer.cu:64-68); a no-op for geometries that use the generic backward kernel.  Training hosts
 * call it once after WeightAlign. */
ESCORT_API int escort_plan_autotune_backward(escort_plan *plan, int num, escort_stream_t stream);
/* apply another plan's tuning (same geometry) instead of measuring again */
ESCORT_API int escort_plan_copy_tuning(escort_plan *dst, const escort_plan *src, escort_stream_t stream);

/* ---- a5-a9: native forward ---------------------------------------------------------------------------------
 * Replaces the whole per-image sequence of ConvolutionLayer::Forward_gpu in SCONV / SCONV_PAR mode
 * (src/caffe/layers/conv_layer.cu:15-26) = forward_gpu_sconv[_par] (base_conv_layer.cpp:749-848:
 * copy_input_data + caffe_gpu_sconv per group) + forward_gpu_bias (:851-856), for the whole batch and all
 * groups in one launch.  bottom: num x C x H x W unpadded NCHW; top: num x M x Ho x Wo; bias (M) nullable;
 * fuse_relu applies max(.,0) in the epilogue (ConvolutionReLULayer, src/caffe/layers/conv_relu_layer.cu:8-30). */
ESCORT_API int escort_sconv_forward(escort_plan *plan, int num, const float *bottom, const float *bias, int fuse_relu,
                         float *top, escort_stream_t stream);

/* ---- a9: backward, restricted to the sparsity mask ------------------------------------------------------
 * Caffe's contract (src/caffe/layers/conv_layer.cu:43-73, base_conv_layer.cpp:859-897): parameter diffs
 * ACCUMULATE, bottom_diff is OVERWRITTEN.
 * backward_data  : bottom_diff = W^T * top_diff           (replaces backward_gpu_gemm + col2im)
 * backward_weight: gradient only at the plan's nonzero positions (replaces weight_gpu_gemm + im2col):
 *                  weight_diff_dense (M x C/g x kh x kw, nullable) gets += at mask positions;
 *                  weight_diff_csr (nullable) gets the same numbers in CSR order, in the reference blob
 *                  layout (group g at offset weight_offset*g), overwritten if accumulate == 0 else +=.
 * bias_backward  : bias_diff[oc] += sum_{n,y,x} top_diff  (replaces backward_gpu_bias gemv). */
ESCORT_API int escort_sconv_backward_data(escort_plan *plan, int num, const float *top_diff, float *bottom_diff,
                               escort_stream_t stream);
ESCORT_API int escort_sconv_backward_weight(escort_plan *plan, int num, const float *bottom, const float *top_diff,
                                 float *weight_diff_dense, float *weight_diff_csr, int accumulate,
                                 escort_stream_t stream);
ESCORT_API int escort_bias_backward(int num, int num_output, int out_spatial, const float *top_diff, float *bias_diff,
                         escort_stream_t stream);

/* Re-gather CSR values from the (updated) dense weights at the plan's fixed positions, into the plan and
 * optionally into the layer's nz_weight_values_ blob (values_csr nullable).  The reference never refreshes
 * its CSR after a solver update; the masked training step needs it. */
ESCORT_API int escort_refresh_values(escort_plan *plan, const float *weights_dense, float *values_csr,
                          escort_stream_t stream);

/* ---- e: multi-GPU gradient exchange ------------------------------------------------------------------------
 * Replaces NCCL<Dtype>::on_gradients_ready (src/caffe/parallel.cpp:238-256): ncclAllReduce(sum) over the
 * flat diff buffer, then scale by `scale` (1/solver_count).  `comm` is an ncclComm_t passed as void*. */
ESCORT_API int escort_allreduce_grads(void *comm, float *flat, size_t count, float scale, escort_stream_t stream);
/* Weight broadcast from `root` before the first step (NCCL<Dtype>::Broadcast, src/caffe/parallel.cpp:189-199). */
ESCORT_API int escort_broadcast(void *comm, float *buf, size_t count, int root, escort_stream_t stream);
/* Communicator bring-up for hosts without an NCCL binding of their own (a Caffe host passes the ncclComm_t it already
 * owns, src/caffe/parallel.cpp:141-171): rank 0 fills a 128-byte ncclUniqueId, ships it to the other ranks by any
 * means, every rank calls init_rank.  One process per GPU; the current CUDA device is the rank's device. */
ESCORT_API int escort_comm_unique_id(void *id128);
ESCORT_API int escort_comm_init_rank(void **comm_out, int nranks, const void *id128, int rank);
ESCORT_API int escort_comm_destroy(void *comm);

/* ---- f1: the layers the reference keeps dense, on tcgen05 ----------------------------------------------------------
 * InnerProductLayer::Forward_gpu (src/caffe/layers/inner_product_layer.cu:9-31, cuBLAS sgemm + bias gemv):
 * top[num x num_output] = bottom[num x K] * weight[num_output x K]^T + bias, optional ReLU, and
 * EscConvolutionLayer::Forward_gpu (src/caffe/layers/esc_conv_layer.cu:21-29, cuDNN IMPLICIT_GEMM) for conv1 / 1x1 /
 * unpruned convolutions: TMA + tcgen05.mma (kind::tf32, fp32 accumulate in TMEM) GEMM kernels, fused bias / ReLU
 * epilogue.  TF32 products: up to 7e-4 relative L2 against fp32 (the sparse path's 1e-4 bar is for the sparse path).
 * inner product: K % 4 == 0, operands 16-byte aligned.  conv: group == 1.  1x1 / stride 1 / no-padding layers whose
 * pixel count and channel count are multiples of 4 run as an implicit GEMM straight from NCHW; every other geometry
 * goes through `workspace` (device): the padded weights and the transposed column buffer,
 * escort_dense_conv_workspace_bytes(geom, num) bytes (always required, so that the caller need not know which path runs). */
ESCORT_API int escort_inner_product_forward(int num, int K, int num_output, const float *bottom, const float *weight,
                                            const float *bias, int fuse_relu, float *top, escort_stream_t stream);
ESCORT_API size_t escort_dense_conv_workspace_bytes(const escort_geom *geom, int num);
ESCORT_API int escort_dense_conv_forward(const escort_geom *geom, int num, const float *bottom, const float *weight,
                                         const float *bias, int fuse_relu, void *workspace, size_t workspace_bytes, float *top,
                                         escort_stream_t stream);
/* the same convolution with the Eltwise SUM that follows a ResNet branch2c fused into its epilogue
 * (src/caffe/layers/eltwise_layer.cu, EltwiseParameter_EltwiseOp_SUM with unit coefficients):
 * top = [relu](conv(bottom) + bias + residual); `residual` has top's shape and may alias top; NULL = plain forward. */
ESCORT_API int escort_dense_conv_forward_residual(const escort_geom *geom, int num, const float *bottom, const float *weight,
                                                  const float *bias, const float *residual, int fuse_relu, void *workspace,
                                                  size_t workspace_bytes, float *top, escort_stream_t stream);

/* ---- f2: glue-layer fusion ------------------------------------------------------------------------------------------
 * conv -> BatchNorm(use_global_stats) -> Scale -> ReLU (the chain around every ResNet-50 sparse conv; the reference runs
 * four layers, src/caffe/net.cpp:531-532 "other time") as ONE forward launch: the per-channel affine
 * y = conv * a[oc] + b[oc] is folded into the plan's nonzero weights (the mask and every record stream keep their
 * positions) and into the bias.  escort_bn_scale_to_affine turns the reference's blobs into (a, b):
 * BatchNormLayer blobs {mean, variance, scale_factor} (src/caffe/layers/batch_norm_layer.cpp:98-106, 139-152) and
 * ScaleLayer blobs {gamma, beta} (may be NULL).  escort_plan_fold_affine: weights_folded = W * a[oc] (dense scratch the
 * caller owns, same size as the weight blob), bias_out = bias_in * a + b, then the plan is refreshed from weights_folded;
 * afterwards escort_sconv_forward(plan, ..., bias_out, fuse_relu = 1, ...) is the whole chain.  All pointers device. */
ESCORT_API int escort_bn_scale_to_affine(int num_output, const float *bn_mean, const float *bn_var, float bn_scale_factor_blob, float eps,
                                         const float *scale_gamma, const float *scale_beta, float *a_out, float *b_out,
                                         escort_stream_t stream);
ESCORT_API int escort_plan_fold_affine(escort_plan *plan, const float *weights_dense, const float *a, const float *b,
                                       const float *bias_in, float *weights_folded, float *bias_out, escort_stream_t stream);
/* the same fold for a layer that stays dense (f1): weights [num_output x row] -> weights_folded (may alias), bias_out = bias_in * a + b */
ESCORT_API int escort_dense_fold_affine(int num_output, long row, const float *weights, const float *a, const float *b,
                                        const float *bias_in, float *weights_folded, float *bias_out, escort_stream_t stream);

/* ---- f3: the LOWERED_SPARSE comparator -----------------------------------------------------------------------------
 * conv_mode 1 of the reference: per image im2col, then per group CSR x dense on cuSPARSE
 * (BaseConvolutionLayer::forward_gpu_gemm, src/caffe/layers/base_conv_layer.cpp:715-745; caffe_gpu_sparse_csrmm =
 * cusparseScsrmm2 + cublasSgeam, src/caffe/util/math_functions.cu:48-62 -- csrmm2 no longer exists, this runs
 * cusparseSpMM), bias as forward_gpu_bias.  rowptr / colidx / values in the reference's blob layout with RAW column
 * indices ((ic * kh + r) * kw + s: WeightAlign without the stretch).  `col_buffer`: group * (C/group) * kh * kw * Ho * Wo
 * floats of scratch (col_buffer_).  Unlike the reference, no density gate: the sparse product always runs.
 * A comparator for the paper's "cuSPARSE" baseline, not a product path; needs libcusparse.so.12 at run time. */
ESCORT_API int escort_lowered_sparse_forward(const escort_geom *geom, int num, const float *bottom, const int *rowptr,
                                             const int *colidx_raw, const float *values, const float *bias, int fuse_relu,
                                             float *col_buffer, float *top, escort_stream_t stream);

/* ---- f4: weight on-disk path ------------------------------------------------------------------------------------
 * Replaces Net::CopyTrainedLayersFrom(const string) (src/caffe/net.cpp:785-821: binary NetParameter -> per layer
 * Blob::FromProto, src/caffe/blob.cpp:466-520 -> WeightAlign): a reader of the protobuf wire format for the fields of
 * src/caffe/proto/caffe.proto that path touches (both `layer` = 100 and the V1 `layers` = 2), host memory only.
 * Blob data pointers stay owned by the model and are writable, so a pruned model can be saved again. */
typedef struct escort_caffemodel escort_caffemodel;
typedef struct escort_layer_info {
  const char *name, *type;          /* valid until close; V1 layers report their enum as "V1:<n>" */
  int num_blobs, is_conv, is_inner_product;
  int num_output, bias_term, group; /* ConvolutionParameter / InnerProductParameter, proto defaults when absent */
  int kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation;
} escort_layer_info;
ESCORT_API int escort_caffemodel_open(const char *path, escort_caffemodel **out);
ESCORT_API int escort_caffemodel_close(escort_caffemodel *m);
ESCORT_API int escort_caffemodel_save(const escort_caffemodel *m, const char *path);
ESCORT_API int escort_caffemodel_num_layers(const escort_caffemodel *m);
/* index of the layer with this name (the match of net.cpp:790-796), or ESCORT_EINVAL */
ESCORT_API int escort_caffemodel_find(const escort_caffemodel *m, const char *layer_name);
ESCORT_API int escort_caffemodel_layer(const escort_caffemodel *m, int layer, escort_layer_info *info);
/* blob `blob` of layer `layer` (0 = weights, 1 = bias): axes (legacy num/channels/height/width or shape.dim), data */
ESCORT_API int escort_caffemodel_blob(escort_caffemodel *m, int layer, int blob, int *ndim, long *shape8, float **data_host,
                                      long *count);
/* magnitude pruning in place (host): the floor(count * sparsity) smallest |w| become 0.  The reference trains its
 * sparsity with an L1 regulariser (src/caffe/solvers/sgd_solver.cpp:161-168) and ships pruned checkpoints (run.sh:13);
 * this produces the same kind of input for WeightAlign from a dense checkpoint. */
ESCORT_API int escort_prune_magnitude(float *weights_host, long count, double sparsity, float *threshold_out, long *nnz_out);

/* ---- misc ---------------------------------------------------------------------------------------------------- */
/* register-resident FFMA microbenchmark used for the FP32 roofline denominator;
 * returns achieved TFLOP/s in *tflops_host, SM count and SM clock (kHz, from device attributes). */
ESCORT_API int escort_measure_fp32_peak(int variant, int iters, double *tflops_host, int *sm_count_host, int *clock_khz_host);
/* diagnostics: the TMEM kernels bound every mbarrier wait (~2 s); a wait that gives up records {code, block, warp,
 * barrier offset, parity, ...} in 16 host-mapped words before trapping.  Copies them to out16. */
ESCORT_API int escort_tmem_debug(int *out16);
ESCORT_API const char *escort_last_error(void);
ESCORT_API const char *escort_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ESCORT_B200_H_ */
