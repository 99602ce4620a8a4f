// core.cu -- plan construction, the reference-ABI compatibility kernel, and the generic (any geometry)
// forward / masked-backward kernels of the Escort direct sparse convolution.
//
// The generic kernels are the always-correct CUDA path for every stride / pad / dilation / group
// combination; the register-blocked tile interpreter in sconv_tile.cu is selected by the plan when the
// geometry fits it.  There is no CPU path.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace escort {

static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  g_last_error = buf;
  return (int)e;
}

// ------------------------------------------------------------------------------------------------------
// a6 (compat): the reference's caffe_gpu_sconv contract -- padded input, stretched colidx.
// One CTA = 128 consecutive output pixels of one (image, out-channel); the CSR row is streamed through
// shared memory in chunks so (colidx, value) are fetched once per CTA (the reference's sconv_shm idea,
// math_functions.cu:264-319), and the whole batch is one launch.
// ------------------------------------------------------------------------------------------------------
static constexpr int kCompatThreads = 128;
static constexpr int kCompatChunk = 512;

template <bool DILATED>
__global__ void __launch_bounds__(kCompatThreads)
    sconv_padded_kernel(const int *__restrict__ rowptr, const int *__restrict__ colidx,
                        const float *__restrict__ values, const float *__restrict__ input, long in_img_stride,
                        const float *__restrict__ bias, int fuse_relu, float *__restrict__ output,
                        long out_img_stride, int num_oc, int Hp, int Wp, int stride_h, int stride_w, int dil_h,
                        int dil_w, int Ho, int Wo) {
  __shared__ int col_s[kCompatChunk];
  __shared__ float val_s[kCompatChunk];
  const int oc = blockIdx.y;
  const int n = blockIdx.z;
  const int pix = blockIdx.x * kCompatThreads + threadIdx.x;
  const bool active = pix < Ho * Wo;
  const int oy = active ? pix / Wo : 0, ox = active ? pix % Wo : 0;
  const float *in = input + (long)n * in_img_stride;
  const float *in_ptr = in + (long)oy * stride_h * Wp + ox * stride_w;
  const int row_start = rowptr[oc], row_end = rowptr[oc + 1];
  float sum = (fuse_relu && bias) ? bias[oc] : 0.f;
  for (int base = row_start; base < row_end; base += kCompatChunk) {
    const int len = min(kCompatChunk, row_end - base);
    for (int i = threadIdx.x; i < len; i += kCompatThreads) {
      col_s[i] = __ldg(colidx + base + i);
      val_s[i] = __ldg(values + base + i);
    }
    __syncthreads();
    if (active) {
      if (!DILATED) {
        for (int i = 0; i < len; ++i) sum = fmaf(val_s[i], __ldg(in_ptr + col_s[i]), sum);
      } else {
        for (int i = 0; i < len; ++i) {
          const int off = col_s[i];
          const int kc = off % Wp, kr = (off / Wp) % Hp, ic = off / (Wp * Hp);
          const int iy = kr * dil_h + oy * stride_h, ix = kc * dil_w + ox * stride_w;
          sum = fmaf(val_s[i], __ldg(in + ((long)ic * Hp + iy) * Wp + ix), sum);
        }
      }
    }
    __syncthreads();
  }
  if (active) {
    if (fuse_relu) sum = fmaxf(sum, 0.f);
    output[(long)n * out_img_stride + (long)oc * Ho * Wo + pix] = sum;
  }
}

// ------------------------------------------------------------------------------------------------------
// generic forward from the UNPADDED tensor: halo handled by predication, bias/ReLU fused.
// meta[j] = {in_off, dy, dx, value}: input element = x[n][...][(oy*sh + dy), (ox*sw + dx)] with
// in_off = ic*H*W + dy*W + dx (so address = base(oy*sh, ox*sw) + in_off).
// ------------------------------------------------------------------------------------------------------
static constexpr int kGenThreads = 128;
static constexpr int kGenChunk = 256;

__global__ void __launch_bounds__(kGenThreads)
    sconv_fwd_generic_kernel(const int *__restrict__ rowptr, const int4 *__restrict__ meta,
                             const float *__restrict__ bottom, const float *__restrict__ bias, int fuse_relu,
                             float *__restrict__ top, int C, int H, int W, int M, int Ho, int Wo, int stride_h,
                             int stride_w) {
  __shared__ int4 meta_s[kGenChunk];
  const int oc = blockIdx.y, n = blockIdx.z;
  const int pix = blockIdx.x * kGenThreads + threadIdx.x;
  const bool active = pix < Ho * Wo;
  const int oy = active ? pix / Wo : 0, ox = active ? pix % Wo : 0;
  const int iy0 = oy * stride_h, ix0 = ox * stride_w;
  const float *in = bottom + (long)n * C * H * W + (long)iy0 * W + ix0;
  const int row_start = rowptr[oc], row_end = rowptr[oc + 1];
  float sum = bias ? bias[oc] : 0.f;
  for (int base = row_start; base < row_end; base += kGenChunk) {
    const int len = min(kGenChunk, row_end - base);
    for (int i = threadIdx.x; i < len; i += kGenThreads) meta_s[i] = __ldg(meta + base + i);
    __syncthreads();
    if (active) {
      for (int i = 0; i < len; ++i) {
        const int4 m = meta_s[i];
        const int iy = iy0 + m.y, ix = ix0 + m.z;
        if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
          sum = fmaf(__int_as_float(m.w), __ldg(in + m.x), sum);
      }
    }
    __syncthreads();
  }
  if (active) {
    if (fuse_relu) sum = fmaxf(sum, 0.f);
    top[((long)n * M + oc) * Ho * Wo + pix] = sum;
  }
}

// ------------------------------------------------------------------------------------------------------
// small maps (LeNet-sized layers, where a persistent tile kernel has fewer units than the GPU has SMs and the
// generic kernel's blocks are a fraction of a warp's worth of pixels): one CTA = one image x a slice of the output
// channels, the WHOLE input image staged once in shared memory; a warp takes (output channel, 64 pixels) items, two
// pixels per lane, and walks the channel's CSR row with warp-uniform 16-byte loads of the same meta records the generic
// kernel uses.  Same accumulation order as the generic kernel (bias first, then CSR order): bit-identical results.
// ------------------------------------------------------------------------------------------------------
static constexpr int kSmallThreads = 256;
static constexpr size_t kSmallMaxBytes = 96 * 1024;

template <bool kHalo>  // false: no padding, every tap of every output is inside the image -- no bounds tests in the loop
__global__ void __launch_bounds__(kSmallThreads)
    sconv_fwd_small_kernel(const int *__restrict__ rowptr, const int4 *__restrict__ meta, const float *__restrict__ bottom,
                           const float *__restrict__ bias, int fuse_relu, float *__restrict__ top, int CHW, int H, int W, int M,
                           int Ho, int Wo, int stride_h, int stride_w, int splits, int oc_per_cta) {
  extern __shared__ __align__(16) float img_s[];
  const int n = blockIdx.x / splits, part = blockIdx.x - n * splits;
  const float *src = bottom + (size_t)n * CHW;
  if ((CHW & 3) == 0 && ((uintptr_t)src & 15) == 0) {
    for (int i = threadIdx.x; i < (CHW >> 2); i += kSmallThreads) reinterpret_cast<float4 *>(img_s)[i] = __ldg(reinterpret_cast<const float4 *>(src) + i);
  } else {
    for (int i = threadIdx.x; i < CHW; i += kSmallThreads) img_s[i] = __ldg(src + i);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int HoWo = Ho * Wo, pblocks = (HoWo + 63) / 64;
  const int oc0 = part * oc_per_cta, oc1 = min(M, oc0 + oc_per_cta);
  const int items = (oc1 - oc0) * pblocks;
  // the warp's window of 32 records: loaded with one coalesced 512-byte access (a record is read by one warp once, so a
  // per-record uniform load would pay an L2 latency per nonzero), read back as broadcasts
  int4 *mw = reinterpret_cast<int4 *>(img_s + ((CHW + 3) & ~3)) + wid * 32;
  for (int it = wid; it < items; it += kSmallThreads / 32) {
    const int oc = oc0 + it / pblocks, pb = it - (it / pblocks) * pblocks;
    const int pa = pb * 64 + lane, pc = pa + 32;
    const bool la = pa < HoWo, lc = pc < HoWo;
    const int oya = la ? pa / Wo : 0, oxa = la ? pa - oya * Wo : 0;
    const int oyc = lc ? pc / Wo : 0, oxc = lc ? pc - oyc * Wo : 0;
    const int iya = oya * stride_h, ixa = oxa * stride_w, iyc = oyc * stride_h, ixc = oxc * stride_w;
    const int ba = iya * W + ixa, bc = iyc * W + ixc;
    float sa = bias ? __ldg(bias + oc) : 0.f, sc = sa;
    const int j1 = __ldg(rowptr + oc + 1);
    int j0 = __ldg(rowptr + oc);
    int4 nxt = make_int4(0, 0, 0, 0);
    if (j0 + lane < j1) nxt = __ldg(meta + j0 + lane);
    for (; j0 < j1; j0 += 32) {
      __syncwarp();
      mw[lane] = nxt;
      __syncwarp();
      if (j0 + 32 + lane < j1) nxt = __ldg(meta + j0 + 32 + lane);  // the next window is in flight during this one
      const int len = min(32, j1 - j0);
#pragma unroll 4
      for (int i = 0; i < len; ++i) {
        const int4 m = mw[i];
        const float w = __int_as_float(m.w);
        if (!kHalo || ((unsigned)(iya + m.y) < (unsigned)H && (unsigned)(ixa + m.z) < (unsigned)W)) sa = fmaf(w, img_s[ba + m.x], sa);
        if (!kHalo || ((unsigned)(iyc + m.y) < (unsigned)H && (unsigned)(ixc + m.z) < (unsigned)W)) sc = fmaf(w, img_s[bc + m.x], sc);
      }
    }
    if (fuse_relu) {
      sa = fmaxf(sa, 0.f);
      sc = fmaxf(sc, 0.f);
    }
    float *out = top + ((size_t)n * M + oc) * HoWo;
    if (la) out[pa] = sa;
    if (lc) out[pc] = sc;
  }
}

// ------------------------------------------------------------------------------------------------------
// generic backward-data: bottom_diff[n][c][iy][ix] = sum over nonzeros (oc,c,kh,kw) of
//   w * top_diff[n][oc][(iy - dy)/sh][(ix - dx)/sw]   when divisible and in range.  Overwrites.
// tmeta (grouped by input channel c) = {oc, dy, dx, value}.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGenThreads)
    sconv_bwd_data_generic_kernel(const int *__restrict__ colptr, const int4 *__restrict__ tmeta,
                                  const float *__restrict__ top_diff, float *__restrict__ bottom_diff, int C, int H,
                                  int W, int M, int Ho, int Wo, int stride_h, int stride_w) {
  __shared__ int4 meta_s[kGenChunk];
  const int c = blockIdx.y, n = blockIdx.z;
  const int pix = blockIdx.x * kGenThreads + threadIdx.x;
  const bool active = pix < H * W;
  const int iy = active ? pix / W : 0, ix = active ? pix % W : 0;
  const float *td = top_diff + (long)n * M * Ho * Wo;
  const int start = colptr[c], end = colptr[c + 1];
  float sum = 0.f;
  for (int base = start; base < end; base += kGenChunk) {
    const int len = min(kGenChunk, end - base);
    for (int i = threadIdx.x; i < len; i += kGenThreads) meta_s[i] = __ldg(tmeta + base + i);
    __syncthreads();
    if (active) {
      for (int i = 0; i < len; ++i) {
        const int4 m = meta_s[i];
        const int ty = iy - m.y, tx = ix - m.z;
        if (ty < 0 || tx < 0) continue;
        int oy = ty, ox = tx;
        if (stride_h != 1) {
          if (ty % stride_h) continue;
          oy = ty / stride_h;
        }
        if (stride_w != 1) {
          if (tx % stride_w) continue;
          ox = tx / stride_w;
        }
        if (oy < Ho && ox < Wo) sum = fmaf(__int_as_float(m.w), __ldg(td + ((long)m.x * Ho + oy) * Wo + ox), sum);
      }
    }
    __syncthreads();
  }
  if (active) bottom_diff[((long)n * C + c) * H * W + pix] = sum;
}

// ------------------------------------------------------------------------------------------------------
// generic backward-weight, restricted to the mask: one warp per nonzero,
//   g = sum_{n,oy,ox} top_diff[n][oc][oy][ox] * bottom[n][ic][oy*sh+dy][ox*sw+dx].
// wmeta (row-major order) = {oc, ic, dy, dx}.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    sconv_bwd_weight_generic_kernel(long nnz, const int4 *__restrict__ wmeta, const int *__restrict__ dense_idx,
                                    const int *__restrict__ csr_pos, const float *__restrict__ bottom,
                                    const float *__restrict__ top_diff, int num, int C, int H, int W, int M, int Ho,
                                    int Wo, int stride_h, int stride_w, float *__restrict__ wd_dense,
                                    float *__restrict__ wd_csr, int accumulate) {
  const long j = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= nnz) return;
  const int4 m = __ldg(wmeta + j);
  const int oc = m.x, ic = m.y, dy = m.z, dx = m.w;
  // valid output range so that the input coordinate is inside the image
  int oy_lo = 0, oy_hi = Ho, ox_lo = 0, ox_hi = Wo;
  while (oy_lo < Ho && oy_lo * stride_h + dy < 0) ++oy_lo;
  while (oy_hi > oy_lo && (oy_hi - 1) * stride_h + dy >= H) --oy_hi;
  while (ox_lo < Wo && ox_lo * stride_w + dx < 0) ++ox_lo;
  while (ox_hi > ox_lo && (ox_hi - 1) * stride_w + dx >= W) --ox_hi;
  const int ny = oy_hi - oy_lo, nx = ox_hi - ox_lo;
  float sum = 0.f;
  if (ny > 0 && nx > 0) {
    const int per_img = ny * nx;
    const long total = (long)num * per_img;
    for (long t = lane; t < total; t += 32) {
      const int n = (int)(t / per_img);
      const int r = (int)(t % per_img);
      const int oy = oy_lo + r / nx, ox = ox_lo + r % nx;
      const float a = __ldg(top_diff + (((long)n * M + oc) * Ho + oy) * Wo + ox);
      const float b = __ldg(bottom + (((long)n * C + ic) * H + oy * stride_h + dy) * W + ox * stride_w + dx);
      sum = fmaf(a, b, sum);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if (lane == 0) {
    if (wd_dense) wd_dense[dense_idx[j]] += sum;
    if (wd_csr) {
      const int p = csr_pos[j];
      wd_csr[p] = accumulate ? wd_csr[p] + sum : sum;
    }
  }
}

// bias_diff[oc] += sum_{n,pix} top_diff
__global__ void __launch_bounds__(256) bias_backward_kernel(int num, int M, int spatial,
                                                            const float *__restrict__ top_diff,
                                                            float *__restrict__ bias_diff) {
  __shared__ float part[8];
  const int oc = blockIdx.x;
  float sum = 0.f;
  const long total = (long)num * spatial;
  for (long t = threadIdx.x; t < total; t += blockDim.x) {
    const int n = (int)(t / spatial);
    const int p = (int)(t % spatial);
    sum += __ldg(top_diff + ((long)n * M + oc) * spatial + p);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += part[i];
    bias_diff[oc] += s;
  }
}

// values refresh: re-gather at fixed positions
__global__ void refresh_kernel(long nnz, const float *__restrict__ w_dense, const int *__restrict__ dense_idx,
                               const int *__restrict__ csr_pos, int4 *__restrict__ meta, float *__restrict__ values_csr) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  const float v = __ldg(w_dense + dense_idx[j]);
  meta[j].w = __float_as_int(v);
  if (values_csr) values_csr[csr_pos[j]] = v;
}
__global__ void refresh_t_kernel(long nnz, const int4 *__restrict__ meta, const int *__restrict__ tsrc,
                                 int4 *__restrict__ tmeta) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz) return;
  tmeta[t].w = meta[tsrc[t]].w;
}

template <typename T>
static int upload(T **dptr, const std::vector<T> &h, cudaStream_t stream) {
  *dptr = nullptr;
  const size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  ESCORT_CUDA(cudaMalloc((void **)dptr, bytes));
  if (!h.empty()) ESCORT_CUDA(cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
  return 0;
}

}  // namespace escort

using namespace escort;

extern "C" const char *escort_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *escort_version(void) { return "escort_b200 0.1 (sm_100a)"; }

extern "C" int escort_sconv_padded(int fuse_relu, int num, const float *input, int ifmap_size, const int *rowptr,
                                   const int *colidx, const float *values, const float *bias, int height, int width,
                                   int pad_h, int pad_w, int stride_h, int stride_w, int dilation_h, int dilation_w,
                                   int kernel_h, int kernel_w, float *output, int num_oc, int num_groups,
                                   escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(input && rowptr && colidx && values && output && num > 0 && num_oc > 0 && num_groups > 0,
                 "escort_sconv_padded: bad arguments");
  ESCORT_REQUIRE(!fuse_relu || bias, "escort_sconv_padded: fuse_relu requires bias (reference sconv_relu_*)");
  const int Ho = out_dim(height, pad_h, kernel_h, stride_h, dilation_h);
  const int Wo = out_dim(width, pad_w, kernel_w, stride_w, dilation_w);
  ESCORT_REQUIRE(Ho > 0 && Wo > 0, "escort_sconv_padded: empty output");
  dim3 grid(ceil_div(Ho * Wo, kCompatThreads), num_oc, num);
  const long in_stride = (long)ifmap_size * num_groups;
  const long out_stride = (long)num_oc * num_groups * Ho * Wo;
  if (dilation_h != 1 || dilation_w != 1)
    sconv_padded_kernel<true><<<grid, kCompatThreads, 0, stream>>>(
        rowptr, colidx, values, input, in_stride, bias, fuse_relu, output, out_stride, num_oc, height + pad_h,
        width + pad_w, stride_h, stride_w, dilation_h, dilation_w, Ho, Wo);
  else
    sconv_padded_kernel<false><<<grid, kCompatThreads, 0, stream>>>(
        rowptr, colidx, values, input, in_stride, bias, fuse_relu, output, out_stride, num_oc, height + pad_h,
        width + pad_w, stride_h, stride_w, 1, 1, Ho, Wo);
  ESCORT_LAUNCH_CHECK();
  return 0;
}

static int build_s2d_plan(escort_plan *p, cudaStream_t stream);

extern "C" int escort_plan_create(const escort_geom *geom, const int *rowptr, const int *colidx, const float *values,
                                  int colidx_is_stretched, escort_plan **plan_out, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(geom && rowptr && colidx && values && plan_out, "escort_plan_create: null argument");
  const escort_geom g = *geom;
  ESCORT_REQUIRE(g.channels > 0 && g.num_output > 0 && g.group > 0 && g.channels % g.group == 0 &&
                     g.num_output % g.group == 0,
                 "escort_plan_create: channels/num_output must be positive multiples of group");
  ESCORT_REQUIRE(g.kernel_h > 0 && g.kernel_w > 0 && g.stride_h > 0 && g.stride_w > 0 && g.dilation_h > 0 &&
                     g.dilation_w > 0 && g.pad_h >= 0 && g.pad_w >= 0 && g.height > 0 && g.width > 0,
                 "escort_plan_create: bad geometry");
  const int Ho = out_dim(g.height, g.pad_h, g.kernel_h, g.stride_h, g.dilation_h);
  const int Wo = out_dim(g.width, g.pad_w, g.kernel_w, g.stride_w, g.dilation_w);
  ESCORT_REQUIRE(Ho > 0 && Wo > 0, "escort_plan_create: empty output");
  const int Mg = g.num_output / g.group, Cg = g.channels / g.group;
  const long weight_offset = (long)Mg * Cg * g.kernel_h * g.kernel_w;
  const int row_offset = Mg + 1;

  // fetch the layer's CSR blobs
  std::vector<int> h_rowptr((size_t)g.num_output + g.group);
  ESCORT_CUDA(cudaMemcpyAsync(h_rowptr.data(), rowptr, h_rowptr.size() * sizeof(int), cudaMemcpyDeviceToHost, stream));
  ESCORT_CUDA(cudaStreamSynchronize(stream));
  std::vector<std::vector<int>> h_col(g.group);
  std::vector<std::vector<float>> h_val(g.group);
  long nnz = 0;
  for (int gi = 0; gi < g.group; ++gi) {
    const int *rp = h_rowptr.data() + (size_t)row_offset * gi;
    ESCORT_REQUIRE(rp[0] == 0, "escort_plan_create: each group's rowptr must start at 0");
    for (int i = 0; i < Mg; ++i) ESCORT_REQUIRE(rp[i + 1] >= rp[i], "escort_plan_create: rowptr not monotone");
    const int n_g = rp[Mg];
    ESCORT_REQUIRE(n_g <= weight_offset, "escort_plan_create: nnz exceeds dense size");
    h_col[gi].resize(n_g);
    h_val[gi].resize(n_g);
    if (n_g) {
      ESCORT_CUDA(cudaMemcpyAsync(h_col[gi].data(), colidx + weight_offset * gi, n_g * sizeof(int),
                                  cudaMemcpyDeviceToHost, stream));
      ESCORT_CUDA(cudaMemcpyAsync(h_val[gi].data(), values + weight_offset * gi, n_g * sizeof(float),
                                  cudaMemcpyDeviceToHost, stream));
    }
    nnz += n_g;
  }
  ESCORT_CUDA(cudaStreamSynchronize(stream));

  auto *nzs = new std::vector<Nz>();
  nzs->reserve(nnz);
  std::vector<int> g_rowptr(g.num_output + 1, 0);
  const int Hp = g.height + g.pad_h, Wp = g.width + g.pad_w;
  for (int gi = 0; gi < g.group; ++gi) {
    const int *rp = h_rowptr.data() + (size_t)row_offset * gi;
    for (int i = 0; i < Mg; ++i) {
      for (int j = rp[i]; j < rp[i + 1]; ++j) {
        const int col = h_col[gi][j];
        int ic, kh, kw;
        if (colidx_is_stretched) {
          kw = col % Wp;
          kh = (col / Wp) % Hp;
          ic = col / (Wp * Hp);
        } else {
          kw = col % g.kernel_w;
          kh = (col / g.kernel_w) % g.kernel_h;
          ic = col / (g.kernel_w * g.kernel_h);
        }
        if (ic < 0 || ic >= Cg || kh >= g.kernel_h || kw >= g.kernel_w || col < 0) {
          delete nzs;
          set_last_error("escort_plan_create: column index out of range (wrong colidx_is_stretched?)");
          return ESCORT_EINVAL;
        }
        Nz z;
        z.oc = gi * Mg + i;
        z.ic = gi * Cg + ic;
        z.kh = kh;
        z.kw = kw;
        z.val = h_val[gi][j];
        z.csr_pos = (int)(weight_offset * gi + j);
        z.dense_idx = (int)((((long)z.oc * Cg + ic) * g.kernel_h + kh) * g.kernel_w + kw);
        nzs->push_back(z);
      }
      g_rowptr[gi * Mg + i + 1] = (int)nzs->size();
    }
  }

  escort_plan *p = new escort_plan();
  memset(p, 0, sizeof(*p));
  p->g = g;
  p->Ho = Ho;
  p->Wo = Wo;
  p->nnz = nnz;
  p->variant = -1;
  p->layout_rank = 0;
  p->host_nz = nzs;
  p->generic_backward = getenv("ESCORT_GENERIC_BACKWARD") ? 1 : 0;
  p->small_maps = getenv("ESCORT_NO_SMALL_MAPS") ? 0 : 1;
  p->mu = new std::mutex();
  cudaGetDevice(&p->device);

  const int H = g.height, W = g.width;
  std::vector<int4> meta(nnz), wmeta(nnz), tmeta(nnz);
  std::vector<int> dense_idx(nnz), csr_pos(nnz), tsrc(nnz), colptr(g.channels + 1, 0);
  for (long j = 0; j < nnz; ++j) {
    const Nz &z = (*nzs)[j];
    const int dy = z.kh * g.dilation_h - g.pad_h, dx = z.kw * g.dilation_w - g.pad_w;
    meta[j] = make_int4(z.ic * H * W + dy * W + dx, dy, dx, __builtin_bit_cast(int, z.val));
    wmeta[j] = make_int4(z.oc, z.ic, dy, dx);
    dense_idx[j] = z.dense_idx;
    csr_pos[j] = z.csr_pos;
    colptr[z.ic + 1]++;
  }
  for (int c = 0; c < g.channels; ++c) colptr[c + 1] += colptr[c];
  {
    std::vector<int> cursor(colptr.begin(), colptr.end() - 1);
    for (long j = 0; j < nnz; ++j) {
      const Nz &z = (*nzs)[j];
      const int t = cursor[z.ic]++;
      tmeta[t] = make_int4(z.oc, z.kh * g.dilation_h - g.pad_h, z.kw * g.dilation_w - g.pad_w,
                           __builtin_bit_cast(int, z.val));
      tsrc[t] = (int)j;
    }
  }
  int rc = 0;
  if ((rc = upload(&p->d_rowptr, g_rowptr, stream)) || (rc = upload(&p->d_meta, meta, stream)) ||
      (rc = upload(&p->d_dense_idx, dense_idx, stream)) || (rc = upload(&p->d_csr_pos, csr_pos, stream)) ||
      (rc = upload(&p->d_colptr, colptr, stream)) || (rc = upload(&p->d_tmeta, tmeta, stream)) ||
      (rc = upload(&p->d_tsrc, tsrc, stream)) || (rc = upload(&p->d_wmeta, wmeta, stream))) {
    escort_plan_destroy(p);
    return rc;
  }
  rc = tile_plan_build(p, -1, stream);
  if (!rc) rc = build_s2d_plan(p, stream);  // stride 2: the space-to-depth sub-plan (stride-1 kernels), default path when it exists
  if (rc) {
    escort_plan_destroy(p);
    return rc;
  }
  ESCORT_CUDA(cudaStreamSynchronize(stream));  // host staging vectors go out of scope
  *plan_out = p;
  return 0;
}

extern "C" int escort_plan_destroy(escort_plan *p) {
  if (!p) return 0;
  cudaFree(p->d_rowptr);
  cudaFree(p->d_meta);
  cudaFree(p->d_dense_idx);
  cudaFree(p->d_csr_pos);
  cudaFree(p->d_colptr);
  cudaFree(p->d_tmeta);
  cudaFree(p->d_tsrc);
  cudaFree(p->d_wmeta);
  if (p->tile) tile_plan_free(p->tile);
  if (p->tm) tmem_plan_free(p->tm);
  if (p->tile_w) tile_plan_free(p->tile_w);
  if (p->bwd) escort_plan_destroy(p->bwd);
  if (p->s2d) escort_plan_destroy(p->s2d);
  cudaFree(p->s2d_buf);
  delete p->host_nz;
  delete p->mu;
  delete p;
  return 0;
}

// Backward data as a forward convolution (stride 1): bottom_diff = conv(top_diff, W'), with
// W'[ic][oc][kh'][kw'] = W[oc][ic][K-1-kh'][K-1-kw'], the same dilation and pad' = dilation*(K-1)-pad, so the tile kernel (and its plan-time
// compiler) serves both directions.  The sub-plan indexes the ORIGINAL dense weight tensor, so escort_refresh_values
// refreshes it from the same weights.  Returns 0 and leaves p->bwd null when the geometry does not qualify.
static int build_bwd_plan(escort_plan *p, cudaStream_t stream) {
  p->bwd_tried = 1;
  const escort_geom &g = p->g;
  if (g.stride_h != 1 || g.stride_w != 1) return 0;  // (stride 2 goes through the space-to-depth sub-plan's own backward plan)
  // dilation carries over (same dilation, pad' = dilation * (K - 1) - pad): the TMEM-window kernels take it, the tile
  // kernels do not -- tile_plan_build then leaves the sub-plan empty and the generic kernel stays
  const int ph = g.dilation_h * (g.kernel_h - 1) - g.pad_h, pw = g.dilation_w * (g.kernel_w - 1) - g.pad_w;
  if (ph < 0 || pw < 0 || p->nnz == 0) return 0;
  escort_plan *q = new escort_plan();
  memset(q, 0, sizeof(*q));
  q->g = g;
  q->g.channels = g.num_output;
  q->g.num_output = g.channels;
  q->g.height = p->Ho;
  q->g.width = p->Wo;
  q->g.pad_h = ph;
  q->g.pad_w = pw;
  q->Ho = out_dim(q->g.height, q->g.pad_h, g.kernel_h, 1, g.dilation_h);
  q->Wo = out_dim(q->g.width, q->g.pad_w, g.kernel_w, 1, g.dilation_w);
  if (q->Ho != g.height || q->Wo != g.width) {
    delete q;
    return 0;
  }
  q->device = p->device;
  q->generic_backward = p->generic_backward;
  q->nnz = p->nnz;
  q->variant = -1;
  q->host_nz = new std::vector<Nz>();
  q->host_nz->reserve(p->nnz);
  std::vector<int> dense_idx;
  dense_idx.reserve(p->nnz);
  for (const Nz &z : *p->host_nz) {
    Nz t = z;
    t.oc = z.ic;
    t.ic = z.oc;
    t.kh = g.kernel_h - 1 - z.kh;
    t.kw = g.kernel_w - 1 - z.kw;
    q->host_nz->push_back(t);
    dense_idx.push_back(z.dense_idx);
  }
  int rc = upload(&q->d_dense_idx, dense_idx, stream);
  if (!rc) rc = tile_plan_build(q, -1, stream);
  if (!rc) {
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
  }
  if (rc || (!q->tile && !q->tm)) {
    escort_plan_destroy(q);
    return rc;
  }
  q->parent = p;
  const escort_plan *root = p;
  while (root->parent) root = root->parent;  // (p may itself be the space-to-depth sub-plan of a stride-2 layer)
  if (root->refreshed && root->d_meta) {
    rc = tile_regather(q, root->d_meta, stream);
    if (rc) {
      escort_plan_destroy(q);
      return rc;
    }
  }
  p->bwd = q;
  return 0;
}

// Stride-2 forward without a stride-2 kernel.  With P = the zero-padded input,
//   out[oy][ox] = sum w[kh][kw] * P[2 oy + kh][2 ox + kw] = sum w[kh][kw] * P_{kh & 1, kw & 1}[oy + (kh >> 1)][ox + (kw >> 1)]
// where P_{py,px}[y][x] = P[2 y + py][2 x + px] are the four parity planes of P: a stride-2 K x K convolution with padding
// is a VALID stride-1 convolution with a ceil(K / 2) kernel over 4 x channels planes of (Ho + KH2 - 1) x (Wo + KW2 - 1)
// (space-to-depth).  The nonzeros map one to one, (ic, kh, kw) -> (4 ic + 2 (kh & 1) + (kw & 1), kh >> 1, kw >> 1), so
// the sub-plan is a forward plan over the same weights (same refresh path as the backward-data sub-plan) and runs on
// the stride-1 TMEM / tile kernels; the parity planes are written by one extra pass over the input (s2d_pad_kernel).
__global__ void s2d_pad_kernel(long total, const float *__restrict__ in, int C, int H, int W, int pad_h, int pad_w, int H2, int W2,
                               float *__restrict__ out) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int x2 = (int)(e % W2);
  long r = e / W2;
  const int y2 = (int)(r % H2);
  r /= H2;
  const int vc = (int)(r % (4 * C));
  const long n = r / (4 * C);
  const int ic = vc >> 2, py = (vc >> 1) & 1, px = vc & 1;
  const int y = 2 * y2 + py - pad_h, x = 2 * x2 + px - pad_w;
  float v = 0.f;
  if ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) v = __ldg(in + ((n * C + ic) * H + y) * W + x);
  out[e] = v;
}

// Without a measurement (no escort_plan_autotune): the parity-plane path wins on the larger layers of the stride-2 sweep
// (channels x height >= 128 x 56: 1.1-1.5x, profiles/r02_sweep_config5_stride2.txt) and loses on the small ones, where
// a window of the 4 x channels planes meets about one tap per output channel.
// the inverse map for the backward data: bottom_diff[n][c][y][x] = dP_{py,px}[y2][x2] with 2 y2 + py = y + pad (a pixel
// beyond the last window of the layer belongs to no parity-plane element: its gradient is zero)
__global__ void d2s_unpad_kernel(long total, const float *__restrict__ dplanes, int C, int H, int W, int pad_h, int pad_w, int H2, int W2,
                                 float *__restrict__ out) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int x = (int)(e % W);
  long r = e / W;
  const int y = (int)(r % H);
  r /= H;
  const int c = (int)(r % C);
  const long n = r / C;
  const int yp = y + pad_h, xp = x + pad_w;
  const int y2 = yp >> 1, x2 = xp >> 1, vc = 4 * c + 2 * (yp & 1) + (xp & 1);
  out[e] = (y2 < H2 && x2 < W2) ? __ldg(dplanes + ((n * 4 * C + vc) * H2 + y2) * W2 + x2) : 0.f;
}

static int s2d_by_default(const escort_plan *p) {
  return p->s2d && (long)(p->g.channels / p->g.group) * p->g.height >= 128L * 56 ? 1 : 0;
}

static int build_s2d_plan(escort_plan *p, cudaStream_t stream) {
  const escort_geom &g = p->g;
  if (g.stride_h != 2 || g.stride_w != 2 || g.dilation_h != 1 || g.dilation_w != 1 || p->nnz == 0) return 0;
  if (getenv("ESCORT_NO_S2D") || p->parent) return 0;
  const int KH2 = (g.kernel_h - 1) / 2 + 1, KW2 = (g.kernel_w - 1) / 2 + 1;
  escort_plan *q = new escort_plan();
  memset(q, 0, sizeof(*q));
  q->g = g;
  q->g.channels = 4 * g.channels;
  q->g.height = p->Ho + KH2 - 1;
  q->g.width = p->Wo + KW2 - 1;
  q->g.kernel_h = KH2;
  q->g.kernel_w = KW2;
  q->g.pad_h = q->g.pad_w = 0;
  q->g.stride_h = q->g.stride_w = 1;
  q->Ho = p->Ho;
  q->Wo = p->Wo;
  q->device = p->device;
  q->generic_backward = p->generic_backward;  // its own backward-data sub-plan serves the layer's stride-2 backward data
  q->tile_w_tried = 1;                          // (no W variant for the parity-plane kernel sizes)
  q->mu = new std::mutex();
  q->nnz = p->nnz;
  q->variant = -1;
  q->host_nz = new std::vector<Nz>();
  q->host_nz->reserve(p->nnz);
  std::vector<int> dense_idx;
  dense_idx.reserve(p->nnz);
  for (const Nz &z : *p->host_nz) {  // (same order as the layer's nonzeros: d_meta[j] is nonzero j of both)
    Nz t = z;
    t.ic = 4 * z.ic + 2 * (z.kh & 1) + (z.kw & 1);
    t.kh = z.kh >> 1;
    t.kw = z.kw >> 1;
    q->host_nz->push_back(t);
    dense_idx.push_back(z.dense_idx);
  }
  int rc = upload(&q->d_dense_idx, dense_idx, stream);
  if (!rc) rc = tile_plan_build(q, -1, stream);
  if (!rc) {
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
  }
  if (rc || (!q->tile && !q->tm)) {
    escort_plan_destroy(q);
    return rc;
  }
  q->parent = p;
  p->s2d = q;
  p->use_s2d = s2d_by_default(p);
  return 0;
}

extern "C" long escort_plan_nnz(const escort_plan *p) { return p ? p->nnz : -1; }

extern "C" const char *escort_plan_kernel_name(const escort_plan *p) {
  if (!p) return "null";
  if (p->use_s2d && p->s2d) return escort_plan_kernel_name(p->s2d);  // (escort_plan_describe says "s2d: ...")
  if (p->tm) return tmem_kernel_name(p->tm);
  if (p->tile) return tile_kernel_name(p->tile);
  const size_t img_bytes = (size_t)p->g.channels * p->g.height * p->g.width * sizeof(float);
  return (img_bytes <= kSmallMaxBytes && p->small_maps) ? "sconv_fwd_small" : "sconv_fwd_generic";
}

extern "C" int escort_plan_set_variant(escort_plan *p, int variant) {
  return escort_plan_set_config(p, variant, 0);
}

extern "C" int escort_plan_get_config(const escort_plan *p, int *variant_host, int *layout_rank_host) {
  ESCORT_REQUIRE(p, "escort_plan_get_config: null plan");
  if (variant_host) *variant_host = p->variant;
  if (layout_rank_host) *layout_rank_host = p->layout_rank;
  return 0;
}

extern "C" int escort_plan_set_config(escort_plan *p, int variant, int layout_rank) {
  ESCORT_REQUIRE(p && layout_rank >= 0, "escort_plan_set_config: bad arguments");
  if (variant == -2) {  // the space-to-depth path of a stride-2 layer (its sub-plan keeps its own configuration)
    ESCORT_REQUIRE(p->s2d, "escort_plan_set_config: variant -2 (space-to-depth) needs a stride-2 layer");
    p->use_s2d = 1;
    p->variant = -2;
    return 0;
  }
  p->use_s2d = variant == -1 ? s2d_by_default(p) : 0;  // an explicit variant means that kernel on the strided input
  p->layout_rank = layout_rank;
  if (p->tile) {
    tile_plan_free(p->tile);
    p->tile = nullptr;
  }
  if (p->tm) {
    tmem_plan_free(p->tm);
    p->tm = nullptr;
  }
  p->variant = variant;
  if (variant == 0) return 0;
  int rc = tile_plan_build(p, variant, 0);
  if (rc) return rc;
  // a stream built after escort_refresh_values starts from the create-time snapshot (host_nz): re-gather the current
  // values from the layer plan's device copy
  const escort_plan *root = p;
  while (root->parent) root = root->parent;  // (the backward-data plan of a space-to-depth sub-plan is two levels down)
  if (root->refreshed && root->d_meta && (p->tile || p->tm)) {
    rc = tile_regather(p, root->d_meta, 0);
    if (rc) return rc;
  }
  ESCORT_CUDA(cudaStreamSynchronize(0));
  if (variant > 0 && !p->tile && !p->tm) {
    set_last_error("escort_plan_set_config: geometry / layout candidate not supported by the requested variant");
    return ESCORT_EINVAL;
  }
  return 0;
}

// Plan-time selection of the forward variant by measurement (the cuDNN "find" idiom): every variant that
// supports the geometry is built with the planner's favourite tiling and timed on a zero-filled scratch batch of
// `num` images on `stream`; the three fastest are then re-timed with the runner-up tilings and the best
// (variant, tiling) is kept.  The backward-data sub-plan (a forward plan over the transposed weights) is tuned the
// same way.  Called by the host layer once after WeightAlign; synchronises the stream.
static int autotune_one(escort_plan *p, int num, cudaStream_t stream) {
  const escort_geom &g = p->g;
  const size_t in_elems = (size_t)num * g.channels * g.height * g.width;
  const size_t out_elems = (size_t)num * g.num_output * p->Ho * p->Wo;
  float *x = nullptr, *y = nullptr;
  ESCORT_CUDA(cudaMalloc((void **)&x, in_elems * sizeof(float)));
  cudaError_t e = cudaMalloc((void **)&y, out_elems * sizeof(float));
  if (e != cudaSuccess) {
    cudaFree(x);
    return cuda_fail(e, "cudaMalloc(autotune scratch)", __FILE__, __LINE__);
  }
  cudaMemsetAsync(x, 0, in_elems * sizeof(float), stream);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto time_config = [&](int v, int rank) -> float {
    if (escort_plan_set_config(p, v, rank) != 0) return -1.f;  // no such layout candidate
    if (v > 0 && !p->tile && !p->tm) return -1.f;
    float ms_best = 1e30f;
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0, stream);
      if (escort_sconv_forward(p, num, x, nullptr, 0, y, stream) != 0) return -1.f;
      cudaEventRecord(e1, stream);
      if (cudaEventSynchronize(e1) != cudaSuccess) return -1.f;
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      if (it > 0 && ms < ms_best) ms_best = ms;
    }
    return ms_best;
  };
  std::vector<std::pair<float, int>> first;  // (ms, variant) with the favourite tiling
  const int nvar = tile_num_variants();
  for (int v = (p->d_rowptr ? 0 : 1); v <= nvar; ++v) {  // (a backward sub-plan has no generic kernel)
    if (v > 0 && !tile_variant_applies(p, v)) continue;
    const float ms = time_config(v, 0);
    if (ms > 0.f) first.push_back({ms, v});
  }
  std::sort(first.begin(), first.end());
  int best_v = 0, best_rank = 0;
  float best_ms = 1e30f;
  if (!first.empty()) {
    best_ms = first[0].first;
    best_v = first[0].second;
  }
  for (size_t i = 0; i < first.size() && i < 3; ++i) {
    const int v = first[i].second;
    if (v == 0) continue;
    for (int rank = 1; rank < 4; ++rank) {
      const float ms = time_config(v, rank);
      if (ms < 0.f) break;
      if (ms < best_ms) {
        best_ms = ms;
        best_v = v;
        best_rank = rank;
      }
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(x);
  cudaFree(y);
  cudaGetLastError();
  if (first.empty() && !p->d_rowptr) return escort_plan_set_config(p, -1, 0);  // a backward sub-plan has no generic kernel: keep its default
  return escort_plan_set_config(p, best_v, best_rank);
}

// Apply the tuning of another plan of the same geometry (forward variant + layout, backward-data sub-plan, backward
// weight variant) without measuring again: layers of one shape (ResNet-50 has 16 branch2b convs in 4 shapes) are
// tuned once.
extern "C" int escort_plan_copy_tuning(escort_plan *dst, const escort_plan *src, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(dst && src, "escort_plan_copy_tuning: null plan");
  ESCORT_REQUIRE(memcmp(&dst->g, &src->g, sizeof(escort_geom)) == 0, "escort_plan_copy_tuning: geometries differ");
  int rc = escort_plan_set_config(dst, src->variant, src->layout_rank);
  if (rc) return rc;
  if (src->s2d && dst->s2d && (rc = escort_plan_set_config(dst->s2d, src->s2d->variant, src->s2d->layout_rank))) return rc;
  if (src->bwd) {
    if (!dst->bwd && !dst->bwd_tried) {
      rc = build_bwd_plan(dst, stream);
      if (rc) return rc;
    }
    if (dst->bwd) {
      rc = escort_plan_set_config(dst->bwd, src->bwd->variant, src->bwd->layout_rank);
      if (rc) return rc;
    }
  }
  if (src->tile_w) {
    dst->tile_w_tried = 1;
    rc = tile_bwdw_build(dst, stream, tile_plan_variant(src->tile_w));
    if (rc) return rc;
    ESCORT_CUDA(cudaStreamSynchronize(stream));
  }
  return 0;
}

extern "C" int escort_plan_autotune(escort_plan *p, int num, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(p && num > 0, "escort_plan_autotune: bad arguments");
  int rc = autotune_one(p, num, stream);
  if (rc || !p->s2d) return rc;
  // stride 2: the best kernel on the strided input (just chosen) against the space-to-depth path with its own tuned sub-plan
  const int direct_v = p->variant, direct_rank = p->layout_rank;
  if ((rc = autotune_one(p->s2d, num, stream))) return rc;
  const escort_geom &g = p->g;
  float *x = nullptr, *y = nullptr;
  ESCORT_CUDA(cudaMalloc((void **)&x, (size_t)num * g.channels * g.height * g.width * sizeof(float)));
  cudaError_t e = cudaMalloc((void **)&y, (size_t)num * g.num_output * p->Ho * p->Wo * sizeof(float));
  if (e != cudaSuccess) {
    cudaFree(x);
    return cuda_fail(e, "cudaMalloc(autotune scratch)", __FILE__, __LINE__);
  }
  cudaMemsetAsync(x, 0, (size_t)num * g.channels * g.height * g.width * sizeof(float), stream);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best[2] = {1e30f, 1e30f};
  for (int mode = 0; mode < 2; ++mode) {
    p->use_s2d = mode;
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0, stream);
      const int frc = escort_sconv_forward(p, num, x, nullptr, 0, y, stream);
      cudaEventRecord(e1, stream);
      float ms = 1e30f;
      if (cudaEventSynchronize(e1) == cudaSuccess && frc == 0) cudaEventElapsedTime(&ms, e0, e1);
      if (it > 0 && ms < best[mode]) best[mode] = ms;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(x);
  cudaFree(y);
  cudaGetLastError();
  p->use_s2d = best[1] < best[0] ? 1 : 0;
  p->variant = p->use_s2d ? -2 : direct_v;
  p->layout_rank = direct_rank;
  return 0;
}

extern "C" int escort_plan_autotune_backward(escort_plan *p, int num, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(p && num > 0, "escort_plan_autotune_backward: bad arguments");
  std::lock_guard<std::mutex> lock(*p->mu);
  if (!p->bwd && !p->bwd_tried && !p->generic_backward) {
    int rc = build_bwd_plan(p, stream);
    if (rc) return rc;
  }
  if (p->bwd) {
    int rc = autotune_one(p->bwd, num, stream);
    if (rc) return rc;
  }
  if (p->s2d && !p->generic_backward) {  // stride 2: the backward-data plan of the space-to-depth sub-plan
    escort_plan *q = p->s2d;
    if (!q->bwd && !q->bwd_tried) {
      int rc = build_bwd_plan(q, stream);
      if (rc) return rc;
    }
    if (q->bwd) {
      int rc = autotune_one(q->bwd, num, stream);
      if (rc) return rc;
    }
  }
  if (p->generic_backward || getenv("ESCORT_BWDW_VARIANT")) return 0;
  return tile_bwdw_autotune(p, num, stream);  // and the backward-weight (W) variant
}

extern "C" int escort_sconv_forward(escort_plan *p, int num, const float *bottom, const float *bias, int fuse_relu,
                                    float *top, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(p && num >= 0, "escort_sconv_forward: bad arguments");
  if (num == 0) return 0;  // empty batch: nothing to do (pointers may be null)
  ESCORT_REQUIRE(bottom && top, "escort_sconv_forward: null tensor");
  if (p->use_s2d && p->s2d) {  // stride 2: parity planes of the padded input, then the stride-1 sub-plan
    const escort_geom &q = p->s2d->g;
    const size_t need = (size_t)num * q.channels * q.height * q.width;
    {
      std::lock_guard<std::mutex> lock(*p->mu);
      if (need > p->s2d_elems) {  // first call / larger batch: (re)allocate the library-owned buffer (synchronises the device)
        cudaFree(p->s2d_buf);
        p->s2d_buf = nullptr;
        p->s2d_elems = 0;
        ESCORT_CUDA(cudaMalloc((void **)&p->s2d_buf, need * sizeof(float)));
        p->s2d_elems = need;
      }
    }
    s2d_pad_kernel<<<(unsigned)((need + 255) / 256), 256, 0, stream>>>((long)need, bottom, p->g.channels, p->g.height, p->g.width,
                                                                     p->g.pad_h, p->g.pad_w, q.height, q.width, p->s2d_buf);
    ESCORT_LAUNCH_CHECK();
    return escort_sconv_forward(p->s2d, num, p->s2d_buf, bias, fuse_relu, top, stream_);
  }
  if (p->tm && tmem_batch_fits(p, num)) return tmem_forward(p, num, bottom, bias, fuse_relu, top, stream);
  if (p->tile) {
    const int rc = tile_forward(p, num, bottom, bias, fuse_relu, top, stream);
    if (rc != ESCORT_ETRYGENERIC) return rc;
  }
  ESCORT_REQUIRE(p->d_rowptr, "escort_sconv_forward: this launch needs the generic kernel, which a backward sub-plan does not have");
  const escort_geom &g = p->g;
  const size_t img_bytes = (size_t)g.channels * g.height * g.width * sizeof(float);
  if (img_bytes <= kSmallMaxBytes && p->small_maps) {  // the whole input image fits in shared memory: the small-map kernel
    static std::once_flag once;
    std::call_once(once, [] {
      cudaFuncSetAttribute(sconv_fwd_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallMaxBytes + 4096 + 16);
      cudaFuncSetAttribute(sconv_fwd_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallMaxBytes + 4096 + 16);
    });
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
    const int splits = std::max(1, std::min(g.num_output, ceil_div(2 * sms, num)));
    const int oc_per_cta = ceil_div(g.num_output, splits);
    const size_t smem = ((img_bytes + 15) & ~(size_t)15) + kSmallThreads / 32 * 32 * sizeof(int4);
    auto kern = (g.pad_h == 0 && g.pad_w == 0) ? sconv_fwd_small_kernel<false> : sconv_fwd_small_kernel<true>;
    kern<<<(unsigned)(num * splits), kSmallThreads, smem, stream>>>(p->d_rowptr, p->d_meta, bottom, bias, fuse_relu, top,
                                                                  g.channels * g.height * g.width, g.height, g.width, g.num_output, p->Ho,
                                                                  p->Wo, g.stride_h, g.stride_w, splits, oc_per_cta);
    ESCORT_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid(ceil_div(p->Ho * p->Wo, kGenThreads), g.num_output, num);
  ESCORT_REQUIRE(num <= 65535, "escort_sconv_forward: batch too large for the generic kernel");
  sconv_fwd_generic_kernel<<<grid, kGenThreads, 0, stream>>>(p->d_rowptr, p->d_meta, bottom, bias, fuse_relu, top,
                                                             g.channels, g.height, g.width, g.num_output, p->Ho,
                                                             p->Wo, g.stride_h, g.stride_w);
  ESCORT_LAUNCH_CHECK();
  return 0;
}

extern "C" int escort_sconv_backward_data(escort_plan *p, int num, const float *top_diff, float *bottom_diff,
                                          escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(p && num >= 0 && num <= 65535, "escort_sconv_backward_data: bad arguments");
  if (num == 0) return 0;
  ESCORT_REQUIRE(top_diff && bottom_diff, "escort_sconv_backward_data: null tensor");
  if (p->s2d && !p->generic_backward) {
    // stride 2: the gradient of the parity planes through the sub-plan's own backward-data plan (a stride-1 forward
    // kernel over its transposed weights), then depth-to-space without the padding
    escort_plan *q = p->s2d;
    const size_t need = (size_t)num * q->g.channels * q->g.height * q->g.width;
    {
      std::lock_guard<std::mutex> lock(*p->mu);
      if (need > p->s2d_elems) {
        cudaFree(p->s2d_buf);
        p->s2d_buf = nullptr;
        p->s2d_elems = 0;
        ESCORT_CUDA(cudaMalloc((void **)&p->s2d_buf, need * sizeof(float)));
        p->s2d_elems = need;
      }
    }
    int rc = escort_sconv_backward_data(q, num, top_diff, p->s2d_buf, stream_);
    if (rc) return rc;
    const long total = (long)num * p->g.channels * p->g.height * p->g.width;
    d2s_unpad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(total, p->s2d_buf, p->g.channels, p->g.height, p->g.width,
                                                                        p->g.pad_h, p->g.pad_w, q->g.height, q->g.width, bottom_diff);
    ESCORT_LAUNCH_CHECK();
    return 0;
  }
  if (!p->bwd && !p->bwd_tried && !p->generic_backward) {
    std::lock_guard<std::mutex> lock(*p->mu);  // first use: one thread builds, the others find it built
    if (!p->bwd && !p->bwd_tried) {
      int rc = build_bwd_plan(p, stream);
      if (rc) return rc;
    }
  }
  if (p->bwd && !p->generic_backward && (p->bwd->tile || (p->bwd->tm && tmem_batch_fits(p->bwd, num)))) {
    int rc = ESCORT_ETRYGENERIC;
    if (p->bwd->tm && tmem_batch_fits(p->bwd, num)) rc = tmem_forward(p->bwd, num, top_diff, nullptr, 0, bottom_diff, stream);
    else if (p->bwd->tile) rc = tile_forward(p->bwd, num, top_diff, nullptr, 0, bottom_diff, stream);
    if (rc != ESCORT_ETRYGENERIC) return rc;
  }
  ESCORT_REQUIRE(p->d_colptr, "escort_sconv_backward_data: this launch needs the generic kernel, which a sub-plan does not have");
  const escort_geom &g = p->g;
  dim3 grid(ceil_div(g.height * g.width, kGenThreads), g.channels, num);
  sconv_bwd_data_generic_kernel<<<grid, kGenThreads, 0, stream>>>(p->d_colptr, p->d_tmeta, top_diff, bottom_diff,
                                                                  g.channels, g.height, g.width, g.num_output, p->Ho,
                                                                  p->Wo, g.stride_h, g.stride_w);
  ESCORT_LAUNCH_CHECK();
  return 0;
}

extern "C" int escort_sconv_backward_weight(escort_plan *p, int num, const float *bottom, const float *top_diff,
                                            float *weight_diff_dense, float *weight_diff_csr, int accumulate,
                                            escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(p && num >= 0, "escort_sconv_backward_weight: bad arguments");
  ESCORT_REQUIRE(weight_diff_dense || weight_diff_csr, "escort_sconv_backward_weight: no output buffer");
  if (num == 0 || p->nnz == 0) return 0;
  ESCORT_REQUIRE(bottom && top_diff, "escort_sconv_backward_weight: null tensor");
  if (!p->tile_w && !p->tile_w_tried && !p->generic_backward) {
    std::lock_guard<std::mutex> lock(*p->mu);
    if (!p->tile_w && !p->tile_w_tried) {
      int rc = tile_bwdw_build(p, stream);
      p->tile_w_tried = 1;
      if (rc) return rc;
      ESCORT_CUDA(cudaStreamSynchronize(stream));
    }
  }
  if (p->tile_w && !p->generic_backward) {
    const int rc = tile_bwdw(p, num, bottom, top_diff, weight_diff_dense, weight_diff_csr, accumulate, stream);
    if (rc != ESCORT_ETRYGENERIC) return rc;
  }
  const escort_geom &g = p->g;
  const int warps = 8;
  const long blocks = (p->nnz + warps - 1) / warps;
  sconv_bwd_weight_generic_kernel<<<(unsigned)blocks, warps * 32, 0, stream>>>(
      p->nnz, p->d_wmeta, p->d_dense_idx, p->d_csr_pos, bottom, top_diff, num, g.channels, g.height, g.width,
      g.num_output, p->Ho, p->Wo, g.stride_h, g.stride_w, weight_diff_dense, weight_diff_csr, accumulate);
  ESCORT_LAUNCH_CHECK();
  return 0;
}

extern "C" int escort_bias_backward(int num, int num_output, int out_spatial, const float *top_diff, float *bias_diff,
                                    escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(top_diff && bias_diff && num >= 0 && num_output > 0 && out_spatial > 0,
                 "escort_bias_backward: bad arguments");
  if (num == 0) return 0;
  bias_backward_kernel<<<num_output, 256, 0, stream>>>(num, num_output, out_spatial, top_diff, bias_diff);
  ESCORT_LAUNCH_CHECK();
  return 0;
}

extern "C" int escort_refresh_values(escort_plan *p, const float *weights_dense, float *values_csr,
                                     escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(p && weights_dense, "escort_refresh_values: bad arguments");
  if (p->nnz == 0) return 0;
  const unsigned blocks = (unsigned)((p->nnz + 255) / 256);
  refresh_kernel<<<blocks, 256, 0, stream>>>(p->nnz, weights_dense, p->d_dense_idx, p->d_csr_pos, p->d_meta, values_csr);
  ESCORT_LAUNCH_CHECK();
  refresh_t_kernel<<<blocks, 256, 0, stream>>>(p->nnz, p->d_meta, p->d_tsrc, p->d_tmeta);
  ESCORT_LAUNCH_CHECK();
  if (!p->bwd && !p->bwd_tried && !p->generic_backward) {
    // a refresh means training: build the backward-data sub-plan now, so that it never starts from stale values
    std::lock_guard<std::mutex> lock(*p->mu);
    if (!p->bwd && !p->bwd_tried) {
      int rc = build_bwd_plan(p, stream);
      if (rc) return rc;
    }
  }
  if (p->bwd) {
    int rc = tile_refresh(p->bwd, weights_dense, stream);
    if (rc) return rc;
  }
  if (p->s2d) {
    int rc = tile_refresh(p->s2d, weights_dense, stream);  // (a backward plan of the sub-plan built later re-gathers from d_meta)
    if (!rc && p->s2d->bwd) rc = tile_refresh(p->s2d->bwd, weights_dense, stream);
    if (rc) return rc;
  }
  p->refreshed = 1;
  if (p->tile || p->tm) return tile_refresh(p, weights_dense, stream);
  return 0;
}

// ---- f2: glue-layer fusion ------------------------------------------------------------------------------------
// conv -> BatchNorm (use_global_stats) -> Scale -> ReLU is the chain every ResNet-50 branch2b conv sits in; the reference
// runs four layers and four passes over the activation (net.cpp:531-532 counts them as "other time").  At inference
// the BatchNorm + Scale pair is one affine map per output channel, y = conv * a[oc] + b[oc], and that folds into the
// plan: the nonzero weights of row oc are multiplied by a[oc] (positions, hence the mask and every kernel's record
// stream, are unchanged) and the bias becomes bias * a + b.  The forward launch with fuse_relu then IS the whole chain:
// no extra kernel, no extra byte of HBM traffic.
__global__ void bn_scale_affine_kernel(int M, const float *__restrict__ mean, const float *__restrict__ var, float scale_factor, float eps,
                                       const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ a,
                                       float *__restrict__ b) {
  const int oc = blockIdx.x * blockDim.x + threadIdx.x;
  if (oc >= M) return;
  // BatchNormLayer::Forward with use_global_stats (src/caffe/layers/batch_norm_layer.cpp:98-106, 139-152):
  // mean = blobs[0] * sf, variance = blobs[1] * sf, top = (x - mean) / sqrt(variance + eps); then ScaleLayer: * gamma + beta
  const float m = mean[oc] * scale_factor, v = var[oc] * scale_factor;
  const float inv = 1.0f / sqrtf(v + eps);
  const float g = gamma ? gamma[oc] : 1.0f;
  a[oc] = g * inv;
  b[oc] = (beta ? beta[oc] : 0.0f) - m * inv * g;
}

__global__ void fold_affine_kernel(long total, long row, int M, const float *__restrict__ w, const float *__restrict__ a,
                                   const float *__restrict__ b, const float *__restrict__ bias_in, float *__restrict__ w_out,
                                   float *__restrict__ bias_out) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < total) w_out[e] = w[e] * __ldg(a + e / row);
  if (e < M && bias_out) bias_out[e] = (bias_in ? bias_in[e] : 0.0f) * a[e] + (b ? b[e] : 0.0f);
}

extern "C" int escort_bn_scale_to_affine(int num_output, const float *bn_mean, const float *bn_var, float bn_scale_factor_blob, float eps,
                                         const float *scale_gamma, const float *scale_beta, float *a_out, float *b_out,
                                         escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(num_output > 0 && bn_mean && bn_var && a_out && b_out, "escort_bn_scale_to_affine: bad arguments");
  const float sf = bn_scale_factor_blob == 0.f ? 0.f : 1.f / bn_scale_factor_blob;  // batch_norm_layer.cpp:100-101
  bn_scale_affine_kernel<<<(num_output + 127) / 128, 128, 0, stream>>>(num_output, bn_mean, bn_var, sf, eps, scale_gamma, scale_beta, a_out,
                                                                        b_out);
  ESCORT_LAUNCH_CHECK();
  return 0;
}

extern "C" int escort_plan_fold_affine(escort_plan *p, const float *weights_dense, const float *a, const float *b, const float *bias_in,
                                       float *weights_folded, float *bias_out, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(p && weights_dense && a && weights_folded, "escort_plan_fold_affine: bad arguments");
  const escort_geom &g = p->g;
  const long row = (long)(g.channels / g.group) * g.kernel_h * g.kernel_w, total = row * g.num_output;
  const long threads = std::max<long>(total, g.num_output);
  fold_affine_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(total, row, g.num_output, weights_dense, a, b, bias_in, weights_folded,
                                                                           bias_out);
  ESCORT_LAUNCH_CHECK();
  // an inference-time refresh: do not build the backward sub-plan for it
  const int tried = p->bwd_tried;
  if (!p->bwd) p->bwd_tried = 1;
  const int rc = escort_refresh_values(p, weights_folded, nullptr, stream_);
  p->bwd_tried = tried;
  return rc;
}

// the same fold for a layer the reference keeps dense (f1): W'[oc][...] = W[oc][...] * a[oc], bias' = bias * a + b
extern "C" int escort_dense_fold_affine(int num_output, long row, const float *weights, const float *a, const float *b, const float *bias_in,
                                        float *weights_folded, float *bias_out, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(num_output > 0 && row > 0 && weights && a && weights_folded, "escort_dense_fold_affine: bad arguments");
  const long total = row * num_output, threads = std::max<long>(total, num_output);
  fold_affine_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(total, row, num_output, weights, a, b, bias_in, weights_folded, bias_out);
  ESCORT_LAUNCH_CHECK();
  return 0;
}
