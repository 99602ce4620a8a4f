// sconv_tile.cu -- the register-blocked "tile interpreter" forward kernel of the Escort direct sparse
// convolution, and the plan-time compiler that turns the layer's CSR into its byte-code.
//
// Design (DESIGN.md has the long version):
//  * A CTA owns (image group, row band of output patches, WO blocks of OT output channels).  Its 8 compute
//    warps are WP pixel-warps x WO channel-block warps; every lane owns one TY x TX output patch and keeps
//    OT x TY x TX accumulators in registers for the whole reduction over input channels.
//  * Input channels stream through shared memory in chunks of CI channels (double buffered; two loader warps
//    copy chunk c+1 from the unpadded NCHW tensor, writing only the interior so the zero halo written once at
//    kernel start stays valid -- the reference's separate padded-copy pass, math_functions.cu:729-766, is gone).
//  * Per (channel block, chunk) the pruned weights are a byte-code segment: LOAD(channel plane) pulls the lane's
//    (TY-1)*S+KH x (TX-1)*S+KW input patch into registers with 128-bit shared loads, then one record per nonzero
//    {weight, handler} runs TY*TX FFMAs on fixed registers via an indirect branch (generated PTX,
//    tools/gen_interp.py).  One shared-memory patch load is reused by every nonzero of that input channel in
//    the OT-channel block: ~OT*KH*KW*density nonzeros x TY*TX FFMAs per load instead of 1 load per FFMA
//    (reference sconv_shm, math_functions.cu:283-310).
//  * Output rows are dealt to channel blocks in nnz-sorted snake order so blocks are nnz-balanced.
//  * Bias and ReLU are fused into the register epilogue.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>

#include "common.cuh"
#include "generated/variant_list.inc"
#define ESCORT_TILE_HOST_ONLY
template <int OT, int TY, int TX, int KH, int KW, int S> struct Interp;  // device side lives in tile_variant.cu
#include "tile_kernel.cuh"

namespace escort {

// ------------------------------------------------------------------------------------------------------
// variant table
// ------------------------------------------------------------------------------------------------------
struct VariantDesc {
  int OT, TY, TX, KH, KW, S, NCW, NLW;
  const char *name;
  const void *kernel;
};

// one translation unit per variant (tile_variant.cu compiled with -DESCORT_VARIANT_ID=k) exports these
#define ESCORT_VARIANT_DECL(ID, OT, TY, TX, KH, KW, S, NCW, NLW) \
  const void *tile_variant_kernel_##ID();                         \
  const char *tile_variant_name_##ID();
ESCORT_VARIANT_LIST(ESCORT_VARIANT_DECL)
static constexpr int kNumVariants = ESCORT_NUM_VARIANTS;
static const VariantDesc *variants() {
  static VariantDesc tab[kNumVariants];
  static bool init = false;
  if (!init) {
#define ESCORT_VARIANT_FILL(ID, OT, TY, TX, KH, KW, S, NCW, NLW) \
  tab[ID] = {OT, TY, TX, KH, KW, S, NCW, NLW, tile_variant_name_##ID(), tile_variant_kernel_##ID()};
    ESCORT_VARIANT_LIST(ESCORT_VARIANT_FILL)
    init = true;
  }
  return tab;
}
#define kVariants (variants())

template <typename T>
static int upload_vec(T **dptr, const std::vector<T> &h, cudaStream_t stream) {
  *dptr = nullptr;
  ESCORT_CUDA(cudaMalloc((void **)dptr, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) ESCORT_CUDA(cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
  return 0;
}

// number of shared-memory wavefronts one warp-wide 128-bit load costs for the given byte addresses
static int lds128_wavefronts(const std::vector<unsigned> &addr) {
  int total = 0;
  for (int q = 0; q < 4; ++q) {  // quarter warps
    int cnt[8] = {0};
    int worst = 0;
    for (int l = 0; l < 8; ++l) {
      const int lane = q * 8 + l;
      if (lane >= (int)addr.size()) break;
      // distinct 16-byte chunk per bank group; identical addresses broadcast
      bool dup = false;
      for (int m = 0; m < l; ++m)
        if (addr[q * 8 + m] == addr[lane]) dup = true;
      if (dup) continue;
      worst = std::max(worst, ++cnt[(addr[lane] >> 4) & 7]);
    }
    total += std::max(worst, 1);
  }
  return total;
}

void tile_plan_free(TilePlan *tp) {
  if (!tp) return;
  cudaFree(tp->d_lanes);
  cudaFree(tp->d_oc_list);
  cudaFree(tp->d_prog);
  cudaFree(tp->d_seg);
  cudaFree(tp->d_dst_off);
  cudaFree(tp->d_prog_pos);
  delete tp;
}

const char *tile_kernel_name(const TilePlan *tp) { return tp->name; }

// Pick the default variant for a geometry (auto mode); returns -1 if the tile kernel does not apply.
static int choose_variant(const escort_geom &g, double density) {
  auto find = [&](int OT, int TY, int TX) {
    for (int i = 0; i < kNumVariants; ++i) {
      const VariantDesc &v = kVariants[i];
      if (v.OT == OT && v.TY == TY && v.TX == TX && v.KH == g.kernel_h && v.KW == g.kernel_w && v.S == g.stride_h) return i;
    }
    return -1;
  };
  (void)density;
  int v = find(8, 4, 4);
  if (v < 0) v = find(8, 2, 4);
  if (v < 0) v = find(4, 4, 4);
  return v;
}

int tile_plan_build(escort_plan *plan, int variant, cudaStream_t stream) {
  plan->tile = nullptr;
  const escort_geom &g = plan->g;
  if (g.dilation_h != 1 || g.dilation_w != 1 || g.stride_h != g.stride_w) return 0;
  if (g.width > 1024 || plan->nnz == 0) return 0;
  const int Cg = g.channels / g.group, Mg = g.num_output / g.group;
  const double density = (double)plan->nnz / ((double)g.num_output * Cg * g.kernel_h * g.kernel_w);
  int vidx;
  if (variant > 0) {
    vidx = variant - 1;
    if (vidx >= kNumVariants) return 0;
    const VariantDesc &v = kVariants[vidx];
    if (v.KH != g.kernel_h || v.KW != g.kernel_w || v.S != g.stride_h) return 0;
  } else {
    vidx = choose_variant(g, density);
    if (vidx < 0) return 0;
  }
  const VariantDesc &V = kVariants[vidx];
  const int OT = V.OT, TY = V.TY, TX = V.TX, KH = V.KH, KW = V.KW, S = V.S;
  const int kComputeWarps = V.NCW;
  const int Ho = plan->Ho, Wo = plan->Wo;
  const int PY = ceil_div(Ho, TY), PX = ceil_div(Wo, TX);
  const int PC = (TX - 1) * S + KW, XW = ceil_div(PC, 4) * 4;
  const int Pmin = ceil_div((PX - 1) * TX * S + XW, 4) * 4;   // every lane's vector over-read stays inside the row
  int dev = 0, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (max_smem <= 0) max_smem = 227 * 1024;
  const size_t smem_budget = (size_t)max_smem - 1024;

  // ---- choose (WP, G, BR): maximise lane utilisation, then prefer more channel-block warps ----
  struct Choice { int WP, G, BR, P, R, CI; double score; } best = {0, 0, 0, 0, 0, 0, -1.0};
  for (int WP = 1; WP <= kComputeWarps; ++WP) {
    if (kComputeWarps % WP) continue;
    const int lanes = WP * 32;
    for (int nb = 1; nb <= PY; ++nb) {
      const int BR = ceil_div(PY, nb);
      if (nb > 1 && ceil_div(PY, nb - 1) == BR) continue;  // same BR as a smaller band count
      const int per_img = BR * PX;
      if (per_img > lanes) continue;
      const int G = std::min(lanes / per_img, 32);
      const int R = (BR * TY - 1) * S + KH;
      // pitch: smallest multiple of 4 >= Pmin with the fewest LDS.128 bank conflicts over the lane layout
      int bestP = Pmin, bestW = 1 << 30;
      for (int P = Pmin; P <= Pmin + 32; P += 4) {
        int wsum = 0;
        for (int w = 0; w < WP; ++w) {
          std::vector<unsigned> addr;
          for (int l = 0; l < 32; ++l) {
            const int slot = w * 32 + l;
            if (slot >= G * per_img) break;
            const int gi = slot / per_img, rem = slot % per_img;
            const int pyb = rem / PX, px = rem % PX;
            addr.push_back((unsigned)(((gi * R + pyb * TY * S) * P + px * TX * S) * 4));
          }
          if (!addr.empty()) wsum += lds128_wavefronts(addr);
        }
        // prefer fewer conflicts, then smaller pitch (4% tolerance per 4 floats)
        const int cost = wsum * 64 + (P - Pmin);
        if (cost < bestW) { bestW = cost; bestP = P; }
      }
      const int P = bestP;
      const size_t plane = (size_t)R * P * 4 * G;
      const size_t tab = (size_t)R * g.width * 2 + 16;
      if (2 * plane + tab > smem_budget) continue;
      if (R * P > 65535) continue;  // dst_off is 16 bit
      int CI = (int)std::min<size_t>((smem_budget - tab) / (2 * plane), (size_t)Cg);
      if (CI < 1) continue;
      CI = std::min(CI, 64);
      const int WO = kComputeWarps / WP;
      const int nblk = ceil_div(Mg, OT);
      const double util = (double)(G * per_img) / lanes * ((double)Ho * Wo / ((double)PY * TY * PX * TX));
      const double band_waste = (double)PY / (nb * BR);          // last band partially empty
      const double wo_eff = (double)std::min(WO, nblk) / WO;      // idle channel-block warps
      const double halo = (double)(BR * TY * S) / R;              // input re-read across bands
      double score = util * band_waste * wo_eff * (0.85 + 0.15 * halo);
      score *= (CI >= 4 ? 1.0 : 0.8);
      score *= 1.0 + 0.02 * std::log2((double)WO);                // tie-break: more input reuse across channels
      if (score > best.score) best = {WP, G, BR, P, R, CI, score};
    }
  }
  if (best.score < 0) return 0;
  const int WP = best.WP, G = best.G, BR = best.BR, P = best.P, R = best.R;
  const int WO = kComputeWarps / WP;
  const int nbands = ceil_div(PY, BR);
  int CI = best.CI;
  const int nchunks = ceil_div(Cg, CI);
  CI = ceil_div(Cg, nchunks);  // even out the chunks
  const int nblk = ceil_div(Mg, OT);
  const int ogroups = ceil_div(nblk, WO);
  const int per_img = BR * PX;
  const int nslots = G * per_img;

  TilePlan *tp = new TilePlan();
  memset((void *)tp, 0, sizeof(*tp));
  tp->vidx = vidx;
  tp->name = V.name;
  tp->OT = OT; tp->TY = TY; tp->TX = TX; tp->KH = KH; tp->KW = KW; tp->S = S;
  TileParams &pr = tp->prm;
  pr.C = g.channels; pr.H = g.height; pr.W = g.width; pr.M = g.num_output; pr.Ho = Ho; pr.Wo = Wo;
  pr.pad_h = g.pad_h; pr.pad_w = g.pad_w; pr.Cg = Cg; pr.Mg = Mg;
  pr.G = G; pr.BR = BR; pr.PX = PX; pr.PY = PY; pr.nbands = nbands; pr.WP = WP; pr.WO = WO; pr.R = R; pr.P = P;
  pr.CI = CI; pr.nchunks = nchunks; pr.nblk = nblk; pr.ogroups = ogroups; pr.nslots = nslots;
  pr.chunk_floats = CI * G * R * P;
  tp->smem_bytes = (size_t)2 * pr.chunk_floats * 4 + (size_t)R * g.width * 2 + 16;

  // ---- lane table ----
  std::vector<int4> lanes(WP * 32, make_int4(0, 0, BR /*invalid*/, 0));
  for (int slot = 0; slot < nslots; ++slot) {
    const int gi = slot / per_img, rem = slot % per_img;
    const int pyb = rem / PX, px = rem % PX;
    lanes[slot] = make_int4(((gi * R + pyb * TY * S) * P + px * TX * S) * 4, gi, pyb, px);
  }
  // ---- loader scatter table: element e (row-major over the band's W-wide input rows) -> smem float offset ----
  std::vector<unsigned short> dst_off((size_t)R * g.width);
  for (int r = 0; r < R; ++r)
    for (int x = 0; x < g.width; ++x) dst_off[(size_t)r * g.width + x] = (unsigned short)(r * P + x + g.pad_w);

  // ---- nnz-balanced channel blocks: rows sorted by nnz (desc), dealt in snake order ----
  const std::vector<Nz> &nz = *plan->host_nz;
  std::vector<int> row_nnz(g.num_output, 0), row_start(g.num_output + 1, 0);
  for (const Nz &z : nz) row_nnz[z.oc]++;
  for (int i = 0; i < g.num_output; ++i) row_start[i + 1] = row_start[i] + row_nnz[i];
  std::vector<int> oc_list((size_t)g.group * nblk * OT, -1);
  std::vector<int> oc_block(g.num_output), oc_slot(g.num_output);
  for (int gi = 0; gi < g.group; ++gi) {
    std::vector<int> rows(Mg);
    std::iota(rows.begin(), rows.end(), gi * Mg);
    std::stable_sort(rows.begin(), rows.end(), [&](int a, int b) { return row_nnz[a] > row_nnz[b]; });
    std::vector<int> fill(nblk, 0);
    for (int i = 0; i < Mg; ++i) {
      const int round = i / nblk, pos = i % nblk;
      const int b = (round & 1) ? (nblk - 1 - pos) : pos;
      const int oc = rows[i];
      oc_block[oc] = b;
      oc_slot[oc] = fill[b];
      oc_list[((size_t)gi * nblk + b) * OT + fill[b]] = oc;
      fill[b]++;
    }
  }
  // ---- compile byte-code: per (group, block, chunk) ----
  const int NC = OT * KH * KW;
  struct Rec { int ic, kh, kw, o; float val; int src; };
  std::vector<std::vector<Rec>> buckets((size_t)g.group * nblk * nchunks);
  for (size_t j = 0; j < nz.size(); ++j) {
    const Nz &z = nz[j];
    const int gi = z.oc / Mg;
    const int icl = z.ic - gi * Cg;
    const int c = icl / CI;
    buckets[((size_t)gi * nblk + oc_block[z.oc]) * nchunks + c].push_back({icl, z.kh, z.kw, oc_slot[z.oc], z.val, (int)j});
  }
  std::vector<uint2> prog;
  prog.reserve(nz.size() + buckets.size() * (CI + 2) + 4);
  std::vector<int> seg(buckets.size());
  std::vector<int> prog_pos(nz.size(), -1);
  const unsigned plane_bytes = (unsigned)G * R * P * 4;
  for (size_t b = 0; b < buckets.size(); ++b) {
    std::vector<Rec> &v = buckets[b];
    const int c = (int)(b % nchunks);
    std::sort(v.begin(), v.end(), [](const Rec &a, const Rec &bb) {
      if (a.ic != bb.ic) return a.ic < bb.ic;
      if (a.kh != bb.kh) return a.kh < bb.kh;
      if (a.kw != bb.kw) return a.kw < bb.kw;
      return a.o < bb.o;
    });
    seg[b] = (int)prog.size();
    int cur_ic = -1;
    for (const Rec &r : v) {
      if (r.ic != cur_ic) {
        cur_ic = r.ic;
        prog.push_back(make_uint2((unsigned)(r.ic - c * CI) * plane_bytes, (unsigned)NC));
      }
      prog_pos[r.src] = (int)prog.size();
      prog.push_back(make_uint2(__builtin_bit_cast(unsigned, r.val), (unsigned)((r.o * KH + r.kh) * KW + r.kw)));
    }
    prog.push_back(make_uint2(0u, (unsigned)(NC + 1)));
  }
  for (int i = 0; i < 4; ++i) prog.push_back(make_uint2(0u, (unsigned)(NC + 1)));  // prefetch slack
  tp->nrecords = prog.size();

  int rc = 0;
  if ((rc = upload_vec(&tp->d_lanes, lanes, stream)) || (rc = upload_vec(&tp->d_oc_list, oc_list, stream)) ||
      (rc = upload_vec(&tp->d_prog, prog, stream)) || (rc = upload_vec(&tp->d_seg, seg, stream)) ||
      (rc = upload_vec(&tp->d_dst_off, dst_off, stream)) || (rc = upload_vec(&tp->d_prog_pos, prog_pos, stream))) {
    tile_plan_free(tp);
    return rc;
  }
  cudaError_t e = cudaStreamSynchronize(stream);  // host vectors die here
  if (e != cudaSuccess) {
    tile_plan_free(tp);
    return cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
  }
  pr.lanes = tp->d_lanes; pr.oc_list = tp->d_oc_list; pr.prog = tp->d_prog; pr.seg = tp->d_seg; pr.dst_off = tp->d_dst_off;
  e = cudaFuncSetAttribute(V.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp->smem_bytes);
  if (e != cudaSuccess) {
    tile_plan_free(tp);
    return cuda_fail(e, "cudaFuncSetAttribute(smem)", __FILE__, __LINE__);
  }
  plan->tile = tp;
  return 0;
}

int tile_forward(escort_plan *plan, int num, const float *bottom, const float *bias, int fuse_relu, float *top,
                 cudaStream_t stream) {
  TilePlan *tp = plan->tile;
  TileParams prm = tp->prm;
  prm.n_igroups = ceil_div(num, prm.G);
  const unsigned grid = (unsigned)((size_t)prm.n_igroups * prm.nbands * plan->g.group * prm.ogroups);
  void *args[] = {(void *)&prm, (void *)&num, (void *)&bottom, (void *)&bias, (void *)&fuse_relu, (void *)&top};
  const VariantDesc &V = kVariants[tp->vidx];
  ESCORT_CUDA(cudaLaunchKernel(V.kernel, dim3(grid), dim3((V.NCW + V.NLW) * 32), args, tp->smem_bytes, stream));
  return 0;
}

__global__ void tile_refresh_kernel(long nnz, const float *__restrict__ w_dense, const int *__restrict__ dense_idx,
                                    const int *__restrict__ prog_pos, uint2 *__restrict__ prog) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  prog[prog_pos[j]].x = __float_as_uint(__ldg(w_dense + dense_idx[j]));
}

int tile_refresh(escort_plan *plan, const float *weights_dense, cudaStream_t stream) {
  TilePlan *tp = plan->tile;
  const unsigned blocks = (unsigned)((plan->nnz + 255) / 256);
  tile_refresh_kernel<<<blocks, 256, 0, stream>>>(plan->nnz, weights_dense, plan->d_dense_idx, tp->d_prog_pos, tp->d_prog);
  ESCORT_LAUNCH_CHECK();
  return 0;
}

// introspection for tests/bench: tiling summary as a string
extern "C" ESCORT_API int escort_plan_describe(const escort_plan *plan, char *buf, int buflen) {
  if (!plan || !buf || buflen <= 0) return ESCORT_EINVAL;
  if (!plan->tile) {
    snprintf(buf, buflen, "generic");
    return 0;
  }
  const TilePlan *tp = plan->tile;
  const TileParams &p = tp->prm;
  snprintf(buf, buflen, "%s G=%d BR=%d nbands=%d WP=%d WO=%d R=%d P=%d CI=%d nchunks=%d nblk=%d ogroups=%d nslots=%d smem=%zu records=%zu",
           tp->name, p.G, p.BR, p.nbands, p.WP, p.WO, p.R, p.P, p.CI, p.nchunks, p.nblk, p.ogroups, p.nslots,
           tp->smem_bytes, tp->nrecords);
  return 0;
}

}  // namespace escort
