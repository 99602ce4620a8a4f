// sconv_tile.cu -- the register-blocked "tile interpreter" forward kernel of the Escort direct sparse
// convolution, and the plan-time compiler that turns the layer's CSR into its byte-code.
//
// Design (DESIGN.md has the long version):
//  * A CTA owns (image group, row band of output patches, WO blocks of OT output channels).  Its 8 compute
//    warps are WP pixel-warps x WO channel-block warps; every lane owns one TY x TX output patch and keeps
//    OT x TY x TX accumulators in registers for the whole reduction over input channels.
//  * Persistent CTAs (one per SM) walk the work units.  Input channels stream through shared memory in chunks of
//    CI channels through an NS-stage mbarrier ring: two loader warps copy chunk c+1.. from the unpadded NCHW tensor
//    (writing only the interior, the zero halo is written once) together with the chunk's byte-code, and run
//    ahead across unit boundaries -- the reference's separate padded-copy pass, math_functions.cu:729-766, is gone.
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
template <int OT, int TY, int TX, int KH, int KW, int S, int PAIR> struct Interp;  // device side: tile_variant.cu
#include "tile_kernel.cuh"

namespace escort {

// ------------------------------------------------------------------------------------------------------
// variant table
// ------------------------------------------------------------------------------------------------------
struct VariantDesc {
  int OT, TY, TX, KH, KW, S, PAIR, NCW, NLW;
  const char *name;
  const void *kernel;
};

// one translation unit per variant (tile_variant.cu compiled with -DESCORT_VARIANT_ID=k) exports these
#define ESCORT_VARIANT_DECL(ID, OT, TY, TX, KH, KW, S, PAIR, NCW, NLW) \
  const void *tile_variant_kernel_##ID();                         \
  const char *tile_variant_name_##ID();
ESCORT_VARIANT_LIST(ESCORT_VARIANT_DECL)
static constexpr int kNumVariants = ESCORT_NUM_VARIANTS;
static const VariantDesc *variants() {
  static VariantDesc tab[kNumVariants];
  static bool init = false;
  if (!init) {
#define ESCORT_VARIANT_FILL(ID, OT, TY, TX, KH, KW, S, PAIR, NCW, NLW) \
  tab[ID] = {OT, TY, TX, KH, KW, S, PAIR, NCW, NLW, tile_variant_name_##ID(), tile_variant_kernel_##ID()};
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
  cudaFree(tp->d_rtab);
  cudaFree(tp->d_dst_off);
  cudaFree(tp->d_prog_pos);
  delete tp;
}

const char *tile_kernel_name(const TilePlan *tp) { return tp->name; }
int tile_num_variants() { return kNumVariants; }
bool tile_variant_applies(const escort_plan *plan, int variant) {
  if (variant < 1 || variant > kNumVariants) return false;
  const VariantDesc &v = kVariants[variant - 1];
  const escort_geom &g = plan->g;
  return v.KH == g.kernel_h && v.KW == g.kernel_w && v.S == g.stride_h && g.stride_h == g.stride_w &&
         g.dilation_h == 1 && g.dilation_w == 1;
}

// Pick the default variant for a geometry (auto mode); returns -1 if the tile kernel does not apply.
static int choose_variant(const escort_geom &g, double density) {
  int best = -1;
  for (int i = 0; i < kNumVariants; ++i) {
    const VariantDesc &v = kVariants[i];
    if (v.KH != g.kernel_h || v.KW != g.kernel_w || v.S != g.stride_h) continue;
    if (best < 0) best = i;  // the generator lists the preferred shape of each kernel size first
  }
  (void)density;
  return best;
}

namespace {
struct Layout {
  int WP, G, BR, P, R, skew, order;
  double score;
};
struct Slot { int gs, pyb, px; };

// enumerate the lane slots of one CTA in a given order; order 0: (image slot, patch row, patch col),
// order q > 0: patch cols in groups of q, image slots interleaved between groups (spreads 128-bit loads over banks)
std::vector<Slot> enumerate_slots(int GP, int BR, int PX, int order) {
  std::vector<Slot> v;
  v.reserve((size_t)GP * BR * PX);
  if (order <= 0) {
    for (int gs = 0; gs < GP; ++gs)
      for (int py = 0; py < BR; ++py)
        for (int px = 0; px < PX; ++px) v.push_back({gs, py, px});
  } else {
    const int q = order;
    for (int py = 0; py < BR; ++py)
      for (int pxh = 0; pxh < PX; pxh += q)
        for (int gs = 0; gs < GP; ++gs)
          for (int px = pxh; px < std::min(PX, pxh + q); ++px) v.push_back({gs, py, px});
  }
  return v;
}
}  // namespace

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
  const int OT = V.OT, TY = V.TY, TX = V.TX, KH = V.KH, KW = V.KW, S = V.S, PAIR = V.PAIR;
  const int NCW = V.NCW;
  const int Ho = plan->Ho, Wo = plan->Wo;
  const int PY = ceil_div(Ho, TY), PX = ceil_div(Wo, TX);
  const int per_vec = 4 / PAIR;
  const int PC = (TX - 1) * S + KW, XW = ceil_div(PC, per_vec) * per_vec;
  // every lane's vector over-read stays inside its row; the pitch keeps 16-byte alignment of every row start
  const int Pmin = ceil_div(std::max((PX - 1) * TX * S + XW, g.width + g.pad_w), per_vec) * per_vec;
  int dev = 0, max_smem = 0, num_sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (max_smem <= 0) max_smem = 227 * 1024;
  if (num_sms <= 0) num_sms = 148;
  const int tab_bytes = ceil_div(std::max(1, ((TY * PY - 1) * S + KH) * g.width) * 2, 128) * 128;
  const long smem_budget = (long)max_smem - kBarBytes - 256;
  const int nblk = ceil_div(Mg, OT);

  // ---- choose (WP, BR, G, pitch, skew, lane order): lane utilisation first, then bank conflicts / reuse ----
  Layout best = {0, 0, 0, 0, 0, 0, 0, -1.0};
  for (int WP = 1; WP <= NCW; ++WP) {
    if (NCW % WP) continue;
    const int lanes = WP * 32;
    const int WO = NCW / WP;
    for (int nb = 1; nb <= PY; ++nb) {
      const int BR = ceil_div(PY, nb);
      if (nb > 1 && ceil_div(PY, nb - 1) == BR) continue;  // same BR as a smaller band count
      const int per_slot = BR * PX;
      if (per_slot > lanes) continue;
      const int GP = std::min(lanes / per_slot, 16);
      const int G = GP * PAIR;
      const int R = (BR * TY - 1) * S + KH;
      if ((long)R * g.width * 2 > tab_bytes) continue;
      const int nslots = GP * per_slot;
      const double util = (double)nslots / lanes * ((double)Ho * Wo / ((double)PY * TY * PX * TX));
      const double band_waste = (double)PY / (nb * BR);
      const double wo_eff = (double)std::min(WO, nblk) / WO;
      const double halo = (double)(BR * TY * S) / R;
      double base_score = util * band_waste * wo_eff * (0.85 + 0.15 * halo) * (1.0 + 0.02 * std::log2((double)WO));
      if (base_score < best.score * 0.98) continue;
      // pitch / skew / lane order with the fewest LDS.128 wavefronts
      int bP = Pmin, bskew = 0, border = 0, bcost = 1 << 30;
      const int orders[4] = {0, 4, 2, 1};
      for (int oi = 0; oi < 4; ++oi) {
        const std::vector<Slot> slots = enumerate_slots(GP, BR, PX, orders[oi]);
        for (int P = Pmin; P <= Pmin + 8 * per_vec; P += per_vec) {
          for (int skew = 0; skew <= 4; skew += 4) {
            const int plane_f = R * P * PAIR + skew;
            int wsum = 0;
            for (int w = 0; w < WP; ++w) {
              std::vector<unsigned> addr;
              for (int l = 0; l < 32 && w * 32 + l < nslots; ++l) {
                const Slot &sl = slots[w * 32 + l];
                addr.push_back((unsigned)((sl.gs * plane_f + (sl.pyb * TY * S * P + sl.px * TX * S) * PAIR) * 4));
              }
              if (!addr.empty()) wsum += lds128_wavefronts(addr);
            }
            const int cost = wsum * 256 + (P - Pmin) * 8 + skew + oi;
            if (cost < bcost) { bcost = cost; bP = P; bskew = skew; border = orders[oi]; }
          }
        }
      }
      const long plane_bytes = ((long)R * bP * PAIR + bskew) * 4 * GP;
      if (3 * plane_bytes > smem_budget - tab_bytes) continue;  // at least 1 channel x 3 stages (program space extra)
      const int ideal = WP * 4;  // 4 wavefronts per LDS.128 per warp is conflict free
      const double conflict = (double)ideal / std::max(ideal, bcost / 256);
      const double score = base_score * (0.6 + 0.4 * conflict);
      if (score > best.score) best = {WP, G, BR, bP, R, bskew, border, score};
    }
  }
  if (best.score < 0) return 0;
  const int WP = best.WP, G = best.G, GP = G / PAIR, BR = best.BR, P = best.P, R = best.R;
  const int WO = NCW / WP;
  const int nbands = ceil_div(PY, BR);
  const int ogroups = ceil_div(nblk, WO);
  const int per_slot = BR * PX;
  const int nslots = GP * per_slot;
  const int plane_f = R * P * PAIR + best.skew;
  const int hdr_bytes = ceil_div(WO * 4, 16) * 16;

  // ---- nnz-balanced channel blocks: rows sorted by nnz (desc), dealt in snake order ----
  const std::vector<Nz> &nz = *plan->host_nz;
  std::vector<int> row_nnz(g.num_output, 0);
  for (const Nz &z : nz) row_nnz[z.oc]++;
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

  // ---- chunking + byte-code: pick CI so that >= 3 stages (input + program) fit; retry smaller on overflow ----
  const int NC = OT * KH * KW;
  struct Rec { int ic, kh, kw, o; float val; int src; };
  const long plane_bytes = (long)plane_f * 4 * GP;
  int CI = (int)std::min<long>((smem_budget - tab_bytes) / (3 * plane_bytes), (long)Cg);
  CI = std::max(1, std::min(CI, 32));
  std::vector<uint4> prog;
  std::vector<int2> rtab;
  std::vector<int> prog_pos;
  int nchunks = 0, NS = 0, in_bytes = 0, stage_bytes = 0, max_region16 = 0;
  for (int attempt = 0; attempt < 8; ++attempt) {
    nchunks = ceil_div(Cg, CI);
    CI = ceil_div(Cg, nchunks);  // even out the chunks
    in_bytes = (int)(plane_bytes * CI);
    std::vector<std::vector<Rec>> buckets((size_t)g.group * nblk * nchunks);
    for (size_t j = 0; j < nz.size(); ++j) {
      const Nz &z = nz[j];
      const int gi = z.oc / Mg, icl = z.ic - gi * Cg, c = icl / CI;
      buckets[((size_t)gi * nblk + oc_block[z.oc]) * nchunks + c].push_back({icl, z.kh, z.kw, oc_slot[z.oc], z.val, (int)j});
    }
    std::vector<unsigned> words;  // u32 stream, regions padded to 16 bytes
    rtab.assign((size_t)g.group * ogroups * nchunks, make_int2(0, 0));
    prog_pos.assign(nz.size(), -1);
    max_region16 = 0;
    for (int gi = 0; gi < g.group; ++gi)
      for (int og = 0; og < ogroups; ++og)
        for (int c = 0; c < nchunks; ++c) {
          const size_t region_start = words.size();  // multiple of 4 words
          words.resize(region_start + hdr_bytes / 4, 0u);
          for (int ow = 0; ow < WO; ++ow) {
            const int blk = og * WO + ow;
            words[region_start + ow] = (unsigned)((words.size() - region_start) * 4);  // byte offset of the segment
            if (blk < nblk) {
              std::vector<Rec> &v = buckets[((size_t)gi * nblk + blk) * nchunks + c];
              std::sort(v.begin(), v.end(), [](const Rec &a, const Rec &b) {
                if (a.ic != b.ic) return a.ic < b.ic;
                if (a.kh != b.kh) return a.kh < b.kh;
                if (a.kw != b.kw) return a.kw < b.kw;
                return a.o < b.o;
              });
              int cur_ic = -1;
              for (const Rec &r : v) {
                if (r.ic != cur_ic) {
                  cur_ic = r.ic;
                  words.push_back((unsigned)((r.ic - c * CI) * (long)GP * plane_f * 4));
                  words.push_back((unsigned)NC);
                }
                prog_pos[r.src] = (int)(words.size() / 2);
                words.push_back(__builtin_bit_cast(unsigned, r.val));
                words.push_back((unsigned)((r.o * KH + r.kh) * KW + r.kw));
              }
            }
            words.push_back(0u);
            words.push_back((unsigned)(NC + 1));  // end of segment
          }
          for (int i = 0; i < 2; ++i) {  // prefetch slack: the interpreter reads two records past the end
            words.push_back(0u);
            words.push_back((unsigned)(NC + 1));
          }
          while (words.size() % 4) words.push_back(0u);
          const int len16 = (int)((words.size() - region_start) / 4);
          rtab[((size_t)gi * ogroups + og) * nchunks + c] = make_int2((int)(region_start / 4), len16);
          max_region16 = std::max(max_region16, len16);
        }
    const int prog_bytes = ceil_div(max_region16 * 16, 128) * 128;
    stage_bytes = ceil_div(in_bytes, 128) * 128 + prog_bytes;
    NS = (int)std::min<long>((smem_budget - tab_bytes) / stage_bytes, 4L);
    if (NS >= 3 || (NS >= 2 && CI == 1)) {
      prog.resize(words.size() / 4);
      memcpy(prog.data(), words.data(), words.size() * 4);
      break;
    }
    if (CI == 1) return 0;  // does not fit at all
    CI = std::max(1, CI * 3 / 4);
    NS = 0;
  }
  if (NS < 2) return 0;
  in_bytes = ceil_div(in_bytes, 128) * 128;

  TilePlan *tp = new TilePlan();
  memset((void *)tp, 0, sizeof(*tp));
  tp->vidx = vidx;
  tp->name = V.name;
  tp->OT = OT; tp->TY = TY; tp->TX = TX; tp->KH = KH; tp->KW = KW; tp->S = S; tp->PAIR = PAIR;
  tp->num_sms = num_sms;
  TileParams &pr = tp->prm;
  pr.C = g.channels; pr.H = g.height; pr.W = g.width; pr.M = g.num_output; pr.Ho = Ho; pr.Wo = Wo;
  pr.pad_h = g.pad_h; pr.pad_w = g.pad_w; pr.Cg = Cg; pr.Mg = Mg; pr.ngroups = g.group;
  pr.G = G; pr.GP = GP; pr.BR = BR; pr.PX = PX; pr.PY = PY; pr.nbands = nbands; pr.WP = WP; pr.WO = WO;
  pr.R = R; pr.P = P; pr.plane_f = plane_f; pr.CI = CI; pr.nchunks = nchunks; pr.nblk = nblk; pr.ogroups = ogroups;
  pr.nslots = nslots; pr.NS = NS; pr.stage0_off = kBarBytes + tab_bytes; pr.stage_bytes = stage_bytes;
  pr.in_bytes = in_bytes; pr.hdr_bytes = hdr_bytes;
  pr.n4 = (R * g.width + 6) / 4 + 1;
  pr.n4_magic = (unsigned)((0x100000000ull + pr.n4 - 1) / pr.n4);
  pr.ci_magic = (unsigned)((0x100000000ull + CI - 1) / CI);
  tp->smem_bytes = (size_t)pr.stage0_off + (size_t)NS * stage_bytes;
  tp->nrecords = prog.size() * 2;

  // ---- lane table ----
  std::vector<int4> lanes(WP * 32, make_int4(0, 0, 0, 0));
  {
    const std::vector<Slot> slots = enumerate_slots(GP, BR, PX, best.order);
    for (int i = 0; i < nslots; ++i) {
      const Slot &sl = slots[i];
      lanes[i] = make_int4((sl.gs * plane_f + (sl.pyb * TY * S * P + sl.px * TX * S) * PAIR) * 4, sl.gs, sl.pyb, sl.px);
    }
  }
  // ---- loader scatter table: element e (row-major over the band's W-wide input rows) -> position offset ----
  std::vector<unsigned short> dst_off((size_t)R * g.width);
  for (int r = 0; r < R; ++r)
    for (int x = 0; x < g.width; ++x) dst_off[(size_t)r * g.width + x] = (unsigned short)(r * P + x + g.pad_w);
  if ((long)R * P > 65535) { delete tp; return 0; }

  int rc = 0;
  if ((rc = upload_vec(&tp->d_lanes, lanes, stream)) || (rc = upload_vec(&tp->d_oc_list, oc_list, stream)) ||
      (rc = upload_vec(&tp->d_prog, prog, stream)) || (rc = upload_vec(&tp->d_rtab, rtab, stream)) ||
      (rc = upload_vec(&tp->d_dst_off, dst_off, stream)) || (rc = upload_vec(&tp->d_prog_pos, prog_pos, stream))) {
    tile_plan_free(tp);
    return rc;
  }
  cudaError_t e = cudaStreamSynchronize(stream);  // host vectors die here
  if (e != cudaSuccess) {
    tile_plan_free(tp);
    return cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
  }
  pr.lanes = tp->d_lanes; pr.oc_list = tp->d_oc_list; pr.prog = tp->d_prog; pr.rtab = tp->d_rtab; pr.dst_off = tp->d_dst_off;
  // the attribute belongs to the kernel, not the plan: several plans share a variant, so always raise it to the
  // device's opt-in maximum instead of this plan's own size
  e = cudaFuncSetAttribute(V.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
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
  int nunits = (int)((size_t)prm.n_igroups * prm.nbands * prm.ngroups * prm.ogroups);
  const unsigned grid = (unsigned)std::min(nunits, tp->num_sms);
  const VariantDesc &V = kVariants[tp->vidx];
  void *args[] = {(void *)&prm, (void *)&num, (void *)&bottom, (void *)&bias, (void *)&fuse_relu, (void *)&top, (void *)&nunits};
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
  tile_refresh_kernel<<<blocks, 256, 0, stream>>>(plan->nnz, weights_dense, plan->d_dense_idx, tp->d_prog_pos,
                                                  reinterpret_cast<uint2 *>(tp->d_prog));
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
  snprintf(buf, buflen,
           "%s G=%d BR=%d nbands=%d WP=%d WO=%d R=%d P=%d plane_f=%d CI=%d nchunks=%d nblk=%d ogroups=%d nslots=%d NS=%d "
           "stage=%dB smem=%zu records=%zu",
           tp->name, p.G, p.BR, p.nbands, p.WP, p.WO, p.R, p.P, p.plane_f, p.CI, p.nchunks, p.nblk, p.ogroups, p.nslots,
           p.NS, p.stage_bytes, tp->smem_bytes, tp->nrecords);
  return 0;
}

}  // namespace escort
