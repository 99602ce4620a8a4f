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

#include <cuda.h>

#include "common.cuh"
#include "generated/variant_list.inc"
#define ESCORT_TILE_HOST_ONLY
template <int VID> struct Interp;  // device side: tile_variant.cu
#include "tile_kernel.cuh"

namespace escort {

// ------------------------------------------------------------------------------------------------------
// variant table
// ------------------------------------------------------------------------------------------------------
struct VariantDesc {
  int OT, TY, TX, KH, KW, S, PAIR, NCW, NLW, NTW, MODE, SHFL;
  const char *name;
  const void *kernel;
  const void *bench;
  const void *bwdw;
};

// one translation unit per variant (tile_variant.cu compiled with -DESCORT_VARIANT_ID=k) exports these
#define ESCORT_VARIANT_DECL(ID, OT, TY, TX, KH, KW, S, PAIR, NCW, NLW, NTW, MODE, SHFL) \
  const void *tile_variant_kernel_##ID();                         \
  const void *tile_variant_bench_##ID();                          \
  const void *tile_variant_bwdw_##ID();                           \
  const char *tile_variant_name_##ID();
ESCORT_VARIANT_LIST(ESCORT_VARIANT_DECL)
static constexpr int kNumVariants = ESCORT_NUM_VARIANTS;
static const VariantDesc *variants() {
  // C++11 magic static: initialised once, thread safe (Caffe runs one host thread per GPU)
  static const struct Table {
    VariantDesc tab[kNumVariants];
    Table() {
#define ESCORT_VARIANT_FILL(ID, OT, TY, TX, KH, KW, S, PAIR, NCW, NLW, NTW, MODE, SHFL) \
  tab[ID] = {OT, TY, TX, KH, KW, S, PAIR, NCW, NLW, NTW, MODE, SHFL, tile_variant_name_##ID(), tile_variant_kernel_##ID(), tile_variant_bench_##ID(), tile_variant_bwdw_##ID()};
      ESCORT_VARIANT_LIST(ESCORT_VARIANT_FILL)
    }
  } table;
  return table.tab;
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
  cudaFree(tp->d_prog_pos);
  cudaFree(tp->d_tapidx);
  cudaFree(tp->d_tap_dense);
  cudaFree(tp->d_tap_csr);
  delete tp;
}

const char *tile_kernel_name(const TilePlan *tp) { return tp->name; }
int tile_plan_variant(const TilePlan *tp) { return tp->vidx + 1; }
int tile_num_variants() { return kNumVariants + tmem_num_variants(); }
bool tile_variant_applies(const escort_plan *plan, int variant) {
  if (variant > kNumVariants) return tmem_variant_applies(plan, variant - kNumVariants - 1);
  if (variant < 1) return false;
  const VariantDesc &v = kVariants[variant - 1];
  const escort_geom &g = plan->g;
  return v.MODE < 5 && v.KH == g.kernel_h && v.KW == g.kernel_w && v.S == g.stride_h && g.stride_h == g.stride_w &&
         g.dilation_h == 1 && g.dilation_w == 1;
}
// Pick the default variant for a geometry (auto mode, no autotune); returns -1 if the tile kernel does not apply.
// The preference lists come from the B200 sweeps in profiles/ (escort_plan_autotune measures instead of guessing).
static int find_variant(const char *name) {
  for (int i = 0; i < kNumVariants; ++i)
    if (strcmp(kVariants[i].name, name) == 0) return i;
  return -1;
}
typedef CUresult (*TmaEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmaEncodeFn tma_encoder();
static int choose_variant(const escort_geom &g, double density, int Ho) {
  const int k = g.kernel_h, s = g.stride_h;
  const bool tma_w = g.width % 4 == 0 && g.pad_w == (k - 1) / 2 && tma_encoder() != nullptr && !getenv("ESCORT_NO_TMA");
  const char *prefs[4] = {nullptr, nullptr, nullptr, nullptr};
  if (k == 3 && g.kernel_w == 3 && s == 1) {
    if (tma_w && Ho >= 14) prefs[0] = "sconv_tile_ssr1_o3_y7_x4_k3x3_s1_w12_r152";
    else if (Ho >= 20) prefs[0] = "sconv_tile_sbr_o3_y7_x4_k3x3_s1_w12_r152";
    else if (Ho >= 14) prefs[0] = "sconv_tile_o4_y4_x4_k3x3_s1_p2_w8_r232";
    else if (density < 0.2) prefs[0] = "sconv_tile_sbr_o4_y4_x4_k3x3_s1_w12_r152";
    else prefs[0] = "sconv_tile_sbr_o5_y4_x4_k3x3_s1_w12_r152";
    prefs[1] = "sconv_tile_sb_o4_y4_x4_k3x3_s1_w12_r152";
  } else if (k == 5 && g.kernel_w == 5 && s == 1) {
    // 100 handlers: code size decides (instruction-cache misses), so two output channels per lane and, where the
    // batch allows, the packed two-image FFMA2 handlers (half the instructions per FMA)
    prefs[0] = Ho >= 13 ? "sconv_tile_rbj_o2_y4_x4_k5x5_s1_p2_w8_r232" : "sconv_tile_sbr_o2_y4_x4_k5x5_s1_w12_r152";
    prefs[1] = "sconv_tile_sbr_o2_y4_x4_k5x5_s1_w12_r152";
  } else if (k == 1 && g.kernel_w == 1 && s == 1) {
    prefs[0] = Ho >= 14 ? "sconv_tile_sb_o6_y7_x4_k1x1_s1_w8_r232" : "sconv_tile_sb_o8_y2_x4_k1x1_s1_w16_r104";
  } else if (k == 3 && g.kernel_w == 3 && s == 2) {
    prefs[0] = Ho >= 14 ? "sconv_tile_sbr_o4_y4_x4_k3x3_s2_w8_r232" : "sconv_tile_sbr_o4_y2_x4_k3x3_s2_w12_r152";
    prefs[1] = "sconv_tile_sb_o4_y4_x4_k3x3_s2_w8_r232";
  }
  for (const char *p : prefs)
    if (p) {
      const int i = find_variant(p);
      if (i >= 0) return i;
    }
  for (int i = 0; i < kNumVariants; ++i) {
    const VariantDesc &v = kVariants[i];
    if (v.KH == g.kernel_h && v.KW == g.kernel_w && v.S == g.stride_h && (v.MODE == 0 || v.MODE == 2 || v.MODE == 4)) return i;  // (never a W variant)
  }
  return -1;
}

// default backward-weight ("W") variant for a geometry, 1-based; 0 if none applies
int tile_bwdw_variant(const escort_plan *plan) {
  const escort_geom &g = plan->g;
  if (g.stride_h != 1 || g.stride_w != 1 || g.dilation_h != 1 || g.dilation_w != 1) return 0;
  const int k = g.kernel_h;
  const bool tma_w = g.width % 4 == 0 && g.pad_w == (k - 1) / 2 && tma_encoder() != nullptr && !getenv("ESCORT_NO_TMA");
  const double density = (double)plan->nnz / ((double)g.num_output * (g.channels / g.group) * g.kernel_h * g.kernel_w);
  const char *pref = nullptr;
  if (k == 3 && g.kernel_w == 3) {
    if (plan->Ho >= 14) pref = tma_w ? "sconv_tile_ws_o3_y7_x4_k3x3_s1_w12_r152" : "sconv_tile_wb_o3_y7_x4_k3x3_s1_w12_r152";
    else pref = density < 0.2 ? "sconv_tile_wb_o4_y4_x4_k3x3_s1_w12_r152" : "sconv_tile_wb_o5_y4_x4_k3x3_s1_w12_r152";
  } else if (k == 5 && g.kernel_w == 5) {
    pref = "sconv_tile_wb_o2_y4_x4_k5x5_s1_w12_r152";
  } else if (k == 1 && g.kernel_w == 1) {
    pref = "sconv_tile_wb_o8_y2_x4_k1x1_s1_w16_r104";
  }
  if (const char *e = getenv("ESCORT_BWDW_VARIANT")) pref = e;
  return pref ? find_variant(pref) + 1 : 0;
}


// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
static TmaEncodeFn resolve_tma_encoder() {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  TmaEncodeFn fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    fn = (TmaEncodeFn)p;
  cudaGetLastError();
  return fn;
}
static TmaEncodeFn tma_encoder() {
  static const TmaEncodeFn fn = resolve_tma_encoder();  // magic static: resolved once, thread safe
  return fn;
}

static thread_local bool g_building_w = false;  // set while tile_bwdw_build runs tile_plan_build for a W variant

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
  const int layout_rank = plan->layout_rank;
  plan->tile = nullptr;
  plan->tm = nullptr;
  const escort_geom &g = plan->g;
  // the TMEM-window kernel (sconv_tmem.cu): explicit variant ids above the tile variants, and the default wherever it applies
  if (variant > kNumVariants) return tmem_plan_build(plan, variant - kNumVariants - 1, layout_rank, stream);
  if (variant <= 0 && !g_building_w) {
    const int tv = tmem_choose_variant(plan);
    if (tv >= 0) {
      const int rc = tmem_plan_build(plan, tv, layout_rank, stream);
      if (rc || plan->tm) return rc;
    }
  }
  if (g.dilation_h != 1 || g.dilation_w != 1 || g.stride_h != g.stride_w) return 0;
  if (g.width > 1024 || plan->nnz == 0) return 0;
  const int Cg = g.channels / g.group, Mg = g.num_output / g.group;
  const double density = (double)plan->nnz / ((double)g.num_output * Cg * g.kernel_h * g.kernel_w);
  int vidx;
  if (variant > 0) {
    vidx = variant - 1;
    if (vidx >= kNumVariants) return 0;
    const VariantDesc &v = kVariants[vidx];
    if (v.MODE < 5 && !tile_variant_applies(plan, variant)) return 0;
    // W (backward-weight) variants are built by tile_bwdw_build only, never as a forward kernel
    if (v.MODE >= 5 && !(g_building_w && v.KH == g.kernel_h && v.KW == g.kernel_w && g.stride_h == 1 && g.stride_w == 1)) return 0;
  } else {
    vidx = choose_variant(g, density, out_dim(g.height, g.pad_h, g.kernel_h, g.stride_h, g.dilation_h));
    if (vidx < 0) return 0;
  }
  const VariantDesc &V = kVariants[vidx];
  const int OT = V.OT, TY = V.TY, TX = V.TX, KH = V.KH, KW = V.KW, S = V.S, PAIR = V.PAIR;
  const int NCW = V.NCW;
  const int Ho = plan->Ho, Wo = plan->Wo;
  const int PY = ceil_div(Ho, TY), PX = ceil_div(Wo, TX);
  const int per_vec = 4 / PAIR;
  const int PC = (TX - 1) * S + KW;
  if ((TX * S * PAIR) % 4 != 0 && PX > 1) return 0;  // the lanes' 128-bit patch loads need 16-byte aligned tile origins
  // columns the patch-aligned load plan touches (128-bit loads with a 64-bit tail; PAIR 2: exactly PC positions)
  const int XW = PAIR == 2 ? PC : (PC % 4 == 0 ? PC : (PC % 4 <= 2 ? PC - PC % 4 + 2 : PC - PC % 4 + 4));
  // TMA staging (cp.async.bulk.tensor with out-of-bounds zero fill = the halo) needs 16-byte global strides and a
  // 16-byte aligned innermost start coordinate, hence the aligned-body row layout
  // (sieve / rows variants are compiled for one patch-load plan: MODE 1, 3 = aligned body, MODE 2, 4 = patch aligned)
  const bool tma_ok = PAIR == 1 && g.width % 4 == 0 && g.pad_w == (KW - 1) / 2 && tma_encoder() != nullptr &&
                      !getenv("ESCORT_NO_TMA");
  const bool plan_b_variant = V.MODE == 2 || V.MODE == 4 || V.MODE == 6;
  // patch-aligned rows cannot come from TMA: a box starting at x = -pad_w faults (illegal instruction, measured on
  // B200) -- the innermost start coordinate must keep the global address 16-byte aligned, hence the aligned-body
  // layout for TMA and the cp.async loader for patch-aligned rows
  const bool use_tma_b = false;
  const bool use_tma = (tma_ok && !plan_b_variant) || use_tma_b;  // (box limits are checked once the tiling is known)
  if ((V.MODE == 1 || V.MODE == 3 || V.MODE == 5) && !use_tma) return 0;
  // aligned body (PAIR 1, TMA): data column 0 sits on a 16-byte boundary, HL = 4 halo columns to its left and the
  // lane base is the tile's first output column; patch aligned: the halo is exactly pad_w positions wide and the lane
  // base is the patch's first column
  const bool aligned_body = use_tma && !use_tma_b;
  const int PADL = aligned_body ? (KW - 1) / 2 : 0;
  const int HL = aligned_body ? (g.pad_w > 0 ? 4 : 0) : g.pad_w;
  const int lane_col0 = aligned_body ? HL : 0;
  // every lane's reads stay inside its row; the pitch keeps 16-byte alignment of every row start
  const int Pmin = ceil_div(std::max(lane_col0 + (PX - 1) * TX * S + (aligned_body ? PC - PADL : XW), HL + g.width + g.pad_w), per_vec) * per_vec;
  int dev = 0, max_smem = 0, num_sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (max_smem <= 0) max_smem = 227 * 1024;
  if (num_sms <= 0) num_sms = 148;
  const int tab_bytes = 0;
  const long smem_budget = (long)max_smem - kBarBytes - 256;
  const int nblk = ceil_div(Mg, OT);

  // ---- choose (WP, BR, G, pitch, slot residue, lane order): lane utilisation first, then bank conflicts / reuse.
  // "skew" = image-slot stride modulo 32 floats (the bank offset between the image slots of a stage); TMA
  // destinations are 128-byte aligned, so it is 0 there ----
  std::vector<Layout> cands;  // one per (WP, BR), best pitch / residue / lane order each
  for (int WP = 1; WP <= NCW; ++WP) {
    if (NCW % WP) continue;
    if (V.SHFL && WP != 1) continue;  // halo shuffles: a tile row's lanes must be consecutive lanes of one warp
    const int lanes = WP * 32;
    const int WO = NCW / WP;
    for (int nb = 1; nb <= PY; ++nb) {
      const int BR = ceil_div(PY, nb);
      if (nb > 1 && ceil_div(PY, nb - 1) == BR) continue;  // same BR as a smaller band count
      const int per_slot = BR * PX;
      if (per_slot > lanes) continue;
      int GP = std::min(lanes / per_slot, 16);
      if (V.SHFL && GP * per_slot >= 32) --GP;  // the TMA shuffle-halo variants keep one lane spare (the zero lane)
      if (GP < 1) continue;
      const int G = GP * PAIR;
      const int R = (BR * TY - 1) * S + KH;
      const int nslots = GP * per_slot;
      const double util = (double)nslots / lanes * ((double)Ho * Wo / ((double)PY * TY * PX * TX));
      const double band_waste = (double)PY / (nb * BR);
      const double wo_eff = (double)std::min(WO, nblk) / WO;
      const double halo = (double)(BR * TY * S) / R;
      double base_score = util * band_waste * wo_eff * (0.85 + 0.15 * halo) * (1.0 + 0.02 * std::log2((double)WO));
      // pitch / skew / lane order with the fewest LDS.128 wavefronts
      int bP = Pmin, bskew = 0, border = 0, bcost = 1 << 30;
      const int orders[4] = {0, 4, 2, 1};
      for (int oi = 0; oi < (V.SHFL ? 1 : 4); ++oi) {
        const std::vector<Slot> slots = enumerate_slots(GP, BR, PX, orders[oi]);
        for (int P = Pmin; P <= Pmin + 8 * per_vec; P += per_vec) {
          for (int skew = 0; skew <= (use_tma ? 0 : 28); skew += 4) {
            int wsum = 0;
            for (int w = 0; w < WP; ++w) {
              std::vector<unsigned> addr;
              for (int l = 0; l < 32 && w * 32 + l < nslots; ++l) {
                const Slot &sl = slots[w * 32 + l];
                addr.push_back((unsigned)((sl.gs * (skew + 8192) + (sl.pyb * TY * S * P + sl.px * TX * S) * PAIR) * 4));  // distinct slots, same residue
              }
              if (!addr.empty()) wsum += lds128_wavefronts(addr);
            }
            const int cost = wsum * 4096 + (P - Pmin) * 64 + skew + oi;
            if (cost < bcost) { bcost = cost; bP = P; bskew = skew; border = orders[oi]; }
          }
        }
      }
      const long plane_bytes = ((long)R * bP * PAIR + 32) * 4 * GP;
      if (3 * plane_bytes > smem_budget - tab_bytes) continue;  // at least 1 channel x 3 stages (program space extra)
      const int ideal = WP * 4;  // 4 wavefronts per LDS.128 per warp is conflict free
      const double conflict = (double)ideal / std::max(ideal, bcost / 4096);
      // a stage should hold several channels (per-chunk hand-shake and interpreter entry/exit are fixed costs) and a
      // unit several channel blocks (every block re-reads the staged input)
      const long ci_est = std::min<long>((smem_budget - tab_bytes) / (3 * plane_bytes), (long)Cg);
      const double chunk_f = std::min(1.0, 0.55 + 0.075 * (double)ci_est);
      const double reuse_f = 1.0 - 0.35 / (double)WO;
      const double score = base_score * (0.6 + 0.4 * conflict) * chunk_f * reuse_f;
      cands.push_back({WP, G, BR, bP, R, bskew, border, score});
    }
  }
  std::stable_sort(cands.begin(), cands.end(), [](const Layout &a, const Layout &b) { return a.score > b.score; });
  if (layout_rank < 0 || layout_rank >= (int)cands.size()) return 0;
  const Layout best = cands[layout_rank];  // rank 0 = the heuristic's favourite; autotune also times the runners-up
  if (use_tma && (best.P > 256 || best.R > 256)) return 0;  // cuTensorMapEncodeTiled: boxDim <= 256 (CI is capped at 32 below)
  const int WP = best.WP, G = best.G, GP = G / PAIR, BR = best.BR, P = best.P, R = best.R;
  const int WO = NCW / WP;
  const int nbands = ceil_div(PY, BR);
  const int ogroups = ceil_div(nblk, WO);
  const int per_slot = BR * PX;
  const int nslots = GP * per_slot;
  const int plane_f = R * P * PAIR;
  // region header: WO segment offsets (+ WO tap bases for the backward-weight variants)
  const int hdr_bytes = ceil_div((V.MODE >= 5 ? 2 : 1) * WO * 4, 16) * 16;

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
  std::vector<int> prog_pos, tapidx;
  int nchunks = 0, NS = 0, in_bytes = 0, stage_bytes = 0, max_region16 = 0, slot_f = 0, max_seg_taps = 0, scratch_bytes = 0;
  if (V.MODE >= 5) {  // short chunks keep the per-warp scratch rows small
    const char *e = getenv("ESCORT_W_CI");
    CI = std::max(1, std::min(CI, e ? atoi(e) : 10));
  }
  for (int attempt = 0; attempt < 8; ++attempt) {
    nchunks = ceil_div(Cg, CI);
    CI = ceil_div(Cg, nchunks);  // even out the chunks
    // image slots are the outer dimension of a stage ([slot][channel][row][col]); the slot stride keeps the bank
    // offset between slots that the layout search assumed (TMA destinations must be 128-byte aligned instead)
    slot_f = CI * plane_f;
    while ((slot_f - best.skew) % 32) slot_f += 4;
    in_bytes = GP * slot_f * 4;
    std::vector<std::vector<Rec>> buckets((size_t)g.group * nblk * nchunks);
    for (size_t j = 0; j < nz.size(); ++j) {
      const Nz &z = nz[j];
      const int gi = z.oc / Mg, icl = z.ic - gi * Cg, c = icl / CI;
      buckets[((size_t)gi * nblk + oc_block[z.oc]) * nchunks + c].push_back({icl, z.kh, z.kw, oc_slot[z.oc], z.val, (int)j});
    }
    std::vector<unsigned> words;  // u32 stream, regions padded to 16 bytes
    rtab.assign((size_t)g.group * ogroups * nchunks, make_int2(0, 0));
    prog_pos.assign(nz.size(), -1);
    tapidx.clear();
    max_region16 = 0;
    max_seg_taps = 0;
    for (int gi = 0; gi < g.group; ++gi)
      for (int og = 0; og < ogroups; ++og)
        for (int c = 0; c < nchunks; ++c) {
          const size_t region_start = words.size();  // multiple of 4 words
          words.resize(region_start + hdr_bytes / 4, 0u);
          for (int ow = 0; ow < WO; ++ow) {
            const int blk = og * WO + ow;
            words[region_start + ow] = (unsigned)((words.size() - region_start) * 4);  // byte offset of the segment
            if (V.MODE == 3 || V.MODE == 4) {
              // rows stream (16-byte quads): H_0 | H_1 R_0.. | H_2 R_1.. | ... | H_END R_(n-1).. | slack quad; H = {plane
              // byte offset (END = ~0), mask words, pad}; R = the KW weights of one nonempty kernel row (absent taps 0),
              // rows of a step in (oc_local, kh) order.  Bit layout as in the sieve stream.
              const int OPW = std::max(1, 32 / (KH * KW)), NW = ceil_div(OT, OPW);
              const int HW = NW <= 3 ? 4 : 8, RW = KW <= 4 ? 4 : 8;
              struct RStep { unsigned off; unsigned m[8]; std::vector<unsigned> w; std::vector<std::pair<int, int>> src; };
              std::vector<RStep> steps;
              if (blk < nblk) {
                std::vector<Rec> &v = buckets[((size_t)gi * nblk + blk) * nchunks + c];
                std::sort(v.begin(), v.end(), [](const Rec &a, const Rec &b) {
                  if (a.ic != b.ic) return a.ic < b.ic;
                  if (a.o != b.o) return a.o < b.o;
                  if (a.kh != b.kh) return a.kh < b.kh;
                  return a.kw < b.kw;
                });
                int cur_ic = -1, cur_row = -1;
                for (const Rec &r : v) {
                  if (r.ic != cur_ic) {
                    cur_ic = r.ic;
                    cur_row = -1;
                    steps.emplace_back();
                    steps.back().off = (unsigned)((r.ic - c * CI) * (long)plane_f * 4);
                    memset(steps.back().m, 0, sizeof(steps.back().m));
                  }
                  RStep &st = steps.back();
                  if (r.o * KH + r.kh != cur_row) {
                    cur_row = r.o * KH + r.kh;
                    st.w.resize(st.w.size() + RW, 0u);
                  }
                  st.m[r.o / OPW] |= 1u << (((r.o % OPW) * KH + r.kh) * KW + r.kw);
                  st.src.push_back({(int)(st.w.size() - RW + r.kw), r.src});
                  st.w[st.w.size() - RW + r.kw] = __builtin_bit_cast(unsigned, r.val);
                }
              }
              auto emit_rhdr = [&](size_t i) {
                const size_t at = words.size();
                words.resize(at + HW, 0u);
                if (i < steps.size()) {
                  words[at] = steps[i].off;
                  for (int k = 0; k < NW; ++k) words[at + 1 + k] = steps[i].m[k];
                } else {
                  words[at] = 0xffffffffu;
                }
              };
              emit_rhdr(0);
              for (size_t i = 0; i < steps.size(); ++i) {
                emit_rhdr(i + 1);
                const size_t at = words.size();
                for (auto &sv : steps[i].src) prog_pos[sv.second] = (int)(at + sv.first);
                words.insert(words.end(), steps[i].w.begin(), steps[i].w.end());
              }
              words.resize(words.size() + 8, 0u);  // the weight prefetch reads up to two quads past the last row
              continue;
            }
            if (V.MODE >= 1) {
              // sieve stream (4-byte words): H_0 | H_1 W_0.. | H_2 W_1.. | ... | H_END W_(n-1)..  with
              // H = {byte offset of the channel plane (END = ~0), mask words}; bit of handler (o, kh, kw) =
              // ((o % OPW) * KH + kh) * KW + kw of word o / OPW; weights of a step in handler order
              const int OPW = std::max(1, 32 / (KH * KW)), NW = ceil_div(OT, OPW);
              struct Step { unsigned off; unsigned m[8]; std::vector<std::pair<unsigned, int>> w; };
              std::vector<Step> steps;
              if (blk < nblk) {
                std::vector<Rec> &v = buckets[((size_t)gi * nblk + blk) * nchunks + c];
                std::sort(v.begin(), v.end(), [](const Rec &a, const Rec &b) {
                  if (a.ic != b.ic) return a.ic < b.ic;
                  if (a.o != b.o) return a.o < b.o;
                  if (a.kh != b.kh) return a.kh < b.kh;
                  return a.kw < b.kw;
                });
                int cur_ic = -1;
                for (const Rec &r : v) {
                  if (r.ic != cur_ic) {
                    cur_ic = r.ic;
                    steps.emplace_back();
                    steps.back().off = (unsigned)((r.ic - c * CI) * (long)plane_f * 4);
                    memset(steps.back().m, 0, sizeof(steps.back().m));
                  }
                  steps.back().m[r.o / OPW] |= 1u << (((r.o % OPW) * KH + r.kh) * KW + r.kw);
                  steps.back().w.push_back({__builtin_bit_cast(unsigned, r.val), r.src});
                }
              }
              auto emit_hdr = [&](size_t i) {
                if (i < steps.size()) {
                  words.push_back(steps[i].off);
                  for (int k = 0; k < NW; ++k) words.push_back(steps[i].m[k]);
                } else {
                  words.push_back(0xffffffffu);
                  for (int k = 0; k < NW; ++k) words.push_back(0u);
                }
              };
              if (V.MODE >= 5) words[region_start + WO + ow] = (unsigned)tapidx.size();
              int seg_taps = 0;
              emit_hdr(0);
              for (size_t i = 0; i < steps.size(); ++i) {
                emit_hdr(i + 1);
                for (auto &wv : steps[i].w) {
                  prog_pos[wv.second] = (int)words.size();
                  words.push_back(wv.first);
                  if (V.MODE >= 5) tapidx.push_back(wv.second);
                  ++seg_taps;
                }
              }
              max_seg_taps = std::max(max_seg_taps, seg_taps);
              words.push_back(0u);  // the weight prefetch reads two words past the last weight
              words.push_back(0u);
              continue;
            }
            if (blk < nblk) {
              std::vector<Rec> &v = buckets[((size_t)gi * nblk + blk) * nchunks + c];
              std::sort(v.begin(), v.end(), [](const Rec &a, const Rec &b) {
                if (a.ic != b.ic) return a.ic < b.ic;
                if (a.kh != b.kh) return a.kh < b.kh;
                if (a.kw != b.kw) return a.kw < b.kw;
                return a.o < b.o;
              });
              // record i = {payload_i, handler_(i+2)}; a header {handler_0, handler_1} opens the segment
              std::vector<std::pair<unsigned, unsigned>> seq;  // (payload, own handler)
              std::vector<int> seq_src;
              int cur_ic = -1;
              for (const Rec &r : v) {
                if (r.ic != cur_ic) {
                  cur_ic = r.ic;
                  seq.push_back({(unsigned)((r.ic - c * CI) * (long)plane_f * 4), (unsigned)(use_tma ? NC + 1 : NC)});
                  seq_src.push_back(-1);
                }
                seq.push_back({__builtin_bit_cast(unsigned, r.val), (unsigned)((r.o * KH + r.kh) * KW + r.kw)});
                seq_src.push_back(r.src);
              }
              seq.push_back({0u, (unsigned)(NC + 2)});  // end of segment
              seq_src.push_back(-1);
              words.push_back(seq[0].second);
              words.push_back(seq.size() > 1 ? seq[1].second : (unsigned)(NC + 2));
              for (size_t i = 0; i < seq.size(); ++i) {
                if (seq_src[i] >= 0) prog_pos[seq_src[i]] = (int)words.size();
                words.push_back(seq[i].first);
                words.push_back(i + 2 < seq.size() ? seq[i + 2].second : (unsigned)(NC + 2));
              }
            } else {
              words.push_back((unsigned)(NC + 2));  // header of an empty segment: straight to the end handler
              words.push_back((unsigned)(NC + 2));
              words.push_back(0u);
              words.push_back((unsigned)(NC + 2));
            }
          }
          for (int i = 0; i < 2; ++i) {  // prefetch slack: the interpreter reads one record past the end
            words.push_back(0u);
            words.push_back((unsigned)(NC + 2));
          }
          while (words.size() % 4) words.push_back(0u);
          const int len16 = (int)((words.size() - region_start) / 4);
          rtab[((size_t)gi * ogroups + og) * nchunks + c] = make_int2((int)(region_start / 4), len16);
          max_region16 = std::max(max_region16, len16);
        }
    const int prog_bytes = ceil_div(max_region16 * 16, 128) * 128;
    stage_bytes = ceil_div(in_bytes, 128) * 128 + prog_bytes;
    scratch_bytes = V.MODE >= 5 ? NCW * (max_seg_taps + 1) * 128 : 0;
    NS = (int)std::min<long>(std::max<long>(smem_budget - tab_bytes - scratch_bytes, 0L) / stage_bytes, 4L);
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
  pr.R = R; pr.P = P; pr.plane_f = plane_f; pr.slot_f = slot_f; pr.use_tma = use_tma ? 1 : 0; pr.HL = HL; pr.CI = CI; pr.nchunks = nchunks; pr.nblk = nblk; pr.ogroups = ogroups;
  pr.nslots = nslots; pr.NS = NS; pr.stage0_off = kBarBytes + tab_bytes; pr.stage_bytes = stage_bytes;
  pr.in_bytes = in_bytes; pr.hdr_bytes = hdr_bytes;
  {
    int lpr = 1, sh = 0;
    while (lpr < g.width && lpr < 32) { lpr <<= 1; ++sh; }
    pr.lpr_shift = sh;
    pr.RO = 32 / lpr;
  }
  pr.scratch_off = pr.stage0_off + NS * stage_bytes;
  pr.scratch_rows = max_seg_taps + 1;  // + the dummy row 0
  tp->smem_bytes = (size_t)pr.stage0_off + (size_t)NS * stage_bytes + (size_t)scratch_bytes;
  tp->nrecords = prog.size() * 2;

  // ---- lane table ----
  std::vector<int4> lanes(WP * 32, make_int4(0, 0, 0, 0));
  {
    const std::vector<Slot> slots = enumerate_slots(GP, BR, PX, best.order);
    for (int i = 0; i < nslots; ++i) {
      const Slot &sl = slots[i];
      const unsigned edge = (sl.px == 0 ? 1u : 0u) | (sl.px == PX - 1 ? 2u : 0u);
      // shuffle-halo variants: the neighbours' lanes, or the spare lane `nslots` (parked on the zero halo block below)
      // for the first / last tile of a row
      const unsigned zl = (unsigned)(nslots & 31);
      const unsigned srcl = sl.px == 0 ? zl : (unsigned)((i - 1) & 31), srcr = sl.px == PX - 1 ? zl : (unsigned)((i + 1) & 31);
      const unsigned base = (unsigned)((sl.gs * slot_f + (sl.pyb * TY * S * P + lane_col0 + sl.px * TX * S) * PAIR) * 4);
      lanes[i] = make_int4((int)(base | (srcl << 20) | (srcr << 25) | (edge << 30)), sl.gs, sl.pyb, sl.px);
    }
    // lanes without a tile read the first 16 bytes of every row: the 4 halo columns TMA fills with zeros (offset 0)
  }
  int rc = 0;
  if ((rc = upload_vec(&tp->d_lanes, lanes, stream)) || (rc = upload_vec(&tp->d_oc_list, oc_list, stream)) ||
      (rc = upload_vec(&tp->d_prog, prog, stream)) || (rc = upload_vec(&tp->d_rtab, rtab, stream)) ||
      (rc = upload_vec(&tp->d_prog_pos, prog_pos, stream)) || (rc = upload_vec(&tp->d_tapidx, tapidx, stream))) {
    tile_plan_free(tp);
    return rc;
  }
  cudaError_t e = cudaStreamSynchronize(stream);  // host vectors die here
  if (e != cudaSuccess) {
    tile_plan_free(tp);
    return cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
  }
  pr.lanes = tp->d_lanes; pr.oc_list = tp->d_oc_list; pr.prog = tp->d_prog; pr.rtab = tp->d_rtab;
  pr.tapidx = tp->d_tapidx;
  if (V.MODE >= 5) {
    e = cudaFuncSetAttribute(V.bwdw, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) {
      tile_plan_free(tp);
      return cuda_fail(e, "cudaFuncSetAttribute(smem)", __FILE__, __LINE__);
    }
  }
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

// Tensor map of `bottom` as a 4-D tensor {W, H, C, N}; one box = the band's R rows x P columns of CI channels of one image,
// starting at x = -HL (out-of-bounds elements read as zero: the halo costs nothing).  Cached per (pointer, batch).
// Returns 1 if the pointer cannot be used with TMA (misaligned): the caller falls back to the generic kernel.
static int tile_tensor_map(TilePlan *tp, const TileParams &prm, const float *bottom, int num, CUtensorMap *out) {
  if ((reinterpret_cast<uintptr_t>(bottom) & 15) != 0) return 1;
  if (tp->tmap_ptr == bottom && tp->tmap_num == num) {
    *out = tp->tmap;
    return 0;
  }
  const cuuint64_t dims[4] = {(cuuint64_t)prm.W, (cuuint64_t)prm.H, (cuuint64_t)prm.C, (cuuint64_t)num};
  const cuuint64_t strides[3] = {(cuuint64_t)prm.W * 4, (cuuint64_t)prm.H * prm.W * 4, (cuuint64_t)prm.C * prm.H * prm.W * 4};
  const cuuint32_t box[4] = {(cuuint32_t)prm.P, (cuuint32_t)prm.R, (cuuint32_t)prm.CI, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = tma_encoder()(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(bottom), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return ESCORT_EINVAL;
  }
  tp->tmap = *out;
  tp->tmap_ptr = bottom;
  tp->tmap_num = num;
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
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (prm.use_tma) {
    const int rc = tile_tensor_map(tp, prm, bottom, num, &tmap);
    if (rc == 1) return ESCORT_ETRYGENERIC;  // misaligned bottom: this launch goes to the generic kernel
    if (rc) return rc;
  }
  void *args[] = {(void *)&prm, (void *)&num, (void *)&bottom, (void *)&bias, (void *)&fuse_relu, (void *)&top,
                  (void *)&nunits, (void *)&tmap};
  ESCORT_CUDA(cudaLaunchKernel(V.kernel, dim3(grid), dim3(V.NTW * 32), args, tp->smem_bytes, stream));
  return 0;
}

__global__ void gather_idx_kernel(long n, const int *__restrict__ tapidx, const int *__restrict__ table, int *__restrict__ out) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = table[tapidx[t]];
}

__global__ void zero_at_kernel(long nnz, const int *__restrict__ idx, float *__restrict__ dst) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nnz) dst[idx[j]] = 0.f;
}

// Backward weight through a W variant's plan (plan->tile_w): gradients of the nonzero positions only, accumulated
// into the dense weight_diff (always +=) and / or the CSR-ordered buffer (+= or overwritten).
int tile_bwdw(escort_plan *plan, int num, const float *bottom, const float *top_diff, float *wd_dense, float *wd_csr,
              int accumulate, cudaStream_t stream) {
  TilePlan *tp = plan->tile_w;
  TileParams prm = tp->prm;
  prm.n_igroups = ceil_div(num, prm.G);
  prm.tap_dense = tp->d_tap_dense;
  prm.tap_csr = tp->d_tap_csr;
  int nunits = (int)((size_t)prm.n_igroups * prm.nbands * prm.ngroups * prm.ogroups);
  const unsigned grid = (unsigned)std::min(nunits, tp->num_sms);
  const VariantDesc &V = kVariants[tp->vidx];
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (prm.use_tma) {
    const int rc = tile_tensor_map(tp, prm, bottom, num, &tmap);
    if (rc == 1) return ESCORT_ETRYGENERIC;
    if (rc) return rc;
  }
  if (!accumulate && wd_csr) {  // the dense diff always accumulates (Caffe's contract); the CSR-ordered copy may be overwritten
    const unsigned blocks = (unsigned)((plan->nnz + 255) / 256);
    zero_at_kernel<<<blocks, 256, 0, stream>>>(plan->nnz, plan->d_csr_pos, wd_csr);
    ESCORT_LAUNCH_CHECK();
  }
  void *args[] = {(void *)&prm, (void *)&num, (void *)&bottom, (void *)&top_diff, (void *)&wd_dense, (void *)&wd_csr,
                  (void *)&nunits, (void *)&tmap};
  ESCORT_CUDA(cudaLaunchKernel(V.bwdw, dim3(grid), dim3(V.NTW * 32), args, tp->smem_bytes, stream));
  return 0;
}

// build plan->tile_w (the backward-weight plan) with the given W variant (1-based; <= 0: the default for the geometry);
// leaves it null if none applies
int tile_bwdw_build(escort_plan *plan, cudaStream_t stream, int variant) {
  const int v = variant > 0 ? variant : tile_bwdw_variant(plan);
  if (v <= 0) return 0;
  if (plan->tile_w) {
    tile_plan_free(plan->tile_w);
    plan->tile_w = nullptr;
  }
  TilePlan *fwd = plan->tile;
  TmemPlan *fwd_tm = plan->tm;  // tile_plan_build clears both forward slots
  const int rank = plan->layout_rank;
  plan->layout_rank = 0;
  g_building_w = true;
  int rc = tile_plan_build(plan, v, stream);
  g_building_w = false;
  plan->tile_w = plan->tile;
  plan->tile = fwd;
  plan->tm = fwd_tm;
  plan->layout_rank = rank;
  if (rc == 0 && plan->tile_w) {
    // per-tap destinations (stream order), so that the flush needs one index load per tap instead of two dependent ones
    TilePlan *tp = plan->tile_w;
    const long n = plan->nnz;
    cudaError_t e = cudaMalloc((void **)&tp->d_tap_dense, std::max<long>(n, 1) * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&tp->d_tap_csr, std::max<long>(n, 1) * sizeof(int));
    if (e == cudaSuccess) {
      const unsigned blocks = (unsigned)((n + 255) / 256);
      gather_idx_kernel<<<blocks, 256, 0, stream>>>(n, tp->d_tapidx, plan->d_dense_idx, tp->d_tap_dense);
      gather_idx_kernel<<<blocks, 256, 0, stream>>>(n, tp->d_tapidx, plan->d_csr_pos, tp->d_tap_csr);
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) {  // never leave a half-built W plan behind: the generic kernel takes over
      tile_plan_free(plan->tile_w);
      plan->tile_w = nullptr;
      return cuda_fail(e, "tile_bwdw_build", __FILE__, __LINE__);
    }
  }
  return rc;
}

// plan-time selection of the backward-weight variant by measurement: every W variant of the layer's kernel size is
// built and timed on scratch tensors of `num` images; the fastest stays in plan->tile_w
int tile_bwdw_autotune(escort_plan *plan, int num, cudaStream_t stream) {
  const escort_geom &g = plan->g;
  if (g.stride_h != 1 || g.stride_w != 1 || g.dilation_h != 1 || g.dilation_w != 1 || plan->nnz == 0) return 0;
  const size_t in_elems = (size_t)num * g.channels * g.height * g.width;
  const size_t out_elems = (size_t)num * g.num_output * plan->Ho * plan->Wo;
  float *x = nullptr, *dy = nullptr, *wd = nullptr;
  ESCORT_CUDA(cudaMalloc((void **)&x, in_elems * sizeof(float)));
  cudaError_t e = cudaMalloc((void **)&dy, out_elems * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void **)&wd, (size_t)g.num_output * (g.channels / g.group) * g.kernel_h * g.kernel_w * sizeof(float));
  if (e != cudaSuccess) {
    cudaFree(x);
    cudaFree(dy);
    return cuda_fail(e, "cudaMalloc(autotune scratch)", __FILE__, __LINE__);
  }
  cudaMemsetAsync(x, 0, in_elems * sizeof(float), stream);
  cudaMemsetAsync(dy, 0, out_elems * sizeof(float), stream);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int best_v = 0;
  float best_ms = 1e30f;
  for (int v = 1; v <= kNumVariants; ++v) {
    const VariantDesc &V = kVariants[v - 1];
    if (V.MODE < 5 || V.KH != g.kernel_h || V.KW != g.kernel_w) continue;
    if (tile_bwdw_build(plan, stream, v) != 0 || !plan->tile_w) continue;
    float ms_best = 1e30f;
    bool ok = true;
    for (int it = 0; it < 3 && ok; ++it) {
      cudaEventRecord(e0, stream);
      ok = tile_bwdw(plan, num, x, dy, wd, nullptr, 1, stream) == 0;
      cudaEventRecord(e1, stream);
      if (cudaEventSynchronize(e1) != cudaSuccess) ok = false;
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      if (it > 0 && ms < ms_best) ms_best = ms;
    }
    if (ok && ms_best < best_ms) {
      best_ms = ms_best;
      best_v = v;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(x);
  cudaFree(dy);
  cudaFree(wd);
  cudaGetLastError();
  plan->tile_w_tried = 1;
  int rc = tile_bwdw_build(plan, stream, best_v);  // best_v == 0: the default
  if (rc) return rc;
  ESCORT_CUDA(cudaStreamSynchronize(stream));
  return 0;
}

__global__ void tile_refresh_kernel(long nnz, const float *__restrict__ w_dense, const int *__restrict__ dense_idx,
                                    const int *__restrict__ prog_pos, unsigned *__restrict__ prog) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  prog[prog_pos[j]] = __float_as_uint(__ldg(w_dense + dense_idx[j]));  // prog_pos = 4-byte word index of the weight
}

int tile_refresh(escort_plan *plan, const float *weights_dense, cudaStream_t stream) {
  if (plan->tm) return tmem_refresh(plan, weights_dense, stream);
  TilePlan *tp = plan->tile;
  const unsigned blocks = (unsigned)((plan->nnz + 255) / 256);
  tile_refresh_kernel<<<blocks, 256, 0, stream>>>(plan->nnz, weights_dense, plan->d_dense_idx, tp->d_prog_pos,
                                                  reinterpret_cast<unsigned *>(tp->d_prog));
  ESCORT_LAUNCH_CHECK();
  return 0;
}

__global__ void tile_regather_kernel(long nnz, const int4 *__restrict__ meta, const int *__restrict__ prog_pos,
                                     unsigned *__restrict__ prog) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nnz) prog[prog_pos[j]] = (unsigned)meta[j].w;
}
// weights of a freshly built stream from the plan's own value copy (d_meta[j].w, which escort_refresh_values keeps current)
int tile_regather(escort_plan *plan, const int4 *meta, cudaStream_t stream) {
  if (plan->tm) return tmem_regather(plan, meta, stream);
  if (!plan->tile) return 0;
  const unsigned blocks = (unsigned)((plan->nnz + 255) / 256);
  tile_regather_kernel<<<blocks, 256, 0, stream>>>(plan->nnz, meta, plan->tile->d_prog_pos, reinterpret_cast<unsigned *>(plan->tile->d_prog));
  ESCORT_LAUNCH_CHECK();
  return 0;
}

// Interpreter-only microbenchmark (tools/interp_bench.py): synthetic byte-code with `per_load` FMA records per LOAD
// record, handlers drawn uniformly; reports FMA-pipe utilisation of the dispatch loop alone.
extern "C" ESCORT_API int escort_interp_bench(int variant, int per_load, int active_warps, int iters, double *ms_host,
                                              double *tflops_host, int *nc_host) {
  if (variant < 1 || variant > kNumVariants || !ms_host) return ESCORT_EINVAL;
  const VariantDesc &V = kVariants[variant - 1];
  const int NC = V.OT * V.KH * V.KW;
  int nfma = 1536;
  std::vector<uint2> prog;
  unsigned rng = 12345u;
  auto next = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
  if (V.MODE >= 3) {
    // rows variants: `per_load` = density in percent; 64 input-channel steps with random masks
    const int KK = V.KH * V.KW, OPW = std::max(1, 32 / KK), NW = ceil_div(V.OT, OPW), nsteps = 64;
    const int HW = NW <= 3 ? 4 : 8, RW = V.KW <= 4 ? 4 : 8;
    std::vector<unsigned> words;
    std::vector<std::vector<unsigned>> masks(nsteps, std::vector<unsigned>(NW, 0u));
    std::vector<int> nrows(nsteps, 0);
    nfma = 0;
    for (int i = 0; i < nsteps; ++i) {
      for (int o = 0; o < V.OT; ++o)
        for (int kh = 0; kh < V.KH; ++kh) {
          bool any = false;
          for (int kw = 0; kw < V.KW; ++kw)
            if ((int)(next() % 100) < per_load) { masks[i][o / OPW] |= 1u << ((o % OPW) * KK + kh * V.KW + kw); any = true; nfma++; }
          if (any) nrows[i]++;
        }
      if (nrows[i] == 0) { masks[i][0] = 1u; nrows[i] = 1; nfma++; }
    }
    auto rhdr = [&](int i) {
      const size_t at = words.size();
      words.resize(at + HW, 0u);
      if (i < nsteps) { words[at] = (next() % 8) * 1024u; for (int k = 0; k < NW; ++k) words[at + 1 + k] = masks[i][k]; }
      else words[at] = 0xffffffffu;
    };
    rhdr(0);
    for (int i = 0; i < nsteps; ++i) {
      rhdr(i + 1);
      for (int k = 0; k < nrows[i] * RW; ++k) words.push_back(k % RW < V.KW ? __builtin_bit_cast(unsigned, 1e-3f) : 0u);
    }
    words.resize(words.size() + 8, 0u);
    for (size_t i = 0; i < words.size(); i += 2) prog.push_back(make_uint2(words[i], words[i + 1]));
  } else
  if (V.MODE >= 1) {
    // sieve variants: `per_load` = density in percent; 64 input-channel steps with random masks
    const int KK = V.KH * V.KW, OPW = std::max(1, 32 / KK), NW = ceil_div(V.OT, OPW), nsteps = 64;
    std::vector<unsigned> words;
    std::vector<std::vector<unsigned>> masks(nsteps, std::vector<unsigned>(NW, 0u));
    std::vector<int> cnt(nsteps, 0);
    nfma = 0;
    for (int i = 0; i < nsteps; ++i) {
      for (int o = 0; o < V.OT; ++o)
        for (int k = 0; k < KK; ++k)
          if ((int)(next() % 100) < per_load) { masks[i][o / OPW] |= 1u << ((o % OPW) * KK + k); cnt[i]++; }
      if (cnt[i] == 0) { masks[i][0] = 1u; cnt[i] = 1; }
      nfma += cnt[i];
    }
    auto hdr = [&](int i) {
      if (i < nsteps) { words.push_back((next() % 8) * 1024u); for (int k = 0; k < NW; ++k) words.push_back(masks[i][k]); }
      else { words.push_back(0xffffffffu); for (int k = 0; k < NW; ++k) words.push_back(0u); }
    };
    hdr(0);
    for (int i = 0; i < nsteps; ++i) {
      hdr(i + 1);
      for (int k = 0; k < cnt[i]; ++k) words.push_back(__builtin_bit_cast(unsigned, 1e-3f));
    }
    for (int i = 0; i < 4 || words.size() % 2; ++i) words.push_back(0u);
    for (size_t i = 0; i < words.size(); i += 2) prog.push_back(make_uint2(words[i], words[i + 1]));
  } else {
  std::vector<std::pair<unsigned, unsigned>> seq;
  for (int i = 0; i < nfma; ++i) {
    if (per_load > 0 && i % per_load == 0) seq.push_back({(unsigned)((next() % 8) * 1024), (unsigned)NC});
    seq.push_back({__builtin_bit_cast(unsigned, 1e-3f), next() % NC});
  }
  seq.push_back({0u, (unsigned)(NC + 2)});
  prog.push_back(make_uint2(seq[0].second, seq[1].second));
  for (size_t i = 0; i < seq.size(); ++i)
    prog.push_back(make_uint2(seq[i].first, i + 2 < seq.size() ? seq[i + 2].second : (unsigned)(NC + 2)));
  prog.push_back(make_uint2(0u, (unsigned)(NC + 2)));
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  uint2 *d_prog = nullptr;
  float *d_out = nullptr;
  ESCORT_CUDA(cudaMalloc((void **)&d_prog, prog.size() * sizeof(uint2)));
  ESCORT_CUDA(cudaMalloc((void **)&d_out, (size_t)sms * V.NTW * 32 * sizeof(float)));
  ESCORT_CUDA(cudaMemcpy(d_prog, prog.data(), prog.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  const size_t smem = 16384 + prog.size() * 8 + 64;
  ESCORT_CUDA(cudaFuncSetAttribute(V.bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int nrec = (int)prog.size();
  if (active_warps <= 0 || active_warps > V.NCW) active_warps = V.NCW;
  void *args[] = {(void *)&d_prog, (void *)&nrec, (void *)&iters, (void *)&active_warps, (void *)&d_out};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0, 0);
    ESCORT_CUDA(cudaLaunchKernel(V.bench, dim3(sms), dim3(V.NTW * 32), args, smem, 0));
    cudaEventRecord(e1, 0);
    ESCORT_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_prog);
  cudaFree(d_out);
  *ms_host = best;
  const double flops = 2.0 * nfma * (double)V.TY * V.TX * V.PAIR * 32.0 * active_warps * sms * (double)iters;
  if (tflops_host) *tflops_host = flops / (best * 1e-3) / 1e12;
  if (nc_host) *nc_host = NC;
  return 0;
}

// introspection for tests/bench: tiling summary as a string
static int describe_tile(const TilePlan *tp, char *buf, int buflen) {
  const TileParams &p = tp->prm;
  return snprintf(buf, buflen,
                  "%s%s G=%d BR=%d nbands=%d WP=%d WO=%d R=%d P=%d plane_f=%d CI=%d nchunks=%d nblk=%d ogroups=%d nslots=%d NS=%d "
                  "stage=%dB smem=%zu records=%zu",
                  tp->name, p.use_tma ? " tma" : "", p.G, p.BR, p.nbands, p.WP, p.WO, p.R, p.P, p.plane_f, p.CI, p.nchunks, p.nblk,
                  p.ogroups, p.nslots, p.NS, p.stage_bytes, tp->smem_bytes, tp->nrecords);
}

// forward kernel + tiling; once the backward plans exist (first backward call / escort_plan_autotune_backward) also
// " | bwd_data: <kernel + tiling>" and " | bwd_weight: <kernel + tiling>"
extern "C" ESCORT_API int escort_plan_describe(const escort_plan *plan, char *buf, int buflen) {
  if (!plan || !buf || buflen <= 0) return ESCORT_EINVAL;
  int n = 0;
  if (plan->use_s2d && plan->s2d) {  // stride 2 through the space-to-depth sub-plan: "<kernel> ... (s2d)"
    const escort_plan *q = plan->s2d;
    n = q->tm ? tmem_describe(q->tm, buf, buflen) : q->tile ? describe_tile(q->tile, buf, buflen) : snprintf(buf, buflen, "generic");
    if (n > 0 && n < buflen - 8) n += snprintf(buf + n, buflen - n, " (s2d)");
  } else {
    n = plan->tm ? tmem_describe(plan->tm, buf, buflen) : plan->tile ? describe_tile(plan->tile, buf, buflen) : snprintf(buf, buflen, "generic");
  }
  if (plan->bwd && (plan->bwd->tile || plan->bwd->tm) && n > 0 && n < buflen - 16) {
    n += snprintf(buf + n, buflen - n, " | bwd_data: ");
    if (n < buflen) n += plan->bwd->tm ? tmem_describe(plan->bwd->tm, buf + n, buflen - n) : describe_tile(plan->bwd->tile, buf + n, buflen - n);
  }
  if (plan->tile_w && n > 0 && n < buflen - 16) {
    n += snprintf(buf + n, buflen - n, " | bwd_weight: ");
    if (n < buflen) n += describe_tile(plan->tile_w, buf + n, buflen - n);
  }
  return 0;
}

}  // namespace escort
