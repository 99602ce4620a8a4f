// sconv_tmem.cu -- host side of the TMEM-window forward kernel: plan-time compilation of the layer's CSR into per
// (CTA pass, input-channel chunk, warp) record lists, and the launch.  Device side: tmem_kernel.cuh.
//
// Layout of the work:
//  * The batch is flattened the way the reference "stretches" its column indices (base_conv_layer.cpp:96-107): rows of
//    pitch PW = W + pad_w, image blocks of H + pad_h rows, so that output position q reads input position
//    q + kh * PW + kw for every tap and one zero column / row serves as the halo of both neighbours.
//  * A work unit = TILE = 32 * T consecutive positions x one pass of NCW channel blocks (OT output channels each).  All
//    four TMEM lane quadrants hold the SAME positions (one window per lane: T + halo columns per input channel); the
//    NCW compute warps differ in their output channels, so a staged input chunk serves NCW * OT channels.
//  * A warp's records for one slot group (CHS input channels) are sorted by (output channel slot, CSR order): per
//    output channel the accumulation order is the reference's sequential CSR order.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <numeric>

#include "common.cuh"
#include "tmem_kernel.cuh"
#include "tmem2_kernel.cuh"

namespace escort {

struct TmVariant {
  int T, OT, NCW, NPW, CREGS, PREGS, LC;
  int kind;  // 0: producer warps fill the windows (sconv_tmem_kernel), 1: the compute warps do (sconv_tmem2_kernel)
  const char *name;
  const void *kernel;
};

// (T positions per lane, OT output channels per compute warp, NCW compute warps, NPW producer warps, compute /
// producer registers, LC: 1 = the compute warps run the global -> shared loader, 0 = the producer warps do).  The
// setmaxnreg split must stay inside the CTA's own register pool: NCW * (CREGS - R0) <= NPW * (R0 - PREGS) with R0 = the
// launch allocation (static_assert in the kernel).  New variants go at the END: ids are cached by hosts.
#define ESCORT_TM_VARIANTS(X)    \
  X(16, 4, 16, 4, 104, 64, 0)    \
  X(32, 2, 16, 4, 104, 64, 0)    \
  X(16, 6, 12, 4, 144, 72, 0)    \
  X(32, 3, 12, 4, 144, 72, 0)    \
  X(16, 8, 8, 4, 216, 72, 0)     \
  X(32, 4, 8, 4, 216, 72, 0)     \
  X(32, 4, 8, 8, 192, 64, 0)     \
  X(16, 4, 16, 4, 104, 64, 1)

#define ESCORT_TM_ROW(T, OT, NCW, NPW, CR, PR, LC) \
  {T, OT, NCW, NPW, CR, PR, LC, 0, "sconv_tmem_t" #T "_o" #OT "_w" #NCW "p" #NPW "l" #LC, (const void *)&sconv_tmem_kernel<T, OT, NCW, NPW, CR, PR, LC>},
// self-fill kernels: (T, OT, NCW); no producer warps, the compute warps load and fill
#define ESCORT_TM2_VARIANTS(X) \
  X(16, 4, 16)                 \
  X(16, 6, 12)                 \
  X(16, 4, 12)                 \
  X(16, 3, 16)                 \
  X(16, 8, 8)
#define ESCORT_TM2_ROW(T, OT, NCW) {T, OT, NCW, 0, 0, 0, 1, 1, "sconv_tmem2_t" #T "_o" #OT "_w" #NCW, (const void *)&sconv_tmem2_kernel<T, OT, NCW>},
static const TmVariant kTmVariants[] = {ESCORT_TM_VARIANTS(ESCORT_TM_ROW) ESCORT_TM2_VARIANTS(ESCORT_TM2_ROW)};
static constexpr int kNumTmVariants = (int)(sizeof(kTmVariants) / sizeof(kTmVariants[0]));

struct TmemPlan {
  int vidx;
  const char *name;
  int T, OT, NCW;
  TmParams prm;
  size_t smem_bytes;
  int *d_oc_list;
  uint4 *d_prog;
  int2 *d_rtab;
  int *d_prog_pos;  // [nnz] row-major nonzero -> 4-byte word index of its weight in d_prog
  size_t nrecords;
  int num_sms;
};

template <typename T>
static int tm_upload(T **dptr, const std::vector<T> &h, cudaStream_t stream) {
  *dptr = nullptr;
  ESCORT_CUDA(cudaMalloc((void **)dptr, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) ESCORT_CUDA(cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
  return 0;
}

// host-mapped debug words shared by all plans of the process (see tm_mbar_wait)
static int *tm_debug_words(int **host_out) {
  static int *h = nullptr, *d = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    if (cudaHostAlloc((void **)&h, 64, cudaHostAllocMapped) == cudaSuccess && h) {
      memset(h, 0, 64);
      if (cudaHostGetDevicePointer((void **)&d, h, 0) != cudaSuccess) d = nullptr;
    }
    cudaGetLastError();
  });
  if (host_out) *host_out = h;
  return d;
}
extern "C" ESCORT_API int escort_tmem_debug(int *out16) {
  int *h = nullptr;
  tm_debug_words(&h);
  if (!h || !out16) return ESCORT_EINVAL;
  memcpy(out16, h, 64);
  return 0;
}

#ifdef ESCORT_TM_TRACE
// trace build only (make TMTRACE=1): copies CTA 0's event log out and resets it; words = warps * events
extern "C" ESCORT_API int escort_tmem_trace(unsigned long long *out, int *counts) {
  ESCORT_CUDA(cudaDeviceSynchronize());
  ESCORT_CUDA(cudaMemcpyFromSymbol(out, g_tm_trace, sizeof(unsigned long long) * kTmTraceWarps * kTmTraceEvents));
  ESCORT_CUDA(cudaMemcpyFromSymbol(counts, g_tm_trace_n, sizeof(int) * kTmTraceWarps));
  static int zeros[kTmTraceWarps];
  ESCORT_CUDA(cudaMemcpyToSymbol(g_tm_trace_n, zeros, sizeof(zeros)));
  return 0;
}
#endif

int tmem_num_variants() { return kNumTmVariants; }
const char *tmem_kernel_name(const TmemPlan *tp) { return tp->name; }

void tmem_plan_free(TmemPlan *tp) {
  if (!tp) return;
  cudaFree(tp->d_oc_list);
  cudaFree(tp->d_prog);
  cudaFree(tp->d_rtab);
  cudaFree(tp->d_prog_pos);
  delete tp;
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

bool tmem_variant_applies(const escort_plan *plan, int tv) {
  if (tv < 0 || tv >= kNumTmVariants) return false;
  const escort_geom &g = plan->g;
  if (g.stride_h != 1 || g.stride_w != 1 || plan->nnz == 0) return false;
  if (plan->Ho > g.height + g.pad_h || plan->Wo > g.width + g.pad_w) return false;
  const int PW = g.width + g.pad_w;
  const int halo = (g.kernel_h - 1) * g.dilation_h * PW + (g.kernel_w - 1) * g.dilation_w;
  const int slotw = round_up(kTmVariants[tv].T + halo, 16);
  if (2 * slotw > 512) return false;                    // at least two windows in flight
  if (g.kernel_h * g.kernel_w > 255) return false;      // per-slot-group tap counts are bytes
  return true;
}

// default TMEM variant for a geometry (auto mode, no autotune), or -1
int tmem_choose_variant(const escort_plan *plan) {
  if (getenv("ESCORT_NO_TMEM")) return -1;
  const escort_geom &g = plan->g;
  const int Cg = g.channels / g.group;
  const double density = (double)plan->nnz / ((double)g.num_output * Cg * g.kernel_h * g.kernel_w);
  // measured on B200 (profiles/r02_tmem_variant_sweep.txt): few taps per staged window (AlexNet 3x3 at 12 %: about one
  // per output channel and input channel) -> the compute warps wait for windows anyway and take over the loader
  const double taps_per_window = density * g.kernel_h * g.kernel_w;
  const int prefs[2] = {taps_per_window < 1.6 ? 7 : 0, 0};
  for (int tv : prefs)
    if (tmem_variant_applies(plan, tv)) return tv;
  for (int tv = 0; tv < kNumTmVariants; ++tv)
    if (tmem_variant_applies(plan, tv)) return tv;
  return -1;
}

int tmem_plan_build(escort_plan *plan, int tv, int layout_rank, cudaStream_t stream) {
  plan->tm = nullptr;
  if (layout_rank > 3) return 0;
  if (!tmem_variant_applies(plan, tv)) return 0;
  const TmVariant &V = kTmVariants[tv];
  const escort_geom &g = plan->g;
  const int T = V.T, OT = V.OT, NCW = V.NCW;
  const int Cg = g.channels / g.group, Mg = g.num_output / g.group;
  const int KH = g.kernel_h, KW = g.kernel_w;
  const int PW = g.width + g.pad_w, IMGR = g.height + g.pad_h, IMG = IMGR * PW;
  const int HALO = (KH - 1) * g.dilation_h * PW + (KW - 1) * g.dilation_w;
  const int SLOTW = round_up(T + HALO, 16);
  // Channels per slot group.  Measured (profiles/r02_tm_knob_sweep.txt): the per-group cost (hand-shake, tap-count
  // decode, ring bookkeeping) dominates the slack a deeper TMEM ring buys -- two slot groups of as many windows as fit
  // beat three or five smaller ones by 5-12 % on every 3x3 layer, and three large shared-memory stages beat four or
  // five smaller ones by 2-4 %.  layout_rank 1..3 are the runner-up layouts the autotuner re-times.
  const int chs_max = std::max(1, std::min(8, 256 / SLOTW));
  int CHS = chs_max;
  int ns_min = 3;
  if (layout_rank == 1) ns_min = 4;
  if (layout_rank == 2) CHS = std::max(1, chs_max - 1);
  if (layout_rank == 3) CHS = std::max(1, std::min(chs_max, 512 / (3 * SLOTW)));  // three slot groups in flight
  if (layout_rank >= 2 && CHS == chs_max) return 0;  // no such layout candidate
  if (const char *e = getenv("ESCORT_TM_CHS")) CHS = std::max(1, std::min(atoi(e), 512 / (2 * SLOTW)));  // tuning knob
  while (CHS > 1 && CHS * KH * KW > 255) --CHS;
  const int NSLOT = V.kind == 1 ? 2 : std::min(512 / (CHS * SLOTW), kTmMaxSlots);
  if (const char *e = getenv("ESCORT_TM_NSMIN")) ns_min = atoi(e);  // tuning knob
  if (NSLOT < 2) return 0;
  const int TILE = 32 * T;
  const int SW = round_up(TILE - T + SLOTW, 32);
  int dev = 0, max_smem = 0, num_sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (max_smem <= 0) max_smem = 227 * 1024;
  if (num_sms <= 0) num_sms = 148;
  const int nblk = ceil_div(Mg, OT), ogroups = ceil_div(nblk, NCW);
  // loader table: one {src, dst} entry per lane and copy step (RO padded rows, or one 32-column block of a wide row)
  int lpr = 1, lpr_shift = 0;
  while (lpr < PW && lpr < 32) { lpr <<= 1; ++lpr_shift; }
  const int RO = 32 / lpr, nxb = (PW + 31) / 32;
  const int ltab_n = ceil_div(SW / PW + 2, RO) * nxb;
  const int ltab_bytes = round_up(ltab_n * 32 * 8, 128);
  const int stage0_off = round_up(kTmBarBytes, 1024);
  const int SWP = SW / 32 * 36;                // staged row incl. skew padding (one pad chunk per 8 chunks)
  const int ostage_bytes = NCW * (TILE / 32 * 36) * 4;
  const long budget = (long)max_smem - stage0_off - ostage_bytes - ltab_bytes - 1024;  // (1 KiB: base alignment slack)
  if (budget <= 0) return 0;

  // ---- nnz-balanced channel blocks: rows sorted by nnz (descending), dealt in snake order ----
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

  // ---- chunking: CI channels per stage (a multiple of CHS); as many stages as fit, at least 3 ----
  struct Rec { int sg, o, ic, kh, kw; unsigned col; float val; int src; };
  int CI = std::min(round_up(Cg, CHS), 8 * CHS);
  std::vector<unsigned> words;
  std::vector<int2> rtab;
  std::vector<int> prog_pos;
  int nchunks = 0, NS = 0, nsg = 0, in_bytes = 0, stage_bytes = 0, hdr_counts_off = 0;
  for (;;) {
    nchunks = ceil_div(Cg, CI);
    CI = round_up(ceil_div(Cg, nchunks), CHS);  // even out the chunks (the fill lead D is bounded by the smallest one)
    nchunks = ceil_div(Cg, CI);
    nsg = CI / CHS;
    in_bytes = round_up(CI * SWP * 4, 128);
    hdr_counts_off = round_up(NCW * 4, 16);
    const int hdr_bytes = round_up(hdr_counts_off + NCW * nsg * 8, 16);
    // bucket the nonzeros by (conv group, pass, chunk, warp)
    std::vector<std::vector<Rec>> buckets((size_t)g.group * ogroups * nchunks * NCW);
    for (size_t j = 0; j < nz.size(); ++j) {
      const Nz &z = nz[j];
      const int gi = z.oc / Mg, icl = z.ic - gi * Cg, c = icl / CI, in_chunk = icl - c * CI;
      const int b = oc_block[z.oc], og = b / NCW, w = b - og * NCW;
      Rec r;
      r.sg = in_chunk / CHS;
      r.o = oc_slot[z.oc];
      r.ic = icl; r.kh = z.kh; r.kw = z.kw;
      r.col = (unsigned)((in_chunk % CHS) * SLOTW + z.kh * g.dilation_h * PW + z.kw * g.dilation_w);
      r.col |= (unsigned)((w & 3) * 32) << 16;  // absolute TMEM address: the warp's lane quadrant (slot / buffer added by the kernel)
      r.val = z.val;
      r.src = (int)j;
      buckets[(((size_t)gi * ogroups + og) * nchunks + c) * NCW + w].push_back(r);
    }
    words.clear();
    rtab.assign((size_t)g.group * ogroups * nchunks, make_int2(0, 0));
    prog_pos.assign(nz.size(), -1);
    int max_region16 = 0;
    for (int gi = 0; gi < g.group; ++gi)
      for (int og = 0; og < ogroups; ++og)
        for (int c = 0; c < nchunks; ++c) {
          const size_t region_start = words.size();  // multiple of 4 words
          words.resize(region_start + hdr_bytes / 4, 0u);
          for (int w = 0; w < NCW; ++w) {
            std::vector<Rec> &v = buckets[(((size_t)gi * ogroups + og) * nchunks + c) * NCW + w];
            std::sort(v.begin(), v.end(), [](const Rec &a, const Rec &b) {
              if (a.sg != b.sg) return a.sg < b.sg;
              if (a.o != b.o) return a.o < b.o;
              if (a.ic != b.ic) return a.ic < b.ic;
              if (a.kh != b.kh) return a.kh < b.kh;
              return a.kw < b.kw;
            });
            words[region_start + w] = (unsigned)((words.size() - region_start) * 4);  // byte offset of the warp's records
            const size_t counts_at = (region_start + hdr_counts_off / 4 + (size_t)w * nsg * 2) * 4;  // byte index: 8 one-byte counts per slot group
            for (const Rec &r : v) {
              reinterpret_cast<unsigned char *>(words.data())[counts_at + r.sg * 8 + r.o]++;
              words.push_back(r.col);
              prog_pos[r.src] = (int)words.size();
              words.push_back(__builtin_bit_cast(unsigned, r.val));
            }
          }
          words.resize(words.size() + 8, 0u);  // the record walk prefetches up to 32 bytes past the last record
          while (words.size() % 4) words.push_back(0u);
          const int len16 = (int)((words.size() - region_start) / 4);
          rtab[((size_t)gi * ogroups + og) * nchunks + c] = make_int2((int)(region_start / 4), len16);
          max_region16 = std::max(max_region16, len16);
        }
    stage_bytes = in_bytes + round_up(max_region16 * 16, 128);
    NS = (int)std::min<long>(budget / stage_bytes, (long)kTmMaxStages);
    if (NS >= ns_min || (NS >= 3 && CI == CHS)) break;
    if (CI == CHS) return 0;  // does not fit
    CI = std::max(CHS, (CI * 3 / 4) / CHS * CHS);
  }

  TmemPlan *tp = new TmemPlan();
  memset((void *)tp, 0, sizeof(*tp));
  tp->vidx = tv;
  tp->name = V.name;
  tp->T = T; tp->OT = OT; tp->NCW = NCW;
  tp->num_sms = num_sms;
  TmParams &pr = tp->prm;
  pr.C = g.channels; pr.H = g.height; pr.W = g.width; pr.M = g.num_output; pr.Ho = plan->Ho; pr.Wo = plan->Wo;
  pr.pad_h = g.pad_h; pr.pad_w = g.pad_w; pr.Cg = Cg; pr.Mg = Mg; pr.ngroups = g.group;
  pr.PW = PW; pr.IMGR = IMGR; pr.IMG = IMG; pr.TILE = TILE; pr.HALO = HALO; pr.SW = SW;
  pr.SLOTW = SLOTW; pr.CHS = CHS; pr.NSLOT = NSLOT; pr.CI = CI; pr.nchunks = nchunks; pr.NS = NS; pr.nsg = nsg;
  pr.nblk = nblk; pr.ogroups = ogroups;
  pr.stage0_off = stage0_off; pr.stage_bytes = stage_bytes; pr.in_bytes = in_bytes; pr.hdr_counts_off = hdr_counts_off;
  pr.ostage_off = stage0_off + NS * stage_bytes;
  pr.ltab_off = pr.ostage_off + ostage_bytes;
  pr.ltab_n = ltab_n;
  pr.lpr_shift = lpr_shift;
  pr.RO = RO;
  pr.SWP = SWP;
  tp->smem_bytes = (size_t)pr.ltab_off + (size_t)ltab_bytes;
  tp->nrecords = nz.size();
  std::vector<uint4> prog(words.size() / 4);
  memcpy(prog.data(), words.data(), words.size() * 4);
  int rc = 0;
  if ((rc = tm_upload(&tp->d_oc_list, oc_list, stream)) || (rc = tm_upload(&tp->d_prog, prog, stream)) ||
      (rc = tm_upload(&tp->d_rtab, rtab, stream)) || (rc = tm_upload(&tp->d_prog_pos, prog_pos, stream))) {
    tmem_plan_free(tp);
    return rc;
  }
  cudaError_t e = cudaStreamSynchronize(stream);  // host vectors die here
  if (e == cudaSuccess) e = cudaFuncSetAttribute(V.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  if (e != cudaSuccess) {
    tmem_plan_free(tp);
    return cuda_fail(e, "tmem_plan_build", __FILE__, __LINE__);
  }
  pr.oc_list = tp->d_oc_list; pr.prog = tp->d_prog; pr.rtab = tp->d_rtab;
  pr.dbg = tm_debug_words(nullptr);
  pr.skip = getenv("ESCORT_TM_SKIP") ? atoi(getenv("ESCORT_TM_SKIP")) : 0;  // measurement only, see tmem_kernel.cuh
  plan->tm = tp;
  return 0;
}

// can this batch run through the kernel's 32-bit position / offset arithmetic?
bool tmem_batch_fits(const escort_plan *plan, int num) {
  const TmParams &p = plan->tm->prm;
  const double lim = 2147483647.0;
  const double in_lim = lim / 4;  // the loader table holds byte offsets
  return (double)num * p.IMG + p.TILE + p.SW < lim && (double)num * p.M * p.Ho * p.Wo < lim && (double)num * p.C * p.H * p.W < in_lim;
}

int tmem_forward(escort_plan *plan, int num, const float *bottom, const float *bias, int fuse_relu, float *top,
                 cudaStream_t stream) {
  TmemPlan *tp = plan->tm;
  TmParams prm = tp->prm;
  prm.ntiles = (int)(((long)num * prm.IMG + prm.TILE - 1) / prm.TILE);
  int nunits = prm.ntiles * prm.ngroups * prm.ogroups;
  const unsigned grid = (unsigned)std::min(nunits, tp->num_sms);
  const TmVariant &V = kTmVariants[tp->vidx];
  void *args[] = {(void *)&prm, (void *)&num, (void *)&bottom, (void *)&bias, (void *)&fuse_relu, (void *)&top, (void *)&nunits};
  ESCORT_CUDA(cudaLaunchKernel(V.kernel, dim3(grid), dim3((V.NCW + V.NPW) * 32), args, tp->smem_bytes, stream));
  return 0;
}

__global__ void tmem_refresh_kernel(long nnz, const float *__restrict__ w_dense, const int *__restrict__ dense_idx,
                                    const int *__restrict__ prog_pos, unsigned *__restrict__ prog) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  prog[prog_pos[j]] = __float_as_uint(__ldg(w_dense + dense_idx[j]));
}
// the same from the plan's own value copy (d_meta[j].w, kept current by escort_refresh_values): used when a plan is
// rebuilt (escort_plan_set_config / autotune) after a refresh, so the new stream never starts from create-time values
__global__ void tmem_regather_kernel(long nnz, const int4 *__restrict__ meta, const int *__restrict__ prog_pos,
                                     unsigned *__restrict__ prog) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  prog[prog_pos[j]] = (unsigned)meta[j].w;
}

int tmem_refresh(escort_plan *plan, const float *weights_dense, cudaStream_t stream) {
  TmemPlan *tp = plan->tm;
  const unsigned blocks = (unsigned)((plan->nnz + 255) / 256);
  tmem_refresh_kernel<<<blocks, 256, 0, stream>>>(plan->nnz, weights_dense, plan->d_dense_idx, tp->d_prog_pos,
                                                  reinterpret_cast<unsigned *>(tp->d_prog));
  ESCORT_LAUNCH_CHECK();
  return 0;
}

int tmem_regather(escort_plan *plan, const int4 *meta, cudaStream_t stream) {
  TmemPlan *tp = plan->tm;
  const unsigned blocks = (unsigned)((plan->nnz + 255) / 256);
  tmem_regather_kernel<<<blocks, 256, 0, stream>>>(plan->nnz, meta, tp->d_prog_pos, reinterpret_cast<unsigned *>(tp->d_prog));
  ESCORT_LAUNCH_CHECK();
  return 0;
}

int tmem_describe(const TmemPlan *tp, char *buf, int buflen) {
  const TmParams &p = tp->prm;
  return snprintf(buf, buflen,
                  "%s PW=%d IMG=%d TILE=%d HALO=%d SLOTW=%d CHS=%d NSLOT=%d SW=%d CI=%d nchunks=%d nblk=%d ogroups=%d NS=%d "
                  "stage=%dB smem=%zu records=%zu",
                  tp->name, p.PW, p.IMG, p.TILE, p.HALO, p.SLOTW, p.CHS, p.NSLOT, p.SW, p.CI, p.nchunks, p.nblk, p.ogroups, p.NS,
                  p.stage_bytes, tp->smem_bytes, tp->nrecords);
}

}  // namespace escort
