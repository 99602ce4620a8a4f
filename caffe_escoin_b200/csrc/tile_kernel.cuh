// tile_kernel.cuh -- device side of the tile-interpreter forward kernel (design notes in sconv_tile.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace escort {

static constexpr int kLoaderUnroll = 8;
static constexpr int kMaxStages = 8;
static constexpr int kBarBytes = 256;  // mbarrier area at the start of dynamic shared memory

struct TileParams {
  // geometry
  int C, H, W, M, Ho, Wo, pad_h, pad_w;
  int Cg, Mg, ngroups;     // channels / outputs per conv group, number of conv groups
  // tiling
  int G, GP;               // images per CTA, image slots per CTA (= G / PAIR)
  int BR, PX, PY, nbands;  // patch rows per band, patch grid, bands per image
  int WP, WO;              // pixel warps x channel-block warps (WP*WO == NCW of the variant)
  int R, P;                // staged rows per plane, row pitch in positions
  int plane_f;             // floats per (channel, image slot) plane = R*P*PAIR + skew
  int slot_f;              // floats per image slot of a stage (>= CI*plane_f); stage = [slot][channel][row][col]
  int HL;                  // halo columns left of data column 0 in a staged row
  int use_tma;             // input chunks staged by cp.async.bulk.tensor (W % 4 == 0) instead of per-element cp.async
  int CI, nchunks;         // channels per chunk, chunks per conv group
  int nblk, ogroups;       // channel blocks per conv group, CTAs' channel-block groups per conv group
  int nslots;              // valid lane slots per CTA (<= WP*32)
  // pipeline
  int NS;                  // stages
  int stage0_off;          // byte offset of stage 0 in dynamic smem
  int stage_bytes;         // input area + program area
  int in_bytes;            // CI * GP * plane_f * 4
  int hdr_bytes;           // program-region header (segment offsets), multiple of 16
  int n_igroups;           // filled per launch
  // tables
  const int4 *lanes;       // [WP*32] {lane_base_bytes, image slot, pyb, px}
  const int *oc_list;      // [ngroups*nblk*OT] global out-channel or -1
  const uint4 *prog;       // program regions (16-byte units)
  const int2 *rtab;        // [ngroups*ogroups*nchunks] {offset, length} of a region in 16-byte units
  int lpr_shift, RO;       // loader: log2(lanes per row), rows per warp-wide copy instruction
  // backward weight ("W" variants): per-warp scratch rows P[tap ordinal][lane] behind the stage ring
  int scratch_off;         // byte offset of the scratch area in dynamic smem
  int scratch_rows;        // rows (taps) per warp = the longest segment of the plan
  const int *tapidx;       // stream-order tap -> row-major nonzero index
  const int *tap_dense;    // stream-order tap -> index in the dense weight tensor
  const int *tap_csr;      // stream-order tap -> position in the reference's CSR blob layout
};

#ifndef ESCORT_TILE_DEVICE_ONLY
struct TilePlan {
  int vidx;                // index into the variant table
  const char *name;
  int OT, TY, TX, KH, KW, S, PAIR;
  TileParams prm;
  size_t smem_bytes;
  int4 *d_lanes;
  int *d_oc_list;
  uint4 *d_prog;
  int2 *d_rtab;
  int *d_prog_pos;         // [nnz] row-major nonzero -> 4-byte word index of its weight in d_prog
  int *d_tapidx;           // W variants: stream-order tap -> row-major nonzero index
  int *d_tap_dense, *d_tap_csr;  // W variants: stream-order tap -> destination in the dense / CSR-ordered gradient
  size_t nrecords;
  int num_sms;
  // the bulk-tensor map of the last (bottom pointer, batch) this plan was launched with: encoded on the host once per
  // pair, not per launch (Caffe's blobs keep their addresses between iterations)
  CUtensorMap tmap;
  const void *tmap_ptr;
  int tmap_num;
};
#endif

#ifndef ESCORT_TILE_HOST_ONLY
// ---- mbarrier helpers (CTA scope) ---------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned addr) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(addr),
      "r"(parity)
      : "memory");
}

struct UnitCoord {
  int og, cg, band, n0;
};
__device__ __forceinline__ UnitCoord decode_unit(const TileParams &p, int u) {
  // u = ((igroup * nbands + band) * ngroups + cg) * ogroups + og  -- og fastest: CTAs that share an input tile run
  // at the same time and hit it in L2
  UnitCoord c;
  c.og = u % p.ogroups; u /= p.ogroups;
  c.cg = u % p.ngroups; u /= p.ngroups;
  c.band = u % p.nbands; u /= p.nbands;
  c.n0 = u * p.G;
  return c;
}

// loader: copy one chunk (CI channels x G images, the band's input rows) into a stage, interior only.
// Every (image, channel) plane band is one contiguous run of nrows*W floats in the NCHW tensor.  It is copied with
// 4-byte cp.async (LDGSTS): global -> shared without staging registers, so a loader warp keeps hundreds of
// requests in flight (the copy is latency bound; memory-level parallelism is what matters) and scatters straight
// into the padded, optionally image-interleaved, shared layout through the dst_off table.  Completion is tracked
// by the stage's "full" mbarrier via cp.async.mbarrier.arrive.noinc, so the loader never waits for its own loads.
__device__ __forceinline__ void cp_async4(unsigned dst_smem, const float *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned dst_smem, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
// One warp copies whole plane bands; a warp-wide copy instruction covers RO consecutive rows (lanes = RO rows x LPR
// columns, LPR = the power of two >= W, capped at 32) or, for W > 32, one 32-column segment of a row.  Per
// instruction the loop needs a row-bound predicate and two pointer increments -- no table, no division.
template <int PAIR>
__device__ __forceinline__ void load_chunk(const TileParams &p, const float *__restrict__ bottom, float *in_s,
                                           unsigned in_s_addr, int n0, int num, int cbase, int cend, int ylo, int nrows,
                                           int rowshift, int lw, int nlw, int lane) {
  const int nplanes = p.CI * p.G;
  const size_t plane_stride = (size_t)p.H * p.W;  // floats between channels
  const int rsub = lane >> p.lpr_shift, xl = lane & ((1 << p.lpr_shift) - 1);
  const bool edge = rowshift > 0 || rowshift + nrows < p.R;
  int g = 0, ci = lw;
  while (ci >= p.CI) { ci -= p.CI; ++g; }
  for (int pl = lw; pl < nplanes; pl += nlw) {
    const int c = cbase + ci, n = n0 + g;
    const int plane_off = (g / PAIR) * p.slot_f + ci * p.plane_f + (g % PAIR);  // floats
    if (edge) {
      // rows outside the image differ between bands: rewrite them as zeros (edge bands of a banded layer only)
      float *plane = in_s + plane_off;
      const int lo_end = rowshift * p.P, hi_beg = (rowshift + nrows) * p.P, tot = p.R * p.P;
      for (int i = lane; i < lo_end; i += 32) plane[i * PAIR] = 0.f;
      for (int i = hi_beg + lane; i < tot; i += 32) plane[i * PAIR] = 0.f;
    }
    if (c < cend && n < num) {
      const float *src0 = bottom + ((size_t)n * p.C + c) * plane_stride + (size_t)ylo * p.W;
      const unsigned dst0 = in_s_addr + 4u * (unsigned)(plane_off + (rowshift * p.P + p.HL) * PAIR);
      for (int x = xl; x < p.W; x += 32) {  // one trip unless W > 32
        const float *src = src0 + rsub * p.W + x;
        unsigned dst = dst0 + 4u * PAIR * (unsigned)(rsub * p.P + x);
        const int sstep = p.RO * p.W;
        const unsigned dstep = 4u * PAIR * (unsigned)(p.RO * p.P);
#pragma unroll 4
        for (int r = rsub; r < nrows; r += p.RO) {
          cp_async4(dst, src);
          src += sstep;
          dst += dstep;
        }
      }
    }
    ci += nlw;
    while (ci >= p.CI) { ci -= p.CI; ++g; }
  }
}

// loader warps: stream input chunks + byte-code regions through the stage ring (shared by the forward and the
// backward-weight kernel)
template <int VID>
__device__ __forceinline__ void tile_loader_loop(const TileParams &p, int num, const float *__restrict__ bottom, int nunits,
                                                 const CUtensorMap &tmap, unsigned char *smem_raw, unsigned smem_base, int wid,
                                                 int lane) {
  using IP = Interp<VID>;
  constexpr int TY = IP::TY, S = IP::S, PAIR = IP::PAIR;
  constexpr int NCW = IP::NCW, NLW = IP::NLW;
  const unsigned full_bar = smem_base, empty_bar = smem_base + 8 * kMaxStages;
    if constexpr (IP::CREGS > 0) {
      // the loader warpgroup hands its registers to the compute warpgroups; its idle warps leave
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
      if (wid >= NCW + NLW) return;
    }
    const int lw = wid - NCW;
    int it = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      const UnitCoord uc = decode_unit(p, u);
      const int y_in0 = uc.band * p.BR * TY * S - p.pad_h;  // smem row r <-> input row y_in0 + r
      const int ylo = max(0, y_in0), yhi = min(p.H, y_in0 + p.R);
      const int nrows = max(0, yhi - ylo), rowshift = ylo - y_in0;
      const int cbase0 = uc.cg * p.Cg, cend = cbase0 + p.Cg;
      const int2 *rt = p.rtab + ((size_t)uc.cg * p.ogroups + uc.og) * p.nchunks;
      for (int c = 0; c < p.nchunks; ++c, ++it) {
        const int s = it % p.NS, k = it / p.NS;
        if (k > 0) mbar_wait(empty_bar + 8 * s, (unsigned)((k - 1) & 1));
        unsigned char *stage = smem_raw + p.stage0_off + (size_t)s * p.stage_bytes;
        const unsigned stage_addr = smem_base + p.stage0_off + (unsigned)s * p.stage_bytes;
        if (p.use_tma) {
          // one bulk-tensor copy per image of the group: box {P, R, CI, 1} at {-HL, y_in0, channel, image} (HL = 4: the
          // innermost start must be 16-byte aligned); rows /
          // columns / images outside the tensor arrive as zeros, bytes are counted on the stage's full barrier
          if (lw == 0 && lane == 0) {
            const unsigned slot_bytes = (unsigned)(p.P * p.R * p.CI) * 4u;
            asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(full_bar + 8 * s),
                         "r"(slot_bytes * (unsigned)p.GP)
                         : "memory");
            for (int gi = 0; gi < p.GP; ++gi) {
              asm volatile(
                  "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, "
                  "%4, %5}], [%6];" ::"r"(stage_addr + (unsigned)(gi * p.slot_f) * 4u),
                  "l"(&tmap), "r"(-p.HL), "r"(y_in0), "r"(cbase0 + c * p.CI), "r"(uc.n0 + gi), "r"(full_bar + 8 * s)
                  : "memory");
            }
          }
        } else {
          load_chunk<PAIR>(p, bottom, reinterpret_cast<float *>(stage), stage_addr, uc.n0, num, cbase0 + c * p.CI, cend,
                           ylo, nrows, rowshift, lw, NLW, lane);
        }
        {  // byte-code region of this (channel-block group, chunk): contiguous 16-byte async copies
          const int2 r = rt[c];
          const uint4 *src = p.prog + r.x;
          const unsigned dst = stage_addr + p.in_bytes;
          for (int i = lw * 32 + lane; i < r.y; i += NLW * 32) cp_async16(dst + 16u * i, src + i);
        }
        // every loader thread arrives once its own cp.asyncs have landed (noinc: counted in the init count)
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_bar + 8 * s) : "memory");
      }
    }
}

template <int VID>
__global__ void __launch_bounds__(Interp<VID>::NTW * 32, 1)
    sconv_tile_kernel(const TileParams p, int num, const float *__restrict__ bottom, const float *__restrict__ bias,
                      int fuse_relu, float *__restrict__ top, int nunits, const __grid_constant__ CUtensorMap tmap) {
  using IP = Interp<VID>;
  constexpr int OT = IP::OT, TY = IP::TY, TX = IP::TX, S = IP::S, PAIR = IP::PAIR;
  constexpr int NCW = IP::NCW, NLW = IP::NLW, NT = IP::NTW * 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  // the broadcast tells ptxas that wid is warp-uniform: everything derived from it (the byte-code segment address,
  // hence every header word loaded from it) is then uniform too, and the skip branches of the sieve / rows variants
  // compile to plain branches without per-step SHFL + R2UR or BSSY/BSYNC reconvergence bookkeeping
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const unsigned full_bar = smem_base, empty_bar = smem_base + 8 * kMaxStages;

  // one-time set-up: barriers, zero halo (all stages), loader scatter table
  if (tid == 0) {
    for (int s = 0; s < p.NS; ++s) {
      mbar_init(full_bar + 8 * s, NLW * 32);
      mbar_init(empty_bar + 8 * s, NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    float4 *st4 = reinterpret_cast<float4 *>(smem_raw + p.stage0_off);
    const int n4 = (p.NS * p.stage_bytes) >> 4;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < n4; i += NT) st4[i] = z;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the TMA path overwrites these bytes
  }
  __syncthreads();

  if (wid >= NCW) {
    // ================= loader warps: stream input chunks + byte-code regions through the stage ring ==========
    tile_loader_loop<VID>(p, num, bottom, nunits, tmap, smem_raw, smem_base, wid, lane);
    return;
  }

  // ================= compute warps ============================================================================
  // (kept deliberately lean: everything that is live across the interpreter block costs a register on top of the
  // accumulators, so unit coordinates and the lane record are recomputed in the epilogue instead of kept)
  if constexpr (IP::CREGS > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(IP::CREGS) : "memory");
  // lane word: bits 0-19 byte offset of the lane's patch in a stage | 20-24 / 25-29 the lanes that hold its left / right
  // neighbour tile (shuffle-halo variants) | 30 / 31 first / last tile of its row
  const unsigned lane_word = (unsigned)p.lanes[(wid % p.WP) * 32 + lane].x;
  const unsigned lane_base_off = lane_word & 0xfffffu, lane_edge = lane_word >> 20;
  const unsigned pitch_bytes = (unsigned)p.P * 4u * PAIR;
  float acc[IP::NACC];
  unsigned s = 0, ph = 0;
  const unsigned ow4 = 4u * (unsigned)(wid / p.WP);
#pragma unroll 1
  for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
#pragma unroll
    for (int i = 0; i < IP::NACC; ++i) acc[i] = 0.f;
#pragma unroll 1
    for (int c = 0; c < p.nchunks; ++c) {
      mbar_wait(full_bar + 8 * s, ph);
      {
        // channel blocks past the end of the layer own an empty segment (a lone END record), so the interpreter
        // block is entered unconditionally and appears exactly once in the kernel
        const unsigned stage = smem_base + p.stage0_off + s * p.stage_bytes;
        const unsigned region = stage + p.in_bytes;
        unsigned seg;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(seg) : "r"(region + ow4));
        IP::run(acc, region + seg, stage + lane_base_off, pitch_bytes, lane_edge);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar + 8 * s);
      if (++s == (unsigned)p.NS) {
        s = 0;
        ph ^= 1u;
      }
    }
    // epilogue: bias + ReLU fused, predicated stores
    const UnitCoord uc = decode_unit(p, u);
    const int pw = wid % p.WP, ow = wid / p.WP;
    const int blk = uc.og * p.WO + ow;
    const int slot = pw * 32 + lane;
    if (blk < p.nblk && slot < p.nslots) {
      const int4 li = p.lanes[slot];
      const int gs = li.y, pyb = li.z, px = li.w;
      const int y0 = (uc.band * p.BR + pyb) * TY, x0 = px * TX;
#pragma unroll
      for (int o = 0; o < OT; ++o) {
        const int oc = p.oc_list[((size_t)uc.cg * p.nblk + blk) * OT + o];
        if (oc < 0) continue;
        const float b = bias ? __ldg(bias + oc) : 0.f;
#pragma unroll
        for (int im = 0; im < PAIR; ++im) {
          const int n = uc.n0 + gs * PAIR + im;
          if (n >= num) continue;
          float *out = top + (((size_t)n * p.M + oc) * p.Ho) * p.Wo;
#pragma unroll
          for (int ty = 0; ty < TY; ++ty) {
            const int y = y0 + ty;
            if (y >= p.Ho) continue;
#pragma unroll
            for (int tx = 0; tx < TX; ++tx) {
              const int x = x0 + tx;
              if (x >= p.Wo) continue;
              float v = acc[((o * TY + ty) * TX + tx) * PAIR + im] + b;
              if (fuse_relu) v = fmaxf(v, 0.f);
              out[(size_t)y * p.Wo + x] = v;
            }
          }
        }
      }
    }
  }
}

// ---- backward weight: same units, same loader, same stage ring as the forward kernel.  A compute lane keeps its
// OT x TY x TX tile of top_diff in registers for the whole unit; per staged chunk the W variant's handler chain leaves
// one partial per executed tap and lane in the warp's scratch rows, the warp sums every row over its lanes and adds
// it to the weight gradient (dense layout and / or the reference's CSR blob layout) with one atomic per tap.
template <int VID>
__global__ void __launch_bounds__(Interp<VID>::NTW * 32, 1)
    sconv_tile_bwdw_kernel(const TileParams p, int num, const float *__restrict__ bottom, const float *__restrict__ top_diff,
                           float *__restrict__ wd_dense, float *__restrict__ wd_csr, int nunits,
                           const __grid_constant__ CUtensorMap tmap) {
  using IP = Interp<VID>;
  if constexpr (IP::MODE >= 5) {
    constexpr int OT = IP::OT, TY = IP::TY, TX = IP::TX;
    constexpr int NCW = IP::NCW, NLW = IP::NLW, NT = IP::NTW * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const unsigned full_bar = smem_base, empty_bar = smem_base + 8 * kMaxStages;
    if (tid == 0) {
      for (int s = 0; s < p.NS; ++s) {
        mbar_init(full_bar + 8 * s, NLW * 32);
        mbar_init(empty_bar + 8 * s, NCW);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
      float4 *st4 = reinterpret_cast<float4 *>(smem_raw + p.stage0_off);
      const int n4 = (p.NS * p.stage_bytes) >> 4;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = tid; i < n4; i += NT) st4[i] = z;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (wid >= NCW) {
      tile_loader_loop<VID>(p, num, bottom, nunits, tmap, smem_raw, smem_base, wid, lane);
      return;
    }
    if constexpr (IP::CREGS > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(IP::CREGS) : "memory");
    const unsigned lane_word = (unsigned)p.lanes[(wid % p.WP) * 32 + lane].x;
    const unsigned lane_base_off = lane_word & 0xfffffu, lane_edge = lane_word >> 20;
    const unsigned pitch_bytes = (unsigned)p.P * 4u;
    float *scratch = reinterpret_cast<float *>(smem_raw + p.scratch_off) + (size_t)wid * p.scratch_rows * 32;
    const unsigned scratch_lane = smem_base + p.scratch_off + ((unsigned)wid * p.scratch_rows * 32 + lane) * 4u;
    float dy[IP::NACC];
    unsigned s = 0, ph = 0;
    const int pw = wid % p.WP, ow = wid / p.WP;
    const unsigned ow4 = 4u * (unsigned)ow;
#pragma unroll 1
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      {  // the lane's tile of top_diff (zero outside the image / the batch / the layer)
        const UnitCoord uc = decode_unit(p, u);
        const int blk = uc.og * p.WO + ow;
        const int slot = pw * 32 + lane;
        const bool lane_ok = blk < p.nblk && slot < p.nslots;
        const int4 li = p.lanes[slot < p.WP * 32 ? slot : 0];
        const int n = uc.n0 + li.y;
        const int y0 = (uc.band * p.BR + li.z) * TY, x0 = li.w * TX;
#pragma unroll
        for (int o = 0; o < OT; ++o) {
          const int oc = lane_ok ? p.oc_list[((size_t)uc.cg * p.nblk + blk) * OT + o] : -1;
          const float *src = top_diff + (((size_t)n * p.M + (oc < 0 ? 0 : oc)) * p.Ho) * p.Wo;
#pragma unroll
          for (int ty = 0; ty < TY; ++ty)
#pragma unroll
            for (int tx = 0; tx < TX; ++tx) {
              const int y = y0 + ty, x = x0 + tx;
              const bool ok = oc >= 0 && n < num && y < p.Ho && x < p.Wo;
              dy[(o * TY + ty) * TX + tx] = ok ? __ldg(src + (size_t)y * p.Wo + x) : 0.f;
            }
        }
      }
#pragma unroll 1
      for (int c = 0; c < p.nchunks; ++c) {
        mbar_wait(full_bar + 8 * s, ph);
        const unsigned stage = smem_base + p.stage0_off + s * p.stage_bytes;
        const unsigned region = stage + p.in_bytes;
        unsigned seg, tapbase;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(seg) : "r"(region + ow4));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tapbase) : "r"(region + 4u * (unsigned)p.WO + ow4));
        const unsigned ntaps = IP::run_w(dy, region + seg, stage + lane_base_off, pitch_bytes, scratch_lane, lane_edge);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar + 8 * s);  // the stage is no longer read: the sums live in the scratch rows
        for (unsigned k = lane; k < ntaps; k += 32) {
          // destinations first: their L2 latency hides behind the row sum
          const int jd = wd_dense ? __ldg(p.tap_dense + tapbase + k) : 0;
          const int jc = wd_csr ? __ldg(p.tap_csr + tapbase + k) : 0;
          const float *row = scratch + (size_t)(k + 1) * 32;  // row 0 is the handler chain's dummy row
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int l = 0; l < 32; l += 4) {  // rotated: one bank per lane; four independent chains
            s0 += row[(l + lane) & 31];
            s1 += row[(l + 1 + lane) & 31];
            s2 += row[(l + 2 + lane) & 31];
            s3 += row[(l + 3 + lane) & 31];
          }
          const float sum = (s0 + s1) + (s2 + s3);
          if (wd_dense) atomicAdd(wd_dense + jd, sum);
          if (wd_csr) atomicAdd(wd_csr + jc, sum);
        }
        __syncwarp();
        if (++s == (unsigned)p.NS) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  }
}

// ---- interpreter-only microbenchmark: no loader, no barriers.  Every active warp walks the same synthetic
// byte-code (already in the interpreter's record format) `iters` times against a zeroed input area; measures the
// dispatch + FMA ceiling of a variant in isolation (tools/interp_bench.py).
template <int VID>
__global__ void __launch_bounds__(Interp<VID>::NTW * 32, 1)
    interp_bench_kernel(const uint2 *__restrict__ prog, int nrec, int iters, int active_warps, float *__restrict__ out) {
  using IP = Interp<VID>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem_raw);
  const int tid = threadIdx.x, wid = tid >> 5;
  constexpr int kIn = 16384;  // zeroed input area in front of the program
  for (int i = tid; i < kIn / 4; i += blockDim.x) reinterpret_cast<float *>(smem_raw)[i] = 0.f;
  for (int i = tid; i < nrec; i += blockDim.x) reinterpret_cast<uint2 *>(smem_raw + kIn)[i] = prog[i];
  __syncthreads();
  if (wid >= IP::NCW) {
    if constexpr (IP::CREGS > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    return;
  }
  if constexpr (IP::CREGS > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(IP::CREGS) : "memory");
  if (wid >= active_warps) return;
  float acc[IP::NACC];
#pragma unroll
  for (int i = 0; i < IP::NACC; ++i) acc[i] = 0.f;
#pragma unroll 1
  for (int it = 0; it < (IP::MODE >= 5 ? 0 : iters); ++it) IP::run(acc, smem_base + kIn, smem_base + (tid & 31) * 16u, 128u * IP::PAIR, 0u);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < IP::NACC; ++i) sum += acc[i];
  out[blockIdx.x * blockDim.x + tid] = sum;
}
#endif  // !ESCORT_TILE_HOST_ONLY

}  // namespace escort
