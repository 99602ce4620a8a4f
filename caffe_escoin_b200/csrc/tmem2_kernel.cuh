// tmem2_kernel.cuh -- second generation of the TMEM-window forward kernel: the compute warps fill their own windows
// ("self-fill").  Same algorithm, staging layout and epilogue as tmem_kernel.cuh (the walk of
// include/caffe/util/sconv.hpp:594-678 with the shift of a nonzero as a TMEM column address); what changed is who does
// what, and how many instructions it takes.
//
// Why (event trace of the first kernel, tools/tm_trace.py, profiles/r02_tm_trace_*.txt; ncu profiles/r02_ncu_tm2_*):
//  * its single producer warp per quadrant needed ~590 clocks per channel window and ~70 clocks per cp.async (issue
//    starved next to four compute warps); the compute warps waited for it 45 % of the time;
//  * 42 warp instructions were issued per nonzero of which 8 were FFMA2: integer divisions per chunk, TMEM addresses
//    rebuilt from kernel parameters per output-channel section, spilled accumulators inside the tap loop.
// Here all warps compute (no setmaxnreg split), the NQ warps of a TMEM lane quadrant share the fill of the NEXT slot
// group into the other half of TMEM (two buffers of 256 columns at columns 0 and 256), one named barrier per slot group
// and quadrant replaces the tm_full / tm_empty mbarrier hand-shakes, records carry ABSOLUTE TMEM addresses (lane
// quadrant of their warp included; the buffer is an immediate), every ring position is tracked incrementally, and the
// walk over one output channel's records of a slot group is a single PTX block (odd nonzero, then pairs with both
// window loads in flight and the next pair's records prefetched).
#pragma once
#include "tmem_kernel.cuh"

namespace escort {
#ifndef ESCORT_TMEM_HOST_ONLY

static constexpr int kTm2BufCols = 256;  // TMEM columns per buffer

// One output channel's records of one slot group: n records of {absolute TMEM address, weight} at shared address rp
// (advanced past them).  Accumulation order = record order (the reference's CSR order).  Reads up to 32 bytes past the
// last record (prefetch; the host pads every region).
#define TM2_T16 "{t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15}"
#define TM2_U16 "{u0,u1,u2,u3,u4,u5,u6,u7,u8,u9,u10,u11,u12,u13,u14,u15}"
#define TM2_PACK(P, R)                                                                                                    \
  "mov.b64 " #P "0, {" #R "0, " #R "1};\n\tmov.b64 " #P "1, {" #R "2, " #R "3};\n\tmov.b64 " #P "2, {" #R "4, " #R "5};\n\t"         \
  "mov.b64 " #P "3, {" #R "6, " #R "7};\n\tmov.b64 " #P "4, {" #R "8, " #R "9};\n\tmov.b64 " #P "5, {" #R "10, " #R "11};\n\t"      \
  "mov.b64 " #P "6, {" #R "12, " #R "13};\n\tmov.b64 " #P "7, {" #R "14, " #R "15};\n\t"
#define TM2_FMA8(W, P)                                                                                                    \
  "fma.rn.f32x2 %0, " #W ", " #P "0, %0;\n\tfma.rn.f32x2 %1, " #W ", " #P "1, %1;\n\tfma.rn.f32x2 %2, " #W ", " #P "2, %2;\n\t"     \
  "fma.rn.f32x2 %3, " #W ", " #P "3, %3;\n\tfma.rn.f32x2 %4, " #W ", " #P "4, %4;\n\tfma.rn.f32x2 %5, " #W ", " #P "5, %5;\n\t"     \
  "fma.rn.f32x2 %6, " #W ", " #P "6, %6;\n\tfma.rn.f32x2 %7, " #W ", " #P "7, %7;\n\t"
template <int BUF>
__device__ __forceinline__ void tm2_section16(unsigned long long (&a)[8], unsigned &rp, unsigned n) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 n2, c0, w0, c1, w1;\n\t"
      ".reg .b32 t<16>, u<16>;\n\t"
      ".reg .b64 x<8>, y<8>, ww, vv;\n\t"
      "and.b32 n2, %9, 1;\n\t"
      "setp.eq.u32 p, n2, 0;\n\t"
      "@p bra TM2_PAIRS;\n\t"
      "ld.shared.v2.u32 {c0, w0}, [%8];\n\t"
      "add.u32 %8, %8, 8;\n\t"
      "add.u32 c0, c0, %10;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 " TM2_T16 ", [c0];\n\t"
      "mov.b64 ww, {w0, w0};\n\t"
      "tcgen05.wait::ld.sync.aligned;\n\t"
      TM2_PACK(x, t)
      TM2_FMA8(ww, x)
      "TM2_PAIRS:\n\t"
      "shr.u32 n2, %9, 1;\n\t"
      "setp.eq.u32 p, n2, 0;\n\t"
      "@p bra TM2_DONE;\n\t"
      "ld.shared.v2.u32 {c0, w0}, [%8];\n\t"
      "ld.shared.v2.u32 {c1, w1}, [%8+8];\n\t"
      "TM2_LOOP:\n\t"
      "add.u32 c0, c0, %10;\n\t"
      "add.u32 c1, c1, %10;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 " TM2_T16 ", [c0];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 " TM2_U16 ", [c1];\n\t"
      "mov.b64 ww, {w0, w0};\n\t"
      "mov.b64 vv, {w1, w1};\n\t"
      "ld.shared.v2.u32 {c0, w0}, [%8+16];\n\t"
      "ld.shared.v2.u32 {c1, w1}, [%8+24];\n\t"
      "add.u32 %8, %8, 16;\n\t"
      "tcgen05.wait::ld.sync.aligned;\n\t"
      TM2_PACK(x, t)
      TM2_PACK(y, u)
      TM2_FMA8(ww, x)
      TM2_FMA8(vv, y)
      "sub.u32 n2, n2, 1;\n\t"
      "setp.ne.u32 p, n2, 0;\n\t"
      "@p bra TM2_LOOP;\n\t"
      "TM2_DONE:\n\t"
      "}"
      : "+l"(a[0]), "+l"(a[1]), "+l"(a[2]), "+l"(a[3]), "+l"(a[4]), "+l"(a[5]), "+l"(a[6]), "+l"(a[7]), "+r"(rp)
      : "r"(n), "n"(BUF * kTm2BufCols)
      : "memory");
}

// ring position of a chunk: unit, chunk inside the unit, shared-memory stage and its phase parity
struct Tm2Pos {
  int u, c;
  unsigned st, ph;
};
__device__ __forceinline__ void tm2_advance(const TmParams &p, Tm2Pos &x) {
  if (++x.c == p.nchunks) {
    x.c = 0;
    x.u += (int)gridDim.x;
  }
  if (++x.st == (unsigned)p.NS) {
    x.st = 0;
    x.ph ^= 1u;
  }
}

// ---- loader: the chunk at ring position `ld` global -> shared memory (4-byte cp.async driven by the per-unit table) --
struct Tm2Load {
  Tm2Pos pos;
  int count;      // chunks issued so far
  int tab_unit;   // unit the loader table was built for
  int chan0;      // first input channel of the unit's conv group
  const int2 *rt; // the unit's row of the region table
  int2 r;         // region {offset, length} of `pos` (fetched one call ahead: a dependent global load otherwise)
};
template <int NCW>
__device__ __forceinline__ void tm2_unit_setup(const TmParams &p, int num, unsigned char *smem_raw, Tm2Load &ls, int wid, int lane) {
  const TmUnit uc = tm_decode_unit(p, ls.pos.u);
  ls.tab_unit = ls.pos.u;
  ls.chan0 = uc.cg * p.Cg;
  ls.rt = p.rtab + ((size_t)uc.cg * p.ogroups + uc.og) * p.nchunks;
  int2 *ltab = reinterpret_cast<int2 *>(smem_raw + p.ltab_off);
  asm volatile("bar.sync 5, %0;" ::"n"(NCW * 32) : "memory");  // every warp is done with the previous unit's table
  const int tile_start = uc.tile * p.TILE;
  const int R0 = tile_start / p.PW;
  const int nrows = (tile_start + p.SW - 1) / p.PW - R0 + 1;
  const int nxb = (p.PW + 31) >> 5;
  const int lpr = 1 << p.lpr_shift;
  for (int e = wid * 32 + lane; e < p.ltab_n * 32; e += NCW * 32) {
    const int j = e >> 5, l = e & 31;
    const int jr = j / nxb, xb = j - jr * nxb;
    const int row = jr * p.RO + (l >> p.lpr_shift), x = (l & (lpr - 1)) + 32 * xb;
    int2 ent = make_int2(-1, -1);
    if (row < nrows && x < p.PW) {
      const int R = R0 + row;
      const int d = R * p.PW + x - tile_start;
      if (d >= 0 && d < p.SW) {
        const int n = R / p.IMGR, yy = R - n * p.IMGR;
        ent.y = (int)((tm_skew((unsigned)d >> 2) << 4) + (((unsigned)d & 3u) << 2));
        if (n < num && yy >= p.pad_h && x >= p.pad_w) ent.x = ((n * p.C * p.H + (yy - p.pad_h)) * p.W + (x - p.pad_w)) * 4;  // byte offset
      }
    }
    ltab[e] = ent;
  }
  asm volatile("bar.sync 5, %0;" ::"n"(NCW * 32) : "memory");
}
template <int NCW>
__device__ __forceinline__ void tm2_issue_load(const TmParams &p, int num, const float *__restrict__ bottom, unsigned char *smem_raw,
                                               unsigned smem_base, Tm2Load &ls, int wid, int lane) {
  TM_EV(7);
  if (ls.pos.u != ls.tab_unit) {
    tm2_unit_setup<NCW>(p, num, smem_raw, ls, wid, lane);
    ls.r = ls.rt[ls.pos.c];
  }
  const unsigned smem_full = smem_base, smem_empty = smem_base + 8 * kTmMaxStages;
  if (ls.count >= p.NS) tm_mbar_wait(smem_empty + 8 * ls.pos.st, ls.pos.ph ^ 1u, 1, p.dbg);
  TM_EV(8);
  const int2 *ltab = reinterpret_cast<const int2 *>(smem_raw + p.ltab_off);
  const unsigned stage_addr = smem_base + p.stage0_off + ls.pos.st * (unsigned)p.stage_bytes;
  const int nch = min(p.CI, p.Cg - ls.pos.c * p.CI);
  const size_t HW4 = (size_t)(p.H * p.W) * 4;
  const char *src0 = reinterpret_cast<const char *>(bottom) + (size_t)(ls.chan0 + ls.pos.c * p.CI) * HW4;
  const unsigned row_bytes = (unsigned)p.SWP * 4u;
#define TM2_CP4(K) asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + (K) * row_bytes), "l"(src + (K) * sstep), "r"(nbytes))
#pragma unroll 1
  for (int j = wid; j < ((p.skip & 4) ? 0 : p.ltab_n); j += NCW) {  // this warp's table steps, all channels of the chunk
    const int2 e = ltab[j * 32 + lane];
    if (e.y >= 0) {
      const char *src = e.x >= 0 ? src0 + (unsigned)e.x : reinterpret_cast<const char *>(bottom);
      const unsigned nbytes = e.x >= 0 ? 4u : 0u;
      const size_t sstep = e.x >= 0 ? HW4 : 0;
      unsigned dst = stage_addr + (unsigned)e.y;
      int ch = nch;
#pragma unroll 1
      for (; ch >= 4; ch -= 4) {  // four copies back to back (ptxas pads a lone LDGSTS with dummy LDS)
        TM2_CP4(0);
        TM2_CP4(1);
        TM2_CP4(2);
        TM2_CP4(3);
        src += 4 * sstep;
        dst += 4 * row_bytes;
      }
      if (ch & 2) {
        TM2_CP4(0);
        TM2_CP4(1);
        src += 2 * sstep;
        dst += 2 * row_bytes;
      }
      if (ch & 1) TM2_CP4(0);
    }
  }
#undef TM2_CP4
  {  // record region of this (pass, chunk): contiguous 16-byte async copies
    const uint4 *src = p.prog + ls.r.x;
    const unsigned dst = stage_addr + p.in_bytes;
#pragma unroll 1
    for (int i = wid * 32 + lane; i < ls.r.y; i += NCW * 32) tm_cp_async16(dst + 16u * i, src + i);
  }
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_full + 8 * ls.pos.st) : "memory");
  ++ls.count;
  tm2_advance(p, ls.pos);
  if (ls.pos.u == ls.tab_unit) ls.r = ls.rt[ls.pos.c];  // next call's region (same unit: same table row)
  TM_EV(9);
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
template <int T, int OT, int NCW>
__global__ void __launch_bounds__(NCW * 32, 1)
    sconv_tmem2_kernel(const TmParams p, int num, const float *__restrict__ bottom, const float *__restrict__ bias, int fuse_relu,
                       float *__restrict__ top, int nunits) {
  static_assert(T == 16, "tm2_section16 is written for 16-column windows");
  static_assert(NCW % 4 == 0 && OT <= 8, "whole quadrants; tap counts are 8 bytes per slot group");
  constexpr int NQ = NCW / 4;  // warps per TMEM lane quadrant
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const unsigned smem_full = smem_base, smem_empty = smem_base + 8 * kTmMaxStages;
  uint32_t *tbase_slot = reinterpret_cast<uint32_t *>(smem_raw + (2 * kTmMaxStages + 8 * kTmMaxSlots) * 8);
  tm_ev_init();
  if (tid == 0) {
    for (int s = 0; s < p.NS; ++s) {
      tm_mbar_init(smem_full + 8 * s, NCW * 32);  // every thread arrives through its cp.asyncs
      tm_mbar_init(smem_empty + 8 * s, NCW);      // every warp releases the stage
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"l"((uint64_t)__cvta_generic_to_shared(tbase_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tm_fence_before();
  __syncthreads();
  tm_fence_after();
  if (*tbase_slot != 0u) {  // the records hold absolute addresses: the CTA owns all 512 columns, so its base is column 0
    if (p.dbg && tid == 0 && atomicCAS(p.dbg, 0, 9) == 0) p.dbg[1] = (int)*tbase_slot;
    __trap();
  }
  const int q = wid & 3, wa = wid >> 2;  // TMEM lane quadrant (= the warp's scheduler), share index inside the quadrant
  const uint32_t tq = (uint32_t)(q * 32) << 16;
  const int nblocks = p.SLOTW >> 4;  // 16-column blocks per channel window
  const int my_units = blockIdx.x < (unsigned)nunits ? (nunits - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total = my_units * p.nchunks;
  Tm2Load ls;
  ls.pos = {(int)blockIdx.x, 0, 0u, 0u};
  ls.count = 0;
  ls.tab_unit = -1;
  const int LA = p.NS - 2;  // chunks the loads run ahead
  for (int k = 0; k < LA && k < total; ++k) tm2_issue_load<NCW>(p, num, bottom, smem_raw, smem_base, ls, wid, lane);

  // this warp's share of one slot group's fill: 16-column blocks wa, wa + NQ, ... of the chn channel windows of the
  // staged rows at `rows`, two blocks (eight 128-bit loads) in flight, into the buffer at column tcol0
  const unsigned lane_chunk = (unsigned)lane * (T / 4);
  const int fk0 = wa / nblocks, fblk0 = wa - fk0 * nblocks;  // (channel, block) of block index wa
  auto fill_group = [&](unsigned rows, int chn, uint32_t tcol0) {
    const int nb = (p.skip & 1) ? 0 : chn * nblocks;
    int k = fk0, blk = fblk0;
#define TM2_LDS128(V, J) asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4+" #J "*16];" : "=r"(V[4 * J]), "=r"(V[4 * J + 1]), "=r"(V[4 * J + 2]), "=r"(V[4 * J + 3]) : "r"(a))
#define TM2_LDBLOCK(V, K, BLK) { const unsigned a = rows + (unsigned)((K) * p.SWP) * 4u + (tm_skew(lane_chunk + 4u * (unsigned)(BLK)) << 4); TM2_LDS128(V, 0); TM2_LDS128(V, 1); TM2_LDS128(V, 2); TM2_LDS128(V, 3); }
#define TM2_ADVANCE(K, BLK) { BLK += NQ; while (BLK >= nblocks) { BLK -= nblocks; ++K; } }
#pragma unroll 1
    for (int b = wa; b < nb; b += 2 * NQ) {
      uint32_t v0[16], v1[16];
      int k1 = k, blk1 = blk;
      TM2_ADVANCE(k1, blk1)
      const bool two = b + NQ < nb;
      TM2_LDBLOCK(v0, k, blk)
      if (two) TM2_LDBLOCK(v1, k1, blk1)
      tm_st16(tcol0 + (unsigned)(k * p.SLOTW + 16 * blk), v0);
      if (two) tm_st16(tcol0 + (unsigned)(k1 * p.SLOTW + 16 * blk1), v1);
      k = k1; blk = blk1;
      TM2_ADVANCE(k, blk)
    }
#undef TM2_ADVANCE
#undef TM2_LDBLOCK
#undef TM2_LDS128
  };
  auto group_sync = [&]() {  // my stores are done and visible; everybody's reads of the other buffer are done
    tm_wait_st();
    tm_fence_before();
    asm volatile("bar.sync %0, %1;" ::"r"(q + 1), "n"(NQ * 32) : "memory");
    tm_fence_after();
  };

  unsigned long long acc[OT][T / 2];
  unsigned buf = 0;  // TMEM buffer of the current slot group
  if (total > 0) {
    tm_mbar_wait(smem_full, 0u, 4, p.dbg);
    fill_group(smem_base + p.stage0_off, min(p.CHS, min(p.CI, p.Cg)), tq);
    group_sync();
  }
  Tm2Pos cur = {(int)blockIdx.x, 0, 0u, 0u};
  const bool fill_first = (wa & 1) == 0;  // half of the quadrant fills while the other half computes
  const unsigned row_bytes = (unsigned)p.SWP * 4u;
#pragma unroll 1
  for (int it = 0; it < total; ++it) {
    if (cur.c == 0) {
#pragma unroll
      for (int o = 0; o < OT; ++o)
#pragma unroll
        for (int k = 0; k < T / 2; ++k) acc[o][k] = 0ull;
    }
    if (ls.count < total) tm2_issue_load<NCW>(p, num, bottom, smem_raw, smem_base, ls, wid, lane);
    // (this chunk's smem_full was waited for before its first slot group was filled)
    const unsigned stage_addr = smem_base + p.stage0_off + cur.st * (unsigned)p.stage_bytes;
    const unsigned region = stage_addr + p.in_bytes;
    unsigned rp;  // this warp's records (8 bytes each: {absolute TMEM address of the window in buffer 0, fp32 weight})
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rp) : "r"(region + 4u * (unsigned)wid));
    rp += region;
    unsigned cp = region + p.hdr_counts_off + (unsigned)(wid * p.nsg) * 8u;  // per slot group: 8 tap counts (one byte per o)
    const int nch = min(p.CI, p.Cg - cur.c * p.CI);
    Tm2Pos nxt = cur;
    tm2_advance(p, nxt);
    const unsigned nstage_addr = smem_base + p.stage0_off + nxt.st * (unsigned)p.stage_bytes;
    const int nnch = min(p.CI, p.Cg - nxt.c * p.CI);  // channels of the next chunk
#pragma unroll 1
    for (int ch0 = 0; ch0 < nch; ch0 += p.CHS) {
      // the next slot group: in this chunk, or the first one of the next chunk (whose stage must have landed)
      unsigned nrows = stage_addr + (unsigned)(ch0 + p.CHS) * row_bytes;
      int nchn = nch - ch0 - p.CHS;
      if (nchn <= 0) {
        nchn = 0;
        if (it + 1 < total) {
          nrows = nstage_addr;
          nchn = nnch;
          TM_EV(1);
          tm_mbar_wait(smem_full + 8 * nxt.st, nxt.ph, 4, p.dbg);
          TM_EV(2);
        }
      }
      nchn = min(nchn, p.CHS);
      const uint32_t ntcol = tq + (buf ^ 1u) * (unsigned)kTm2BufCols;
      if (fill_first && nchn > 0) {
        fill_group(nrows, nchn, ntcol);
        TM_EV(14);
      }
      unsigned cnt_lo, cnt_hi;
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cnt_lo), "=r"(cnt_hi) : "r"(cp));
      cp += 8;
      if ((cnt_lo | cnt_hi) != 0u && !(p.skip & 2)) {
        if (buf == 0u) {
#pragma unroll
          for (int o = 0; o < OT; ++o) tm2_section16<0>(acc[o], rp, ((o < 4 ? cnt_lo : cnt_hi) >> (8 * (o & 3))) & 0xffu);
        } else {
#pragma unroll
          for (int o = 0; o < OT; ++o) tm2_section16<1>(acc[o], rp, ((o < 4 ? cnt_lo : cnt_hi) >> (8 * (o & 3))) & 0xffu);
        }
      }
      TM_EV(4);
      if (!fill_first && nchn > 0) {
        fill_group(nrows, nchn, ntcol);
        TM_EV(14);
      }
      group_sync();
      TM_EV(3);
      buf ^= 1u;
    }
    __syncwarp();
    if (lane == 0) tm_mbar_arrive(smem_empty + 8 * cur.st);
    if (cur.c == p.nchunks - 1) {
      // ---- epilogue: bias + ReLU, lane-major tile -> (skewed) shared memory -> coalesced predicated stores ----
      TM_EV(5);
      const TmUnit uc = tm_decode_unit(p, cur.u);
      const int blk = uc.og * NCW + wid;
      if (blk < p.nblk) {
        const int HoWo = p.Ho * p.Wo;
        int ooff[T];  // where position i * 32 + lane of the tile goes (image offset + pixel), -1 = no such output
        {
          const int pos = uc.tile * p.TILE + lane;  // the host checks num * IMG < 2^31
          int n = pos / p.IMG;
          const int r = pos - n * p.IMG;
          int y = r / p.PW, x = r - y * p.PW;
          const int dy = 32 / p.PW, dx = 32 - dy * p.PW;
#pragma unroll
          for (int i = 0; i < T; ++i) {
            ooff[i] = (n < num && y < p.Ho && x < p.Wo) ? (n * p.M) * HoWo + y * p.Wo + x : -1;
            x += dx;
            y += dy;
            if (x >= p.PW) {
              x -= p.PW;
              ++y;
            }
            while (y >= p.IMGR) {
              y -= p.IMGR;
              ++n;
            }
          }
        }
        const unsigned ost = smem_base + p.ostage_off + (unsigned)wid * (unsigned)(36 * T * 4);  // TILE floats + skew padding
#pragma unroll
        for (int o = 0; o < OT; ++o) {
          const int oc = p.oc_list[((size_t)uc.cg * p.nblk + blk) * OT + o];
          if (oc < 0) continue;
          const float b = bias ? __ldg(bias + oc) : 0.f;
#pragma unroll
          for (int j = 0; j < T / 4; ++j) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              v[2 * e] = __uint_as_float((unsigned)(acc[o][2 * j + e] & 0xffffffffull)) + b;
              v[2 * e + 1] = __uint_as_float((unsigned)(acc[o][2 * j + e] >> 32)) + b;
            }
            if (fuse_relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            const unsigned ch = tm_skew((unsigned)lane * (T / 4) + (unsigned)j);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ost + (ch << 4)), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
          }
          __syncwarp();
          float *out = top + (size_t)oc * HoWo;
#pragma unroll
          for (int i = 0; i < T; ++i) {
            const unsigned pl = (unsigned)(i * 32 + lane);
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(ost + (tm_skew(pl >> 2) << 4) + ((pl & 3u) << 2)) : "memory");
            if (ooff[i] >= 0) out[ooff[i]] = v;
          }
          __syncwarp();
        }
      }
      TM_EV(6);
    }
    cur = nxt;
  }
  tm_fence_before();
  __syncthreads();
  tm_ev_fini();
  if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(0u) : "memory");
}
#endif  // !ESCORT_TMEM_HOST_ONLY

}  // namespace escort
