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
  for (int k = 0; k < LA && k < total; ++k) tm2_issue_load<NCW, 5>(p, num, bottom, smem_raw, smem_base, ls, wid, lane);

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
    if (ls.count < total) tm2_issue_load<NCW, 5>(p, num, bottom, smem_raw, smem_base, ls, wid, lane);
    // (this chunk's smem_full was waited for before its first slot group was filled)
    const unsigned stage_addr = smem_base + p.stage0_off + cur.st * (unsigned)p.stage_bytes;
    const unsigned region = stage_addr + p.in_bytes;
    unsigned rp;  // this warp's records (8 bytes each: {absolute TMEM address of the window in buffer 0, fp32 weight})
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rp) : "r"(region + 4u * (unsigned)wid));
    rp += region;
    // With registers to spare (fewer than 16 warps) the next two records are carried across sections (tm2_section16's
    // invariant: +5-9 %); at 128 registers the four extra live values make ptxas spill accumulators (-5 %).
    constexpr bool CARRY = NCW < 16;
    unsigned c0 = 0, w0 = 0, c1 = 0, w1 = 0;
    if (CARRY) {
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(c0), "=r"(w0) : "r"(rp));
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+8];" : "=r"(c1), "=r"(w1) : "r"(rp));
    }
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
          for (int o = 0; o < OT; ++o) {
            const unsigned n = ((o < 4 ? cnt_lo : cnt_hi) >> (8 * (o & 3))) & 0xffu;
            if constexpr (CARRY) tm2_section16<0>(acc[o], rp, n, c0, w0, c1, w1);
            else tm2_section16_r(acc[o], rp, n, 0u);
          }
        } else {
#pragma unroll
          for (int o = 0; o < OT; ++o) {
            const unsigned n = ((o < 4 ? cnt_lo : cnt_hi) >> (8 * (o & 3))) & 0xffu;
            if constexpr (CARRY) tm2_section16<1>(acc[o], rp, n, c0, w0, c1, w1);
            else tm2_section16_r(acc[o], rp, n, (unsigned)kTm2BufCols);
          }
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
