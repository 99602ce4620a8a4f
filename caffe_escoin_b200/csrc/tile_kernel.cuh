// tile_kernel.cuh -- device side of the tile-interpreter forward kernel (see sconv_tile.cu for the design).
#pragma once
#include "common.cuh"

namespace escort {

static constexpr int kLoaderUnroll = 8;

struct TileParams {
  // geometry
  int C, H, W, M, Ho, Wo, pad_h, pad_w;
  int Cg, Mg;              // channels / outputs per conv group
  // tiling
  int G, BR, PX, PY, nbands;
  int WP, WO;              // pixel warps x channel-block warps (WP*WO == NCW of the variant)
  int R, P;                // smem rows per plane, pitch (floats)
  int CI, nchunks;         // channels per chunk, chunks per conv group
  int nblk;                // channel blocks per conv group
  int ogroups;             // ceil(nblk / WO) per conv group
  int nslots;              // valid lane slots per CTA (<= WP*32)
  int chunk_floats;        // CI*G*R*P
  int n_igroups;
  // tables
  const int4 *lanes;       // [WP*32] {lane_base_bytes, g, pyb, px}
  const int *oc_list;      // [group*nblk*OT] global out-channel or -1
  const uint2 *prog;       // records
  const int *seg;          // [group*nblk*nchunks] record offset of segment
  const unsigned short *dst_off;  // [R*W] smem float offset of element e of a plane band (row-major, W wide)
};

#ifndef ESCORT_TILE_DEVICE_ONLY
struct TilePlan {
  int vidx;                // index into the variant table
  const char *name;
  int OT, TY, TX, KH, KW, S;
  TileParams prm;
  size_t smem_bytes;
  dim3 grid;
  int4 *d_lanes;
  int *d_oc_list;
  uint2 *d_prog;
  int *d_seg;
  unsigned short *d_dst_off;
  int *d_prog_pos;         // [nnz] row-major nonzero -> record index (for value refresh)
  size_t nrecords;
};
#endif

#ifndef ESCORT_TILE_HOST_ONLY
// ------------------------------------------------------------------------------------------------------
// loader: copy one chunk (CI channels x G images, the band's input rows) into a smem buffer, interior only
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_chunk(const TileParams &p, const float *__restrict__ bottom, float *buf,
                                           const unsigned short *dst_off_s, int n0, int num, int cbase, int cend,
                                           int ylo, int nrows, int rowshift, int wid, int nw, int lane) {
  const int L = nrows * p.W;
  const int nplanes = p.CI * p.G;
  for (int pl = wid; pl < nplanes; pl += nw) {
    const int ci = pl / p.G, g = pl - ci * p.G;
    const int c = cbase + ci, n = n0 + g;
    if (c >= cend || n >= num) continue;
    const float *src = bottom + (((size_t)n * p.C + c) * p.H + ylo) * p.W;
    float *dst = buf + (size_t)pl * p.R * p.P + rowshift * p.P;
    for (int e0 = 0; e0 < L; e0 += 32 * kLoaderUnroll) {
      float v[kLoaderUnroll];
#pragma unroll
      for (int u = 0; u < kLoaderUnroll; ++u) {
        const int e = e0 + u * 32 + lane;
        v[u] = (e < L) ? __ldg(src + e) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kLoaderUnroll; ++u) {
        const int e = e0 + u * 32 + lane;
        if (e < L) dst[dst_off_s[e]] = v[u];
      }
    }
  }
}

template <int OT, int TY, int TX, int KH, int KW, int S>
__global__ void __launch_bounds__((Interp<OT, TY, TX, KH, KW, S>::NCW + Interp<OT, TY, TX, KH, KW, S>::NLW) * 32, 1)
    sconv_tile_kernel(const TileParams p, int num, const float *__restrict__ bottom, const float *__restrict__ bias,
                      int fuse_relu, float *__restrict__ top) {
  using IP = Interp<OT, TY, TX, KH, KW, S>;
  constexpr int kComputeWarps = IP::NCW, kLoaderWarps = IP::NLW, kTileThreads = (IP::NCW + IP::NLW) * 32;
  extern __shared__ float4 smem_f4[];
  float *smem = reinterpret_cast<float *>(smem_f4);
  float *buf0 = smem;
  float *buf1 = smem + p.chunk_floats;
  unsigned short *dst_off_s = reinterpret_cast<unsigned short *>(smem + 2 * (size_t)p.chunk_floats);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // unit decode: blockIdx.x = ((igroup * nbands + band) * group + cg) * ogroups + og   (og fastest: CTAs that share
  // the same input tile are launched together and hit it in L2)
  int u = blockIdx.x;
  const int og = u % p.ogroups; u /= p.ogroups;
  const int ngroups = p.C / p.Cg;
  const int cg = u % ngroups; u /= ngroups;
  const int band = u % p.nbands; u /= p.nbands;
  const int n0 = u * p.G;

  // band input rows: smem row r <-> input row y_in0 + r
  const int y_in0 = band * p.BR * TY * S - p.pad_h;
  const int ylo = max(0, y_in0);
  const int yhi = min(p.H, y_in0 + p.R);
  const int nrows = max(0, yhi - ylo);
  const int rowshift = ylo - y_in0;

  // zero both buffers once (halo), stage the loader's scatter table
  {
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n4 = (2 * p.chunk_floats) >> 2;
    for (int i = tid; i < n4; i += kTileThreads) smem_f4[i] = z;
    const int ntab = p.R * p.W;
    for (int i = tid; i < ntab; i += kTileThreads) dst_off_s[i] = p.dst_off[i];
  }
  __syncthreads();
  const int cbase0 = cg * p.Cg, cend = cbase0 + p.Cg;
  load_chunk(p, bottom, buf0, dst_off_s, n0, num, cbase0, cend, ylo, nrows, rowshift, wid, kComputeWarps + kLoaderWarps,
             lane);
  __syncthreads();

  const bool is_loader = wid >= kComputeWarps;
  // compute-warp role
  const int pw = wid % p.WP;            // pixel warp
  const int ow = wid / p.WP;            // channel-block warp (valid for compute warps)
  const int blk = og * p.WO + ow;       // channel block within the conv group
  const bool blk_valid = !is_loader && blk < p.nblk;
  int4 li = make_int4(0, 0, 0, 0);
  if (!is_loader) li = p.lanes[pw * 32 + lane];

  float acc[IP::NACC];
#pragma unroll
  for (int i = 0; i < IP::NACC; ++i) acc[i] = 0.f;

  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned pitch_bytes = (unsigned)p.P * 4u;
  const int *seg = p.seg + ((size_t)cg * p.nblk + (blk_valid ? blk : 0)) * p.nchunks;

  for (int c = 0; c < p.nchunks; ++c) {
    if (is_loader) {
      if (c + 1 < p.nchunks)
        load_chunk(p, bottom, ((c + 1) & 1) ? buf1 : buf0, dst_off_s, n0, num, cbase0 + (c + 1) * p.CI, cend, ylo,
                   nrows, rowshift, wid - kComputeWarps, kLoaderWarps, lane);
    } else if (blk_valid) {
      const unsigned lane_base = smem_base + ((c & 1) ? (unsigned)p.chunk_floats * 4u : 0u) + (unsigned)li.x;
      IP::run(acc, p.prog + seg[c], lane_base, pitch_bytes);
    }
    __syncthreads();
  }

  // epilogue: bias + ReLU fused, predicated stores
  if (blk_valid) {
    const int g = li.y, pyb = li.z, px = li.w;
    const int n = n0 + g;
    const int y0 = (band * p.BR + pyb) * TY, x0 = px * TX;
    const int slot = pw * 32 + lane;
    if (slot < p.nslots && n < num && pyb < p.BR) {
#pragma unroll
      for (int o = 0; o < OT; ++o) {
        const int oc = p.oc_list[((size_t)cg * p.nblk + blk) * OT + o];
        if (oc < 0) continue;
        const float b = bias ? __ldg(bias + oc) : 0.f;
        float *out = top + (((size_t)n * p.M + oc) * p.Ho) * p.Wo;
#pragma unroll
        for (int ty = 0; ty < TY; ++ty) {
          const int y = y0 + ty;
          if (y >= p.Ho) continue;
#pragma unroll
          for (int tx = 0; tx < TX; ++tx) {
            const int x = x0 + tx;
            if (x >= p.Wo) continue;
            float v = acc[(o * TY + ty) * TX + tx] + b;
            if (fuse_relu) v = fmaxf(v, 0.f);
            out[(size_t)y * p.Wo + x] = v;
          }
        }
      }
    }
  }
}


#endif  // !ESCORT_TILE_HOST_ONLY

}  // namespace escort
