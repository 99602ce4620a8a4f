// common.cuh -- shared declarations of the escort_b200 library (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <string>
#include <vector>

#include "escort_b200.h"

namespace escort {

void set_last_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define ESCORT_CUDA(call)                                                          \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) return ::escort::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define ESCORT_LAUNCH_CHECK() ESCORT_CUDA(cudaGetLastError())

#define ESCORT_REQUIRE(cond, msg)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::escort::set_last_error(std::string(msg)); \
      return ESCORT_EINVAL;                       \
    }                                             \
  } while (0)

// internal: the tile kernel cannot take this launch (e.g. a bottom pointer TMA cannot address): use the generic kernel
#define ESCORT_ETRYGENERIC (-100)

inline int out_dim(int in, int pad, int k, int s, int d) { return (in + 2 * pad - (d * (k - 1) + 1)) / s + 1; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// One nonzero in layer-global coordinates (all groups flattened), host side.
struct Nz {
  int oc;        // global output channel
  int ic;        // global input channel
  int kh, kw;
  float val;
  int csr_pos;   // position in the reference's blob layout (weight_offset*g + j)
  int dense_idx; // index into the dense M x C/g x kh x kw weight tensor
};

struct TilePlan;  // sconv_tile.cu
struct TmemPlan;  // sconv_tmem.cu

}  // namespace escort

// The opaque plan (declared in escort_b200.h).
struct escort_plan {
  escort_geom g;
  int Ho, Wo;
  int device;
  long nnz;
  int variant;  // -1 auto
  int layout_rank;  // which of the tile planner's layout candidates (0 = best heuristic score)
  // ---- generic forward: row-major (global rows) ----
  int *d_rowptr;    // num_output + 1
  int4 *d_meta;     // nnz: {in_off, dy, dx, val bits}; in_off = ic*H*W + dy*W + dx  (may be negative)
  // ---- bookkeeping for refresh / gradient scatter ----
  int *d_dense_idx; // nnz (row-major order) -> dense weight index
  int *d_csr_pos;   // nnz (row-major order) -> position in reference CSR blob layout
  // ---- backward data: grouped by input channel ----
  int *d_colptr;    // channels + 1
  int4 *d_tmeta;    // nnz: {oc, dy, dx, val bits}
  int *d_tsrc;      // nnz: transposed slot -> row-major nonzero index
  // ---- backward weight ----
  int4 *d_wmeta;    // nnz (row-major order): {oc, ic, dy, dx}
  // ---- tile-interpreter forward (sconv_tile.cu) ----
  escort::TilePlan *tile;
  // ---- TMEM-window forward (sconv_tmem.cu); at most one of tile / tm is set
  escort::TmemPlan *tm;
  int refreshed;        // escort_refresh_values has run: rebuilt streams re-gather their weights from d_meta
  escort_plan *parent;  // backward-data sub-plan -> the layer's plan (owner of d_meta)
  std::vector<escort::Nz> *host_nz;  // kept for re-planning with another variant
  // ---- backward data through the tile kernel: dX = conv(dY, W^T flipped), a forward plan over the transposed,
  // 180-degree rotated weights (stride 1 only).  Built on first use; owns only g / nnz / host_nz / d_dense_idx / tile.
  escort_plan *bwd;
  int bwd_tried;
  // ---- stride-2 forward as a stride-1 convolution over the space-to-depth of the padded input (core.cu build_s2d_plan):
  // a forward sub-plan over 4 x channels parity planes with a ceil(K / 2) kernel, and the library-owned transformed input
  escort_plan *s2d;
  int use_s2d;          // the forward goes through s2d (default when the sub-plan exists; escort_plan_autotune measures both)
  float *s2d_buf;       // [num][4 * channels][Ho + KH2 - 1][Wo + KW2 - 1], grown on demand
  size_t s2d_elems;
  // ---- backward weight through the tile kernel's "W" variants (stride 1 only), built on first use
  escort::TilePlan *tile_w;
  int tile_w_tried;
  // ---- knobs read ONCE at plan creation (never getenv on the launch path) and the lock of the lazy builds
  int generic_backward;   // ESCORT_GENERIC_BACKWARD: keep the backward on the generic kernels (tests)
  int small_maps;         // 0 with ESCORT_NO_SMALL_MAPS: variant 0 is always the generic kernel, never the small-map one (tests)
  std::mutex *mu;         // guards the first-use builds of bwd / tile_w when several host threads share a plan
};

namespace escort {
// sconv_tile.cu
int tile_plan_build(escort_plan *plan, int variant, cudaStream_t stream);
void tile_plan_free(TilePlan *tp);
int tile_forward(escort_plan *plan, int num, const float *bottom, const float *bias, int fuse_relu, float *top,
                 cudaStream_t stream);
int tile_refresh(escort_plan *plan, const float *weights_dense, cudaStream_t stream);
int tile_bwdw_build(escort_plan *plan, cudaStream_t stream, int variant = 0);
int tile_bwdw_autotune(escort_plan *plan, int num, cudaStream_t stream);
int tile_bwdw(escort_plan *plan, int num, const float *bottom, const float *top_diff, float *wd_dense, float *wd_csr,
              int accumulate, cudaStream_t stream);
const char *tile_kernel_name(const TilePlan *tp);
int tile_plan_variant(const TilePlan *tp);  // 1-based variant id of a built plan
int tile_num_variants();
bool tile_variant_applies(const escort_plan *plan, int variant);  // variant = 1-based index
int tile_regather(escort_plan *plan, const int4 *meta, cudaStream_t stream);
// sconv_tmem.cu (variant ids follow the tile variants: tile_num_tile_variants() + 1 + tv)
int tmem_num_variants();
bool tmem_variant_applies(const escort_plan *plan, int tv);
int tmem_choose_variant(const escort_plan *plan);
int tmem_plan_build(escort_plan *plan, int tv, int layout_rank, cudaStream_t stream);
void tmem_plan_free(TmemPlan *tp);
bool tmem_batch_fits(const escort_plan *plan, int num);
int tmem_forward(escort_plan *plan, int num, const float *bottom, const float *bias, int fuse_relu, float *top,
                 cudaStream_t stream);
int tmem_refresh(escort_plan *plan, const float *weights_dense, cudaStream_t stream);
int tmem_regather(escort_plan *plan, const int4 *meta, cudaStream_t stream);
const char *tmem_kernel_name(const TmemPlan *tp);
int tmem_describe(const TmemPlan *tp, char *buf, int buflen);
}  // namespace escort
