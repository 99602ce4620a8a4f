// sconv_tile.cu -- register-blocked tile-interpreter forward kernel (placeholder: generic path only).
#include "common.cuh"
namespace escort {
struct TilePlan { int dummy; };
int tile_plan_build(escort_plan *plan, int variant, cudaStream_t stream) { plan->tile = nullptr; return 0; }
void tile_plan_free(TilePlan *tp) { delete tp; }
int tile_forward(escort_plan *, int, const float *, const float *, int, float *, cudaStream_t) { return ESCORT_EINVAL; }
int tile_refresh(escort_plan *, const float *, cudaStream_t) { return 0; }
const char *tile_kernel_name(const TilePlan *) { return "sconv_tile"; }
}  // namespace escort
