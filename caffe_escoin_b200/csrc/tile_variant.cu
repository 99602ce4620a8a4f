// tile_variant.cu -- one tile-interpreter kernel instantiation per translation unit (compiled with
// -DESCORT_VARIANT_ID=k so the generated inline-PTX blocks build in parallel).
#define ESCORT_TILE_DEVICE_ONLY
#define ESCORT_STR2(x) #x
#define ESCORT_STR(x) ESCORT_STR2(x)
#define ESCORT_CAT2(a, b) a##b
#define ESCORT_CAT(a, b) ESCORT_CAT2(a, b)
#include ESCORT_STR(ESCORT_CAT(generated/interp_v, ESCORT_VARIANT_ID).inc)
#include "tile_kernel.cuh"

namespace escort {
const void *ESCORT_CAT(tile_variant_kernel_, ESCORT_VARIANT_ID)() {
  return (const void *)&sconv_tile_kernel<ESCORT_VARIANT_ID>;
}
const void *ESCORT_CAT(tile_variant_bwdw_, ESCORT_VARIANT_ID)() {
  return (const void *)&sconv_tile_bwdw_kernel<ESCORT_VARIANT_ID>;
}
const void *ESCORT_CAT(tile_variant_bench_, ESCORT_VARIANT_ID)() {
  return (const void *)&interp_bench_kernel<ESCORT_VARIANT_ID>;
}
const char *ESCORT_CAT(tile_variant_name_, ESCORT_VARIANT_ID)() { return Interp<ESCORT_VARIANT_ID>::name(); }
}  // namespace escort
