// spread_tiled.cu -- tiled spread / interpolation kernels ("method 2").  Placeholder until the
// tiled kernels land: reports "unsupported" so the dispatcher uses the point-driven kernels.
#include "common.cuh"

bool tiled_supported(const b200_plan*, int) { return false; }
void tiled_free(b200_plan*) {}
int spread_tiled(b200_plan*, const float2*, const float*, float2*, int, cudaStream_t) {
  b200_set_error("tiled spread not built");
  return B200_ESTATE;
}
int interp_tiled(b200_plan*, const float2*, float2*, int, float, const float2*, cudaStream_t) {
  b200_set_error("tiled interp not built");
  return B200_ESTATE;
}
