// placeholder until the sum-factorised kernel lands
#include "igab_internal.cuh"
namespace igab {
bool eval_tensor_sf_supported(const PatchDev&, const int*, int, const EvalOutDev&) { return false; }
igab_status launch_eval_tensor_sf(const PatchDev&, const PointSource&, long long, int, const EvalOutDev&, cudaStream_t) {
  set_error("no sum-factorised kernel");
  return IGAB_UNSUPPORTED;
}
}  // namespace igab
