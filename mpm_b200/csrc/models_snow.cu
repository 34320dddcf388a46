// Substep kernels instantiated for MMSnow (include/mpm_b200/MaterialModel.cuh) with both SVD policies.
#include "substep.cuh"

namespace mpm {
const ModelOps* model_snow(int svd_mode) {
  return svd_mode == MPM_SVD_EXACT ? ModelImpl<MMSnow<Particle, ExactOps>>::ops("MMSnow<ExactOps>")
                                   : ModelImpl<MMSnow<Particle, FastOps>>::ops("MMSnow<FastOps>");
}
}  // namespace mpm
