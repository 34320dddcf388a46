// Substep kernels instantiated for MMFixedCorotated (include/mpm_b200/MaterialModel.cuh) with both SVD policies.
#include "substep.cuh"

namespace mpm {
const ModelOps* model_fixed_corotated(int svd_mode) {
  return svd_mode == MPM_SVD_EXACT ? ModelImpl<MMFixedCorotated<Particle, ExactOps>>::ops("MMFixedCorotated<ExactOps>")
                                   : ModelImpl<MMFixedCorotated<Particle, FastOps>>::ops("MMFixedCorotated<FastOps>");
}
}  // namespace mpm
