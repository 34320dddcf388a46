// Substep kernels instantiated for MMJelly (include/mpm_b200/MaterialModel.cuh) with both SVD policies.
#include "substep.cuh"

namespace mpm {
const ModelOps* model_jelly(int svd_mode) {
  return svd_mode == MPM_SVD_EXACT ? ModelImpl<MMJelly<Particle, ExactOps>>::ops("MMJelly<ExactOps>")
                                   : ModelImpl<MMJelly<Particle, FastOps>>::ops("MMJelly<FastOps>");
}
}  // namespace mpm
