// A user-defined material compiled into the library through include/mpm_b200/plugin.cuh — what a
// user of the reference's plugin surface would write, in the reference's Eigen expression style
// (compare include/MaterialModel.cuh:118-150 of the reference): a hardening corotated solid without
// plasticity.  Physically this is the reference's MMJelly, which lets tests/test_gpu_plugin.py check
// it against the CPU checker's restatement of that class; the code below shares nothing with the
// shipped MMJelly but the MaterialModelBase fields.
//
// Also registers a second tuple that differs in the INTERPOLATION KERNEL and TRANSFER SCHEME types
// (thin subclasses), which routes a handle through the generic concept-driven kernels.
#include <mpm_b200/plugin.cuh>

template <class Particle>
class UserHardeningSolid : public MaterialModelBase<Particle> {
 public:
  real mu0;
  real lambda0;
  real hardening;

  UserHardeningSolid() = default;
  UserHardeningSolid(real volume, real density, real E, real Nu, real hardening_) : MaterialModelBase<Particle>(volume, density), hardening(hardening_) {
    mu0 = E / (2 * (1 + Nu));
    lambda0 = E * Nu / ((1 + Nu) * (1 - 2 * Nu));
  }

  __device__ Mat computePF(Particle const& particle) const {
    Mat R, S;
    linalg::polar_decomposition_device(particle.F, R, S);
    real e = std::exp(hardening * (1.0 - particle.Jp));
    real mu = mu0 * e;
    real lambda = lambda0 * e;
    real J = particle.Jp;
    return (2.0 * mu * (particle.F - R) * particle.F.transpose()) + lambda * ((J - 1.0) * J) * Mat::Identity();
  }

  __device__ void endOfStepMutation(Particle& particle) const {
    Mat& F = particle.F;
    real oldJ = F.determinant();
    real newJ = clamp(particle.Jp * oldJ / F.determinant(), 0.6, 20.0);
    particle.Jp = newJ;
  }
};
static_assert(sizeof(UserHardeningSolid<MLS_APIC_Particle>) == 20, "five floats, passed to mpm_create_raw");

MPM_B200_REGISTER_MATERIAL(16, UserHardeningSolid<MLS_APIC_Particle>)

// the same arithmetic behind user-owned kernel / scheme types: no staged kernels for these
class UserQuadraticKernel : public QuadraticInterpolationKernel {};
class UserScheme : public MLS_APIC_Scheme<UserQuadraticKernel> {};
MPM_B200_REGISTER_TUPLE(17, UserHardeningSolid<MLS_APIC_Particle>, UserQuadraticKernel, UserScheme)
