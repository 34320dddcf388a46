// mpm_b200 plugin surface — material models.
//
// Concept-compatible with the reference's include/MaterialModel.cuh:18-150: a MaterialModel is a
// trivially copyable class (it is memcpy'd to the device, reference src/mpm.cu:198-201) with public
// `particleVolume`, `particleMass` and
//     __device__ Mat  computePF(Particle const&) const;        // P(F) F^T of the elastic energy
//     __device__ void endOfStepMutation(Particle&) const;      // per-particle plasticity after G2P
// The substep kernels (mpm_b200/csrc) are templates over this type and call nothing else, so a
// user-defined material is compiled in by adding one instantiation (include/mpm_b200/plugin.cuh).
//
// Shipped models and their reference counterparts:
//   MMFixedCorotated  include/MaterialModel.cuh:39-64   (4 floats)
//   MMSnow            include/MaterialModel.cuh:66-115  (7 floats = MpmMaterial at the C ABI)
//   MMJelly           include/MaterialModel.cuh:118-150 (5 floats; never instantiated by the reference)
// Second template parameter = arithmetic policy of the 3x3 SVD (include/mpm_b200/linalg.cuh):
// ExactOps follows the reference operation for operation (double exp, double lambda term, svd3 bit
// for bit); FastOps stays in f32, takes the polar rotation from a Newton iteration and skips the
// plasticity SVD of particles inside the elastic range.  Deviations are bounded by the tests.
#pragma once
#include "linalg.cuh"
#include "types.cuh"

__device__ __forceinline__ real clamp(const real& number, const real& lower, const real& upper) {
  return fmaxf(fminf(number, upper), lower);  // reference include/MaterialModel.cuh:14-16
}

template <class Particle>
class MaterialModelBase {
 public:
  real particleVolume;
  real particleMass;

  MaterialModelBase() = default;
  CUDA_HOSTDEV MaterialModelBase(real volume, real density) : particleVolume(volume) { particleMass = density * volume; }
};

namespace mpm {

// 2 mu (F - R) F^T + lam I, the shape every corotated model below shares
__device__ __forceinline__ Mat corotated_PF(const Mat& F, const Mat& R, real two_mu, real lam_term) {
  Mat PF = mul_abt(two_mu * (F - R), F);
#pragma unroll
  for (int i = 0; i < 3; ++i) PF.m[i][i] += lam_term;
  return PF;
}
// hardening factor exp(h (1 - Jp)) and lambda (Jp - 1) Jp: double like the reference (ExactOps) or f32
template <class Ops>
__device__ __forceinline__ void hardened_lame(real mu0, real lambda0, real hardening, real Jp, real& two_mu, real& lam_term) {
  real e;
  if constexpr (Ops::kExact) e = (real)exp((double)hardening * (1.0 - (double)Jp));
  else e = (hardening == 0.0f) ? 1.0f : __expf(hardening * (1.0f - Jp));
  const real mu = mu0 * e, lambda = lambda0 * e;
  two_mu = 2.0f * mu;
  if constexpr (Ops::kExact) lam_term = (real)((double)lambda * (((double)Jp - 1.0) * (double)Jp));
  else lam_term = lambda * ((Jp - 1.0f) * Jp);
}

// true when every singular value of F lies strictly inside (lo, hi) and det F > 0, i.e. when the
// clamp of MMSnow::endOfStepMutation changes nothing: with C = F^T F, both C - lo^2 I and
// hi^2 I - C are positive definite (Sylvester's criterion, three leading minors each).
__device__ __forceinline__ bool pd3(float a11, float a12, float a13, float a22, float a23, float a33) {
  const float m2 = a11 * a22 - a12 * a12;
  const float det = a11 * (a22 * a33 - a23 * a23) - a12 * (a12 * a33 - a13 * a23) + a13 * (a12 * a23 - a13 * a22);
  return a11 > 0.0f && m2 > 0.0f && det > 0.0f;
}
__device__ __forceinline__ bool within_elastic_range(const Mat& F, float lo, float hi) {
  if (!(F.determinant() > 0.0f)) return false;
  float c[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j) c[i][j] = F.m[0][i] * F.m[0][j] + F.m[1][i] * F.m[1][j] + F.m[2][i] * F.m[2][j];
  const float l2 = lo * lo;
  if (!pd3(c[0][0] - l2, c[0][1], c[0][2], c[1][1] - l2, c[1][2], c[2][2] - l2)) return false;
  if (hi > 1.0e15f) return true;  // "rubber": no upper clamp (hi^2 would overflow)
  const float h2 = hi * hi;
  return pd3(h2 - c[0][0], -c[0][1], -c[0][2], h2 - c[1][1], -c[1][2], h2 - c[2][2]);
}

}  // namespace mpm

// "Neo-hookean based" fixed-corotated elasticity [Stomakhin et al. 2012]
template <class Particle, class Ops = mpm::ExactOps>
class MMFixedCorotated : public MaterialModelBase<Particle> {
 public:
  real mu0;  // Lame parameters
  real lambda0;

  MMFixedCorotated() = default;
  CUDA_HOSTDEV MMFixedCorotated(real volume, real density, real E, real Nu) : MaterialModelBase<Particle>(volume, density) {
    mu0 = E / (2 * (1 + Nu));
    lambda0 = E * Nu / ((1 + Nu) * (1 - 2 * Nu));
  }

  // "J" is the particle's Jp, as in the reference (include/MaterialModel.cuh:56-61)
  __device__ __forceinline__ Mat computePF(Particle const& particle) const {
    const Mat R = linalg::polar_rotation<Ops>(particle.F);
    const real J = particle.Jp;
    real lam_term;
    if constexpr (Ops::kExact) lam_term = (real)((double)lambda0 * (((double)J - 1.0) * (double)J));
    else lam_term = lambda0 * ((J - 1.0f) * J);
    return mpm::corotated_PF(particle.F, R, 2.0f * mu0, lam_term);
  }
  __device__ __forceinline__ void endOfStepMutation(Particle&) const {}
};

// fixed-corotated + hardening + singular-value clamp plasticity [Stomakhin et al. 2013]
template <class Particle, class Ops = mpm::ExactOps>
class MMSnow : public MMFixedCorotated<Particle, Ops> {
 public:
  real hardening;
  real plast_clamp_lower;
  real plast_clamp_higher;

  MMSnow() = default;
  CUDA_HOSTDEV MMSnow(real volume, real density = 400, real E = 1.4e5, real Nu = 0.2, real hardening = 10,
                      real plast_clamp_lower = 1.0 - 2.5e-2, real plast_clamp_higher = 1.0 + 7.5e-3)
      : MMFixedCorotated<Particle, Ops>(volume, density, E, Nu),
        hardening(hardening),
        plast_clamp_lower(plast_clamp_lower),
        plast_clamp_higher(plast_clamp_higher) {}

  // reference include/MaterialModel.cuh:85-93
  __device__ __forceinline__ Mat computePF(Particle const& particle) const {
    const Mat R = linalg::polar_rotation<Ops>(particle.F);
    real two_mu, lam_term;
    mpm::hardened_lame<Ops>(this->mu0, this->lambda0, hardening, particle.Jp, two_mu, lam_term);
    return mpm::corotated_PF(particle.F, R, two_mu, lam_term);
  }

  // reference include/MaterialModel.cuh:95-114.  FastOps skips the SVD for particles inside the
  // elastic range: there the reference only re-synthesises F = U S V^T and Jp * det F / det F from the
  // SVD's own round-off (~1e-6), so leaving F and Jp untouched is within the FAST tolerance.
  __device__ __forceinline__ void endOfStepMutation(Particle& particle) const {
    Mat R;
    (void)endOfStepMutationR(particle, R);
  }

  // ---- optional hooks of the hand-over pipeline (mpm_b200/csrc/g2p_tile.cuh) ----
  // The same mutation, also returning the polar rotation of the MUTATED F when the SVD ran: with
  // F' = U clamp(S) V^T and clamp(S) > 0 that rotation is U V^T, so the stress of the next substep
  // (computePF_R) needs no second SVD.  The reference runs svd3 again on F' in its next P2G and gets
  // the same rotation up to svd3's own accuracy (~1e-6); a handle created with MPM_PIPE_CLASSIC keeps
  // the reference's two decompositions per particle-step.  Returns false when no SVD ran (FastOps,
  // particle inside the elastic range): R is then undefined.
  __device__ __forceinline__ bool endOfStepMutationR(Particle& particle, Mat& R) const {
    Mat& F = particle.F;
    if constexpr (!Ops::kExact) {
      if (mpm::within_elastic_range(F, plast_clamp_lower, plast_clamp_higher)) {
        particle.Jp = clamp(particle.Jp, 0.6f, 20.0f);  // the outer clamp of the Jp update still applies
        return false;
      }
    }
    Mat U, V;
    float sig[3];
    mpm::svd3<Ops>(F, U, sig, V);
#pragma unroll
    for (int i = 0; i < 3; ++i) sig[i] = clamp(sig[i], plast_clamp_lower, plast_clamp_higher);
    const real oldJ = linalg::determinant(F);
    Mat US;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) US.m[i][j] = U.m[i][j] * sig[j];
    F = mul_abt(US, V);
    const real Fdet = linalg::determinant(F);
    particle.Jp = clamp(particle.Jp * oldJ / Fdet, 0.6f, 20.0f);
    R = mul_abt(U, V);
    return sig[2] > 0.0f;  // (a clamp floor <= 0 would leave an inverted or singular F': no shortcut then)
  }
  // computePF with the polar rotation of particle.F supplied by the caller
  __device__ __forceinline__ Mat computePF_R(Particle const& particle, Mat const& R) const {
    real two_mu, lam_term;
    mpm::hardened_lame<Ops>(this->mu0, this->lambda0, hardening, particle.Jp, two_mu, lam_term);
    return mpm::corotated_PF(particle.F, R, two_mu, lam_term);
  }
};

// fixed-corotated + hardening, no plasticity
template <class Particle, class Ops = mpm::ExactOps>
class MMJelly : public MMFixedCorotated<Particle, Ops> {
 public:
  real hardening;

  MMJelly() = default;
  CUDA_HOSTDEV MMJelly(real volume, real density = 1000, real E = 1.0e5, real Nu = 0.3, real hardening = 10)
      : MMFixedCorotated<Particle, Ops>(volume, density, E, Nu), hardening(hardening) {}

  // reference include/MaterialModel.cuh:133-140
  __device__ __forceinline__ Mat computePF(Particle const& particle) const {
    const Mat R = linalg::polar_rotation<Ops>(particle.F);
    real two_mu, lam_term;
    mpm::hardened_lame<Ops>(this->mu0, this->lambda0, hardening, particle.Jp, two_mu, lam_term);
    return mpm::corotated_PF(particle.F, R, two_mu, lam_term);
  }
  // reference include/MaterialModel.cuh:142-149: Jp * det F / det F with the same F, then the clamp
  __device__ __forceinline__ void endOfStepMutation(Particle& particle) const {
    const real oldJ = particle.F.determinant();
    particle.Jp = clamp(particle.Jp * oldJ / particle.F.determinant(), 0.6f, 20.0f);
  }
};
