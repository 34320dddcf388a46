// mpm_b200 plugin surface — particle <-> grid transfer schemes.
//
// Concept-compatible with the reference's include/TransferScheme.h:6-143.  One instance per
// particle (registers); the generic substep kernels (mpm_b200/csrc/kernels.cuh) call, in this order,
//   P2G:  p2g_prepare_particle(particle, par, kernel, material); get_range_begin();
//         p2g_node_contribution(particle, dist, mass, i, j, k, out)          per stencil node
//   G2P:  g2p_prepare_particle(particle, par, kernel); get_range_begin();
//         g2p_node_contribution(particle, dist, node, i, j, k)               per stencil node
//         g2p_finish_particle(particle, par);  material.endOfStepMutation(particle)
// with dist = x_node - x_particle in world units (reference src/mpm.cu:51-60, 150-160).
// The staged production kernels (p2g_sched.cuh, g2p_tile.cuh) are specialisations of exactly this
// arithmetic for MLS_APIC_Scheme<QuadraticInterpolationKernel>.
#pragma once
#include "InterpolationKernel.cuh"
#include "types.cuh"

// reference include/TransferScheme.h:6-29: dx = 1.0 / N narrowed to f32, dx_inv = 1.0 / dx from that f32
struct SimulationParameters {
  float dt;
  u32 N;  // cubic unit domain
  real N_real;
  real dx;
  real dx_inv;
  CUDA_HOSTDEV SimulationParameters(float dt_, u32 N_)
      : dt(dt_), N(N_), N_real(static_cast<real>(N_)), dx((real)(1.0 / N_)), dx_inv((real)(1.0 / (double)((real)(1.0 / N_)))) {}
};

class TransferSchemeBase {
 public:
  Veci range_begin;
  // only after p2g_prepare_particle() or g2p_prepare_particle()
  CUDA_HOSTDEV Veci get_range_begin() { return range_begin; }
};

// MLS-MPM / APIC [Hu et al. 2018] (reference include/TransferScheme.h:57-143)
template <class InterpolationKernel>
class MLS_APIC_Scheme : public TransferSchemeBase {
 public:
  Mat Dinv;
  Mat affine;
  WeightMat<InterpolationKernel::size()> weights;

  template <class MaterialModel>
  __device__ __forceinline__ void p2g_prepare_particle(MLS_APIC_Particle const& particle, SimulationParameters const& par,
                                                       InterpolationKernel const& interpolationKernel,
                                                       MaterialModel const& materialModel) {
    prepare_weights(particle, par, interpolationKernel);
    const Mat PF = materialModel.computePF(particle);
    // stress = -Dinv dt vol PF, scaled left to right like the reference expression
    const Mat stress = (((-Dinv) * par.dt) * materialModel.particleVolume) * PF;
    affine = stress + materialModel.particleMass * particle.C;
  }

  // momentum (xyz) and mass (w) the particle adds to one node
  __device__ __forceinline__ void p2g_node_contribution(MLS_APIC_Particle const& particle, Vec const& dist_part2node,
                                                        real particle_mass, int i, int j, int k, Vec4& contribution) {
    const Vec a = affine * dist_part2node;
    const real weight = weights(0, i) * weights(1, j) * weights(2, k);
    contribution[0] = weight * (particle.v[0] * particle_mass + a[0]);
    contribution[1] = weight * (particle.v[1] * particle_mass + a[1]);
    contribution[2] = weight * (particle.v[2] * particle_mass + a[2]);
    contribution[3] = weight * particle_mass;
  }

  __device__ __forceinline__ void g2p_prepare_particle(MLS_APIC_Particle& particle, SimulationParameters const& par,
                                                       InterpolationKernel const& interpolationKernel) {
    prepare_weights(particle, par, interpolationKernel);
    particle.C = Mat::Zero();
    particle.v = Vec::Zero();
  }

  // v += w v_i ;  C += (w v_i) (d^T Dinv)
  __device__ __forceinline__ void g2p_node_contribution(MLS_APIC_Particle& particle, Vec const& dist_part2node,
                                                        Vec4 const& grid_node, int i, int j, int k) {
    const real weight = weights(0, i) * weights(1, j) * weights(2, k);
    const Vec wv = weight * grid_node.head3();
    particle.v += wv;
    Vec dD;  // row vector d^T Dinv
#pragma unroll
    for (int c = 0; c < 3; ++c) dD(c) = dist_part2node(0) * Dinv(0, c) + dist_part2node(1) * Dinv(1, c) + dist_part2node(2) * Dinv(2, c);
    particle.C += outer(wv, dD);
  }

  // F <- (I + dt C) F
  __device__ __forceinline__ void g2p_finish_particle(MLS_APIC_Particle& particle, SimulationParameters const& par) {
    particle.F = (Mat::Identity() + par.dt * particle.C) * particle.F;
  }

 private:
  __device__ __forceinline__ void prepare_weights(MLS_APIC_Particle const& particle, SimulationParameters const& par,
                                                  InterpolationKernel const& interpolationKernel) {
    weights = interpolationKernel.weights_per_direction(particle.x, par.dx_inv, range_begin);
    if (InterpolationKernel::d_is_const()) Dinv = interpolationKernel.D_inv_const(par.dx_inv);
    else Dinv = interpolationKernel.D_inv(particle.x, range_begin, weights, par.dx);
  }
};
