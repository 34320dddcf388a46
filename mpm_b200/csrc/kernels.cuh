// The GENERIC P2G / G2P kernels: one thread per particle, written against the plugin concepts only (include/mpm_b200/*.cuh), so
// any MaterialModel / InterpolationKernel / TransferScheme tuple runs through them.
// Reference behaviour: src/mpm.cu:14-178.  The staged production kernels for the shipped tuple are
// p2g_sched.cuh and g2p_tile.cuh.
#pragma once
#include "common.cuh"
#include "f32x2.cuh"

namespace mpm {

constexpr int kParticleBlock = 128;

// ---- in-register particle view <-> streams -----------------------------------------------------
__device__ __forceinline__ Particle load_particle(const float* __restrict__ c, uint8_t material_type) {
  Particle q;
  q.material_type = material_type;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    q.x(a) = c[(SX + a) * kTile];
    q.v(a) = c[(SV + a) * kTile];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      q.F(r, cc) = c[(SF + 3 * r + cc) * kTile];
      q.C(r, cc) = c[(SC + 3 * r + cc) * kTile];
    }
  q.Jp = c[SJ * kTile];
  return q;
}
__device__ __forceinline__ void store_particle(const Particle& q, float* __restrict__ c) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    c[(SX + a) * kTile] = q.x(a);
    c[(SV + a) * kTile] = q.v(a);
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      c[(SF + 3 * r + cc) * kTile] = q.F(r, cc);
      c[(SC + 3 * r + cc) * kTile] = q.C(r, cc);
    }
  c[SJ * kTile] = q.Jp;
}

// ---- stage (2), generic: P2G -------------------------------------------------------------------
// One thread per particle, the loop nest of the reference (src/mpm.cu:14-74) over the plugin
// concepts.  Particles are in cell-sorted order, so a warp's reductions land on a handful of
// neighbouring nodes; each node is one aligned float4 -> a single red.global.add.v4.f32
// (REDG.E.ADD.F32x4) instead of the reference's four scalar atomics (src/mpm.cu:66-70).  Nodes outside
// the domain, or outside the x-planes this handle holds, are skipped.
template <class Material, class Kernel, class Scheme>
__global__ void __launch_bounds__(kParticleBlock)
p2g_generic_kernel(Soa p, size_t count, const Material* __restrict__ mats, float4* __restrict__ grid, KParams k, Kernel kernel) {
  const size_t pi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pi >= count) return;
  const Particle particle = load_particle(p.col(pi), p.mat[pi]);
  const Material material = mats[particle.material_type];
  const SimulationParameters par = k.par();
  Scheme ts;
  ts.p2g_prepare_particle(particle, par, kernel, material);
  const Veci rb = ts.get_range_begin();
  constexpr int S = (int)Kernel::size();
  for (int a = 0; a < 3; ++a)
    if (rb(a) + S < 0 || rb(a) >= k.N) return;  // completely outside the domain (src/mpm.cu:31-35)
  const long long NN = (long long)k.N * k.N;
  Vec dist;
#pragma unroll
  for (int i = 0; i < S; ++i) {
    const int gx = rb(0) + i;
    dist(0) = (real)gx * par.dx - particle.x(0);
#pragma unroll
    for (int j = 0; j < S; ++j) {
      const int gy = rb(1) + j;
      dist(1) = (real)gy * par.dx - particle.x(1);
#pragma unroll
      for (int kz = 0; kz < S; ++kz) {
        const int gz = rb(2) + kz;
        dist(2) = (real)gz * par.dx - particle.x(2);
        Vec4 cb;
        ts.p2g_node_contribution(particle, dist, material.particleMass, i, j, kz, cb);
        const bool ok = gx >= max(0, k.x0) && gx < min(k.N, k.x0 + k.nxl) && (unsigned)gy < (unsigned)k.N && (unsigned)gz < (unsigned)k.N;
        if (ok) atomicAdd(grid + ((long long)(gx - k.x0) * NN + (long long)gy * k.N + gz), make_float4(cb[0], cb[1], cb[2], cb[3]));
      }
    }
  }
}

// ---- stage (3): grid update -------------------------------------------------------------------
// momentum -> velocity, gravity, sticky walls / separating floor (reference src/mpm.cu:76-107).
// One float4 per thread, fully coalesced; mass is left untouched (the reference overwrites it
// with 1.0, SURVEY.md F8 — nothing downstream reads it).
__device__ __forceinline__ float4 grid_node_update(float4 c, int xi, int yi, int zi, const KParams& k) {
  if (c.w > 0.0f) {
    c.x = c.x / c.w;
    c.y = c.y / c.w;
    c.z = c.z / c.w;
    c.y += k.dt * -9.81f;
    const float boundary = 0.05f;
    const float hi = 1.0f - boundary;
    const float X = (float)xi / (float)k.N, Y = (float)yi / (float)k.N, Z = (float)zi / (float)k.N;
    if (X < boundary || X > hi || Y > hi || Z < boundary || Z > hi) {
      c.x = 0.f;
      c.y = 0.f;
      c.z = 0.f;
    }
    if (Y < boundary) c.y = fmaxf(0.0f, c.y);
  }
  return c;
}
// ---- stage (4), generic: G2P -------------------------------------------------------------------
// One thread per particle, the loop nest of the reference (src/mpm.cu:109-178) over the plugin
// concepts: gather, F update, plasticity, advection x += dt v (no position clamp).
template <class Material, class Kernel, class Scheme>
__global__ void __launch_bounds__(kParticleBlock)
g2p_generic_kernel(Soa p, size_t count, const Material* __restrict__ mats, const float4* __restrict__ grid, KParams k, Kernel kernel) {
  const size_t pi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pi >= count) return;
  float* col = p.col(pi);
  Particle particle = load_particle(col, p.mat[pi]);
  const Material material = mats[particle.material_type];
  const SimulationParameters par = k.par();
  Scheme ts;
  ts.g2p_prepare_particle(particle, par, kernel);
  const Veci rb = ts.get_range_begin();
  constexpr int S = (int)Kernel::size();
  for (int a = 0; a < 3; ++a)
    if (rb(a) + S < 0 || rb(a) >= k.N) return;  // untouched, like the reference's early return: nothing was stored yet
  const long long NN = (long long)k.N * k.N;
  Vec dist;
#pragma unroll
  for (int i = 0; i < S; ++i) {
    const int gx = rb(0) + i;
    dist(0) = (real)gx * par.dx - particle.x(0);
#pragma unroll
    for (int j = 0; j < S; ++j) {
      const int gy = rb(1) + j;
      dist(1) = (real)gy * par.dx - particle.x(1);
#pragma unroll
      for (int kz = 0; kz < S; ++kz) {
        const int gz = rb(2) + kz;
        dist(2) = (real)gz * par.dx - particle.x(2);
        const bool ok = gx >= max(0, k.x0) && gx < min(k.N, k.x0 + k.nxl) && (unsigned)gy < (unsigned)k.N && (unsigned)gz < (unsigned)k.N;
        if (!ok) continue;
        const float4 g = __ldg(grid + ((long long)(gx - k.x0) * NN + (long long)gy * k.N + gz));
        ts.g2p_node_contribution(particle, dist, Vec4{{g.x, g.y, g.z, g.w}}, i, j, kz);
      }
    }
  }
  ts.g2p_finish_particle(particle, par);
  material.endOfStepMutation(particle);
  particle.x = particle.x + par.dt * particle.v;
  store_particle(particle, col);
}

}  // namespace mpm
