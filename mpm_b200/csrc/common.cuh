// Shared device-side definitions: particle streams in HBM, kernel parameters.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>
#include <utility>

#include "../../include/mpm_b200.h"
#include "../../include/mpm_b200/InterpolationKernel.cuh"
#include "../../include/mpm_b200/MaterialModel.cuh"
#include "../../include/mpm_b200/TransferScheme.cuh"
#include "../../include/mpm_b200/linalg.cuh"
#include "../../include/mpm_b200/types.cuh"

namespace mpm {

using Particle = ::MLS_APIC_Particle;
using DefaultKernel = ::QuadraticInterpolationKernel;
using DefaultScheme = ::MLS_APIC_Scheme<DefaultKernel>;

// ---- particle streams in HBM ------------------------------------------------------------------
// 25 float streams per particle, stored TILE-MAJOR: particles are grouped in tiles of kTile = 256
// consecutive slots and a tile holds its 25 streams back to back, [tile][stream][lane].
//   * a CTA that works on one tile (the production P2G and G2P kernels) addresses every stream as
//     base + stream * 1 KB: one address computation per thread, immediate offsets after that;
//   * a warp reading one stream of 32 consecutive particles touches exactly one 128-byte line;
//   * any range of stream rows of a tile is ONE contiguous span, so a tile's inputs arrive in shared
//     memory with a single 1-D bulk copy (cp.async.bulk), no tensor map.
// Stream order: v(3), C(9) | x(3) | F(9), Jp.  G2P reads rows 12..24 (x, F, Jp) and writes all rows;
// P2G reads rows 0..24, or rows 0..14 (v, A, x) when G2P handed the affine matrix over (see SC).
// Plus u32 id (upload order, to un-permute on download) and u8 material per slot.
enum : int { SV = 0, SC = 3, SX = 12, SF = 15, SJ = 24, NSTREAM = 25 };
// Rows SC..SC+8 hold either the APIC matrix C (the particle state of the reference) or, between two
// substeps of one mpm_advance call, dx * affine = dx * (-Dinv dt vol P(F)F^T + m C) of the NEXT P2G,
// computed by G2P while F and C are in its registers ("hand-over").  Row-major [r][c] either way.
constexpr int kTile = 256;
constexpr int kTileFloats = NSTREAM * kTile;

struct Soa {
  float* f;         // capacity / kTile tiles of kTileFloats floats
  uint32_t* id;     // capacity
  uint8_t* mat;     // capacity
  size_t capacity;  // multiple of kTile
  // slot i of stream 0; stream s of the same particle is s * kTile floats further
  __host__ __device__ __forceinline__ float* col(size_t i) const { return f + (i >> 8) * (size_t)kTileFloats + (i & 255); }
  __host__ __device__ __forceinline__ float* tile(size_t t) const { return f + t * (size_t)kTileFloats; }
};
static_assert(kTile == 256, "Soa::col shifts by 8");

// replaces SimulationParameters (reference include/TransferScheme.h:6-29) on the device, plus the slab
struct KParams {
  float dt;
  float dx;      // (float)(1.0 / N)
  float dx_inv;  // (float)(1.0 / (double)dx)  -- not exactly N for non powers of two
  float dinv;    // (4 * dx_inv) * dx_inv : the diagonal of D^-1 (InterpolationKernel.cuh, D_inv_const)
  int N;
  int x0;   // first x-plane held in the local grid
  int nxl;  // x-planes held locally (owned + ghost)
  int x_own_begin, x_own_end;
  __host__ __device__ SimulationParameters par() const { return SimulationParameters(dt, (u32)N); }
};

// Small read-backs of the substep path (counters the host decides by) are STORED into pinned host memory by
// this kernel instead of copied by cudaMemcpyAsync: a device -> host copy of a few bytes would queue in the
// copy engine behind a 7 GB read-back of the particle state that is in flight on another stream
// (mpm_download_particles_aos_async) and stall the substeps for its whole duration.
static __global__ void __launch_bounds__(32) store_words_to_host_kernel(const uint32_t* __restrict__ src, volatile uint32_t* dst, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
  __threadfence_system();
}
inline cudaError_t readback_words(void* host_pinned, const void* dev, int n_words, cudaStream_t stream) {
  store_words_to_host_kernel<<<1, 32, 0, stream>>>(static_cast<const uint32_t*>(dev), static_cast<volatile uint32_t*>(host_pinned), n_words);
  return cudaGetLastError();
}

constexpr uint32_t kDeadId = 0xffffffffu;  // tombstone of a particle that migrated to another rank

// counters the kernels keep for the host (mpm_get_diagnostics); device memory, one per handle
struct DeviceDiag {
  unsigned int jp_not_one;      // uploaded particles with Jp != 1 (fixed-corotated handles skip the Jp stream while 0)
  unsigned int escaped;         // slab handles: particles whose stencil left the local grid between re-bins (mass lost)
  unsigned int nonfinite;       // particles with a non-finite position seen at the last re-bin
  unsigned int out_of_domain;   // particles whose stencil lies completely outside the domain at the last re-bin
};

// quadratic B-spline weights and base node (include/mpm_b200/InterpolationKernel.cuh); base uses C
// truncation like the reference's cast<int>()
__device__ __forceinline__ void bspline(float x, float dx_inv, int& base, float& fx, float w[3]) {
  const float g = x * dx_inv;
  base = (int)(g - 0.5f);
  fx = g - (float)base;
  const float d0 = 1.5f - fx, d1 = fx - 1.0f, d2 = fx - 0.5f;
  w[0] = 0.5f * (d0 * d0);
  w[1] = 0.75f - (d1 * d1);
  w[2] = 0.5f * (d2 * d2);
}

// the particle's whole stencil lies outside the domain: P2G and G2P leave it alone (reference
// src/mpm.cu:31-35, 128-132)
__device__ __forceinline__ bool stencil_outside(const int base[3], int N) {
  return base[0] + 3 < 0 || base[0] >= N || base[1] + 3 < 0 || base[1] >= N || base[2] + 3 < 0 || base[2] >= N;
}

// dx * affine of the MLS-APIC P2G (include/mpm_b200/TransferScheme.cuh, p2g_prepare_particle) for
// the constant-D quadratic kernel: dx * ((-Dinv dt vol) PF + m C).  Shared by the P2G kernel and by
// the G2P kernel that hands it over.
template <class Material>
__device__ __forceinline__ Mat p2g_affine_dx(const Particle& particle, const Material& material, const KParams& k) {
  const Mat PF = material.computePF(particle);
  const float kk = (((-k.dinv) * k.dt) * material.particleVolume) * k.dx;
  const float s_c = material.particleMass * k.dx;
  Mat A;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A.m[i][j] = fmaf(kk, PF.m[i][j], s_c * particle.C.m[i][j]);
  return A;
}

// the affine matrix from a stress the caller already has
template <class Material>
__device__ __forceinline__ Mat p2g_affine_dx_from_PF(const Mat& PF, const Particle& particle, const Material& material, const KParams& k) {
  const float kk = (((-k.dinv) * k.dt) * material.particleVolume) * k.dx;
  const float s_c = material.particleMass * k.dx;
  Mat A;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A.m[i][j] = fmaf(kk, PF.m[i][j], s_c * particle.C.m[i][j]);
  return A;
}

// does the material offer the rotation-sharing hooks (MaterialModel.cuh, MMSnow)?
template <class M, class = void>
struct HasRotationHooks : std::false_type {};
template <class M>
struct HasRotationHooks<M, std::void_t<decltype(std::declval<const M&>().endOfStepMutationR(std::declval<Particle&>(), std::declval<Mat&>())),
                                       decltype(std::declval<const M&>().computePF_R(std::declval<const Particle&>(), std::declval<const Mat&>()))>>
    : std::true_type {};

// Materials as the kernels see them.  Single-material handles (the common case) read the material
// from the kernel parameters (constant bank, no load latency: template flag ONE_MAT); the others
// index the device array by the particle's material_type like the reference (src/mpm.cu:23, 119).
template <class Material>
struct MatTable {
  Material one;         // material 0
  const Material* all;  // all n materials in device memory
  int n;
  template <bool ONE_MAT>
  __device__ __forceinline__ Material get(const uint8_t* __restrict__ mat_ids, size_t pi) const {
    if constexpr (ONE_MAT) return one;
    else return all[mat_ids[pi]];
  }
};

}  // namespace mpm
