// Shared device-side definitions: SoA particle streams, kernel parameters, material math.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpm_b200.h"
#include "svd3.cuh"

namespace mpm {

// ---- particle streams in HBM ------------------------------------------------------------------
// 25 float streams (x3, F9 row-major, Jp, v3, C9 row-major), each `stride` floats long and
// 128-byte aligned, so a warp reading stream s for 32 consecutive particles touches exactly one
// 128 B line.  Seen as a 2-D tensor [NSTREAM][stride], what G2P reads (x, F, Jp) is rows 0..12 and
// what P2G reads is rows 0..24: one TMA box each.  Plus u32 id (upload order, for un-permuting on
// download) and u8 material.
enum : int { SX = 0, SF = 3, SJ = 12, SV = 13, SC = 16, NSTREAM = 25 };

struct Soa {
  float* f;       // NSTREAM * stride floats
  uint32_t* id;   // stride
  uint8_t* mat;   // stride
  size_t stride;  // multiple of 32
  __host__ __device__ __forceinline__ float* s(int stream) const { return f + (size_t)stream * stride; }
};

// replaces SimulationParameters (reference include/TransferScheme.h:6-29) on the device
struct KParams {
  float dt;
  float dx;      // (float)(1.0 / N)
  float dx_inv;  // (float)(1.0 / (double)dx)  -- not exactly N for non powers of two
  float dinv;    // (4 * dx_inv) * dx_inv : the diagonal of D^-1 (InterpolationKernel.cuh:71-73)
  int N;
  int x0;   // first x-plane held in the local grid
  int nxl;  // x-planes held locally (owned + ghost)
  int x_own_begin, x_own_end;
};

constexpr uint32_t kDeadId = 0xffffffffu;  // tombstone of a particle that migrated to another rank

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fmaxf(fminf(x, hi), lo); }

// quadratic B-spline weights and base node (reference include/InterpolationKernel.cuh:57-69);
// base uses C truncation like the reference's cast<int>()
__device__ __forceinline__ void bspline(float x, float dx_inv, int& base, float& fx, float w[3]) {
  const float g = x * dx_inv;
  base = (int)(g - 0.5f);
  fx = g - (float)base;
  const float d0 = 1.5f - fx, d1 = fx - 1.0f, d2 = fx - 0.5f;
  w[0] = 0.5f * (d0 * d0);
  w[1] = 0.75f - (d1 * d1);
  w[2] = 0.5f * (d2 * d2);
}

__device__ __forceinline__ MpmMaterial load_material(const MpmMaterial* __restrict__ mats, int idx) {
  const float* p = reinterpret_cast<const float*>(mats + idx);
  MpmMaterial m;
  m.particleVolume = __ldg(p + 0);
  m.particleMass = __ldg(p + 1);
  m.mu0 = __ldg(p + 2);
  m.lambda0 = __ldg(p + 3);
  m.hardening = __ldg(p + 4);
  m.plast_clamp_lower = __ldg(p + 5);
  m.plast_clamp_higher = __ldg(p + 6);
  return m;
}

// P(F) F^T of the fixed-corotated model with snow hardening
// (reference MMSnow::computePF, include/MaterialModel.cuh:85-93; MMFixedCorotated :56-61).
// "J" is the plastic scalar Jp, as in the reference.  EXACT evaluates exp and the lambda term
// in double like the reference; FAST stays in f32 and skips exp when hardening == 0.
template <int MODEL, class O, bool EXACT>
__device__ __forceinline__ Mat3 compute_PF(const Mat3& F, float Jp, const MpmMaterial& m) {
  Mat3 R;
  if constexpr (EXACT) R = polar_rotation<O>(F); else R = polar_rotation_newton(F);
  float mu = m.mu0, lambda = m.lambda0;
  if (MODEL == MPM_MODEL_SNOW) {
    float e;
    if (EXACT) {
      e = (float)exp((double)m.hardening * (1.0 - (double)Jp));
    } else {
      e = (m.hardening == 0.0f) ? 1.0f : __expf(m.hardening * (1.0f - Jp));
    }
    mu *= e;
    lambda *= e;
  }
  const float two_mu = 2.0f * mu;
  float lam_term;
  if (EXACT) {
    lam_term = (float)((double)lambda * (((double)Jp - 1.0) * (double)Jp));
  } else {
    lam_term = lambda * ((Jp - 1.0f) * Jp);
  }
  Mat3 D;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) D.m[i][j] = two_mu * (F.m[i][j] - R.m[i][j]);
  Mat3 PF = mul_abt(D, F);
#pragma unroll
  for (int i = 0; i < 3; ++i) PF.m[i][i] += lam_term;
  return PF;
}

// true when every singular value of F lies strictly inside (lo, hi) and det F > 0, i.e. when the
// clamp of MMSnow::endOfStepMutation changes nothing: with C = F^T F, both C - lo^2 I and
// hi^2 I - C are positive definite (Sylvester's criterion, three leading minors each).
__device__ __forceinline__ bool pd3(float a11, float a12, float a13, float a22, float a23, float a33) {
  const float m2 = a11 * a22 - a12 * a12;
  const float det = a11 * (a22 * a33 - a23 * a23) - a12 * (a12 * a33 - a13 * a23) + a13 * (a12 * a23 - a13 * a22);
  return a11 > 0.0f && m2 > 0.0f && det > 0.0f;
}
__device__ __forceinline__ bool snow_within_elastic_range(const Mat3& F, float lo, float hi) {
  if (!(det3(F) > 0.0f)) return false;
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

// snow plasticity (reference MMSnow::endOfStepMutation, include/MaterialModel.cuh:95-114).
// FAST mode skips the SVD for particles inside the elastic range: there the reference only
// re-synthesises F = U S V^T and Jp * det F / det F from the SVD's own round-off (~1e-6), so
// leaving F and Jp untouched is within the FAST tolerance (tests/test_gpu_substep.py); EXACT mode
// always runs the full sequence, bit for bit.
template <class O>
__device__ __forceinline__ void snow_plasticity(Mat3& F, float& Jp, const MpmMaterial& m) {
  if constexpr (!O::kExact) {
    if (snow_within_elastic_range(F, m.plast_clamp_lower, m.plast_clamp_higher)) {
      Jp = clampf(Jp, 0.6f, 20.0f);  // the outer clamp of the Jp update still applies
      return;
    }
  }
  Mat3 U, V;
  float sig[3];
  svd3<O>(F, U, sig, V);
#pragma unroll
  for (int i = 0; i < 3; ++i) sig[i] = clampf(sig[i], m.plast_clamp_lower, m.plast_clamp_higher);
  const float oldJ = det3(F);
  Mat3 US;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) US.m[i][j] = U.m[i][j] * sig[j];
  F = mul_abt(US, V);
  const float Fdet = det3(F);
  Jp = clampf(Jp * oldJ / Fdet, 0.6f, 20.0f);
}

}  // namespace mpm
