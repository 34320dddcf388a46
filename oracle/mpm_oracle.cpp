// TEST INFRASTRUCTURE ONLY (oracle/).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this; the product path (mpm_b200/) never does.
//
// CPU restatement ("port") of the reference's MLS-MPM substep, one function per reference
// function, each citing the /root/reference file:line it follows.  The reference has no CPU
// substep (its P2G / grid update / G2P exist only as CUDA kernels, src/mpm.cu:14-178), so this
// file transcribes the *arithmetic* of those kernels and of the plugin headers into plain
// scalar C++ with the same operation order, evaluated without FMA contraction
// (-ffp-contract=off).  Parity pins (see tests/test_oracle.py):
//   * svd3 restatement: bit-exact against the reference's own header compiled for the host
//     (oracle/_ref/libref_svd3.so) on 10^6+ random / degenerate / inverted matrices, and
//     against the two gtest matrices of tests/test_linalg.cu:27,33 (L1 bounds :10-14, :42-46).
//   * substep: checked against the reference's own plugin headers + kernel bodies compiled
//     through a test-only Eigen shim (oracle/_ref/libref_mpm.so) where that library exists.
// Two documented deviations from the reference source: the G2P guard uses
// the particle count (reference: N^3, src/mpm.cu:113-114) and the P2G upper stencil clip is
// N - base (reference: N + base, src/mpm.cu:42-44).  Neither is reachable in the test scenes.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// Layouts crossing the boundary
// ---------------------------------------------------------------------------------------------
// MLS_APIC_Particle, include/types.h:24-35 + include/TransferScheme.h:46-54: 104 B, matrices
// column-major (Eigen default).
struct Particle {
  uint8_t material_type;
  uint8_t pad_[3];
  float x[3];
  float v[3];
  float F[9];  // F(r,c) = F[3*c + r]
  float C[9];
  float Jp;
};
static_assert(sizeof(Particle) == 104, "reference particle is 104 bytes");

// MMSnow<Particle>, include/MaterialModel.cuh:23-24,47-48,70-72: 7 floats.
struct Material {
  float particleVolume, particleMass, mu0, lambda0, hardening, plast_clamp_lower, plast_clamp_higher;
};
static_assert(sizeof(Material) == 28, "reference material is 28 bytes");

// SimulationParameters, include/TransferScheme.h:6-29.
struct Params {
  float dt;
  uint32_t N;
  float N_real, dx, dx_inv;
  Params(float dt_, uint32_t N_) : dt(dt_), N(N_), N_real((float)N_) {
    dx = (float)(1.0 / (double)N_);    // dx(1.0/(N)), :26
    dx_inv = (float)(1.0 / (double)dx);  // dx_inv(1.0/dx), :27
  }
};

struct M3 {
  float a[3][3];  // row-major a[r][c]
};

inline M3 load_colmajor(const float* p) {
  M3 m;
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) m.a[r][c] = p[3 * c + r];
  return m;
}
inline void store_colmajor(const M3& m, float* p) {
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) p[3 * c + r] = m.a[r][c];
}
// Eigen fixed-size coefficient product: ((a0*b0 + a1*b1) + a2*b2), no contraction on the host.
inline float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
  float p0 = a0 * b0, p1 = a1 * b1, p2 = a2 * b2;
  float s = p0 + p1;
  return s + p2;
}
inline M3 mul(const M3& A, const M3& B) {
  M3 R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      R.a[i][j] = dot3(A.a[i][0], B.a[0][j], A.a[i][1], B.a[1][j], A.a[i][2], B.a[2][j]);
  return R;
}
inline M3 mul_bt(const M3& A, const M3& B) {  // A * B^T
  M3 R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      R.a[i][j] = dot3(A.a[i][0], B.a[j][0], A.a[i][1], B.a[j][1], A.a[i][2], B.a[j][2]);
  return R;
}

// ---------------------------------------------------------------------------------------------
// svd3 — restatement of include/svd3_cuda.h:35-1043 (McAdams et al. TR1690).
// The reference spells every step out on 30+ scalar unions; here the three Jacobi conjugations
// and the three QR Givens steps are one routine each, applied with rotated roles.  Every float
// operation is performed in the same order on the same operands as the reference, so results
// are bit-identical (pinned by tests/test_oracle.py against oracle/_ref).
// ---------------------------------------------------------------------------------------------
const float kTiny = 1.e-20f;                          // gtiny_number, :30
const float kSmall = 1.e-12f;                         // gsmall_number, :29
const float kFourGammaSquared = 5.8284273147583007813f;  // :31
const uint32_t kSinPi8Bits = 1053028117u;             // gsine_pi_over_eight, :26
const uint32_t kCosPi8Bits = 1064076127u;             // gcosine_pi_over_eight, :27

inline float bits_to_float(uint32_t u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline float rsqrt_rn(float x) { return (float)(1.0 / std::sqrt((double)x)); }  // __frsqrt_rn

// One approximate-Givens Jacobi conjugation on the symmetric matrix S (lower triangle) for the
// rotation plane whose off-diagonal entry is s21; s3x are the entries coupling to the third
// axis.  Quaternion roles (qx,qy,qz) rotate with the plane.  svd3_cuda.h:106-199 (plane 1-2),
// :204-303 (2-3, roles rotated once), :309-405 (3-1, rotated twice).
inline void jacobi_conjugation(float& s11, float& s21, float& s31, float& s22, float& s32, float& s33,
                               float& qx, float& qy, float& qz, float& qs) {
  float sh = s21 * 0.5f;
  float t5 = s11 - s22;
  float t2 = sh * sh;
  bool nz = (t2 >= kTiny);
  sh = nz ? sh : 0.0f;
  float ch = nz ? t5 : 1.0f;
  float t1 = sh * sh;
  t2 = ch * ch;
  float t3 = t1 + t2;
  float t4 = rsqrt_rn(t3);
  sh = t4 * sh;
  ch = t4 * ch;
  t1 = kFourGammaSquared * t1;
  bool big = (t2 <= t1);
  sh = big ? bits_to_float(kSinPi8Bits) : sh;
  ch = big ? bits_to_float(kCosPi8Bits) : ch;
  t1 = sh * sh;
  t2 = ch * ch;
  float c = t2 - t1;
  float s = ch * sh;
  s = s + s;
  // conjugation Q^T S Q (:144-172)
  t3 = t1 + t2;
  s33 = s33 * t3;
  s31 = s31 * t3;
  s32 = s32 * t3;
  s33 = s33 * t3;
  t1 = s * s31;
  t2 = s * s32;
  s31 = c * s31;
  s32 = c * s32;
  s31 = t2 + s31;
  s32 = s32 - t1;
  t2 = s * s;
  t1 = s22 * t2;
  t3 = s11 * t2;
  t4 = c * c;
  s11 = s11 * t4;
  s22 = s22 * t4;
  s11 = s11 + t1;
  s22 = s22 + t3;
  t4 = t4 - t2;
  t2 = s21 + s21;
  s21 = s21 * t4;
  t4 = c * s;
  t2 = t2 * t4;
  t5 = t5 * t4;
  s11 = s11 + t2;
  s21 = s21 - t5;
  s22 = s22 - t2;
  // cumulative rotation in quaternion form (:184-197)
  t1 = sh * qx;
  t2 = sh * qy;
  t3 = sh * qz;
  sh = sh * qs;
  qs = ch * qs;
  qx = ch * qx;
  qy = ch * qy;
  qz = ch * qz;
  qz = qz + sh;
  qs = qs - t3;
  qx = qx + t2;
  qy = qy - t1;
}

// rsqrt with one Newton step as the reference writes it (:426-432 and the QR steps).
inline float rsqrt_refined(float t2) {
  float t1 = rsqrt_rn(t2);
  float t4 = t1 * 0.5f;
  float t3 = t1 * t4;
  t3 = t1 * t3;
  t3 = t2 * t3;
  t1 = t1 + t4;
  t1 = t1 - t3;
  return t1;
}

// One Givens step of the QR factorisation, :719-817 (pivot a11/a21), :821-919, :923-1021.
// Computes (c, s) from pivot app and sub-diagonal aqp.
inline void qr_givens(float app, float aqp, float& c, float& s) {
  float sh = aqp * aqp;
  sh = (sh >= kSmall) ? aqp : 0.0f;
  float t5 = 0.0f;
  float ch = t5 - app;
  ch = std::fmax(ch, app);
  ch = std::fmax(ch, kSmall);
  bool pos = (app >= t5);
  float t1 = ch * ch;
  float t2 = sh * sh;
  t2 = t1 + t2;
  t1 = rsqrt_refined(t2);
  t1 = t1 * t2;
  ch = ch + t1;
  if (!pos) std::swap(ch, sh);
  t1 = ch * ch;
  t2 = sh * sh;
  t2 = t1 + t2;
  t1 = rsqrt_refined(t2);
  ch = ch * t1;
  sh = sh * t1;
  c = ch * ch;
  s = sh * sh;
  c = c - s;
  s = sh * ch;
  s = s + s;
}
inline void rot_pair(float c, float s, float& p, float& q) {  // p' = c p + s q ; q' = c q - s p
  float t1 = s * p;
  float t2 = s * q;
  p = c * p;
  q = c * q;
  p = p + t2;
  q = q - t1;
}

// A = U * diag(S) * V^T.  All matrices row-major a[r][c].
void svd3(const M3& Ain, M3& U, float S[3], M3& V) {
  float a[3][3];
  std::memcpy(a, Ain.a, sizeof(a));
  // normal equations S = A^T A, lower triangle (:63-97)
  auto ata = [&](int i, int j) {
    float r = a[0][i] * a[0][j];
    float t = a[1][i] * a[1][j];
    r = t + r;
    t = a[2][i] * a[2][j];
    r = t + r;
    return r;
  };
  float s11 = ata(0, 0), s21 = ata(1, 0), s31 = ata(2, 0), s22 = ata(1, 1), s32 = ata(2, 1), s33 = ata(2, 2);
  float qs = 1.f, qx = 0.f, qy = 0.f, qz = 0.f;
  for (int sweep = 0; sweep < 4; ++sweep) {  // :104
    jacobi_conjugation(s11, s21, s31, s22, s32, s33, qx, qy, qz, qs);
    jacobi_conjugation(s22, s32, s21, s33, s31, s11, qy, qz, qx, qs);
    jacobi_conjugation(s33, s31, s32, s11, s21, s22, qz, qx, qy, qs);
  }
  // normalise quaternion (:417-437)
  float t2 = qs * qs;
  float t1 = qx * qx;
  t2 = t1 + t2;
  t1 = qy * qy;
  t2 = t1 + t2;
  t1 = qz * qz;
  t2 = t1 + t2;
  t1 = rsqrt_refined(t2);
  qs = qs * t1;
  qx = qx * t1;
  qy = qy * t1;
  qz = qz * t1;
  // quaternion -> V (:443-468)
  float v[3][3];
  {
    float x2 = qx * qx, y2 = qy * qy, z2 = qz * qz;
    float v11 = qs * qs;
    float v22 = v11 - x2;
    float v33 = v22 - y2;
    v33 = v33 + z2;
    v22 = v22 + y2;
    v22 = v22 - z2;
    v11 = v11 + x2;
    v11 = v11 - y2;
    v11 = v11 - z2;
    float dx2 = qx + qx, dy2 = qy + qy, dz2 = qz + qz;
    float v32 = qs * dx2;
    float v13 = qs * dy2;
    float v21 = qs * dz2;
    float p1 = qy * dx2;
    float p2 = qz * dy2;
    float p3 = qx * dz2;
    float v12 = p1 - v21;
    float v23 = p2 - v32;
    float v31 = p3 - v13;
    v21 = p1 + v21;
    v32 = p2 + v32;
    v13 = p3 + v13;
    v[0][0] = v11; v[0][1] = v12; v[0][2] = v13;
    v[1][0] = v21; v[1][1] = v22; v[1][2] = v23;
    v[2][0] = v31; v[2][1] = v32; v[2][2] = v33;
  }
  // B = A * V (:474-526), row by row
  for (int r = 0; r < 3; ++r) {
    float a1 = a[r][0], a2 = a[r][1], a3 = a[r][2];
    float b1 = v[0][0] * a1;
    float b2 = v[0][1] * a1;
    float b3 = v[0][2] * a1;
    float t = v[1][0] * a2;
    b1 = b1 + t;
    t = v[2][0] * a3;
    b1 = b1 + t;
    t = v[1][1] * a2;
    b2 = b2 + t;
    t = v[2][1] * a3;
    b2 = b2 + t;
    t = v[1][2] * a2;
    b3 = b3 + t;
    t = v[2][2] * a3;
    b3 = b3 + t;
    a[r][0] = b1; a[r][1] = b2; a[r][2] = b3;
  }
  // sort singular values by column norm (:532-707)
  auto colnorm = [&](int c) {
    float r = a[0][c] * a[0][c];
    float t = a[1][c] * a[1][c];
    r = r + t;
    t = a[2][c] * a[2][c];
    r = r + t;
    return r;
  };
  float n1 = colnorm(0), n2 = colnorm(1), n3 = colnorm(2);
  auto swap_cols = [&](int ca, int cb, int cneg, float& na, float& nb) {
    bool sw = na < nb;
    if (sw) {
      for (int r = 0; r < 3; ++r) {
        std::swap(a[r][ca], a[r][cb]);
        std::swap(v[r][ca], v[r][cb]);
      }
      std::swap(na, nb);
    }
    float f = sw ? -2.f : 0.f;  // :588-601: 1 + (-2 & mask)
    f = 1.f + f;
    for (int r = 0; r < 3; ++r) {
      a[r][cneg] = a[r][cneg] * f;
      v[r][cneg] = v[r][cneg] * f;
    }
  };
  swap_cols(0, 1, 1, n1, n2);
  swap_cols(0, 2, 0, n1, n3);
  swap_cols(1, 2, 2, n2, n3);
  // QR by three Givens rotations (:713-1021)
  float u[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
  float c, s;
  qr_givens(a[0][0], a[1][0], c, s);
  for (int j = 0; j < 3; ++j) rot_pair(c, s, a[0][j], a[1][j]);
  for (int i = 0; i < 3; ++i) rot_pair(c, s, u[i][0], u[i][1]);
  qr_givens(a[0][0], a[2][0], c, s);
  for (int j = 0; j < 3; ++j) rot_pair(c, s, a[0][j], a[2][j]);
  for (int i = 0; i < 3; ++i) rot_pair(c, s, u[i][0], u[i][2]);
  qr_givens(a[1][1], a[2][1], c, s);
  for (int j = 0; j < 3; ++j) rot_pair(c, s, a[1][j], a[2][j]);
  for (int i = 0; i < 3; ++i) rot_pair(c, s, u[i][1], u[i][2]);
  std::memcpy(U.a, u, sizeof(u));
  std::memcpy(V.a, v, sizeof(v));
  S[0] = a[0][0];
  S[1] = a[1][1];
  S[2] = a[2][2];
}

// linalg::determinant, src/linalg.cu:47-52
inline float determinant(const M3& M) {
  float sub1 = M.a[1][0] * M.a[2][1] - M.a[1][1] * M.a[2][0];
  float sub2 = M.a[1][0] * M.a[2][2] - M.a[1][2] * M.a[2][0];
  float sub3 = M.a[1][1] * M.a[2][2] - M.a[1][2] * M.a[2][1];
  float r = M.a[0][0] * sub3 - M.a[0][1] * sub2;
  return r + M.a[0][2] * sub1;
}

// linalg::polar_decomposition_device, src/linalg.cu:18-33: R = U V^T, S = V Z V^T.
inline void polar(const M3& A, M3& R, M3& Sym) {
  M3 U, V;
  float sig[3];
  svd3(A, U, sig, V);
  R = mul_bt(U, V);
  M3 VZ;  // V * Z with Z a full matrix holding the diagonal: exact zeros elsewhere
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) VZ.a[i][j] = V.a[i][j] * sig[j];
  Sym = mul_bt(VZ, V);
}

inline float clampf(float x, float lo, float hi) { return std::fmax(std::fmin(x, hi), lo); }  // MaterialModel.cuh:14-16

// QuadraticInterpolationKernel::weights_per_direction, include/InterpolationKernel.cuh:57-69.
// w[axis][node], base = truncated (x*dx_inv - 0.5).
inline void weights_per_direction(const float x[3], float dx_inv, int base[3], float w[3][3]) {
  for (int a = 0; a < 3; ++a) {
    float g = x[a] * dx_inv;
    base[a] = (int)(g - 0.5f);
    float fx = g - (float)base[a];
    float d0 = 1.5f - fx, d1 = fx - 1.0f, d2 = fx - 0.5f;
    w[a][0] = 0.5f * (d0 * d0);
    w[a][1] = 0.75f - (d1 * d1);
    w[a][2] = 0.5f * (d2 * d2);
  }
}
// D_inv_const, InterpolationKernel.cuh:71-73: Identity * 4.0 * dx_inv * dx_inv (scalar promoted to f32).
inline float dinv_scalar(float dx_inv) { return (4.0f * dx_inv) * dx_inv; }

enum ModelKind { kSnow = 0, kFixedCorotated = 1, kJelly = 2 };

// MMSnow::computePF, include/MaterialModel.cuh:85-93; MMFixedCorotated::computePF, :56-61;
// MMJelly::computePF, :133-140 (the same expression as MMSnow's).
inline M3 computePF(const M3& F, float Jp, const Material& m, int kind) {
  M3 R, Sym;
  polar(F, R, Sym);
  float mu = m.mu0, lambda = m.lambda0;
  if (kind == kSnow || kind == kJelly) {
    float e = (float)std::exp((double)m.hardening * (1.0 - (double)Jp));  // :88, :136
    mu = m.mu0 * e;
    lambda = m.lambda0 * e;
  }
  float two_mu = (float)(2.0 * (double)mu);
  float lam_term = (float)((double)lambda * (((double)Jp - 1.0) * (double)Jp));
  M3 D;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) D.a[i][j] = two_mu * (F.a[i][j] - R.a[i][j]);
  M3 PF = mul_bt(D, F);
  for (int i = 0; i < 3; ++i) PF.a[i][i] = PF.a[i][i] + lam_term;
  return PF;
}

// MMSnow::endOfStepMutation, include/MaterialModel.cuh:95-114 (no-op for MMFixedCorotated, :63);
// MMJelly::endOfStepMutation, :142-149: Jp * det F / det F of the SAME F, then the clamp.
inline void endOfStepMutation(M3& F, float& Jp, const Material& m, int kind) {
  if (kind == kJelly) {
    float oldJ = determinant(F);
    Jp = clampf(Jp * oldJ / determinant(F), 0.6f, 20.0f);
    return;
  }
  if (kind != kSnow) return;
  M3 U, V;
  float sig[3];
  svd3(F, U, sig, V);
  for (int i = 0; i < 3; ++i) sig[i] = clampf(sig[i], m.plast_clamp_lower, m.plast_clamp_higher);
  float oldJ = determinant(F);
  M3 US;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) US.a[i][j] = U.a[i][j] * sig[j];
  F = mul_bt(US, V);
  float Fdet = determinant(F);
  Jp = clampf(Jp * oldJ / Fdet, 0.6f, 20.0f);
}

inline bool outside(const int base[3], int N) {  // src/mpm.cu:31-35, :127-131
  for (int a = 0; a < 3; ++a)
    if (base[a] + 3 < 0 || base[a] >= N) return true;
  return false;
}

// particleToGrid, src/mpm.cu:14-74 + MLS_APIC_Scheme::p2g_prepare_particle / p2g_node_contribution,
// include/TransferScheme.h:66-100.  grid: N^3 float4 (px,py,pz,m), idx = N*N*i + N*j + k.
// magnitudes: accumulate w (|m v| + sum_k |A_ck d_k|) instead (the per-node error scale of SURVEY.md 8(c)(3))
void p2g(const Particle* ps, size_t count, const Material* mats, const Params& par, int kind, float* grid, bool magnitudes = false) {
  const int N = (int)par.N;
  const float dinv = dinv_scalar(par.dx_inv);
#pragma omp parallel for schedule(static)
  for (long long pi = 0; pi < (long long)count; ++pi) {
    const Particle& p = ps[pi];
    const Material& m = mats[p.material_type];
    int base[3];
    float w[3][3];
    weights_per_direction(p.x, par.dx_inv, base, w);
    M3 F = load_colmajor(p.F), C = load_colmajor(p.C);
    M3 PF = computePF(F, p.Jp, m, kind);
    float k = ((-dinv) * par.dt) * m.particleVolume;  // -Dinv * dt * vol, TransferScheme.h:83
    M3 A;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) A.a[i][j] = k * PF.a[i][j] + m.particleMass * C.a[i][j];  // :85
    if (outside(base, N)) continue;
    int lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::max(0, -base[a]);
      hi[a] = std::min(3, N - base[a]);  // F7: reference writes N + base
    }
    for (int i = lo[0]; i < hi[0]; ++i) {
      unsigned ig = (unsigned)(base[0] + i);
      float d0 = (float)ig * par.dx - p.x[0];
      for (int j = lo[1]; j < hi[1]; ++j) {
        unsigned jg = (unsigned)(base[1] + j);
        float d1 = (float)jg * par.dx - p.x[1];
        for (int kk = lo[2]; kk < hi[2]; ++kk) {
          unsigned kg = (unsigned)(base[2] + kk);
          float d2 = (float)kg * par.dx - p.x[2];
          float weight = w[0][i] * w[1][j] * w[2][kk];
          float out[4];
          for (int c = 0; c < 3; ++c) {
            float ad = dot3(A.a[c][0], d0, A.a[c][1], d1, A.a[c][2], d2);
            out[c] = weight * (p.v[c] * m.particleMass + ad);
          }
          out[3] = weight * m.particleMass;
          if (magnitudes)  // the terms' magnitudes: m v and A d may cancel, a rounding error does not
            for (int c = 0; c < 3; ++c)
              out[c] = weight * (std::fabs(p.v[c] * m.particleMass) + std::fabs(A.a[c][0] * d0) + std::fabs(A.a[c][1] * d1) + std::fabs(A.a[c][2] * d2));
          float* cell = grid + 4 * ((size_t)N * N * ig + (size_t)N * jg + kg);
          for (int c = 0; c < 4; ++c) {
#pragma omp atomic
            cell[c] += out[c];
          }
        }
      }
    }
  }
}

// gridOpKernel, src/mpm.cu:76-107.
void grid_update(float* grid, const Params& par, int nx, int x_offset) {
  const int N = (int)par.N;
  const float gravity = -9.81f;  // src/mpm.cu:6
  const float boundary = 0.05f;  // :92
  const float hi = 1 - boundary;
#pragma omp parallel for schedule(static)
  for (long long idx = 0; idx < (long long)nx * N * N; ++idx) {
    int xi = (int)(idx / ((long long)N * N)) + x_offset;
    int yi = (int)((idx / N) % N);
    int zi = (int)(idx % N);
    float* cell = grid + 4 * idx;
    if (cell[3] > 0.0f) {
      float m = cell[3];
      cell[0] /= m;
      cell[1] /= m;
      cell[2] /= m;
      cell[3] /= cell[3];  // F8: mass <- 1, kept for fidelity
      cell[1] += par.dt * gravity;
      float x = (float)xi / N, y = (float)yi / N, z = (float)zi / N;
      if (x < boundary || x > hi || y > hi || z < boundary || z > hi) {
        cell[0] = 0.f;
        cell[1] = 0.f;
        cell[2] = 0.f;
      }
      if (y < boundary) cell[1] = std::fmax(0.0f, cell[1]);
    }
  }
}

// gridToParticle, src/mpm.cu:109-178 + g2p_prepare_particle / g2p_node_contribution /
// g2p_finish_particle, include/TransferScheme.h:102-142.
void g2p(const float* grid, Particle* ps, size_t count, const Material* mats, const Params& par, int kind) {
  const int N = (int)par.N;
  const float dinv = dinv_scalar(par.dx_inv);
#pragma omp parallel for schedule(static)
  for (long long pi = 0; pi < (long long)count; ++pi) {  // F6: reference guards with N^3
    Particle p = ps[pi];
    const Material& m = mats[p.material_type];
    int base[3];
    float w[3][3];
    weights_per_direction(p.x, par.dx_inv, base, w);
    float v[3] = {0.f, 0.f, 0.f};
    M3 C;
    std::memset(&C, 0, sizeof(C));
    if (outside(base, N)) continue;
    int lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::max(0, -base[a]);
      hi[a] = std::min(3, N - base[a]);
    }
    for (int i = lo[0]; i < hi[0]; ++i) {
      unsigned ig = (unsigned)(base[0] + i);
      float d0 = (float)ig * par.dx - p.x[0];
      for (int j = lo[1]; j < hi[1]; ++j) {
        unsigned jg = (unsigned)(base[1] + j);
        float d1 = (float)jg * par.dx - p.x[1];
        for (int kk = lo[2]; kk < hi[2]; ++kk) {
          unsigned kg = (unsigned)(base[2] + kk);
          float d2 = (float)kg * par.dx - p.x[2];
          const float* cell = grid + 4 * ((size_t)N * N * ig + (size_t)N * jg + kg);
          float weight = w[0][i] * w[1][j] * w[2][kk];
          float wv[3] = {weight * cell[0], weight * cell[1], weight * cell[2]};
          float r[3] = {d0 * dinv, d1 * dinv, d2 * dinv};  // dist^T * Dinv (diagonal)
          for (int c = 0; c < 3; ++c) {
            v[c] += wv[c];
            for (int q = 0; q < 3; ++q) C.a[c][q] += wv[c] * r[q];
          }
        }
      }
    }
    // F <- (I + dt*C) * F, TransferScheme.h:141
    M3 G;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) G.a[i][j] = (i == j ? 1.0f : 0.0f) + par.dt * C.a[i][j];
    M3 F = mul(G, load_colmajor(p.F));
    float Jp = p.Jp;
    endOfStepMutation(F, Jp, m, kind);
    for (int c = 0; c < 3; ++c) {
      p.v[c] = v[c];
      p.x[c] += par.dt * v[c];  // src/mpm.cu:175
    }
    store_colmajor(F, p.F);
    store_colmajor(C, p.C);
    p.Jp = Jp;
    ps[pi] = p;
  }
}

}  // namespace

extern "C" {

void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int oracle_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// MaterialModelBase / MMFixedCorotated / MMSnow constructors, include/MaterialModel.cuh:26-29,
// :50-54, :74-84, with the argument conversions of src/main.cu:36-42 (doubles -> real).
void oracle_make_material(double volume, double density, double E, double Nu, double hardening, double clamp_lo,
                          double clamp_hi, float* out7) {
  float vol = (float)volume, rho = (float)density, e = (float)E, nu = (float)Nu;
  Material m;
  m.particleVolume = vol;
  m.particleMass = rho * vol;
  m.mu0 = e / (2 * (1 + nu));
  m.lambda0 = e * nu / ((1 + nu) * (1 - 2 * nu));
  m.hardening = (float)hardening;
  m.plast_clamp_lower = (float)clamp_lo;
  m.plast_clamp_higher = (float)clamp_hi;
  std::memcpy(out7, &m, sizeof(m));
}

void oracle_params(float dt, uint32_t N, float* dx, float* dx_inv) {
  Params p(dt, N);
  *dx = p.dx;
  *dx_inv = p.dx_inv;
}

// row-major 3x3 batches
void oracle_svd3_batch(const float* A, float* U, float* S, float* V, size_t n) {
#pragma omp parallel for schedule(static)
  for (long long q = 0; q < (long long)n; ++q) {
    M3 a, u, v;
    std::memcpy(a.a, A + 9 * q, 36);
    svd3(a, u, S + 3 * q, v);
    std::memcpy(U + 9 * q, u.a, 36);
    std::memcpy(V + 9 * q, v.a, 36);
  }
}
void oracle_polar_batch(const float* A, float* R, float* Sym, size_t n) {
#pragma omp parallel for schedule(static)
  for (long long q = 0; q < (long long)n; ++q) {
    M3 a, r, s;
    std::memcpy(a.a, A + 9 * q, 36);
    polar(a, r, s);
    std::memcpy(R + 9 * q, r.a, 36);
    std::memcpy(Sym + 9 * q, s.a, 36);
  }
}
float oracle_determinant(const float* A) {
  M3 a;
  std::memcpy(a.a, A, 36);
  return determinant(a);
}
void oracle_weights(const float* x, float dx_inv, int* base, float* w9) {
  float w[3][3];
  weights_per_direction(x, dx_inv, base, w);
  std::memcpy(w9, w, 36);
}

// Cell key of north-star stage (1): N^2*bi + N*bj + bk on the clamped base node.
void oracle_cell_keys(const void* particles, size_t count, float dt, uint32_t N, uint32_t* keys) {
  const Particle* ps = (const Particle*)particles;
  Params par(dt, N);
  for (size_t i = 0; i < count; ++i) {
    int base[3];
    float w[3][3];
    weights_per_direction(ps[i].x, par.dx_inv, base, w);
    uint32_t b[3];
    for (int a = 0; a < 3; ++a) b[a] = (uint32_t)std::min(std::max(base[a], 0), (int)N - 1);
    keys[i] = N * N * b[0] + N * b[1] + b[2];
  }
}
// stable sort by key: perm[r] = original index of the particle at sorted rank r
void oracle_sort_perm(const uint32_t* keys, size_t count, uint32_t* perm) {
  std::iota(perm, perm + count, 0u);
  std::stable_sort(perm, perm + count, [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
}

void oracle_reset_grid(float* grid, uint32_t N) { std::memset(grid, 0, sizeof(float) * 4 * (size_t)N * N * N); }
void oracle_p2g(const void* particles, size_t count, const float* mats, float dt, uint32_t N, int kind, float* grid) {
  Params par(dt, N);
  p2g((const Particle*)particles, count, (const Material*)mats, par, kind, grid);
}
// sum over particles of the magnitudes of the terms of each contribution, per node and channel: what a
// per-node P2G error is measured against
void oracle_p2g_magnitudes(const void* particles, size_t count, const float* mats, float dt, uint32_t N, int kind, float* grid) {
  Params par(dt, N);
  p2g((const Particle*)particles, count, (const Material*)mats, par, kind, grid, true);
}
void oracle_grid_update(float* grid, float dt, uint32_t N) {
  Params par(dt, N);
  grid_update(grid, par, (int)N, 0);
}
void oracle_g2p(const float* grid, void* particles, size_t count, const float* mats, float dt, uint32_t N, int kind) {
  Params par(dt, N);
  g2p(grid, (Particle*)particles, count, (const Material*)mats, par, kind);
}
// Simulation::advance, src/mpm.cu:323-329, repeated n_steps times on a caller-owned grid buffer.
void oracle_advance(void* particles, size_t count, const float* mats, float dt, uint32_t N, int kind, float* grid,
                    int n_steps) {
  Params par(dt, N);
  for (int s = 0; s < n_steps; ++s) {
    std::memset(grid, 0, sizeof(float) * 4 * (size_t)N * N * N);
    p2g((const Particle*)particles, count, (const Material*)mats, par, kind, grid);
    grid_update(grid, par, (int)N, 0);
    g2p(grid, (Particle*)particles, count, (const Material*)mats, par, kind);
  }
}

}  // extern "C"
