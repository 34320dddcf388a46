// mpm_b200 plugin surface — common types.
//
// Replaces the reference's include/types.h:5-35 (Eigen fixed-size aliases Vec / Mat / Vec4 / Veci,
// ParticleBase) and the MLS_APIC_Particle of include/TransferScheme.h:46-54 for code that runs
// inside the kernels.  The reference builds these on Eigen 3.3.7; here they are small register
// aggregates with just the operations the plugin concepts use, so a material / interpolation
// kernel / transfer scheme written against the reference's headers compiles against these:
//   Mat::Zero() Mat::Identity() A(i,j) A+B A-B A*B A*v s*A A*s -A A.transpose() A.determinant()
//   Vec::Zero() Vec::Constant(c) v(i) v[i] u+v u-v s*v -v, Vec4 with head3()
// Scalars of any arithmetic type are narrowed to `real` where they meet a matrix, which is what
// Eigen does with the reference's `2.0 * mu0 * (F - R)` style expressions.
//
// These are the IN-KERNEL views of a particle (registers).  The 104-byte AoS record that crosses
// the C ABI is MpmParticle (include/mpm_b200.h); in HBM the particles live as SoA streams.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifdef __CUDACC__
#define CUDA_HOSTDEV __host__ __device__
#define MPM_INL __host__ __device__ __forceinline__
#else
#define CUDA_HOSTDEV
#define MPM_INL inline
#endif

using real = float;
using u8 = uint8_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i32 = int32_t;
using i64 = int64_t;
using f32 = float;
using f64 = double;

struct Veci {
  int v[3];
  MPM_INL int& operator()(int i) { return v[i]; }
  MPM_INL int operator()(int i) const { return v[i]; }
  MPM_INL int& operator[](int i) { return v[i]; }
  MPM_INL int operator[](int i) const { return v[i]; }
};

struct Vec {
  real v[3];
  MPM_INL real& operator()(int i) { return v[i]; }
  MPM_INL real operator()(int i) const { return v[i]; }
  MPM_INL real& operator[](int i) { return v[i]; }
  MPM_INL real operator[](int i) const { return v[i]; }
  MPM_INL static Vec Zero() { return Vec{{0.f, 0.f, 0.f}}; }
  MPM_INL static Vec Constant(real c) { return Vec{{c, c, c}}; }
  MPM_INL Vec& operator+=(const Vec& o) {
    v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2];
    return *this;
  }
  MPM_INL real dot(const Vec& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
};
MPM_INL Vec operator+(const Vec& a, const Vec& b) { return Vec{{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; }
MPM_INL Vec operator-(const Vec& a, const Vec& b) { return Vec{{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; }
MPM_INL Vec operator-(const Vec& a) { return Vec{{-a.v[0], -a.v[1], -a.v[2]}}; }
template <class S>
MPM_INL Vec operator*(S s, const Vec& a) {
  const real f = (real)s;
  return Vec{{f * a.v[0], f * a.v[1], f * a.v[2]}};
}
template <class S>
MPM_INL Vec operator*(const Vec& a, S s) { return s * a; }

struct Vec4 {
  real v[4];
  MPM_INL real& operator()(int i) { return v[i]; }
  MPM_INL real operator()(int i) const { return v[i]; }
  MPM_INL real& operator[](int i) { return v[i]; }
  MPM_INL real operator[](int i) const { return v[i]; }
  MPM_INL static Vec4 Zero() { return Vec4{{0.f, 0.f, 0.f, 0.f}}; }
  MPM_INL Vec head3() const { return Vec{{v[0], v[1], v[2]}}; }
};

struct Mat {
  real m[3][3];  // row-major m[r][c] (registers; the AoS record at the C ABI is column-major like Eigen)
  MPM_INL real& operator()(int r, int c) { return m[r][c]; }
  MPM_INL real operator()(int r, int c) const { return m[r][c]; }
  MPM_INL static Mat Zero() {
    Mat r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) r.m[i][j] = 0.f;
    return r;
  }
  MPM_INL static Mat Identity() {
    Mat r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) r.m[i][j] = (i == j) ? 1.f : 0.f;
    return r;
  }
  MPM_INL Mat transpose() const {
    Mat r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) r.m[i][j] = m[j][i];
    return r;
  }
  // cofactor expansion along the first row, the order of the reference's linalg::determinant
  // (src/linalg.cu:47-52)
  MPM_INL real determinant() const {
    const real sub1 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
    const real sub2 = m[1][0] * m[2][2] - m[1][2] * m[2][0];
    const real sub3 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
    return m[0][0] * sub3 - m[0][1] * sub2 + m[0][2] * sub1;
  }
  MPM_INL Mat& operator+=(const Mat& o) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) m[i][j] += o.m[i][j];
    return *this;
  }
};
MPM_INL Mat operator+(const Mat& a, const Mat& b) {
  Mat r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j];
  return r;
}
MPM_INL Mat operator-(const Mat& a, const Mat& b) {
  Mat r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] - b.m[i][j];
  return r;
}
MPM_INL Mat operator-(const Mat& a) {
  Mat r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = -a.m[i][j];
  return r;
}
MPM_INL Mat operator*(const Mat& a, const Mat& b) {
  Mat r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
MPM_INL Vec operator*(const Mat& a, const Vec& x) {
  Vec r;
#pragma unroll
  for (int i = 0; i < 3; ++i) r.v[i] = a.m[i][0] * x.v[0] + a.m[i][1] * x.v[1] + a.m[i][2] * x.v[2];
  return r;
}
template <class S>
MPM_INL Mat operator*(S s, const Mat& a) {
  const real f = (real)s;
  Mat r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = f * a.m[i][j];
  return r;
}
template <class S>
MPM_INL Mat operator*(const Mat& a, S s) { return s * a; }
// a b^T
MPM_INL Mat outer(const Vec& a, const Vec& b) {
  Mat r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.v[i] * b.v[j];
  return r;
}
// a b^T with a, b 3x3 (a * b.transpose() without forming the transpose)
MPM_INL Mat mul_abt(const Mat& a, const Mat& b) {
  Mat r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[j][0] + a.m[i][1] * b.m[j][1] + a.m[i][2] * b.m[j][2];
  return r;
}

// minimal particle (reference include/types.h:24-35)
struct ParticleBase {
  u8 material_type;
  Vec x;  // position
  Vec v;  // velocity
  Mat F;  // deformation gradient
};
// particle of the MLS-APIC transfer (reference include/TransferScheme.h:46-54)
struct MLS_APIC_Particle : public ParticleBase {
  Mat C;    // affine momentum
  real Jp;  // plastic volume ratio, drives the snow hardening
};
