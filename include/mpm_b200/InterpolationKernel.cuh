// mpm_b200 plugin surface — interpolation kernels.
//
// Concept-compatible with the reference's include/InterpolationKernel.cuh:12-74:
//     static constexpr u32  size();          // nodes per direction of the stencil
//     static constexpr bool d_is_const();    // D^-1 independent of the particle position
//     WeightMat<size()> weights_per_direction(Vec const& x, real dx_inv, Veci& range_begin) const;
//     Mat D_inv(Vec const& x, Veci const& range_begin, WeightMat<size()> const& w, real dx) const;
//     Mat D_inv_const(real dx_inv) const;
// The generic substep kernels (mpm_b200/csrc/kernels.cuh) loop over size()^3 nodes through these
// calls only; the staged production kernels are specialisations for QuadraticInterpolationKernel.
#pragma once
#include "types.cuh"

// 3 x N weights, one row per axis (the reference's Eigen::Matrix<real, 3, N>)
template <u32 N>
struct WeightMat {
  real w[3][N];
  MPM_INL real& operator()(int axis, int i) { return w[axis][i]; }
  MPM_INL real operator()(int axis, int i) const { return w[axis][i]; }
};

// adjugate / determinant; only the generic D_inv needs it
MPM_INL Mat inverse3(const Mat& A) {
  Mat c;
  c.m[0][0] = A.m[1][1] * A.m[2][2] - A.m[1][2] * A.m[2][1];
  c.m[0][1] = A.m[0][2] * A.m[2][1] - A.m[0][1] * A.m[2][2];
  c.m[0][2] = A.m[0][1] * A.m[1][2] - A.m[0][2] * A.m[1][1];
  c.m[1][0] = A.m[1][2] * A.m[2][0] - A.m[1][0] * A.m[2][2];
  c.m[1][1] = A.m[0][0] * A.m[2][2] - A.m[0][2] * A.m[2][0];
  c.m[1][2] = A.m[0][2] * A.m[1][0] - A.m[0][0] * A.m[1][2];
  c.m[2][0] = A.m[1][0] * A.m[2][1] - A.m[1][1] * A.m[2][0];
  c.m[2][1] = A.m[0][1] * A.m[2][0] - A.m[0][0] * A.m[2][1];
  c.m[2][2] = A.m[0][0] * A.m[1][1] - A.m[0][1] * A.m[1][0];
  const real det = A.m[0][0] * c.m[0][0] + A.m[0][1] * c.m[1][0] + A.m[0][2] * c.m[2][0];
  return (1.0f / det) * c;
}

template <u32 N, bool D_is_const = false>
class InterpolationKernelBase {
 public:
  CUDA_HOSTDEV static constexpr u32 size() { return N; }
  CUDA_HOSTDEV static constexpr bool d_is_const() { return D_is_const; }

  // D = sum_nodes w d d^T over the stencil, inverted (reference include/InterpolationKernel.cuh:21-50):
  // the fallback for kernels whose D depends on the particle position
  CUDA_HOSTDEV Mat D_inv(Vec const& x_particle, Veci const& range_begin, WeightMat<N> const& weights, real dx) const {
    Mat D = Mat::Zero();
    Vec d;
    for (u32 i = 0; i < N; ++i) {
      d(0) = (range_begin(0) + (int)i) * dx - x_particle(0);
      for (u32 j = 0; j < N; ++j) {
        d(1) = (range_begin(1) + (int)j) * dx - x_particle(1);
        for (u32 k = 0; k < N; ++k) {
          d(2) = (range_begin(2) + (int)k) * dx - x_particle(2);
          const real weight = weights(0, i) * weights(1, j) * weights(2, k);
          D += outer(weight * d, d);
        }
      }
    }
    return inverse3(D);
  }
};

// quadratic B-spline (reference include/InterpolationKernel.cuh:55-74)
class QuadraticInterpolationKernel : public InterpolationKernelBase<3, true> {
 public:
  // base node by C truncation like the reference's cast<int>(); fx in [0.5, 1.5)
  CUDA_HOSTDEV WeightMat<3> weights_per_direction(Vec const& x_particle, real dx_inv, Veci& range_begin) const {
    WeightMat<3> w;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const real g = x_particle(a) * dx_inv;
      range_begin(a) = (int)(g - 0.5f);
      const real fx = g - (real)range_begin(a);
      const real d0 = 1.5f - fx, d1 = fx - 1.0f, d2 = fx - 0.5f;
      w(a, 0) = 0.5f * (d0 * d0);
      w(a, 1) = 0.75f - (d1 * d1);
      w(a, 2) = 0.5f * (d2 * d2);
    }
    return w;
  }
  // (4 dx_inv) dx_inv on the diagonal, evaluated left to right like `Identity() * 4.0 * dx_inv * dx_inv`
  CUDA_HOSTDEV Mat D_inv_const(real dx_inv) const { return D_inv_scalar(dx_inv) * Mat::Identity(); }
  CUDA_HOSTDEV static real D_inv_scalar(real dx_inv) { return (4.0f * dx_inv) * dx_inv; }
};
