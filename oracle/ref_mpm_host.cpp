// TEST INFRASTRUCTURE ONLY (oracle/).  Host build of the reference's OWN substep code:
//   include/types.h, linalg.h, svd3_cuda.h, MaterialModel.cuh, InterpolationKernel.cuh,
//   TransferScheme.h                      — included unmodified from $(REF)/include
//   src/linalg.cu:18-53 (device half)     — extracted at build time into a temp dir (gen_linalg.inc)
//   src/mpm.cu:6-8,14-178 (three kernels) — extracted at build time into a temp dir (gen_kernels.inc)
// compiled against the Eigen shim in oracle/shim (Eigen itself is not in this image), with the
// CUDA execution model replaced by a serial loop: one call of the kernel function per
// (blockIdx, threadIdx).  Nothing from the reference is copied into the repository; the .inc
// files live in a temporary directory during the build only.  Result: oracle/_ref/libref_mpm_{snow,fc,jelly}.so, the
// "reference itself run here" that pins oracle/mpm_oracle.cpp (tests/test_oracle.py).
// Caveats (documented in DESIGN.md): host float arithmetic without FMA contraction, whereas the
// reference's nvcc build contracts; the kernels are launched only for particle counts <= N^3
// (the unmodified G2P guard `pi > N*N*N`, src/mpm.cu:114) and for particles away from the upper
// faces (the P2G clip bug, src/mpm.cu:42-44).
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

// ---- CUDA surface the reference sources use, mapped to the host -------------------------------
struct ref_dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local ref_dim3 threadIdx, blockIdx, blockDim, gridDim;
#define __global__
#define __shared__ static thread_local
#define __forceinline__ inline
#define __host__
// CUDA's global min/max overload set (crt/math_functions.hpp): mixed signedness -> unsigned,
// float -> fminf/fmaxf
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
static inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
static inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
static inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float atomicAdd(float* a, float v) { float o = *a; *a = o + v; return o; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __frsqrt_rn(float x) { return (float)(1.0 / std::sqrt((double)x)); }

#include <Eigen/Dense>  // oracle/shim
#include "types.h"
// linalg.h declares these only under __CUDACC__ (include/linalg.h:9-13)
namespace linalg {
void polar_decomposition_device(const Mat&, Mat&, Mat&);
void svd_decomposition(const Mat&, Mat&, Mat&, Mat&);
real determinant(const Mat&);
}
#include "linalg.h"
#include "svd3_cuda.h"
#include "MaterialModel.cuh"
#include "InterpolationKernel.cuh"
#include "TransferScheme.h"

namespace linalg {
#include "gen_linalg.inc"
}

// the four aliases of include/mpm.cuh:24-27 (that header itself needs boost/libigl/GL)
#if defined(REF_FIXED_COROTATED)
using Particle = MLS_APIC_Particle;
using MaterialModel = MMFixedCorotated<Particle>;
#elif defined(REF_JELLY)
using Particle = MLS_APIC_Particle;
using MaterialModel = MMJelly<Particle>;
#else
using Particle = MLS_APIC_Particle;
using MaterialModel = MMSnow<Particle>;
#endif
using InterpolationKernel = QuadraticInterpolationKernel;
using TransferScheme = MLS_APIC_Scheme<InterpolationKernel>;

#include "gen_kernels.inc"

static_assert(sizeof(Particle) == 104, "particle layout");
static_assert(sizeof(Vec4) == 16, "grid node layout");

namespace {
std::vector<MaterialModel> make_models(const float* mats7, int n) {
  std::vector<MaterialModel> v;
  for (int i = 0; i < n; ++i) {
    const float* m = mats7 + 7 * i;
#if defined(REF_FIXED_COROTATED)
    MaterialModel mm(m[0], 1.0f, 1.0f, 0.25f);
#elif defined(REF_JELLY)
    MaterialModel mm(m[0], 1.0f, 1.0f, 0.25f, m[4]);
#else
    MaterialModel mm(m[0], 1.0f, 1.0f, 0.25f, m[4], m[5], m[6]);
#endif
    mm.particleVolume = m[0];
    mm.particleMass = m[1];
    mm.mu0 = m[2];
    mm.lambda0 = m[3];
    v.push_back(mm);
  }
  return v;
}
}  // namespace

#if defined(REF_FIXED_COROTATED)
#define REFNAME(x) ref_fc_##x
#elif defined(REF_JELLY)
#define REFNAME(x) ref_jelly_##x
#else
#define REFNAME(x) ref_snow_##x
#endif

#pragma GCC visibility push(default)
extern "C" {

size_t REFNAME(sizeof_material)() { return sizeof(MaterialModel); }

// MaterialModel constructor exactly as src/main.cu:36-42 calls it (snow build only)
void REFNAME(make_material)(double volume, double density, double E, double Nu, double hardening, double lo,
                            double hi, float* out7) {
#if defined(REF_FIXED_COROTATED)
  MaterialModel mm(volume, density, E, Nu);
  float tmp[7] = {mm.particleVolume, mm.particleMass, mm.mu0, mm.lambda0, (float)hardening, (float)lo, (float)hi};
  std::memcpy(out7, tmp, sizeof(tmp));
#elif defined(REF_JELLY)
  MaterialModel mm(volume, density, E, Nu, hardening);
  float tmp[7] = {mm.particleVolume, mm.particleMass, mm.mu0, mm.lambda0, mm.hardening, (float)lo, (float)hi};
  std::memcpy(out7, tmp, sizeof(tmp));
#else
  MaterialModel mm(volume, density, E, Nu, hardening, lo, hi);
  static_assert(sizeof(MaterialModel) == 28, "material layout");
  std::memcpy(out7, &mm, 28);
#endif
}

void REFNAME(params)(float dt, uint32_t N, float* dx, float* dx_inv) {
  SimulationParameters p(dt, N);
  *dx = p.dx;
  *dx_inv = p.dx_inv;
}

// Simulation::particleToGridTransfer launch shape, src/mpm.cu:217-222
void REFNAME(p2g)(const void* particles, size_t count, const float* mats7, int n_mats, float dt, uint32_t N,
                  float* grid) {
  std::vector<MaterialModel> models = make_models(mats7, n_mats);
  SimulationParameters par(dt, N);
  InterpolationKernel kern;
  blockDim.x = 64;
  unsigned blocks = (unsigned)std::ceil(float(count) / 64.0f);
  for (unsigned b = 0; b < blocks; ++b)
    for (unsigned t = 0; t < 64; ++t) {
      blockIdx.x = b;
      threadIdx.x = t;
      particleToGrid((Particle*)particles, (Vec4*)grid, models.data(), (int)count, &par, &kern);
    }
}
// Simulation::gridOperations launch shape, src/mpm.cu:224-228
void REFNAME(grid_update)(float* grid, float dt, uint32_t N) {
  blockDim.x = N;
  for (unsigned x = 0; x < N; ++x)
    for (unsigned y = 0; y < N; ++y)
      for (unsigned z = 0; z < N; ++z) {
        blockIdx.x = x;
        blockIdx.y = y;
        threadIdx.x = z;
        gridOpKernel((Vec4*)grid, (int)(N * N * N), dt);
      }
}
// Simulation::gridToParticleTransfer launch shape, src/mpm.cu:230-235; only threads with
// pi < count are run (the device array holds exactly count particles).
int REFNAME(g2p)(const float* grid, void* particles, size_t count, const float* mats7, int n_mats, float dt,
                 uint32_t N) {
  if (count > (size_t)N * N * N) return -1;  // unmodified guard would freeze the tail (src/mpm.cu:114)
  std::vector<MaterialModel> models = make_models(mats7, n_mats);
  SimulationParameters par(dt, N);
  InterpolationKernel kern;
  blockDim.x = 64;
  for (size_t pi = 0; pi < count; ++pi) {
    blockIdx.x = (unsigned)(pi / 64);
    threadIdx.x = (unsigned)(pi % 64);
    gridToParticle((Vec4*)grid, (Particle*)particles, models.data(), par, kern);
  }
  return 0;
}
// Simulation::advance, src/mpm.cu:323-329
int REFNAME(advance)(void* particles, size_t count, const float* mats7, int n_mats, float dt, uint32_t N, float* grid,
                     int n_steps) {
  for (int s = 0; s < n_steps; ++s) {
    std::memset(grid, 0, sizeof(float) * 4 * (size_t)N * N * N);
    REFNAME(p2g)(particles, count, mats7, n_mats, dt, N, grid);
    REFNAME(grid_update)(grid, dt, N);
    if (REFNAME(g2p)(grid, particles, count, mats7, n_mats, dt, N)) return -1;
  }
  return 0;
}
// linalg::polar_decomposition_device on row-major 3x3 inputs (tests/test_linalg.cu:49-75)
void REFNAME(polar)(const float* A, float* R, float* S) {
  Mat a, r, s;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a(i, j) = A[3 * i + j];
  linalg::polar_decomposition_device(a, r, s);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      R[3 * i + j] = r(i, j);
      S[3 * i + j] = s(i, j);
    }
}

}  // extern "C"
#pragma GCC visibility pop
