// TEST INFRASTRUCTURE ONLY (oracle/): never linked, imported or executed by the product path.
//
// Host build of the reference's own svd3 (include/svd3_cuda.h:35-1043), compiled from the
// header WHERE IT LIES under /root/reference (never copied into this repo) into
// oracle/_ref/libref_svd3.so by oracle/Makefile.  The header is CUDA-only as shipped; the
// five mappings below are the complete host adaptation (SURVEY.md §8(c) probe):
//   __device__ / __forceinline__  -> nothing / inline
//   __fadd_rn / __fsub_rn         -> IEEE a+b / a-b (build uses -ffp-contract=off: no FMA)
//   __frsqrt_rn(x)                -> correctly rounded 1/sqrt(x) via double
//   max                           -> fmaxf
// Used (a) to pin oracle/mpm_oracle.cpp's own svd3 restatement bit-for-bit, (b) to generate
// tests/golden/svd3_golden.npz, (c) as the reference answer for the CUDA svd3 parity tests.
#include <cmath>
#include <cstddef>

#define __device__
#define __forceinline__ inline
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __frsqrt_rn(float x) { return (float)(1.0 / std::sqrt((double)x)); }
static inline float max(float a, float b) { return std::fmax(a, b); }

// found through -I/root/reference/include (the header's own <cuda.h> comes from the toolkit)
#include "svd3_cuda.h"

extern "C" {

// A, U, V are row-major 3x3 (a11 a12 a13 a21 ...), S = (s11, s22, s33); n matrices, packed.
void ref_svd3_batch(const float* A, float* U, float* S, float* V, size_t n) {
#pragma omp parallel for schedule(static)
  for (long long q = 0; q < (long long)n; ++q) {
    const float* a = A + 9 * q;
    float* u = U + 9 * q;
    float* s = S + 3 * q;
    float* v = V + 9 * q;
    svd(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8],
        u[0], u[1], u[2], u[3], u[4], u[5], u[6], u[7], u[8],
        s[0], s[1], s[2],
        v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
  }
}

}  // extern "C"
