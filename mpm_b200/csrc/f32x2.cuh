// Packed float pairs for the sm_100 FFMA2 / FMUL2 / FADD2 instructions (PTX fma.rn.f32x2 etc.).
//
// Measured on the B200 (tools/microbench.cu): FFMA issues 1 warp-instruction/clk/SMSP, FFMA2
// one per 2 clk — the same flop rate, but HALF the issue slots.  P2G and G2P are issue-bound
// before they are HBM-bound, so their stencil accumulations run on pairs.
#pragma once
#include <cuda_runtime.h>

namespace mpm {

typedef unsigned long long f2;  // two floats in one 64-bit register pair: (lo, hi)

__device__ __forceinline__ f2 pack2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f2 dup2(float a) { return pack2(a, a); }
__device__ __forceinline__ float lo2(f2 v) { return __uint_as_float((unsigned int)v); }
__device__ __forceinline__ float hi2(f2 v) { return __uint_as_float((unsigned int)(v >> 32)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

}  // namespace mpm
