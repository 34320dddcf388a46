// How fast can the particle streams move?  G2P reads 12 and writes 24 float streams per particle.
// Compares the SoA layout (25 arrays of P floats) with AoSoA blocks ([P/B][25][B] floats).
#include <cuda_runtime.h>
#include <cstdio>
template <class F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
  return best;
}
// NR streams read, NW streams written (streams 0..NR-1 read; written streams start at W0)
template <int NR, int NW, int W0>
__global__ void soa_kernel(float* f, size_t stride, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
#pragma unroll
  for (int s = 0; s < NR; ++s) a += __ldcs(f + s * stride + i);
#pragma unroll
  for (int s = 0; s < NW; ++s) __stcs(f + (size_t)((W0 + s) % 25) * stride + i, a + s);
}
// misaligned variant: every warp's 128-byte access straddles two lines (tiles start at arbitrary particles)
template <int NR, int NW, int W0>
__global__ void soa_off_kernel(float* f, size_t stride, size_t n, int off) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x + off;
  if (i >= n) return;
  float a = 0.f;
#pragma unroll
  for (int s = 0; s < NR; ++s) a += __ldcs(f + s * stride + i);
#pragma unroll
  for (int s = 0; s < NW; ++s) __stcs(f + (size_t)((W0 + s) % 25) * stride + i, a + s);
}
// 252-particle tiles processed by 256-thread blocks (4 idle threads), like the staged kernels
template <int NR, int NW, int W0>
__global__ void soa_tile_kernel(float* f, size_t stride, size_t n) {
  if (threadIdx.x >= 252) return;
  const size_t i = (size_t)blockIdx.x * 252 + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
#pragma unroll
  for (int s = 0; s < NR; ++s) a += __ldcs(f + s * stride + i);
#pragma unroll
  for (int s = 0; s < NW; ++s) __stcs(f + (size_t)((W0 + s) % 25) * stride + i, a + s);
}
template <int NR, int NW, int W0, int B>
__global__ void aosoa_kernel(float* f, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* blk = f + (i / B) * (25 * B) + (i % B);
  float a = 0.f;
#pragma unroll
  for (int s = 0; s < NR; ++s) a += __ldcs(blk + s * B);
#pragma unroll
  for (int s = 0; s < NW; ++s) __stcs(blk + ((W0 + s) % 25) * B, a + s);
}
int main() {
  const size_t n = 1 << 26, stride = n;
  float* f; cudaMalloc(&f, sizeof(float) * 25 * stride); cudaMemset(f, 0, sizeof(float) * 25 * stride);
  const unsigned nb = (unsigned)(n / 256);
  auto rep = [&](const char* name, float ms, int nr, int nw) { printf("%-34s %.3f ms  %.0f GB/s\n", name, ms, (nr + nw) * 4.0 * n / ms / 1e6); };
  rep("SoA   read 12 write 24 (G2P)", time_ms([&] { soa_kernel<12, 24, 0><<<nb, 256>>>(f, stride, n); }), 12, 24);
  rep("SoA   read 25 write 0  (P2G)", time_ms([&] { soa_kernel<25, 0, 0><<<nb, 256>>>(f, stride, n); }), 25, 0);
  rep("SoA   read 12 write 12", time_ms([&] { soa_kernel<12, 12, 12><<<nb, 256>>>(f, stride, n); }), 12, 12);
  rep("SoA   read 1 write 1", time_ms([&] { soa_kernel<1, 1, 1><<<nb, 256>>>(f, stride, n); }), 1, 1);
  rep("SoA   read 1 write 2", time_ms([&] { soa_kernel<1, 2, 1><<<nb, 256>>>(f, stride, n); }), 1, 2);
  rep("SoA   read 0 write 24", time_ms([&] { soa_kernel<0, 24, 0><<<nb, 256>>>(f, stride, n); }), 0, 24);
  rep("SoA +5 misaligned r12 w24", time_ms([&] { soa_off_kernel<12, 24, 0><<<nb, 256>>>(f, stride, n, 5); }), 12, 24);
  rep("SoA +16 misaligned r12 w24", time_ms([&] { soa_off_kernel<12, 24, 0><<<nb, 256>>>(f, stride, n, 16); }), 12, 24);
  rep("SoA 252-tiles r12 w24", time_ms([&] { soa_tile_kernel<12, 24, 0><<<(unsigned)(n / 252 + 1), 256>>>(f, stride, n); }), 12, 24);
  rep("SoA +5 misaligned r12 w6", time_ms([&] { soa_off_kernel<12, 6, 0><<<nb, 256>>>(f, stride, n, 5); }), 12, 6);
  rep("SoA aligned r12 w6", time_ms([&] { soa_kernel<12, 6, 0><<<nb, 256>>>(f, stride, n); }), 12, 6);
  rep("AoSoA32  read 12 write 24", time_ms([&] { aosoa_kernel<12, 24, 0, 32><<<nb, 256>>>(f, n); }), 12, 24);
  rep("AoSoA128 read 12 write 24", time_ms([&] { aosoa_kernel<12, 24, 0, 128><<<nb, 256>>>(f, n); }), 12, 24);
  rep("AoSoA256 read 12 write 24", time_ms([&] { aosoa_kernel<12, 24, 0, 256><<<nb, 256>>>(f, n); }), 12, 24);
  rep("AoSoA256 read 25 write 0", time_ms([&] { aosoa_kernel<25, 0, 0, 256><<<nb, 256>>>(f, n); }), 25, 0);
  rep("AoSoA1024 read 12 write 24", time_ms([&] { aosoa_kernel<12, 24, 0, 1024><<<nb, 256>>>(f, n); }), 12, 24);
  return 0;
}
