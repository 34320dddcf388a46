// Microbenchmark: cp.reduce.async.bulk (shared -> global f32 add) against per-lane red.global.add.v4.f32
// for the P2G flush pattern.  Each block flushes NRUNS runs x 27 nodes of 16 B.
//   mode 0: per-lane REDG.128, 3 lanes per run hitting 3 consecutive nodes (what p2g_sched does)
//   mode 1: one bulk reduce of 48 B per (run, ij): 9 per run
//   mode 2: one bulk reduce of NRUNS*16 B per (c, ij): 27 per block (z-contiguous runs)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bulkred_bench tools/bulkred_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
constexpr int NRUNS = 40;
constexpr int N = 256;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_red_add(float* g, const void* s, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(s)), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(256, 4) k(float4* grid, int mode, int nblocks) {
  __shared__ __align__(128) float4 stage[27 * 48];
  const int tid = threadIdx.x;
  // block b handles a z-column segment: row (x, y) from b, z0 from b
  const long long b = blockIdx.x;
  const int z0 = (int)((b * 33) % 200) + 20;
  const long long row = (b * 33) / 200;  // advancing rows
  const int y = (int)(row % 200) + 20, x = (int)((row / 200) % 200) + 20;
  for (int i = tid; i < 27 * 48; i += 256) stage[i] = make_float4(1.f, 2.f, 3.f, 4.f);
  __syncthreads();
  const long long NN = (long long)N * N;
  if (mode == 0) {
    for (int u = tid; u < 3 * NRUNS; u += 256) {
      const int r = u / 3, c = u - 3 * r;
      float4* gp = grid + ((long long)x * NN + (long long)y * N + z0 + r + c);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) atomicAdd(gp + i * NN + j * N, stage[(c * 9 + i * 3 + j) * 48 + r]);
    }
  } else if (mode == 1) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int u = tid; u < 9 * NRUNS; u += 256) {
      const int r = u / 9, ij = u - 9 * r;
      float4* gp = grid + ((long long)(x + ij / 3) * NN + (long long)(y + ij % 3) * N + z0 + r);
      bulk_red_add((float*)gp, &stage[(ij * 48 + r) * 1], 48);  // (layout does not matter for timing)
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  } else {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid < 27) {
      const int c = tid / 9, ij = tid % 9;
      float4* gp = grid + ((long long)(x + ij / 3) * NN + (long long)(y + ij % 3) * N + z0 + c);
      bulk_red_add((float*)gp, &stage[tid * 48], NRUNS * 16);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
}
int main() {
  float4* grid;
  const size_t nodes = (size_t)N * N * N;
  cudaMalloc(&grid, nodes * 16);
  const int nblocks = 262144;  // as many as the P2G kernel has tiles at 2^26 particles
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(grid, 0, nodes * 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<<<nblocks, 256>>>(grid, mode, nblocks);
    cudaDeviceSynchronize();
    cudaMemset(grid, 0, nodes * 16);
    cudaEventRecord(e0);
    for (int it = 0; it < 5; ++it) k<<<nblocks, 256>>>(grid, mode, nblocks);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // check: total sum of .x must be 5 * nblocks * NRUNS * 27
    printf("mode %d: %.3f ms per launch (%s)\n", mode, ms / 5, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
