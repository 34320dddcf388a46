// Stand-alone check of the two tiled TMA loads the staged kernels use.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include "../mpm_b200/csrc/tma.cuh"
using namespace mpm;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k2d(const __grid_constant__ CUtensorMap tm, float* out, int c0, int c1, int n) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + n * 4);
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_arrive_expect_tx(bar, n * 4); tma_load_2d(sm, &tm, c0, c1, bar); }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i];
}
__global__ void k4d(const __grid_constant__ CUtensorMap tm, float* out, int c1, int c2, int c3, int n) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + n * 4);
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_arrive_expect_tx(bar, n * 4); tma_load_4d(sm, &tm, 0, c1, c2, c3, bar); }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i];
}
int main() {
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)ptr;
  printf("entry %p q=%d\n", ptr, (int)q);
  // ---- 2D: [25][stride]
  const int stride = 4096, rows = 25, T = 256;
  std::vector<float> h(rows * stride);
  for (int r = 0; r < rows; ++r) for (int c = 0; c < stride; ++c) h[r * stride + c] = r * 10000 + c;
  float *d, *out; cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 1 << 20);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  for (int R : {12, 13, 25}) {
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)stride, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)stride * 4};
    cuuint32_t box[2] = {(cuuint32_t)T, (cuuint32_t)R}; cuuint32_t es[2] = {1, 1};
    CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int n = T * R;
    cudaFuncSetAttribute(k2d, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 4 + 64);
    for (int c0 : {0, 256, 36, 4000}) {
      k2d<<<1, 128, n * 4 + 64>>>(tm, out, c0, 0, n);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> o(n); cudaMemcpy(o.data(), out, n * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < R; ++r) for (int c = 0; c < T; ++c) { float want = (c0 + c < stride) ? h[r * stride + c0 + c] : 0.f; if (o[r * T + c] != want) ++bad; }
      printf("2D rows %d c0 %d: enc rc %d, sync %s, mismatches %d\n", R, c0, (int)rc, cudaGetErrorString(e), bad);
      if (e != cudaSuccess) return 1;
    }
  }
  // ---- 4D grid [nx][N][N][4]
  for (int N : {32, 256}) {
    const int nx = 8, LT = 48;
    std::vector<float> g((size_t)nx * N * N * 4);
    for (size_t i = 0; i < g.size(); ++i) g[i] = (float)(i % 1000003);
    float* dg; cudaMalloc(&dg, g.size() * 4); cudaMemcpy(dg, g.data(), g.size() * 4, cudaMemcpyHostToDevice);
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[4] = {4, (cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)nx}; cuuint64_t strides[3] = {16, 16ull * N, 16ull * N * N};
    cuuint32_t box[4] = {4, LT, 5, 5}; cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dg, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int n = 25 * LT * 4;
    cudaFuncSetAttribute(k4d, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 4 + 64);
    for (int z0 : {-1, 3, N - 10}) {
      const int y0 = 2, x0 = -1;
      k4d<<<1, 128, n * 4 + 64>>>(tm, out, z0, y0, x0, n);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> o(n); cudaMemcpy(o.data(), out, n * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) for (int t = 0; t < LT; ++t) for (int c = 0; c < 4; ++c) {
        const int x = x0 + i, y = y0 + j, z = z0 + t;
        float want = (x >= 0 && x < nx && y >= 0 && y < N && z >= 0 && z < N) ? g[(((size_t)x * N + y) * N + z) * 4 + c] : 0.f;
        if (o[((i * 5 + j) * LT + t) * 4 + c] != want) ++bad;
      }
      printf("4D N %d z0 %d: enc rc %d, sync %s, mismatches %d\n", N, z0, (int)rc, cudaGetErrorString(e), bad);
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
