// Kernels of the handle itself (included by mpm_sim.cu only): AoS <-> SoA conversion at the C-ABI
// boundary, the synthetic dense-block generator, the grid update, the linalg test hooks.
#pragma once
#include "kernels.cuh"

namespace mpm {

// ---- particle <-> AoS conversion (boundary only, off the hot path) ---------------------------
__device__ __forceinline__ void aos_record_to_slot(const MpmParticle& q, float* __restrict__ c) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    c[(SX + a) * kTile] = q.x[a];
    c[(SV + a) * kTile] = q.v[a];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      c[(SF + 3 * r + cc) * kTile] = q.F[3 * cc + r];  // AoS is column-major
      c[(SC + 3 * r + cc) * kTile] = q.C[3 * cc + r];
    }
  c[SJ * kTile] = q.Jp;
}

// aos[0..count) -> slots [offset, offset + count), ids first_id + slot
__global__ void aos_to_soa_kernel(const MpmParticle* __restrict__ aos, Soa p, size_t count, uint32_t first_id, size_t offset,
                                  DeviceDiag* __restrict__ diag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const MpmParticle& q = aos[i];
  i += offset;
  aos_record_to_slot(q, p.col(i));
  p.id[i] = first_id + (uint32_t)i;
  p.mat[i] = q.material_type;
  if (q.Jp != 1.0f) diag->jp_not_one = 1u;
}

// slot r <- aos[id[r] - first_id]: new particle data into the existing (cell-sorted) slots
__global__ void aos_overwrite_kernel(const MpmParticle* __restrict__ aos, Soa p, size_t count, uint32_t first_id,
                                     DeviceDiag* __restrict__ diag) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const MpmParticle& q = aos[p.id[i] - first_id];
  aos_record_to_slot(q, p.col(i));
  p.mat[i] = q.material_type;
  if (q.Jp != 1.0f) diag->jp_not_one = 1u;
}

// ---- tile-wise conversion through shared memory ----------------------------------------------------
// A thread that reads or writes "its" 104-byte record touches 26 words 104 bytes apart from its
// neighbour's: every warp instruction hits 32 sectors for 128 useful bytes.  Here a CTA moves one
// 256-particle tile: the SoA side is read / written row by row (coalesced), the AoS side record by
// record with 26 consecutive lanes on the 26 words of one record (contiguous 104 bytes), and the
// transposition happens in shared memory ([particle][27]: odd stride, no bank conflicts either way).
constexpr int kRecWords = (int)(sizeof(MpmParticle) / 4);  // 26
static_assert(sizeof(MpmParticle) == 104, "record layout");
// word w of the AoS record <-> stream row: word 0 = material_type (+pad), 1..3 x, 4..6 v, 7..15 F (column-major),
// 16..24 C (column-major), 25 Jp
__device__ __forceinline__ int aos_word_of_row(int row) {
  if (row >= SX && row < SX + 3) return 1 + (row - SX);
  if (row >= SV && row < SV + 3) return 4 + (row - SV);
  if (row >= SF && row < SF + 9) {
    const int e = row - SF;  // row-major 3 r + c -> column-major 3 c + r
    return 7 + 3 * (e % 3) + e / 3;
  }
  if (row >= SC && row < SC + 9) {
    const int e = row - SC;
    return 16 + 3 * (e % 3) + e / 3;
  }
  return 25;  // SJ
}

// particle r of the tile -> aos[id[r] - first_id] (by_id: restores upload order) or aos[r]
__global__ void __launch_bounds__(kTile) soa_to_aos_kernel(Soa p, size_t count, MpmParticle* __restrict__ aos, uint32_t first_id, bool by_id) {
  __shared__ uint32_t rec[kTile * (kRecWords + 1)];
  __shared__ uint32_t dst[kTile];
  const int tid = threadIdx.x;
  const size_t i = (size_t)blockIdx.x * kTile + tid;
  const size_t n_here = min((size_t)kTile, count - (size_t)blockIdx.x * kTile);
  if (i < count) {
    const float* __restrict__ c = p.tile(blockIdx.x) + tid;
    uint32_t* r = rec + tid * (kRecWords + 1);
    r[0] = (uint32_t)p.mat[i];  // pad bytes zero
#pragma unroll
    for (int row = 0; row < NSTREAM; ++row) r[aos_word_of_row(row)] = __float_as_uint(c[row * kTile]);
    dst[tid] = by_id ? (p.id[i] - first_id) : (uint32_t)i;
  }
  __syncthreads();
  uint32_t* __restrict__ out = reinterpret_cast<uint32_t*>(aos);
  for (int e = tid; e < (int)n_here * kRecWords; e += kTile) {
    const int q = e / kRecWords, w = e - q * kRecWords;
    out[(size_t)dst[q] * kRecWords + w] = rec[q * (kRecWords + 1) + w];
  }
}

// cell keys straight from the uploaded AoS records (the pairs the re-bin sorts), and the Jp != 1 flag
__global__ void __launch_bounds__(256) aos_keys_kernel(const MpmParticle* __restrict__ aos, size_t count, KParams k, uint32_t* __restrict__ keys,
                                                       uint32_t* __restrict__ vals, DeviceDiag* __restrict__ diag) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const MpmParticle& q = aos[i];
  int b[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float fx, w[3];
    bspline(q.x[a], k.dx_inv, b[a], fx, w);
    b[a] = min(max(b[a], 0), k.N - 1);
  }
  const int bx = min(max(b[0] - k.x0, 0), k.nxl - 1);
  keys[i] = (uint32_t)((bx * k.N + b[1]) * k.N + b[2]);
  vals[i] = (uint32_t)i;
  if (q.Jp != 1.0f) diag->jp_not_one = 1u;
}

// slot r of the (sorted) SoA <- aos[perm[r]]: conversion and cell-order permutation in one pass, so that
// an upload never moves the particles through the SoA twice; ids: upload order (first_id + index)
__global__ void __launch_bounds__(kTile) aos_gather_to_soa_kernel(const MpmParticle* __restrict__ aos, const uint32_t* __restrict__ perm, Soa p,
                                                                  size_t count, uint32_t first_id) {
  __shared__ uint32_t rec[kTile * (kRecWords + 1)];
  __shared__ uint32_t src[kTile];
  const int tid = threadIdx.x;
  const size_t i = (size_t)blockIdx.x * kTile + tid;
  const size_t n_here = min((size_t)kTile, count - (size_t)blockIdx.x * kTile);
  if (i < count) src[tid] = perm[i];
  __syncthreads();
  const uint32_t* __restrict__ in = reinterpret_cast<const uint32_t*>(aos);
  for (int e = tid; e < (int)n_here * kRecWords; e += kTile) {
    const int q = e / kRecWords, w = e - q * kRecWords;
    rec[q * (kRecWords + 1) + w] = in[(size_t)src[q] * kRecWords + w];
  }
  __syncthreads();
  if (i < count) {
    float* __restrict__ c = p.tile(blockIdx.x) + tid;
    const uint32_t* r = rec + tid * (kRecWords + 1);
#pragma unroll
    for (int row = 0; row < NSTREAM; ++row) c[row * kTile] = __uint_as_float(r[aos_word_of_row(row)]);
    p.id[i] = first_id + src[tid];
    p.mat[i] = (uint8_t)(r[0] & 0xffu);
  }
}

__global__ void positions_kernel(Soa p, size_t count, float* __restrict__ xyz, uint32_t first_id, bool by_id) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float* c = p.col(i);
  const size_t o = (by_id ? (size_t)(p.id[i] - first_id) : i) * 3;
  xyz[o + 0] = c[(SX + 0) * kTile];
  xyz[o + 1] = c[(SX + 1) * kTile];
  xyz[o + 2] = c[(SX + 2) * kTile];
}

// ---- synthetic dense block; lowbias32 counter hash, mirrored in tests/ ------
__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

// Optional "stress" of the synthetic block (bench.py --stress): a shear velocity field and a
// perturbed deformation gradient, so that the data-dependent paths of the kernels (Newton polar
// iterations, plasticity SVD, cell crossings, shorter P2G runs) are exercised.
struct BlockStress {
  float shear;    // v = shear * (y - 0.5, 0, 0.3 * (x - 0.5))  [1/s]
  float f_noise;  // F = I + f_noise * hash noise in [-1, 1)
};

__global__ void generate_block_kernel(Soa p, unsigned long long first_id, unsigned long long count, uint32_t seed_hash,
                                      float lo, float hi, uint8_t material, KParams k, bool whole_domain,
                                      unsigned long long* __restrict__ n_out, size_t capacity, BlockStress st) {
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const unsigned long long id = first_id + t;
  float x[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const uint32_t h = lowbias32((uint32_t)(id * 3ull + (unsigned long long)a) ^ seed_hash);
    const float u = (float)(h >> 8) * (1.0f / 16777216.0f);
    x[a] = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), u));  // no FMA: bit-identical to the host generator
  }
  size_t slot;
  if (whole_domain) {
    slot = (size_t)t;
  } else {
    int b;
    float fx, w[3];
    bspline(x[0], k.dx_inv, b, fx, w);
    b = min(max(b, 0), k.N - 1);
    if (b < k.x_own_begin || b >= k.x_own_end) return;
    slot = (size_t)atomicAdd(n_out, 1ull);
    if (slot >= capacity) return;
  }
  float* c = p.col(slot);
  // individually rounded operations: bit-identical to the numpy mirror the tests keep of this generator
  float v[3] = {0.f, 0.f, 0.f};
  if (st.shear != 0.f) {
    v[0] = __fmul_rn(st.shear, __fsub_rn(x[1], 0.5f));
    v[2] = __fmul_rn(__fmul_rn(0.3f, st.shear), __fsub_rn(x[0], 0.5f));
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    c[(SX + a) * kTile] = x[a];
    c[(SV + a) * kTile] = v[a];
  }
#pragma unroll
  for (int e = 0; e < 9; ++e) {
    float f = (e % 4 == 0) ? 1.0f : 0.0f;
    if (st.f_noise != 0.f) {
      const uint32_t h = lowbias32((uint32_t)(id * 9ull + (unsigned long long)e) ^ (seed_hash * 0x9e3779b9u + 77u));
      const float u = __fsub_rn(__fmul_rn((float)(h >> 8), 2.0f / 16777216.0f), 1.0f);
      f = __fadd_rn(f, __fmul_rn(st.f_noise, u));
    }
    c[(SF + e) * kTile] = f;
    c[(SC + e) * kTile] = 0.0f;
  }
  c[SJ * kTile] = 1.0f;
  p.id[slot] = (uint32_t)id;
  p.mat[slot] = material;
}

__global__ void __launch_bounds__(256) grid_update_kernel(float4* __restrict__ grid, KParams k, int plane_begin, int plane_end) {
  const long long NN = (long long)k.N * k.N;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x + (long long)plane_begin * NN;
  if (idx >= (long long)plane_end * NN) return;
  const float4 c = grid[idx];
  if (c.w > 0.0f) {
    const int xi = (int)(idx / NN) + k.x0;
    const int rem = (int)(idx % NN);
    grid[idx] = grid_node_update(c, xi, rem / k.N, rem % k.N, k);
  }
}

// ---- linalg test hooks (reference tests/test_linalg.cu:49-55) ---------------------------------
template <class O>
__global__ void svd3_batch_kernel(const float* __restrict__ A, float* U, float* S, float* V, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Mat3 a, u, v;
  float s[3];
  for (int e = 0; e < 9; ++e) a.m[e / 3][e % 3] = A[9 * i + e];
  svd3<O>(a, u, s, v);
  for (int e = 0; e < 9; ++e) {
    U[9 * i + e] = u.m[e / 3][e % 3];
    V[9 * i + e] = v.m[e / 3][e % 3];
  }
  for (int e = 0; e < 3; ++e) S[3 * i + e] = s[e];
}
template <class O>
__global__ void polar_batch_kernel(const float* __restrict__ A, float* R, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Mat3 a;
  for (int e = 0; e < 9; ++e) a.m[e / 3][e % 3] = A[9 * i + e];
  const Mat3 r = linalg::polar_rotation<O>(a);  // the same routine the material models use for this mode
  for (int e = 0; e < 9; ++e) R[9 * i + e] = r.m[e / 3][e % 3];
}
__global__ void det_batch_kernel(const float* __restrict__ A, float* det, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Mat3 a;
  for (int e = 0; e < 9; ++e) a.m[e / 3][e % 3] = A[9 * i + e];
  det[i] = linalg::determinant(a);
}
// Sum_nodes w d d^T of the quadratic kernel through the GENERIC D_inv (InterpolationKernel.cuh):
// must equal D_inv_const (SURVEY.md 8(c) pin 6).  out = 9 floats per position, row-major.
__global__ void dinv_batch_kernel(const float* __restrict__ x, float* out, size_t n, float dx, float dx_inv) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const QuadraticInterpolationKernel kern;
  const Vec xp{{x[3 * i], x[3 * i + 1], x[3 * i + 2]}};
  Veci rb;
  const WeightMat<3> w = kern.weights_per_direction(xp, dx_inv, rb);
  const Mat D = kern.D_inv(xp, rb, w, dx);
  for (int e = 0; e < 9; ++e) out[9 * i + e] = D.m[e / 3][e % 3];
}

}  // namespace mpm
