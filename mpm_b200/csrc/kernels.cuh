// Substep kernels: P2G scatter, grid update, G2P gather (sm_100a).
// Reference behaviour: src/mpm.cu:14-178 + include/TransferScheme.h:66-142.
#pragma once
#include "common.cuh"

namespace mpm {

constexpr int kParticleBlock = 128;

// ---- particle <-> AoS conversion (boundary only, off the hot path) ---------------------------
__global__ void aos_to_soa_kernel(const MpmParticle* __restrict__ aos, Soa p, size_t count, uint32_t first_id) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const MpmParticle& q = aos[i];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    p.s(SX + a)[i] = q.x[a];
    p.s(SV + a)[i] = q.v[a];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      p.s(SF + 3 * r + c)[i] = q.F[3 * c + r];  // AoS is column-major
      p.s(SC + 3 * r + c)[i] = q.C[3 * c + r];
    }
  p.s(SJ)[i] = q.Jp;
  p.id[i] = first_id + (uint32_t)i;
  p.mat[i] = q.material_type;
}

// writes particle r to aos[id[r] - first_id]: restores upload order
__global__ void soa_to_aos_kernel(Soa p, size_t count, MpmParticle* __restrict__ aos, uint32_t first_id, bool by_id) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  MpmParticle q;
  q.material_type = p.mat[i];
  q.pad_[0] = q.pad_[1] = q.pad_[2] = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    q.x[a] = p.s(SX + a)[i];
    q.v[a] = p.s(SV + a)[i];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      q.F[3 * c + r] = p.s(SF + 3 * r + c)[i];
      q.C[3 * c + r] = p.s(SC + 3 * r + c)[i];
    }
  q.Jp = p.s(SJ)[i];
  aos[by_id ? (size_t)(p.id[i] - first_id) : i] = q;  // slab handles: current (cell-sorted) order
}

__global__ void positions_kernel(Soa p, size_t count, float* __restrict__ xyz, uint32_t first_id, bool by_id) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const size_t o = (by_id ? (size_t)(p.id[i] - first_id) : i) * 3;
  xyz[o + 0] = p.s(SX + 0)[i];
  xyz[o + 1] = p.s(SX + 1)[i];
  xyz[o + 2] = p.s(SX + 2)[i];
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

__global__ void generate_block_kernel(Soa p, unsigned long long first_id, unsigned long long count, uint32_t seed_hash,
                                      float lo, float hi, uint8_t material, KParams k, bool whole_domain,
                                      unsigned long long* __restrict__ n_out, size_t capacity) {
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
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    p.s(SX + a)[slot] = x[a];
    p.s(SV + a)[slot] = 0.0f;
  }
#pragma unroll
  for (int e = 0; e < 9; ++e) {
    p.s(SF + e)[slot] = (e % 4 == 0) ? 1.0f : 0.0f;
    p.s(SC + e)[slot] = 0.0f;
  }
  p.s(SJ)[slot] = 1.0f;
  p.id[slot] = (uint32_t)id;
  p.mat[slot] = material;
}

// ---- stage (2): P2G ---------------------------------------------------------------------------
// One thread per particle, particles in cell-sorted order so that a warp's 27 vector
// reductions land on a handful of neighbouring grid nodes (L2 atomic locality).  Each node is
// one aligned float4 -> a single red.global.add.v4.f32 (REDG.E.ADD.F32x4) per node instead of
// the reference's four scalar atomics (src/mpm.cu:66-70).  The momentum term m v + A d is affine
// in the node offset, so it is carried incrementally (3 adds per node instead of 9 FMAs).
template <int MODEL, class O, bool EXACT>
__global__ void __launch_bounds__(kParticleBlock)
p2g_kernel(Soa p, size_t count, const MpmMaterial* __restrict__ mats, float4* __restrict__ grid, KParams k) {
  const size_t pi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pi >= count) return;
  float x[3], v[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    x[a] = p.s(SX + a)[pi];
    v[a] = p.s(SV + a)[pi];
  }
  Mat3 F, C;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      F.m[r][c] = p.s(SF + 3 * r + c)[pi];
      C.m[r][c] = p.s(SC + 3 * r + c)[pi];
    }
  const float Jp = p.s(SJ)[pi];
  const MpmMaterial m = load_material(mats, p.mat[pi]);

  int base[3];
  float fx[3], w[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a) bspline(x[a], k.dx_inv, base[a], fx[a], w[a]);

  // affine = -Dinv*dt*vol*PF + m*C   (TransferScheme.h:83-85)
  const Mat3 PF = compute_PF<MODEL, O, EXACT>(F, Jp, m);
  const float kk = ((-k.dinv) * k.dt) * m.particleVolume;
  Mat3 A;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A.m[i][j] = kk * PF.m[i][j] + m.particleMass * C.m[i][j];

  // particles completely outside the domain (src/mpm.cu:31-35)
#pragma unroll
  for (int a = 0; a < 3; ++a)
    if (base[a] + 3 < 0 || base[a] >= k.N) return;

  bool ok[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int g = base[a] + i;
      ok[a][i] = (g >= 0) && (g < k.N);
      if (a == 0) ok[a][i] = ok[a][i] && (g >= k.x0) && (g < k.x0 + k.nxl);
    }

  // q(node) = m v + A (x_node - x) = q0 + i*colx + j*coly + k*colz
  float d0[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) d0[a] = (float)base[a] * k.dx - x[a];
  float q0[3], cx[3], cy[3], cz[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    q0[c] = v[c] * m.particleMass + (A.m[c][0] * d0[0] + A.m[c][1] * d0[1] + A.m[c][2] * d0[2]);
    cx[c] = A.m[c][0] * k.dx;
    cy[c] = A.m[c][1] * k.dx;
    cz[c] = A.m[c][2] * k.dx;
  }
  const long long NN = (long long)k.N * k.N;
  float4* gbase = grid + ((long long)(base[0] - k.x0) * NN + (long long)base[1] * k.N + base[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float qi[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) qi[c] = q0[c] + (float)i * cx[c];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float qj[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) qj[c] = qi[c] + (float)j * cy[c];
      const float wij = w[0][i] * w[1][j];
#pragma unroll
      for (int kz = 0; kz < 3; ++kz) {
        const float wt = wij * w[2][kz];
        float4 out;
        out.x = wt * (qj[0] + (float)kz * cz[0]);
        out.y = wt * (qj[1] + (float)kz * cz[1]);
        out.z = wt * (qj[2] + (float)kz * cz[2]);
        out.w = wt * m.particleMass;
        if (ok[0][i] && ok[1][j] && ok[2][kz]) atomicAdd(gbase + ((long long)i * NN + j * k.N + kz), out);
      }
    }
  }
}

// ---- stage (3): grid update -------------------------------------------------------------------
// momentum -> velocity, gravity, sticky walls / separating floor (reference src/mpm.cu:76-107).
// One float4 per thread, fully coalesced; mass is left untouched (the reference overwrites it
// with 1.0, SURVEY.md F8 — nothing downstream reads it).
__global__ void __launch_bounds__(256) grid_update_kernel(float4* __restrict__ grid, KParams k, int plane_begin, int plane_end) {
  const long long NN = (long long)k.N * k.N;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x + (long long)plane_begin * NN;
  if (idx >= (long long)plane_end * NN) return;
  float4 c = grid[idx];
  if (c.w > 0.0f) {
    const int xi = (int)(idx / NN) + k.x0;
    const int rem = (int)(idx % NN);
    const int yi = rem / k.N, zi = rem % k.N;
    c.x = c.x / c.w;
    c.y = c.y / c.w;
    c.z = c.z / c.w;
    c.y += k.dt * -9.81f;
    const float boundary = 0.05f;
    const float hi = 1.0f - boundary;
    const float X = (float)xi / (float)k.N, Y = (float)yi / (float)k.N, Z = (float)zi / (float)k.N;
    if (X < boundary || X > hi || Y > hi || Z < boundary || Z > hi) {
      c.x = 0.f;
      c.y = 0.f;
      c.z = 0.f;
    }
    if (Y < boundary) c.y = fmaxf(0.0f, c.y);
    grid[idx] = c;
  }
}

// ---- stage (4): G2P ---------------------------------------------------------------------------
// One thread per particle: 27 float4 node reads (sorted order -> L1/L2 hits), APIC C, F update,
// plasticity, advection (reference src/mpm.cu:109-178, TransferScheme.h:102-142).
// Tuning switches (tools/ab.py measures them on the same box; defaults = the fastest measured):
#ifndef MPM_G2P_VARIANT
#define MPM_G2P_VARIANT 0   // 0: per-node accumulation, 1: separable along z
#endif
#ifndef MPM_G2P_PREFETCH
#define MPM_G2P_PREFETCH 0  // 1: issue the F loads before the gather
#endif
#ifndef MPM_G2P_MINBLK
#define MPM_G2P_MINBLK 0    // __launch_bounds__ min blocks per SM (0 = unconstrained)
#endif
#ifndef MPM_G2P_BLOCK
#define MPM_G2P_BLOCK 128
#endif
constexpr int kG2pBlock = MPM_G2P_BLOCK;

template <int MODEL, class O>
__global__ void
#if MPM_G2P_MINBLK > 0
__launch_bounds__(kG2pBlock, MPM_G2P_MINBLK)
#else
__launch_bounds__(kG2pBlock)
#endif
g2p_kernel(Soa p, size_t count, const MpmMaterial* __restrict__ mats, const float4* __restrict__ grid, KParams k) {
  const size_t pi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pi >= count) return;
  float x[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) x[a] = p.s(SX + a)[pi];
  Mat3 F;
#if MPM_G2P_PREFETCH
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) F.m[r][c] = p.s(SF + 3 * r + c)[pi];
#endif
  int base[3];
  float fx[3], w[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a) bspline(x[a], k.dx_inv, base[a], fx[a], w[a]);
#pragma unroll
  for (int a = 0; a < 3; ++a)
    if (base[a] + 3 < 0 || base[a] >= k.N) return;  // untouched, like the reference's early return

  bool ok[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int g = base[a] + i;
      ok[a][i] = (g >= 0) && (g < k.N);
      if (a == 0) ok[a][i] = ok[a][i] && (g >= k.x0) && (g < k.x0 + k.nxl);
    }
  float d[3][3];  // node - particle distance per axis (world units)
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int i = 0; i < 3; ++i) d[a][i] = (float)(base[a] + i) * k.dx - x[a];

  const long long NN = (long long)k.N * k.N;
  const float4* gbase = grid + ((long long)(base[0] - k.x0) * NN + (long long)base[1] * k.N + base[2]);
  float v[3] = {0.f, 0.f, 0.f};
  Mat3 B;  // sum_i w v_i d_i^T, scaled by dinv at the end
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) B.m[r][c] = 0.f;
#if MPM_G2P_VARIANT == 0
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float wij = w[0][i] * w[1][j];
#pragma unroll
      for (int kz = 0; kz < 3; ++kz) {
        if (ok[0][i] && ok[1][j] && ok[2][kz]) {
          const float4 g = __ldg(gbase + ((long long)i * NN + j * k.N + kz));
          const float wt = wij * w[2][kz];
          const float wv[3] = {wt * g.x, wt * g.y, wt * g.z};
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            v[c] += wv[c];
            B.m[c][0] += wv[c] * d[0][i];
            B.m[c][1] += wv[c] * d[1][j];
            B.m[c][2] += wv[c] * d[2][kz];
          }
        }
      }
    }
  }
#else
  // Separable accumulation: for each (i,j) row, s0 = sum_k wz_k v_k and s1 = sum_k wz_k dz_k v_k
  // (the three k-nodes are one contiguous 48 B run), then one rank-1 update per row.  Nodes
  // outside the domain / slab load as zero instead of branching.
  float wzd[3], wxd[3], wyd[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    wxd[q] = w[0][q] * d[0][q];
    wyd[q] = w[1][q] * d[1][q];
    wzd[q] = w[2][q] * d[2][q];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float4* row = gbase + ((long long)i * NN + (long long)j * k.N);
      const bool okr = ok[0][i] && ok[1][j];
      float s0[3] = {0.f, 0.f, 0.f}, s1[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int kz = 0; kz < 3; ++kz) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (okr && ok[2][kz]) g = __ldg(row + kz);
        s0[0] += w[2][kz] * g.x; s0[1] += w[2][kz] * g.y; s0[2] += w[2][kz] * g.z;
        s1[0] += wzd[kz] * g.x;  s1[1] += wzd[kz] * g.y;  s1[2] += wzd[kz] * g.z;
      }
      const float wij = w[0][i] * w[1][j];
      const float wdx = wxd[i] * w[1][j];
      const float wdy = w[0][i] * wyd[j];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        v[c] += wij * s0[c];
        B.m[c][0] += wdx * s0[c];
        B.m[c][1] += wdy * s0[c];
        B.m[c][2] += wij * s1[c];
      }
    }
  }
#endif
  Mat3 C, G;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      C.m[r][c] = B.m[r][c] * k.dinv;
      G.m[r][c] = ((r == c) ? 1.0f : 0.0f) + k.dt * C.m[r][c];
    }
#if !MPM_G2P_PREFETCH
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) F.m[r][c] = p.s(SF + 3 * r + c)[pi];
#endif
  F = mul_ab(G, F);  // F <- (I + dt C) F
  if (MODEL == MPM_MODEL_SNOW) {
    float Jp = p.s(SJ)[pi];
    const MpmMaterial m = load_material(mats, p.mat[pi]);
    snow_plasticity<O>(F, Jp, m);
    p.s(SJ)[pi] = Jp;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    p.s(SX + a)[pi] = x[a] + k.dt * v[a];
    p.s(SV + a)[pi] = v[a];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      p.s(SF + 3 * r + c)[pi] = F.m[r][c];
      p.s(SC + 3 * r + c)[pi] = C.m[r][c];
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
  Mat3 r;  // the same routine compute_PF uses for this mode
  if constexpr (O::kExact) r = polar_rotation<O>(a); else r = polar_rotation_newton(a);
  for (int e = 0; e < 9; ++e) R[9 * i + e] = r.m[e / 3][e % 3];
}
__global__ void det_batch_kernel(const float* __restrict__ A, float* det, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Mat3 a;
  for (int e = 0; e < 9; ++e) a.m[e / 3][e % 3] = A[9 * i + e];
  det[i] = det3(a);
}

}  // namespace mpm
