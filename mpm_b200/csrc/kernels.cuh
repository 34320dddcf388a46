// Substep kernels: P2G scatter, grid update, G2P gather (sm_100a).
// Reference behaviour: src/mpm.cu:14-178 + include/TransferScheme.h:66-142.
#pragma once
#include "common.cuh"
#include "f32x2.cuh"

namespace mpm {

constexpr int kParticleBlock = 128;

// ---- particle <-> AoS conversion (boundary only, off the hot path) ---------------------------
// aos[0..count) -> slots [offset, offset + count), ids first_id + slot
__global__ void aos_to_soa_kernel(const MpmParticle* __restrict__ aos, Soa p, size_t count, uint32_t first_id, size_t offset = 0) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const MpmParticle& q = aos[i];
  i += offset;
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

// slot r <- aos[id[r] - first_id]: new particle data into the existing (cell-sorted) slots
__global__ void aos_overwrite_kernel(const MpmParticle* __restrict__ aos, Soa p, size_t count, uint32_t first_id) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const MpmParticle& q = aos[p.id[i] - first_id];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    p.s(SX + a)[i] = q.x[a];
    p.s(SV + a)[i] = q.v[a];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      p.s(SF + 3 * r + c)[i] = q.F[3 * c + r];
      p.s(SC + 3 * r + c)[i] = q.C[3 * c + r];
    }
  p.s(SJ)[i] = q.Jp;
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
// One thread per particle (reference src/mpm.cu:109-178, TransferScheme.h:102-142).
//
// The kernel is issue-bound before it is HBM-bound, so the gather is written for few issue slots:
//   * separable along z: per (i,j) row s0 = sum_k wz_k v_k and s1 = sum_k wz_k dz_k v_k (the three
//     k-nodes are one contiguous 48 B run), then one rank-1 update of (v, B) per row;
//   * the x,y components travel as packed pairs (FFMA2, f32x2.cuh), z as scalars;
//   * warps whose 32 particles all have their whole stencil inside the local grid (every warp away
//     from the domain faces) take a path with no per-node predicates; the others take the generic
//     per-node path below, which clips like the reference.
#ifndef MPM_G2P_MINBLK
#define MPM_G2P_MINBLK 8    // __launch_bounds__ min blocks per SM (64 regs: occupancy beats ILP here, tools/ab.py)
#endif
#ifndef MPM_G2P_BLOCK
#define MPM_G2P_BLOCK 128
#endif
#ifndef MPM_G2P_STREAMING
#define MPM_G2P_STREAMING 1  // particle streams are touched once per kernel: evict-first loads / streaming stores
#endif
#ifndef MPM_G2P_PREFETCH
#define MPM_G2P_PREFETCH 0   // 1: issue the F loads together with the x loads
#endif
#if MPM_G2P_STREAMING
#define MPM_LDP(ptr) __ldcs(ptr)
#define MPM_STP(ptr, val) __stcs(ptr, val)
#else
#define MPM_LDP(ptr) (*(ptr))
#define MPM_STP(ptr, val) (*(ptr) = (val))
#endif
constexpr int kG2pBlock = MPM_G2P_BLOCK;

// generic gather with per-node clipping (domain faces, slab edges); B = sum_i w v_i d_i^T.
// Deliberately not inlined and fed by value: the rare clipped warps pay a call, the interior path
// keeps its registers.
struct G2pGather {
  float v[3];
  float B[3][3];
};
__device__ __noinline__ G2pGather g2p_gather_clipped(const float4* __restrict__ grid, KParams k, float x0, float x1, float x2) {
  const float x[3] = {x0, x1, x2};
  int base[3];
  float fx[3], w[3][3], d[3][3];
  for (int a = 0; a < 3; ++a) {
    bspline(x[a], k.dx_inv, base[a], fx[a], w[a]);
    for (int i = 0; i < 3; ++i) d[a][i] = (float)(base[a] + i) * k.dx - x[a];
  }
  G2pGather o;
  for (int c = 0; c < 3; ++c) {
    o.v[c] = 0.f;
    for (int a = 0; a < 3; ++a) o.B[c][a] = 0.f;
  }
  const long long NN = (long long)k.N * k.N;
  const float4* gbase = grid + ((long long)(base[0] - k.x0) * NN + (long long)base[1] * k.N + base[2]);
  for (int i = 0; i < 3; ++i) {
    const int gx = base[0] + i;
    if (gx < 0 || gx >= k.N || gx < k.x0 || gx >= k.x0 + k.nxl) continue;
    for (int j = 0; j < 3; ++j) {
      const int gy = base[1] + j;
      if (gy < 0 || gy >= k.N) continue;
      const float wij = w[0][i] * w[1][j];
      for (int kz = 0; kz < 3; ++kz) {
        const int gz = base[2] + kz;
        if (gz < 0 || gz >= k.N) continue;
        const float4 g = __ldg(gbase + ((long long)i * NN + j * k.N + kz));
        const float wt = wij * w[2][kz];
        const float wv[3] = {wt * g.x, wt * g.y, wt * g.z};
        for (int c = 0; c < 3; ++c) {
          o.v[c] += wv[c];
          o.B[c][0] += wv[c] * d[0][i];
          o.B[c][1] += wv[c] * d[1][j];
          o.B[c][2] += wv[c] * d[2][kz];
        }
      }
    }
  }
  return o;
}

template <int MODEL, class O>
__global__ void __launch_bounds__(kG2pBlock, MPM_G2P_MINBLK)
g2p_kernel(Soa p, size_t count, const MpmMaterial* __restrict__ mats, const float4* __restrict__ grid, KParams k) {
  const size_t pi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = pi < count;
  float x[3] = {0.f, 0.f, 0.f};
  if (live) {
#pragma unroll
    for (int a = 0; a < 3; ++a) x[a] = MPM_LDP(p.s(SX + a) + pi);
  }
  Mat3 F;
#if MPM_G2P_PREFETCH
  if (live) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) F.m[r][c] = MPM_LDP(p.s(SF + 3 * r + c) + pi);
  }
#endif
  int base[3];
  float fx[3], w[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a) bspline(x[a], k.dx_inv, base[a], fx[a], w[a]);
  bool inside = live, interior = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    inside = inside && !(base[a] + 3 < 0 || base[a] >= k.N);  // else untouched, like the reference's early return
    interior = interior && base[a] >= 0 && base[a] + 2 < k.N;
  }
  interior = interior && base[0] >= k.x0 && base[0] + 2 < k.x0 + k.nxl;
  const bool fast = __all_sync(0xffffffffu, interior || !inside);
  if (!inside) return;

  float v[3];
  Mat3 B;  // sum_i w v_i d_i^T, scaled by dinv at the end
  if (fast) {
    float d[3][3];  // node - particle distance per axis (world units)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int i = 0; i < 3; ++i) d[a][i] = (float)(base[a] + i) * k.dx - x[a];
    const long long NN = (long long)k.N * k.N;
    const float4* gbase = grid + ((long long)(base[0] - k.x0) * NN + (long long)base[1] * k.N + base[2]);
    float wzd[3], wxd[3], wyd[3];
    f2 WZ[3], WZD[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      wxd[q] = w[0][q] * d[0][q];
      wyd[q] = w[1][q] * d[1][q];
      wzd[q] = w[2][q] * d[2][q];
      WZ[q] = dup2(w[2][q]);
      WZD[q] = dup2(wzd[q]);
    }
    f2 Vxy = pack2(0.f, 0.f), B0xy = Vxy, B1xy = Vxy, B2xy = Vxy;  // B?xy = (B[0][?], B[1][?])
    float vz = 0.f, B0z = 0.f, B1z = 0.f, B2z = 0.f;                // B?z  = B[2][?]
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float4* row = gbase + ((long long)i * NN + (long long)j * k.N);
        const float4 g0 = __ldg(row), g1 = __ldg(row + 1), g2 = __ldg(row + 2);
        f2 s0 = mul2(WZ[0], pack2(g0.x, g0.y));
        f2 s1 = mul2(WZD[0], pack2(g0.x, g0.y));
        float s0z = w[2][0] * g0.z, s1z = wzd[0] * g0.z;
        s0 = fma2(WZ[1], pack2(g1.x, g1.y), s0);
        s1 = fma2(WZD[1], pack2(g1.x, g1.y), s1);
        s0z = fmaf(w[2][1], g1.z, s0z);
        s1z = fmaf(wzd[1], g1.z, s1z);
        s0 = fma2(WZ[2], pack2(g2.x, g2.y), s0);
        s1 = fma2(WZD[2], pack2(g2.x, g2.y), s1);
        s0z = fmaf(w[2][2], g2.z, s0z);
        s1z = fmaf(wzd[2], g2.z, s1z);
        const float wij = w[0][i] * w[1][j], wdx = wxd[i] * w[1][j], wdy = w[0][i] * wyd[j];
        const f2 WIJ = dup2(wij);
        Vxy = fma2(WIJ, s0, Vxy);
        vz = fmaf(wij, s0z, vz);
        B0xy = fma2(dup2(wdx), s0, B0xy);
        B0z = fmaf(wdx, s0z, B0z);
        B1xy = fma2(dup2(wdy), s0, B1xy);
        B1z = fmaf(wdy, s0z, B1z);
        B2xy = fma2(WIJ, s1, B2xy);
        B2z = fmaf(wij, s1z, B2z);
      }
    }
    v[0] = lo2(Vxy); v[1] = hi2(Vxy); v[2] = vz;
    B.m[0][0] = lo2(B0xy); B.m[1][0] = hi2(B0xy); B.m[2][0] = B0z;
    B.m[0][1] = lo2(B1xy); B.m[1][1] = hi2(B1xy); B.m[2][1] = B1z;
    B.m[0][2] = lo2(B2xy); B.m[1][2] = hi2(B2xy); B.m[2][2] = B2z;
  } else {
    const G2pGather o = g2p_gather_clipped(grid, k, x[0], x[1], x[2]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v[c] = o.v[c];
#pragma unroll
      for (int a = 0; a < 3; ++a) B.m[c][a] = o.B[c][a];
    }
  }
  Mat3 C, G;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      C.m[r][c] = B.m[r][c] * k.dinv;
      G.m[r][c] = ((r == c) ? 1.0f : 0.0f) + k.dt * C.m[r][c];
#if !MPM_G2P_PREFETCH
      F.m[r][c] = MPM_LDP(p.s(SF + 3 * r + c) + pi);
#endif
    }
  F = mul_ab(G, F);  // F <- (I + dt C) F
  if (MODEL == MPM_MODEL_SNOW) {
    float Jp = MPM_LDP(p.s(SJ) + pi);
    const MpmMaterial m = load_material(mats, p.mat[pi]);
    snow_plasticity<O>(F, Jp, m);
    MPM_STP(p.s(SJ) + pi, Jp);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    MPM_STP(p.s(SX + a) + pi, x[a] + k.dt * v[a]);
    MPM_STP(p.s(SV + a) + pi, v[a]);
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      MPM_STP(p.s(SF + 3 * r + c) + pi, F.m[r][c]);
      MPM_STP(p.s(SC + 3 * r + c) + pi, C.m[r][c]);
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
