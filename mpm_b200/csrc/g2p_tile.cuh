// Stage (4), production kernel: G2P as a persistent, bulk-copy-fed pipeline
// (reference behaviour: src/mpm.cu:109-178, TransferScheme.h:102-142), for the shipped transfer
// tuple MLS_APIC_Scheme<QuadraticInterpolationKernel> and any MaterialModel.
//
// Why: one thread per particle gathering 27 float4 nodes straight from L2 is latency-bound — 64
// registers hold ~8 of the 27 loads, so every particle pays 3-4 serial L2 round trips on top of
// the two DRAM round trips for x and F (profiles/r01_ncu_v3_g2p_direct.txt: issue active 48 %,
// long-scoreboard stalls dominate).  Here the memory system works ahead of the arithmetic:
//
//   * Work unit = one 256-particle tile of the SoA (common.cuh).  What G2P reads of it — the stream
//     rows x, F (and Jp) — is ONE contiguous span of the tile-major layout: a single 1-D bulk copy
//     (cp.async.bulk, the TMA path without a tensor map) into a shared-memory stage.
//   * CTA = 8 warps (one thread per particle of the tile) over a ring of kG2pStages stage buffers
//     with one "full" mbarrier each.  Thread 0 requests the first kG2pStages tiles; afterwards the
//     warp that releases a stage LAST (a shared-memory counter per stage tells it so) requests the
//     tile that goes there next, so the copy of tile it+S runs while tile it is computed and
//     no warp ever waits for another one, only for data.  After the first ring fill, tiles are handed
//     out by a global counter: the B200's two dies do not see the same memory latency and a static
//     share per CTA ends when the slowest SM is done.
//   * Every thread gathers its 27 nodes from global memory through L1 — the fastest form measured,
//     because a cell-sorted warp's LDG.128 touches one or two 128-byte lines (1-2 L1 wavefronts)
//     while an LDS.128 of a staged node box always costs four (profiles/r01_ab2_session6.txt).
//     Warps whose particles all have their stencil inside the local grid take a path with no
//     per-node predicates, in separable FFMA2 form; the others clip per node like the reference.
//   * EMIT (hand-over, common.cuh): after the F update and the plasticity, the affine matrix of the
//     NEXT substep's P2G is computed here — F, C, Jp and the material are in registers — and stored
//     in the C rows, so that P2G reads 15 instead of 25 streams and does not evaluate the material.
#pragma once
#include <cuda.h>

#include "kernels.cuh"
#include "p2g_sched.cuh"
#include "tma.cuh"

namespace mpm {

constexpr int kG2pStages = 3;  // (2 and 4 measured in round 1: no better)
constexpr int kG2pThreads = kTile;
constexpr int kG2pRows = NSTREAM - SX;  // stream rows x, F, Jp = 13, contiguous in a tile
static_assert(SX == 12 && SF == 15 && SJ == 24, "G2P reads stream rows 12..24 as one span");

#define MPM_STP(ptr, val) __stcs(ptr, val)  // particle streams are touched once per kernel: streaming stores

struct G2pTileLayout {
  // Jp is the last row: materials that never read it get a 12-row copy
  __host__ __device__ static constexpr size_t stage_bytes(bool with_jp) { return (size_t)(with_jp ? kG2pRows : kG2pRows - 1) * kTile * 4; }
  // stages + tile indices + barriers + release counters + slack for 128-byte alignment of the first stage
  __host__ __device__ static constexpr size_t bytes(bool with_jp) { return kG2pStages * (stage_bytes(with_jp) + 32) + 128; }
};

// generic gather with per-node clipping (domain faces, slab edges); B = sum_i w v_i d_i^T.
// Deliberately not inlined and fed by value: the rare clipped warps pay a call, the interior path
// keeps its registers.
struct G2pGather {
  float v[3];
  float B[3][3];
};
static __device__ __noinline__ G2pGather g2p_gather_clipped(const float4* __restrict__ grid, KParams k, float x0, float x1, float x2) {
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

// separable FFMA2 gather of the 27 stencil nodes starting at tp (row_stride / plane_stride in nodes):
// per (i, j) row s0 = sum_k wz_k v_k and s1 = sum_k wz_k dz_k v_k (the three k-nodes are one
// contiguous 48 B run), then one rank-1 update of (v, B) per row; x, y components travel as packed
// pairs (f32x2.cuh), z as scalars
struct G2pAcc {
  float v[3];
  Mat B;  // sum_i w v_i d_i^T
};
__device__ __forceinline__ void g2p_gather27(const float4* __restrict__ tp, long long row_stride, long long plane_stride,
                                             const float w[3][3], const float d[3][3], G2pAcc& o) {
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
      const float4* row = tp + (i * plane_stride + j * row_stride);
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
      // (forming these products as packed pairs from pre-duplicated weights saves the dup moves but costs
      // 12 registers: 2.02 -> 2.24 ms, profiles/r02_ab3_g2p.txt)
      const float wij = w[0][i] * w[1][j], wdx = wxd[i] * w[1][j], wdy = w[0][i] * wyd[j];
      const f2 WIJ = dup2(wij), WDX = dup2(wdx), WDY = dup2(wdy);
      Vxy = fma2(WIJ, s0, Vxy);
      vz = fmaf(wij, s0z, vz);
      B0xy = fma2(WDX, s0, B0xy);
      B0z = fmaf(wdx, s0z, B0z);
      B1xy = fma2(WDY, s0, B1xy);
      B1z = fmaf(wdy, s0z, B1z);
      B2xy = fma2(WIJ, s1, B2xy);
      B2z = fmaf(wij, s1z, B2z);
    }
  }
  o.v[0] = lo2(Vxy); o.v[1] = hi2(Vxy); o.v[2] = vz;
  o.B.m[0][0] = lo2(B0xy); o.B.m[1][0] = hi2(B0xy); o.B.m[2][0] = B0z;
  o.B.m[0][1] = lo2(B1xy); o.B.m[1][1] = hi2(B1xy); o.B.m[2][1] = B1z;
  o.B.m[0][2] = lo2(B2xy); o.B.m[1][2] = hi2(B2xy); o.B.m[2][2] = B2z;
}

// one particle of a staged tile: ps = this particle's column of the stage (rows x, F, Jp), col = its
// column of the tile in HBM
template <class Material, bool ONE_MAT, bool EMIT>
__device__ __forceinline__ void g2p_tile_compute(const Soa& p, const MatTable<Material>& mats, const float4* __restrict__ grid, const KParams& k,
                                                 const float* __restrict__ ps, float* __restrict__ col, size_t pi, bool live, bool with_jp,
                                                 bool count_moved, unsigned& moved) {
  float x[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) x[a] = live ? ps[a * kTile] : 0.f;
  int base[3];
  float fx[3], w[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a) bspline(x[a], k.dx_inv, base[a], fx[a], w[a]);
  const bool valid = live && !stencil_outside(base, k.N);  // else untouched (reference early return)
  const int bxl = base[0] - k.x0;                           // x-plane in the local grid
  bool interior = bxl >= 0 && bxl + 2 < k.nxl;
#pragma unroll
  for (int a = 0; a < 3; ++a) interior = interior && base[a] >= 0 && base[a] + 2 < k.N;
  if (!valid) return;

  G2pAcc acc;
  if (interior) {  // whole stencil inside the local grid: unclipped gather
    float d[3][3];  // node - particle distance per axis (world units)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int i = 0; i < 3; ++i) d[a][i] = (float)(base[a] + i) * k.dx - x[a];
    const long long NN = (long long)k.N * k.N;
    g2p_gather27(grid + (bxl * NN + (long long)base[1] * k.N + base[2]), k.N, NN, w, d, acc);
  } else {
    const G2pGather o = g2p_gather_clipped(grid, k, x[0], x[1], x[2]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      acc.v[c] = o.v[c];
#pragma unroll
      for (int a = 0; a < 3; ++a) acc.B.m[c][a] = o.B[c][a];
    }
  }
  Particle part;
  part.material_type = 0;
  Mat G;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      part.C.m[r][c] = acc.B.m[r][c] * k.dinv;
      G.m[r][c] = ((r == c) ? 1.0f : 0.0f) + k.dt * part.C.m[r][c];
      part.F.m[r][c] = ps[(SF - SX + 3 * r + c) * kTile];
    }
  part.F = G * part.F;  // F <- (I + dt C) F  (g2p_finish_particle)
  constexpr bool kJp = MaterialTraits<Material>::kMutatesJp;
  part.Jp = 1.0f;
  if (kJp) part.Jp = ps[(SJ - SX) * kTile];
  else if (EMIT && with_jp) part.Jp = col[SJ * kTile];  // read-only Jp != 1 of a fixed-corotated handle
  const Material m = mats.template get<ONE_MAT>(p.mat, pi);
  // hand-over: a material whose end-of-step hook decomposes F anyway (MMSnow) passes the rotation on to the
  // stress of the next substep: one SVD per particle-step instead of the reference's two
  constexpr bool kShareR = EMIT && HasRotationHooks<Material>::value;
  Mat R;
  bool have_R = false;
  if constexpr (kShareR) have_R = m.endOfStepMutationR(part, R); else m.endOfStepMutation(part);
  if (kJp) MPM_STP(col + SJ * kTile, part.Jp);
  bool crossed = false;  // did the advection take the particle into another cell (rebin_permille, mpm_b200.h)
  int nbase[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float xn = x[a] + k.dt * acc.v[a];
    part.x(a) = xn;
    part.v(a) = acc.v[a];
    nbase[a] = (int)(xn * k.dx_inv - 0.5f);
    crossed = crossed || (nbase[a] != base[a]);
    MPM_STP(col + (SX + a) * kTile, xn);
    MPM_STP(col + (SV + a) * kTile, acc.v[a]);
  }
  if (count_moved) moved += crossed ? 1u : 0u;
  Mat out = part.C;
  if (EMIT) {
    // the next P2G skips a particle whose stencil has left the domain, and so does every G2P after
    // it: such a particle keeps C (the state the reference would hold), everyone else gets dx * affine
    if (!stencil_outside(nbase, k.N)) {
      if constexpr (kShareR) {
        if (have_R) out = p2g_affine_dx_from_PF(m.computePF_R(part, R), part, m, k);
        else out = p2g_affine_dx(part, m, k);
      } else {
        out = p2g_affine_dx(part, m, k);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      MPM_STP(col + (SF + 3 * r + c) * kTile, part.F.m[r][c]);
      MPM_STP(col + (SC + 3 * r + c) * kTile, out.m[r][c]);
    }
}

// Persistent CTAs: the first kG2pStages tiles of CTA b are b, b + gridDim.x, ...; afterwards tiles come
// from tile_counters[parity] (the other counter is reset for the next launch).
template <class Material, bool ONE_MAT, bool EMIT>
__global__ void __launch_bounds__(kG2pThreads, 4)
g2p_tile_kernel(Soa p, const MatTable<Material> mats, const float4* __restrict__ grid, KParams k, size_t count,
                unsigned long long* __restrict__ moved_total, unsigned int* __restrict__ tile_counters, int parity,
                const DeviceDiag* __restrict__ diag) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);  // bulk-copy destinations: 128 B aligned
  constexpr bool kJp = MaterialTraits<Material>::kMutatesJp;
  constexpr size_t kStage = G2pTileLayout::stage_bytes(kJp);
  uint32_t* tile_of = reinterpret_cast<uint32_t*>(smem + kG2pStages * kStage);  // tile index in each stage, 0xffffffff = end
  uint64_t* full = reinterpret_cast<uint64_t*>(tile_of + 2 * kG2pStages);
  uint32_t* rel = reinterpret_cast<uint32_t*>(full + kG2pStages);              // warps done with each stage
  const int tid = threadIdx.x;
  const uint32_t n_tiles = (uint32_t)((count + kTile - 1) / kTile);
  const bool with_jp = kJp || (EMIT && diag->jp_not_one != 0);

  // one bulk copy: rows x, F (, Jp) of tile t into stage s
  auto issue = [&](uint32_t t, int s) {
    tile_of[s] = t;
    mbar_arrive_expect_tx(full + s, (uint32_t)kStage);
    bulk_g2s_hint(smem + s * kStage, p.tile(t) + SX * kTile, (uint32_t)kStage, full + s, l2_evict_first_policy());
  };
  if (tid == 0) {
    for (int s = 0; s < kG2pStages; ++s) {
      mbar_init(full + s, 1);
      rel[s] = 0;
    }
    mbar_fence_init();
    for (int s = 0; s < kG2pStages; ++s) {
      const uint32_t t = blockIdx.x + (uint32_t)s * gridDim.x;
      if (t < n_tiles) {
        issue(t, s);
      } else {  // end marker
        tile_of[s] = 0xffffffffu;
        mbar_arrive(full + s);
      }
    }
    if (blockIdx.x == 0) tile_counters[parity ^ 1] = 0;  // the next launch's counter
  }
  __syncthreads();

  unsigned moved = 0;  // particles of this thread that changed cell in this substep
  for (int it = 0;; ++it) {
    const int s = it % kG2pStages;
    mbar_wait(full + s, (uint32_t)((it / kG2pStages) & 1));
    const uint32_t t = tile_of[s];
    if (t == 0xffffffffu) break;  // no more tiles for this CTA
    const size_t pi = (size_t)t * kTile + tid;
    g2p_tile_compute<Material, ONE_MAT, EMIT>(p, mats, grid, k, reinterpret_cast<const float*>(smem + s * kStage) + tid, p.tile(t) + tid, pi,
                                              pi < count, with_jp, moved_total != nullptr, moved);
    __syncwarp();  // the warp is done with stage s
    if ((tid & 31) == 0) {
      __threadfence_block();
      if (atomicAdd(&rel[s], 1u) == kTile / 32 - 1) {  // last warp out refills the stage
        rel[s] = 0;
        __threadfence_block();
        const uint32_t tn = (uint32_t)kG2pStages * gridDim.x + atomicAdd(&tile_counters[parity], 1u);
        if (tn < n_tiles) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the warps' reads of the stage before the copy engine's writes
          issue(tn, s);
        } else {
          tile_of[s] = 0xffffffffu;
          mbar_arrive(full + s);
        }
      }
    }
  }
  if (moved_total) {
    moved = __reduce_add_sync(0xffffffffu, moved);
    if ((tid & 31) == 0 && moved) atomicAdd(moved_total, (unsigned long long)moved);
  }
}

}  // namespace mpm
