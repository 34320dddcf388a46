// Stage (4), production kernel: G2P as a persistent, TMA-fed pipeline
// (reference behaviour: src/mpm.cu:109-178, TransferScheme.h:102-142).
//
// Why: one thread per particle gathering 27 float4 nodes straight from L2 is latency-bound — 64
// registers hold ~8 of the 27 loads, so every particle pays 3-4 serial L2 round trips on top of
// the two DRAM round trips for x and F (profiles/r01_ncu_v3_g2p_direct.txt: issue active 48 %,
// long-scoreboard stalls dominate).  Here the memory system works ahead of the arithmetic:
//
//   * Work unit = a tile (tiles.cuh): <= 252 cell-sorted particles of one grid row.  The grid nodes
//     they can touch form one box (x-1..x+3) x (y-1..y+3) x (z_first-1 .. +LT): ONE 4-D tiled TMA
//     load (cp.async.bulk.tensor) of 5 x 5 x LT float4.  The particle streams G2P reads (x, F and,
//     for snow, Jp) are rows 0..12 of the [25][stride] stream tensor: ONE 2-D TMA load.  Nodes
//     outside the local grid arrive as zeros, which is exactly the reference's stencil clipping
//     for a gather (src/mpm.cu:137-142), so domain faces and slab edges need no special case.
//   * CTA = 8 warps (one thread per particle of the tile) over a ring of kG2pStages stage buffers
//     with one "full" mbarrier each.  Thread 0 requests the first kG2pStages tiles; afterwards the
//     warp that releases a stage LAST (a shared-memory counter per stage tells it so) requests the
//     tile that goes there next, so the two TMA loads of tile it+S run while tile it is computed and
//     no warp ever waits for another one, only for data.  (A dedicated producer warp did the same
//     job at first — MPM_G2P_SELFFEED=0 — but made the CTA 288 threads: 3 CTAs per SM instead of 4.)
//   * A particle finds its stencil at box[(di+i)*5 + (dj+j)][t+k] where (di, dj, t) is its current
//     base node relative to the box origin.  The box has one cell of slack on every side, so a
//     particle that drifted at most one cell in any direction since the re-bin is still inside;
//     the rest take the generic global-memory gather, which clips like the reference.
//   * The gather is the separable FFMA2 form (kernels.cuh).
//   * Node source (MPM_G2P_GATHER): 1 (default) = every thread gathers its 27 nodes from global
//     memory through L1 — the fastest form measured, because a cell-sorted warp's LDG.128 touches one
//     or two 128-byte lines (1-2 L1 wavefronts) while an LDS.128 always costs four; 0 = the TMA node
//     box described above; 2 = a per-warp brick in shared memory (DESIGN.md 3.1 has the numbers).
#pragma once
#include <cuda.h>

#include "kernels.cuh"
#include "tiles.cuh"
#include "tma.cuh"

namespace mpm {

#ifndef MPM_G2P_BOXW
#define MPM_G2P_BOXW 5  // node box of a tile: BOXW x BOXW rows (5 = one row of slack either side, 3 = exactly the stencil rows)
#endif
constexpr int kBoxW = MPM_G2P_BOXW;
constexpr int kBoxSlack = (kBoxW - 3) / 2;
constexpr int kBoxRows = kBoxW * kBoxW;
#ifndef MPM_G2P_STAGES
#define MPM_G2P_STAGES 3
#endif
#ifndef MPM_G2P_TILE_MINBLK
#define MPM_G2P_TILE_MINBLK 3
#endif
#ifndef MPM_G2P_GATHER
#define MPM_G2P_GATHER 1  // node source: 0 = 5x5xLT TMA box per tile, 1 = global memory per thread, 2 = per-warp brick
#endif
#define MPM_G2P_BOX (MPM_G2P_GATHER == 0)
// With the per-thread global gather nothing ties a tile to one grid row, so tiles are simply the
// 256-particle blocks of the sorted order: full lanes (row tiles average 228 of 256 particles on the
// benchmark block) and no tile descriptors.  The shared-memory node sources need row tiles.
#ifndef MPM_G2P_DYNAMIC
#define MPM_G2P_DYNAMIC 1  // 1: after the first ring fill, tiles are handed out by a global counter (needs SELFFEED)
#endif
#ifndef MPM_G2P_FLAT_TILES
#define MPM_G2P_FLAT_TILES (MPM_G2P_GATHER == 1)
#endif
constexpr int kWarpBrickX = 4, kWarpBrickY = 4, kWarpBrickZ = 16;  // nodes; 4 KB of shared memory per warp
constexpr int kG2pStages = MPM_G2P_STAGES;
#ifndef MPM_G2P_SELFFEED
#define MPM_G2P_SELFFEED 1  // 1: no producer warp — the last warp to release a stage refills it (256 threads, 64 registers, 4 CTAs/SM)
#endif
constexpr int kG2pThreads = MPM_G2P_SELFFEED ? kTile : kTile + 32;  // consumers (+ one producer warp)

struct TileHeader {  // written by the producer next to each stage
  int x0b, y0b, z0b;   // box origin in local grid coordinates (may be -1)
  int n;               // particles in the tile
  unsigned int start;  // first particle slot
  int off;             // start & 3: column of the first particle in the stream box
  int pad_[2];
};

template <int MODEL>
struct G2pTileLayout {
  static constexpr int kStreams = (MODEL == MPM_MODEL_SNOW) ? 13 : 12;  // x3, F9 (, Jp) = stream rows 0..kStreams-1
  __host__ __device__ static constexpr size_t box_bytes(int LT) { return MPM_G2P_BOX ? (size_t)kBoxRows * LT * 16 : 0; }
  __host__ __device__ static constexpr size_t stream_bytes() { return (size_t)kStreams * kTile * 4; }
  __host__ __device__ static constexpr size_t stage_bytes(int LT) { return box_bytes(LT) + stream_bytes(); }
  __host__ __device__ static constexpr size_t brick_bytes() {
    return MPM_G2P_GATHER == 2 ? (size_t)(kTile / 32) * kWarpBrickX * kWarpBrickY * kWarpBrickZ * 16 : 0;
  }
  // stages + per-warp bricks + headers + barriers + slack for 128-byte alignment of the first stage
  __host__ __device__ static constexpr size_t bytes(int LT) {
    return kG2pStages * (stage_bytes(LT) + sizeof(TileHeader) + 16) + brick_bytes() + 128;
  }
};

// separable FFMA2 gather of the 27 stencil nodes starting at tp (row_stride / plane_stride in nodes)
struct G2pAcc {
  float v[3];
  Mat3 B;  // sum_i w v_i d_i^T
};
template <bool GLOBAL>
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
      float4 g0, g1, g2;
      if (GLOBAL) {
        g0 = __ldg(row); g1 = __ldg(row + 1); g2 = __ldg(row + 2);
      } else {
        g0 = row[0]; g1 = row[1]; g2 = row[2];
      }
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
  o.v[0] = lo2(Vxy); o.v[1] = hi2(Vxy); o.v[2] = vz;
  o.B.m[0][0] = lo2(B0xy); o.B.m[1][0] = hi2(B0xy); o.B.m[2][0] = B0z;
  o.B.m[0][1] = lo2(B1xy); o.B.m[1][1] = hi2(B1xy); o.B.m[2][1] = B1z;
  o.B.m[0][2] = lo2(B2xy); o.B.m[1][2] = hi2(B2xy); o.B.m[2][2] = B2z;
}

template <int MODEL, class O, int LT, bool COUNT_MOVED>
__device__ __forceinline__ void g2p_tile_compute(const Soa& p, const MpmMaterial* __restrict__ mats, const MpmMaterial& mat0, bool one_mat,
                                                 const float4* __restrict__ grid,
                                                 const KParams& k, const TileHeader& h, const float4* __restrict__ box,
                                                 float4* __restrict__ wtile, const float* __restrict__ ps, int tid, unsigned& moved) {
  const bool live = tid < h.n;
  const size_t pi = (size_t)h.start + tid;
  uint8_t mat_id = 0;
  if (MODEL == MPM_MODEL_SNOW && live && !one_mat) mat_id = p.mat[pi];
  ps += h.off + tid;  // this particle's column of the stream box
  float x[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) x[a] = live ? ps[a * kTile] : 0.f;
  int base[3];
  float fx[3], w[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a) bspline(x[a], k.dx_inv, base[a], fx[a], w[a]);
  bool valid = live;
#pragma unroll
  for (int a = 0; a < 3; ++a) valid = valid && !(base[a] + 3 < 0 || base[a] >= k.N);  // else untouched (reference early return)
  const int bxl = base[0] - k.x0;  // x-plane in the local grid
  float d[3][3];  // node - particle distance per axis (world units)
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int i = 0; i < 3; ++i) d[a][i] = (float)(base[a] + i) * k.dx - x[a];

  G2pAcc acc;
  bool gathered = false;
#if MPM_G2P_GATHER == 0
  // ---- nodes from the tile's TMA box: stencil = box rows (di..di+2, dj..dj+2), nodes t..t+2 ----
  {
    const int di = bxl - h.x0b, dj = base[1] - h.y0b, t = base[2] - h.z0b;
    if (valid && (unsigned)di <= (unsigned)(kBoxW - 3) && (unsigned)dj <= (unsigned)(kBoxW - 3) && (unsigned)t <= (unsigned)(LT - 3)) {
      g2p_gather27<false>(box + (di * kBoxW + dj) * LT + t, LT, kBoxW * LT, w, d, acc);
      gathered = true;
    }
  }
#elif MPM_G2P_GATHER == 2
  // ---- nodes staged per warp: the 32 particles of a warp are neighbours in the sorted order, so
  // their stencils cover a small brick.  Bounding box by redux, every lane fetches its share of
  // the brick (one batch of independent loads = ONE L2 round trip per warp instead of the 6-7
  // register-limited batches of a per-thread gather), then the stencil is read from shared memory.
  {
    constexpr int XT = kWarpBrickX, YT = kWarpBrickY, ZT = kWarpBrickZ;
    const int big = 0x3fffffff;
    const int mnx = __reduce_min_sync(0xffffffffu, valid ? bxl : big), mxx = __reduce_max_sync(0xffffffffu, valid ? bxl : -big);
    const int mny = __reduce_min_sync(0xffffffffu, valid ? base[1] : big), mxy = __reduce_max_sync(0xffffffffu, valid ? base[1] : -big);
    const int mnz = __reduce_min_sync(0xffffffffu, valid ? base[2] : big), mxz = __reduce_max_sync(0xffffffffu, valid ? base[2] : -big);
    const int ex = mxx - mnx + 3, ey = mxy - mny + 3, ez = mxz - mnz + 3;  // brick extent in nodes
    if (mxx >= mnx && ex <= XT && ey <= YT && ez <= ZT) {
      const int lane = tid & 31;
      const long long NN = (long long)k.N * k.N;
      constexpr int kIt = XT * YT * ZT / 32;
      float4 g[kIt];
      // all loads first (independent, predicated, no branches), then all stores: one round trip
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int slot = it * 32 + lane;
        const int iz = slot % ZT, iy = (slot / ZT) % YT, ix = slot / (ZT * YT);
        const int gx = mnx + ix, gy = mny + iy, gz = mnz + iz;
        const bool need = ix < ex && iy < ey && iz < ez;
        // outside the local grid: zero = the reference's clipping
        const bool inb = need && (unsigned)gx < (unsigned)k.nxl && (unsigned)gy < (unsigned)k.N && (unsigned)gz < (unsigned)k.N;
        const float4* src = grid + (inb ? (gx * NN + (long long)gy * k.N + gz) : 0ll);
        g[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inb) g[it] = __ldg(src);
      }
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int slot = it * 32 + lane;
        const int iz = slot % ZT, iy = (slot / ZT) % YT, ix = slot / (ZT * YT);
        if (ix < ex && iy < ey && iz < ez) wtile[slot] = g[it];
      }
      __syncwarp();
      if (valid) {
        g2p_gather27<false>(wtile + ((bxl - mnx) * YT + (base[1] - mny)) * ZT + (base[2] - mnz), ZT, YT * ZT, w, d, acc);
        gathered = true;
      }
    }
  }
#endif
  if (!valid) return;
#if defined(MPM_G2P_EXP) && (MPM_G2P_EXP & 1)  // experiment: no gather at all
  gathered = true;
  for (int c = 0; c < 3; ++c) { acc.v[c] = x[c]; for (int a = 0; a < 3; ++a) acc.B.m[c][a] = w[c][a]; }
#endif
  if (!gathered) {
    bool interior = bxl >= 0 && bxl + 2 < k.nxl;
#pragma unroll
    for (int a = 0; a < 3; ++a) interior = interior && base[a] >= 0 && base[a] + 2 < k.N;
    if (interior) {  // whole stencil inside the local grid: unclipped gather from global memory
      const long long NN = (long long)k.N * k.N;
      g2p_gather27<true>(grid + (bxl * NN + (long long)base[1] * k.N + base[2]), k.N, NN, w, d, acc);
    } else {
      const G2pGather o = g2p_gather_clipped(grid, k, x[0], x[1], x[2]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        acc.v[c] = o.v[c];
#pragma unroll
        for (int a = 0; a < 3; ++a) acc.B.m[c][a] = o.B[c][a];
      }
    }
  }
  const float* v = acc.v;
  const Mat3& B = acc.B;
  Mat3 C, G, F;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      C.m[r][c] = B.m[r][c] * k.dinv;
      G.m[r][c] = ((r == c) ? 1.0f : 0.0f) + k.dt * C.m[r][c];
      F.m[r][c] = ps[(SF + 3 * r + c) * kTile];
    }
  F = mul_ab(G, F);  // F <- (I + dt C) F
  if (MODEL == MPM_MODEL_SNOW) {
    float Jp = ps[SJ * kTile];
    // single-material handles read the clamps straight from the kernel parameters (uniform branch)
    const MpmMaterial m = one_mat ? mat0 : load_material(mats, mat_id);
    snow_plasticity<O>(F, Jp, m);
    MPM_STP(p.s(SJ) + pi, Jp);
  }
  bool crossed = false;  // did the advection take the particle into another cell (rebin_permille, mpm_b200.h)
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float xn = x[a] + k.dt * v[a];
    if (COUNT_MOVED) crossed = crossed || ((int)(xn * k.dx_inv - 0.5f) != base[a]);
    MPM_STP(p.s(SX + a) + pi, xn);
    MPM_STP(p.s(SV + a) + pi, v[a]);
  }
  if (COUNT_MOVED) moved += crossed ? 1u : 0u;
#if defined(MPM_G2P_EXP) && (MPM_G2P_EXP & 2)  // experiment: 6 of the 24 output streams only
  if (F.m[0][0] + C.m[1][1] + F.m[2][2] + C.m[0][2] == 12345.f) MPM_STP(p.s(SF) + pi, 0.f);
#else
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      MPM_STP(p.s(SF + 3 * r + c) + pi, F.m[r][c]);
      MPM_STP(p.s(SC + 3 * r + c) + pi, C.m[r][c]);
    }
#endif
}

// Persistent CTAs: tile `it` of this CTA = blockIdx.x + it * gridDim.x.
template <int MODEL, class O, int LT, bool COUNT_MOVED>
__global__ void __launch_bounds__(kG2pThreads, MPM_G2P_SELFFEED ? 4 : MPM_G2P_TILE_MINBLK)
g2p_tile_kernel(Soa p, const MpmMaterial* __restrict__ mats, const MpmMaterial mat0, const bool one_mat, const float4* __restrict__ grid, KParams k,
                const TileDesc* __restrict__ tiles, const uint32_t* __restrict__ n_tiles_ptr,
                const __grid_constant__ CUtensorMap tm_grid, const __grid_constant__ CUtensorMap tm_streams,
                unsigned long long* __restrict__ moved_total, size_t count, unsigned int* __restrict__ tile_counters, int parity) {
  using L = G2pTileLayout<MODEL>;
  static_assert(SX == 0 && SF == 3 && SJ == 12, "G2P reads stream rows 0..12 as one TMA box");
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);  // TMA destinations: 128 B aligned
  constexpr size_t kStage = L::stage_bytes(LT);
  float4* bricks = reinterpret_cast<float4*>(smem + kG2pStages * kStage);
  TileHeader* hdr = reinterpret_cast<TileHeader*>(smem + kG2pStages * kStage + L::brick_bytes());
  uint64_t* full = reinterpret_cast<uint64_t*>(hdr + kG2pStages);
  uint64_t* empty = full + kG2pStages;
  const int tid = threadIdx.x;
  const uint32_t n_tiles = MPM_G2P_FLAT_TILES ? (uint32_t)((count + kTile - 1) / kTile) : *n_tiles_ptr;
  if (tid == 0) {
    for (int s = 0; s < kG2pStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, kTile / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  // one TMA request: header + stream rows (+ node box) of tile t into stage s
  auto issue = [&](uint32_t t, int s) {
    TileDesc d;
    if (MPM_G2P_FLAT_TILES) {
      d.start = t * (uint32_t)kTile;
      d.n = (uint32_t)min((size_t)kTile, count - (size_t)d.start);
      d.kfirst = d.klast = 0;
    } else {
      d = tiles[t];
    }
    const uint32_t row = d.kfirst / (uint32_t)k.N;
    TileHeader h;
    h.z0b = (int)(d.kfirst - row * (uint32_t)k.N) - 1;
    h.x0b = (int)(row / (uint32_t)k.N);
    h.y0b = (int)(row - (uint32_t)h.x0b * (uint32_t)k.N) - kBoxSlack;
    h.x0b -= kBoxSlack;
    h.n = (int)d.n;
    h.start = d.start;
    h.off = (int)(d.start & 3u);
    hdr[s] = h;
    unsigned char* st = smem + s * kStage;
    mbar_arrive_expect_tx(full + s, (uint32_t)kStage);
    if (MPM_G2P_BOX) tma_load_4d(st, &tm_grid, 0, h.z0b, h.y0b, h.x0b, full + s);
    tma_load_2d(st + L::box_bytes(LT), &tm_streams, (int)(d.start & ~3u), 0, full + s);
  };
#if MPM_G2P_SELFFEED
  // No producer warp: thread 0 fills the ring once, afterwards the LAST warp to finish with a stage
  // (a shared-memory counter tells it so) requests the tile that goes there next.  Nobody waits.
  uint32_t* rel = reinterpret_cast<uint32_t*>(empty);  // the "empty" barriers are unused: one counter per stage
  if (tid == 0) {
    tma_prefetch_desc(&tm_streams);
    for (int s = 0; s < kG2pStages; ++s) {
      rel[2 * s] = 0;
      const uint32_t t = blockIdx.x + (uint32_t)s * gridDim.x;
      if (t < n_tiles) {
        issue(t, s);
      } else if (MPM_G2P_DYNAMIC) {  // end marker
        hdr[s].n = -1;
        mbar_arrive(full + s);
      }
    }
    if (MPM_G2P_DYNAMIC && blockIdx.x == 0) tile_counters[parity ^ 1] = 0;  // the next launch's counter
  }
  __syncthreads();
#else
  if (tid >= kTile) {  // ---- producer warp: one elected lane feeds the ring ----
    if (tid == kTile) {
      tma_prefetch_desc(&tm_grid);
      tma_prefetch_desc(&tm_streams);
      int it = 0;
      for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = it % kG2pStages;
        if (it >= kG2pStages) mbar_wait(empty + s, (uint32_t)(((it / kG2pStages) - 1) & 1));
        issue(t, s);
      }
    }
    return;
  }
#endif
  // ---- consumer warps ----
  int it = 0;
  unsigned moved = 0;  // particles of this thread that changed cell in this substep
  for (uint32_t t = blockIdx.x; MPM_G2P_DYNAMIC || t < n_tiles; t += gridDim.x, ++it) {
    const int s = it % kG2pStages;
    mbar_wait(full + s, (uint32_t)((it / kG2pStages) & 1));
    if (MPM_G2P_DYNAMIC && hdr[s].n < 0) break;  // no more tiles for this CTA
    const unsigned char* st = smem + s * kStage;
    g2p_tile_compute<MODEL, O, LT, COUNT_MOVED>(p, mats, mat0, one_mat, grid, k, hdr[s], reinterpret_cast<const float4*>(st),
                                   bricks + (tid >> 5) * (kWarpBrickX * kWarpBrickY * kWarpBrickZ),
                                   reinterpret_cast<const float*>(st + L::box_bytes(LT)), tid, moved);
    __syncwarp();  // stage s and the warp's brick are free again
#if MPM_G2P_SELFFEED
    if ((tid & 31) == 0) {
      __threadfence_block();
      if (atomicAdd(&rel[2 * s], 1u) == kTile / 32 - 1) {
        rel[2 * s] = 0;
        __threadfence_block();
        // static: the tile kG2pStages rounds ahead; dynamic: the next tile nobody has taken yet
        const uint32_t tn = MPM_G2P_DYNAMIC ? (uint32_t)kG2pStages * gridDim.x + atomicAdd(&tile_counters[parity], 1u)
                                            : t + (uint32_t)kG2pStages * gridDim.x;
        if (tn < n_tiles) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the warps' reads of the stage before the copy engine's writes
          issue(tn, s);
        } else if (MPM_G2P_DYNAMIC) {
          hdr[s].n = -1;
          mbar_arrive(full + s);
        }
      }
    }
#else
    if ((tid & 31) == 0) mbar_arrive(empty + s);
#endif
  }
  if (COUNT_MOVED) {
    moved = __reduce_add_sync(0xffffffffu, moved);
    if ((tid & 31) == 0 && moved) atomicAdd(moved_total, (unsigned long long)moved);
  }
}

}  // namespace mpm
