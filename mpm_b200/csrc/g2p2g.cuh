// Stages (4)+(2) fused: G2P of substep s and P2G of substep s+1 in ONE pass over the particles
// (reference behaviour: src/mpm.cu:109-178 followed by src/mpm.cu:14-74 of the next advance()).
//
// Why: run back to back, G2P writes x, v, F, C, Jp (100 B/particle) and the next P2G reads the same
// 100 B straight back — 40 % of the particle traffic of a substep is a round trip through HBM of
// values that were in registers.  Fused, a particle is gathered from grid A (velocities of substep
// s), updated, stored once, and its substep-(s+1) contribution is scattered into grid B from the
// registers.  Real traffic falls from 252 to 152 B/particle-step; the arithmetic per particle is
// the same sequence of operations (same functions: g2p_gather27, snow_plasticity, p2g_prepare,
// p2g_list_runs, p2g_scatter_runs), so results are those of the separate kernels up to the order
// of the atomic additions.
//
// Second effect: the separate G2P is latency-bound (long-scoreboard stalls) and the separate P2G is
// issue-bound; in one persistent kernel the warps of different CTAs sit in different phases and
// fill each other's stalls.
//
// Structure: persistent CTAs over the tile list (tiles.cuh), three warp roles.
//   * 1 producer warp: one elected lane feeds a ring of TMA stage buffers with the x, F, Jp stream
//     rows of each tile (as in g2p_tile_kernel).
//   * 8 gather warps (one thread per particle of the tile): G2P from grid A, particle update and
//     store, then the P2G payload of the particle for the NEXT substep (p2g_prepare) into a ring of
//     payload buffers in shared memory.  They never wait for the scatter warps unless the ring is full.
//   * kFusedNS scatter warps: take a full payload buffer, list its runs of equal base node by length
//     and scatter them into grid B with three threads per run (p2g_sched.cuh), then free the buffer.
// Gather warps are bound by the latency of the 27 node loads, scatter warps by instruction issue: on
// one SM the two roles fill each other's stalls, and no warp ever stands at a CTA-wide barrier.  (A
// first version that ran both phases in the same 8 warps behind two barriers per tile took 5.4 ms
// per substep at 2^26 particles, slower than the separate kernels; its gather half alone 2.7 ms.)
// Hand-over is by mbarriers: pfull[b] (8 arrivals, one per gather warp) / pempty[b] (kFusedNS).
#pragma once
#include "g2p_tile.cuh"
#include "p2g_sched.cuh"

namespace mpm {

#ifndef MPM_FUSED_STAGES
#define MPM_FUSED_STAGES 3
#endif
#ifndef MPM_FUSED_MINBLK
#define MPM_FUSED_MINBLK 2
#endif
#ifndef MPM_FUSED_PBUF
#define MPM_FUSED_PBUF 2  // payload buffers in the ring between gather and scatter warps (>= 2)
#endif
#ifndef MPM_FUSED_POLL
#define MPM_FUSED_POLL 1   // 1: one polling lane per warp on the mbarriers, 0: all lanes
#endif
#ifndef MPM_FUSED_SLEEP
#define MPM_FUSED_SLEEP 64  // ns between probes of the scatter warps waiting for a payload buffer
#endif
#if MPM_FUSED_POLL
#define MPM_FWAIT(bar, parity, ns) mbar_wait_warp(bar, parity, ns)
#else
#define MPM_FWAIT(bar, parity, ns) mbar_wait(bar, parity)
#endif
#ifndef MPM_FUSED_NS
#define MPM_FUSED_NS 2    // scatter warps per CTA: 1, 2, 4 or 8 (2: 352 threads, 88 registers, no spills)
#endif
constexpr int kFusedStages = MPM_FUSED_STAGES;
constexpr int kFusedPbuf = MPM_FUSED_PBUF;
constexpr int kFusedNS = MPM_FUSED_NS;
constexpr int kFusedThreads = kTile + 32 + 32 * kFusedNS;  // gather warps, producer warp, scatter warps
static_assert(kTile == kP2gBlock || MPM_P2G_BLOCK != 256, "one tile = one P2G block");
static_assert(kFusedPbuf >= 2 && (kFusedNS == 1 || kFusedNS == 2 || kFusedNS == 4 || kFusedNS == 8), "see above");

struct ScatterBarrier {  // the scatter warps only
  __device__ __forceinline__ void operator()() const { asm volatile("bar.sync 1, %0;" ::"n"(32 * kFusedNS) : "memory"); }
};

// phase R for the scatter warps: the tile's 8 groups of 32 keys are read back from shared memory,
// kTile / (32 kFusedNS) groups per warp; otherwise p2g_list_runs.  ts = thread index among the
// scatter warps; next_hist = histogram of the next payload buffer, zeroed between the barriers.
__device__ __forceinline__ int fused_list_runs(P2gSmem& sm, int ts, uint32_t* next_hist) {
  constexpr int NG = kTile / (32 * kFusedNS);
  const int lane = ts & 31, w = ts >> 5;
  const ScatterBarrier bar;
  bool listed[NG];
  int len[NG];
  uint32_t slot[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const uint32_t key = sm.key[(w * NG + g) * 32 + lane];
    const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = (lane == 0) || (prev != key);
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    const uint32_t rest = (lane == 31) ? 0u : (heads >> (lane + 1));
    len[g] = rest ? __ffs(rest) : (32 - lane);
    listed[g] = head && key != kInvalidKey;
    slot[g] = 0;
    if (listed[g]) slot[g] = atomicAdd(&sm.hist[32 - len[g]], 1u);
  }
  bar();
  uint32_t incl = sm.hist[lane];
  if (ts < 32) next_hist[ts] = 0;
  const uint32_t cnt = incl;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const int n_runs = (int)__shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const uint32_t bin_base = __shfl_sync(0xffffffffu, incl - cnt, listed[g] ? (32 - len[g]) : 0);
    if (listed[g]) sm.runs[bin_base + slot[g]] = (uint16_t)(((w * NG + g) * 32 + lane) | ((len[g] - 1) << kRunPosBits));
  }
  bar();
  return n_runs;
}

template <int MODEL>
struct FusedLayout {
  static constexpr int kStreams = G2pTileLayout<MODEL>::kStreams;
  __host__ __device__ static constexpr size_t stage_bytes() { return (size_t)kStreams * kTile * 4; }
  __host__ __device__ static constexpr size_t bytes() {
    return kFusedStages * (stage_bytes() + sizeof(TileHeader) + 16) + kFusedPbuf * (sizeof(P2gSmem) + 16) + 128 + 16;
  }
};

template <int MODEL, class O, bool EXACT, bool ONE_MAT>
__global__ void __launch_bounds__(kFusedThreads, MPM_FUSED_MINBLK)
g2p2g_kernel(Soa p, const MpmMaterial* __restrict__ mats, const MpmMaterial mat0, const float4* __restrict__ grid_in,
             float4* __restrict__ grid_out, KParams k, const TileDesc* __restrict__ tiles, const uint32_t* __restrict__ n_tiles_ptr,
             const __grid_constant__ CUtensorMap tm_streams) {
  using L = FusedLayout<MODEL>;
  static_assert(SX == 0 && SF == 3 && SJ == 12, "G2P reads stream rows 0..12 as one TMA box");
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);  // TMA destinations: 128 B aligned
  constexpr size_t kStage = L::stage_bytes();
  P2gSmem* pbuf = reinterpret_cast<P2gSmem*>(smem + kFusedStages * kStage);
  TileHeader* hdr = reinterpret_cast<TileHeader*>(pbuf + kFusedPbuf);
  uint64_t* full = reinterpret_cast<uint64_t*>(hdr + kFusedStages);
  uint64_t* empty = full + kFusedStages;
  uint64_t* pfull = empty + kFusedStages;
  uint64_t* pempty = pfull + kFusedPbuf;
  const int tid = threadIdx.x;
  const uint32_t n_tiles = *n_tiles_ptr;
  if (tid == 0) {
    for (int s = 0; s < kFusedStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, kTile / 32);
    }
    for (int b = 0; b < kFusedPbuf; ++b) {
      mbar_init(pfull + b, kTile / 32);
      mbar_init(pempty + b, kFusedNS);
    }
    mbar_fence_init();
  }
  if (tid < 32)
    for (int b = 0; b < kFusedPbuf; ++b) pbuf[b].hist[tid] = 0;
  __syncthreads();

#if defined(MPM_FUSED_EXP) && (MPM_FUSED_EXP & 8)  // experiment: no hand-over at all
  if (tid >= kTile + 32) return;
#define MPM_FUSED_NOHAND 1
#endif
  if (tid >= kTile + 32) {  // ---- scatter warps: P2G phases R and 1 of substep s+1, tile by tile ----
    const int ts = tid - (kTile + 32);
    int it = 0;
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int b = it % kFusedPbuf;
      MPM_FWAIT(pfull + b, (uint32_t)((it / kFusedPbuf) & 1), MPM_FUSED_SLEEP);
      P2gSmem& sm = pbuf[b];
#if defined(MPM_FUSED_EXP) && (MPM_FUSED_EXP & 2)  // experiment: scatter warps only hand the buffer back
      if (sm.key[ts] == 0x12345u) grid_out[ts] = sm.pay[ts][0];
#elif defined(MPM_FUSED_EXP) && (MPM_FUSED_EXP & 4)  // experiment: run listing only
      const int n_runs = fused_list_runs(sm, ts, pbuf[(it + 1) % kFusedPbuf].hist);
      if (n_runs == 0x12345) grid_out[ts] = sm.pay[ts][0];
#else
      const int n_runs = fused_list_runs(sm, ts, pbuf[(it + 1) % kFusedPbuf].hist);
      p2g_scatter_runs(sm, n_runs, ts, grid_out, k, 32 * kFusedNS);
#endif
      __syncwarp();  // the warp is done with payload buffer b
      if ((ts & 31) == 0) mbar_arrive(pempty + b);
    }
    return;
  }
  if (tid >= kTile) {  // ---- producer warp: one elected lane feeds the ring ----
    if (tid == kTile) {
      tma_prefetch_desc(&tm_streams);
      int it = 0;
      for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = it % kFusedStages;
        if (it >= kFusedStages) mbar_wait(empty + s, (uint32_t)(((it / kFusedStages) - 1) & 1));
        const TileDesc d = tiles[t];
        TileHeader h;
        h.x0b = h.y0b = h.z0b = 0;
        h.n = (int)d.n;
        h.start = d.start;
        h.off = (int)(d.start & 3u);
        hdr[s] = h;
        mbar_arrive_expect_tx(full + s, (uint32_t)kStage);
        tma_load_2d(smem + s * kStage, &tm_streams, (int)(d.start & ~3u), 0, full + s);
      }
    }
    return;
  }

  // ---- gather warps ----
  int it = 0;
  for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
    const int s = it % kFusedStages;
    MPM_FWAIT(full + s, (uint32_t)((it / kFusedStages) & 1), 0);
    const TileHeader h = hdr[s];
    const int b = it % kFusedPbuf;
    P2gSmem& sm = pbuf[b];
    P2gPayload o;
    o.key = kInvalidKey;
    {
      const float* ps = reinterpret_cast<const float*>(smem + s * kStage) + h.off + tid;  // this particle's column
      const bool live = tid < h.n;
      const size_t pi = (size_t)h.start + tid;
      // ---------------- G2P of substep s ----------------
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
      Mat3 F;
      float Jp = 1.0f;
      G2pAcc acc;
      if (valid) {
        const int bxl = base[0] - k.x0;  // x-plane in the local grid
        float d[3][3];                   // node - particle distance per axis (world units)
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int i = 0; i < 3; ++i) d[a][i] = (float)(base[a] + i) * k.dx - x[a];
        bool interior = bxl >= 0 && bxl + 2 < k.nxl;
#pragma unroll
        for (int a = 0; a < 3; ++a) interior = interior && base[a] >= 0 && base[a] + 2 < k.N;
        if (interior) {  // whole stencil inside the local grid: unclipped gather
          const long long NN = (long long)k.N * k.N;
          g2p_gather27<true>(grid_in + (bxl * NN + (long long)base[1] * k.N + base[2]), k.N, NN, w, d, acc);
        } else {
          const G2pGather o = g2p_gather_clipped(grid_in, k, x[0], x[1], x[2]);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            acc.v[c] = o.v[c];
#pragma unroll
            for (int a = 0; a < 3; ++a) acc.B.m[c][a] = o.B[c][a];
          }
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) F.m[r][c] = ps[(SF + 3 * r + c) * kTile];
        if (MODEL == MPM_MODEL_SNOW) Jp = ps[SJ * kTile];
      }
      __syncwarp();  // the warp is done with stage s: hand it back to the producer
      if ((tid & 31) == 0) mbar_arrive(empty + s);
      if (valid) {
        MpmMaterial m;
        if constexpr (ONE_MAT) m = mat0; else m = load_material(mats, p.mat[pi]);
        Mat3 C, G;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            C.m[r][c] = acc.B.m[r][c] * k.dinv;
            G.m[r][c] = ((r == c) ? 1.0f : 0.0f) + k.dt * C.m[r][c];
          }
        F = mul_ab(G, F);  // F <- (I + dt C) F
        if (MODEL == MPM_MODEL_SNOW) {
          snow_plasticity<O>(F, Jp, m);
          MPM_STP(p.s(SJ) + pi, Jp);
        }
        float v[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          v[a] = acc.v[a];
          x[a] = x[a] + k.dt * v[a];
          MPM_STP(p.s(SX + a) + pi, x[a]);
          MPM_STP(p.s(SV + a) + pi, v[a]);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            MPM_STP(p.s(SF + 3 * r + c) + pi, F.m[r][c]);
            MPM_STP(p.s(SC + 3 * r + c) + pi, C.m[r][c]);
          }
        // ---------------- P2G of substep s+1, phase 0: payload from the registers ----------------
        o = p2g_prepare<MODEL, O, EXACT>(x, v, F, C, (MODEL == MPM_MODEL_SNOW) ? Jp : 1.0f, m, k);
      }
    }
    // the payload buffer is free again once the scatter warps are done with tile it - kFusedPbuf
#ifndef MPM_FUSED_NOHAND
    if (it >= kFusedPbuf) MPM_FWAIT(pempty + b, (uint32_t)(((it / kFusedPbuf) - 1) & 1), 0);
#endif
    if (o.key != kInvalidKey) {
      sm.pay[tid][0] = make_float4(o.f[0], o.f[1], o.f[2], o.mass);
      sm.pay[tid][1] = make_float4(o.q0[0], o.q0[1], o.q0[2], o.cx[0]);
      sm.pay[tid][2] = make_float4(o.cx[1], o.cx[2], o.cy[0], o.cy[1]);
      sm.pay[tid][3] = make_float4(o.cy[2], o.cz[0], o.cz[1], o.cz[2]);
    }
    sm.key[tid] = o.key;
#ifndef MPM_FUSED_NOHAND
    __syncwarp();  // payload and keys of this warp's 32 particles are written
    if ((tid & 31) == 0) mbar_arrive(pfull + b);
#endif
  }
}

}  // namespace mpm
