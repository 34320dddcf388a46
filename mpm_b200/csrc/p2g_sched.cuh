// Stage (2), production kernel: P2G with in-register pre-reduction of same-cell particle runs and
// a length-sorted schedule (reference behaviour: src/mpm.cu:14-74, TransferScheme.h:66-100), for the
// shipped transfer tuple MLS_APIC_Scheme<QuadraticInterpolationKernel> and any MaterialModel.
//
// A direct scatter (p2g_generic_kernel, kernels.cuh) is bound by reduction lanes: 27 REDG per particle.
// Particles are cell-sorted, so consecutive particles mostly share the base node and therefore all
// 27 target nodes.  One CTA per 256-particle tile of the SoA (common.cuh):
//   phase 0  one thread per particle: load the particle's streams (one address + immediate
//            offsets), stress via material.computePF, affine matrix; write a payload record (fractional
//            position, m v + A d at the three base z-nodes, dx*A columns) to shared memory.
//            HANDOVER: the G2P kernel of the previous substep already left dx*A in the C rows, so only
//            v, A, x (15 of the 25 streams) are read and the material is not evaluated here.
//   phase R  warp-local run detection with ballots: a run = maximal stretch of consecutive
//            particles OF ONE WARP with equal base node (<= 32 particles).  Every run head drops
//            its run into a histogram bin by length (one shared-memory integer atomic); after one
//            barrier every warp scans the 32 bins in registers and the heads write the block's run
//            list in order of DESCENDING LENGTH.
//   phase 1  three threads per run, one per stencil z-node (9 nodes, 36 accumulators in
//            registers), taken from the sorted list: the 32 lanes of a warp get runs of (nearly)
//            equal length, so the accumulation loop does not diverge — with runs in memory order
//            a warp waits for its longest run and half the issue slots are lost
//            (profiles/r01_ncu_v2_runs_p2g.txt).  Then ONE red.global.add.v4.f32 per node per run.
// Correctness does not depend on the order being perfectly sorted (a stale order only shortens
// the runs).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "kernels.cuh"

namespace mpm {

constexpr int kP2gMinBlocks = 4;    // CTAs per SM the kernel is compiled for (64 registers; 48 spill, 80 lose a CTA: profiles/r01_ab2, r02_ab5)
constexpr int kP2gWarpSortMin = 4;  // descents in a warp's key sequence from which the warp sorts
#define MPM_LDP(ptr) __ldcs(ptr)    // particle streams are touched once per kernel: evict-first loads
constexpr int kP2gBlock = kTile;
constexpr int kRunPosBits = 8;  // run list entry = first particle | (length - 1) << kRunPosBits
constexpr uint32_t kInvalidKey = 0xffffffffu;
constexpr int kKeyBias = 4;  // base node >= -3 for particles that are not skipped
constexpr int kP2gMaxN = 1023 - kKeyBias;  // 10 bits per axis in the packed run key

// per-particle payload record: 5 x float4 = 80 B, consecutive records start 20 banks apart, so the
// 128-bit reads of 8 different runs are conflict-free unless two of them sit 8 particles apart.
// The scatter thread of z-node c reads slots 0, 1 and 2 + c only (three 128-bit loads):
//   [0]     = (fx, fy, fz, cx.x)          fractional position in cells, [0.5, 1.5)
//   [1]     = (cx.y, cx.z, cy.x, cy.y)    cx, cy: change of q per node step along x, y (dx * A columns)
//   [2 + c] = (cy.z, q_c.xyz)             q_c = m v + A (x_node - x) at node (0, 0, c) of the stencil
struct P2gSmem {
  float4 pay[kP2gBlock][5];
  uint32_t key[kP2gBlock];
  float mass[kP2gBlock];
  uint32_t hist[32];          // bin b = runs of length 32 - b
  uint16_t runs[kP2gBlock];   // first particle | (length - 1) << 8, longest first
};

// quadratic B-spline weights from the fractional position (InterpolationKernel.cuh);
// w2 = w0 + (f - 1) is the same polynomial with two operations fewer
__device__ __forceinline__ void bspline_w(float f, float w[3]) {
  const float a0 = 1.5f - f, a1 = f - 1.0f;
  w[0] = (0.5f * a0) * a0;
  w[1] = fmaf(-a1, a1, 0.75f);
  w[2] = w[0] + a1;
}

// phase 0 tail: payload of one particle from its fractional position f (cells, relative to the base
// node), velocity and dx * affine.
//   q(node) = m v + A (x_node - x) = q0 + i cx + j cy + k cz,  x_base - x = -dx f
__device__ __forceinline__ void p2g_store_payload(P2gSmem& sm, int tid, const float f[3], const float v[3], const Mat& Ad, float mass) {
  float q0[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) q0[c] = v[c] * mass - (Ad.m[c][0] * f[0] + Ad.m[c][1] * f[1] + Ad.m[c][2] * f[2]);
  sm.pay[tid][0] = make_float4(f[0], f[1], f[2], Ad.m[0][0]);
  sm.pay[tid][1] = make_float4(Ad.m[1][0], Ad.m[2][0], Ad.m[0][1], Ad.m[1][1]);
  sm.pay[tid][2] = make_float4(Ad.m[2][1], q0[0], q0[1], q0[2]);
  sm.pay[tid][3] = make_float4(Ad.m[2][1], q0[0] + Ad.m[0][2], q0[1] + Ad.m[1][2], q0[2] + Ad.m[2][2]);
  sm.pay[tid][4] = make_float4(Ad.m[2][1], fmaf(2.0f, Ad.m[0][2], q0[0]), fmaf(2.0f, Ad.m[1][2], q0[1]), fmaf(2.0f, Ad.m[2][2], q0[2]));
  sm.mass[tid] = mass;
}

// base node and packed run key of a position (kInvalidKey: the whole stencil lies outside the domain)
__device__ __forceinline__ uint32_t p2g_key_of(const float x[3], const KParams& k, int base[3], float f[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float g = x[a] * k.dx_inv;
    base[a] = (int)(g - 0.5f);  // C truncation like the reference's cast<int>()
    f[a] = g - (float)base[a];
  }
  if (stencil_outside(base, k.N)) return kInvalidKey;
  return ((uint32_t)(base[0] + kKeyBias) << 20) | ((uint32_t)(base[1] + kKeyBias) << 10) | (uint32_t)(base[2] + kKeyBias);
}

// Stale order.  Between two re-bins the particles of a warp drift into neighbouring cells and the runs
// of equal keys fragment (a sheared block at 0.1 cells per substep: 2.5 x the runs after 4 substeps).
// A warp whose keys are no longer ascending sorts its 32 (key, lane) pairs with a bitonic network of
// shuffles (15 compare-exchange steps) and writes its payload records in that order, which restores
// one run per cell and warp; warps that are still in order (every warp right after a re-bin) pay one
// shuffle and a vote.  pos = payload slot of this lane's particle within the warp, skey = the key at
// slot `lane`.  scratch: kP2gBlock uint16 that nobody else uses before the next CTA barrier.
// (out of line: the common, still-ordered path jumps over nothing and keeps its registers)
static __device__ __noinline__ uint2 p2g_warp_sort(uint32_t key, int tid, uint16_t* scratch) {
  const int lane = tid & 31;
  uint32_t kk = key;
  int src = lane;
#pragma unroll
  for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      const uint32_t pk = __shfl_xor_sync(0xffffffffu, kk, j);
      const int ps = __shfl_xor_sync(0xffffffffu, src, j);
      const bool up = (lane & k2) == 0, lower = (lane & j) == 0;
      const bool partner_less = pk < kk || (pk == kk && ps < src);  // (key, lane) pairs are distinct: a strict order
      if ((lower == up) ? partner_less : !partner_less) {
        kk = pk;
        src = ps;
      }
    }
  scratch[(tid & ~31) + src] = (uint16_t)lane;  // this lane now holds the pair that came from lane `src`: its rank is `lane`
  __syncwarp();
  const uint32_t pos = scratch[tid];
  __syncwarp();
  return make_uint2(pos, kk);
}
template <bool SORT>
__device__ __forceinline__ void p2g_warp_order(uint32_t key, int tid, uint16_t* scratch, int& pos, uint32_t& skey) {
  const int lane = tid & 31;
  pos = lane;
  skey = key;
  if constexpr (SORT) {
    // (one or two stragglers do not pay for a sort: the warp's sort delays its whole CTA at the next barrier,
    // profiles/r02_ab4_p2g_warpsort.txt)
    const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    if (__popc(__ballot_sync(0xffffffffu, lane > 0 && prev > key)) >= kP2gWarpSortMin) {
      const uint2 r = p2g_warp_sort(key, tid, scratch);
      pos = (int)r.x;
      skey = r.y;
    }
  }
}

// phase R: warp-local runs of equal keys, listed block-wide in order of descending length.
// Expects sm.hist zeroed and sm.pay / sm.key of this thread written; two barriers inside.
// Returns the number of runs.
__device__ __forceinline__ int p2g_list_runs(P2gSmem& sm, uint32_t key, int tid) {
  const int lane = tid & 31;
  const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
  const bool head = (lane == 0) || (prev != key);
  const uint32_t heads = __ballot_sync(0xffffffffu, head);
  const uint32_t rest = (lane == 31) ? 0u : (heads >> (lane + 1));
  const int len = rest ? __ffs(rest) : (32 - lane);  // meaningful for heads
  const bool listed = head && key != kInvalidKey;     // skipped particles / the tail of the last block scatter nothing
  uint32_t slot = 0;
  if (listed) slot = atomicAdd(&sm.hist[32 - len], 1u);
  __syncthreads();
  uint32_t incl = sm.hist[lane];
  const uint32_t cnt = incl;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const int n_runs = (int)__shfl_sync(0xffffffffu, incl, 31);
  const uint32_t bin_base = __shfl_sync(0xffffffffu, incl - cnt, listed ? (32 - len) : 0);
  if (listed) sm.runs[bin_base + slot] = (uint16_t)(tid | ((len - 1) << kRunPosBits));
  __syncthreads();
  return n_runs;
}

// phase 1: three threads per run, one per stencil z-node (the 9 (x, y) nodes of that z = 36
// accumulators in registers).  The three lanes of a run then reduce into three CONSECUTIVE float4
// nodes (z is the fastest grid index): 48 contiguous bytes per run and reduction instruction.
template <bool ONE_MASS>
__device__ __forceinline__ void p2g_scatter_runs(const P2gSmem& sm, int n_runs, int tid, float4* __restrict__ grid, const KParams& k,
                                                 float mass_one) {
  const long long NN = (long long)k.N * k.N;
  const int gx_lo = max(0, k.x0), gx_hi = min(k.N, k.x0 + k.nxl);
  for (int u = tid; u < 3 * n_runs; u += kP2gBlock) {
    const int r = u / 3, c = u - 3 * r;
    const uint32_t run = sm.runs[r];
    const int s0 = (int)(run & ((1u << kRunPosBits) - 1u)), s1 = s0 + (int)(run >> kRunPosBits) + 1;
    const uint32_t rk = sm.key[s0];
    float4 acc[3][3];  // [x node][y node]
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) acc[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    // this thread's z weight as a + b (f - c)^2
    const float fc = (float)c;
    const float wc = 1.5f - 0.5f * fc, wa = (c == 1) ? 0.75f : 0.0f, wb = (c == 1) ? -1.0f : 0.5f;
#pragma unroll 1
    for (int s = s0; s < s1; ++s) {
      const float4 r0 = sm.pay[s][0], r1 = sm.pay[s][1], r2 = sm.pay[s][2 + c];
      const float dz = r0.z - wc;
      const float wzc = fmaf(wb * dz, dz, wa);
      float wx[3], wy[3];
      bspline_w(r0.x, wx);
      bspline_w(r0.y, wy);
      const float mass = ONE_MASS ? mass_one : sm.mass[s];
      float q[3] = {r2.y, r2.z, r2.w};
      const float cx[3] = {r0.w, r1.x, r1.y};
      const float cy[3] = {r1.z, r1.w, r2.x};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float wi = wzc * wx[i];
        float qj[3] = {q[0], q[1], q[2]};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float wt = wi * wy[j];
          acc[i][j].x = fmaf(wt, qj[0], acc[i][j].x);
          acc[i][j].y = fmaf(wt, qj[1], acc[i][j].y);
          acc[i][j].z = fmaf(wt, qj[2], acc[i][j].z);
          acc[i][j].w = fmaf(wt, mass, acc[i][j].w);
          if (j < 2) {
#pragma unroll
            for (int d = 0; d < 3; ++d) qj[d] += cy[d];
          }
        }
        if (i < 2) {
#pragma unroll
          for (int d = 0; d < 3; ++d) q[d] += cx[d];
        }
      }
    }
    // flush: one vector reduction per node of this z
    const int bx = (int)(rk >> 20) - kKeyBias, by = (int)((rk >> 10) & 1023u) - kKeyBias, bz = (int)(rk & 1023u) - kKeyBias;
    const int gz = bz + c;
    if (gz < 0 || gz >= k.N) continue;
    float4* gp = grid + ((long long)(bx - k.x0) * NN + (long long)by * k.N + gz);
    if (bx >= gx_lo && bx + 2 < gx_hi && (unsigned)by <= (unsigned)(k.N - 3)) {  // whole 3 x 3 patch inside
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float4* row = gp + i * NN;
#pragma unroll
        for (int j = 0; j < 3; ++j) atomicAdd(row + j * k.N, acc[i][j]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int gx = bx + i;
        if (gx < gx_lo || gx >= gx_hi) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int gy = by + j;
          if (gy < 0 || gy >= k.N) continue;
          atomicAdd(gp + (i * NN + (long long)j * k.N), acc[i][j]);
        }
      }
    }
  }
}

// does the material read or write Jp at all?  (fixed-corotated reads it but never changes it, and
// handles whose uploads all had Jp == 1 skip the stream: DeviceDiag::jp_not_one)
template <class Material>
struct MaterialTraits {
  static constexpr bool kMutatesJp = true;
};
template <class P, class O>
struct MaterialTraits<MMFixedCorotated<P, O>> {
  static constexpr bool kMutatesJp = false;
};

// sort_keys != nullptr: also emit the cell keys (and the identity permutation) of a re-bin that follows
// in the same substep — the positions are in registers here anyway, which saves the sort its own pass
// over them (cell_key_kernel, sort.cuh; same key: clamped base node, x local to the slab, z fastest) —
// and count non-finite / out-of-domain particles for mpm_get_diagnostics.
// SORT: the launch was told that the order has gone stale (the host decides from the cell crossings G2P
// counts since the last re-bin): warps re-order their payload records first (p2g_warp_order)
template <class Material, bool ONE_MAT, bool HANDOVER, bool SORT>
__global__ void __launch_bounds__(kP2gBlock, kP2gMinBlocks)
p2g_sched_kernel(Soa p, size_t count, const MatTable<Material> mats, float4* __restrict__ grid, KParams k,
                 uint32_t* __restrict__ sort_keys, uint32_t* __restrict__ sort_vals, uint32_t first, DeviceDiag* __restrict__ diag) {
  __shared__ P2gSmem sm;
  const int tid = threadIdx.x;
  const size_t pi = (size_t)blockIdx.x * kP2gBlock + tid;
  if (tid < 32) sm.hist[tid] = 0;
  __syncthreads();

  // ---------------- phase 0: per-particle payload ----------------
  // (all loads are issued before anything waits for one of them: one DRAM round trip per CTA)
  const float* __restrict__ col = p.tile(blockIdx.x) + tid;
  const bool live = pi < count;
  float x[3] = {0.f, 0.f, 0.f}, v[3] = {0.f, 0.f, 0.f};
  Mat A = Mat::Zero();  // C, or dx * affine when handed over
  Particle part;
  part.material_type = 0;
  part.Jp = 1.0f;
  if (live) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      x[a] = MPM_LDP(col + (SX + a) * kTile);
      v[a] = MPM_LDP(col + (SV + a) * kTile);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) A.m[r][c] = MPM_LDP(col + (SC + 3 * r + c) * kTile);
    if constexpr (!HANDOVER) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) part.F.m[r][c] = MPM_LDP(col + (SF + 3 * r + c) * kTile);
      if (MaterialTraits<Material>::kMutatesJp || diag->jp_not_one) part.Jp = MPM_LDP(col + SJ * kTile);
    }
  }
  int base[3] = {0, 0, 0};
  float f[3] = {0.f, 0.f, 0.f};
  const uint32_t key = live ? p2g_key_of(x, k, base, f) : kInvalidKey;
  int pos;
  uint32_t skey;
  p2g_warp_order<SORT>(key, tid, sm.runs, pos, skey);
  if (live) {
    const Material m = mats.template get<ONE_MAT>(p.mat, pi);
    if constexpr (!HANDOVER) {
      part.C = A;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        part.x(a) = x[a];
        part.v(a) = v[a];
      }
      A = p2g_affine_dx(part, m, k);
    }
    p2g_store_payload(sm, (tid & ~31) + pos, f, v, A, m.particleMass);
    if (key != kInvalidKey) {  // slab handles: a stencil that leaves the planes held here loses mass
      const int lo = max(base[0], 0), hi = min(base[0] + 2, k.N - 1);
      if (lo < k.x0 || hi >= k.x0 + k.nxl) atomicAdd(&diag->escaped, 1u);
    }
    if (sort_keys) {
      int b[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) b[a] = min(max(base[a], 0), k.N - 1);
      const int bx = min(max(b[0] - k.x0, 0), k.nxl - 1);
      sort_keys[pi] = (uint32_t)((bx * k.N + b[1]) * k.N + b[2]);
      sort_vals[pi] = first + (uint32_t)pi;  // `first`: slot of this launch's first particle (split launches)
      if (!isfinite(x[0] + x[1] + x[2])) atomicAdd(&diag->nonfinite, 1u);
      else if (key == kInvalidKey) atomicAdd(&diag->out_of_domain, 1u);
    }
  }
  sm.key[tid] = skey;  // the key of the record in slot tid
  const int n_runs = p2g_list_runs(sm, skey, tid);
  p2g_scatter_runs<ONE_MAT>(sm, n_runs, tid, grid, k, mats.one.particleMass);
}

}  // namespace mpm
