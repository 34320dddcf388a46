// Stage (2), production kernel: P2G with in-register pre-reduction of same-cell particle runs and
// a length-sorted schedule (reference behaviour: src/mpm.cu:14-74, TransferScheme.h:66-100).
//
// The direct scatter (p2g_kernel, kernels.cuh) is bound by reduction lanes: 27 REDG per particle.
// Particles are cell-sorted, so consecutive particles mostly share the base node and therefore all
// 27 target nodes.  Per block of kP2gBlock consecutive particles:
//   phase 0  one thread per particle: load the 25 streams (coalesced), stress via polar/svd3,
//            affine matrix; write a 16-float payload (fractional position, mass, m v + A d at the
//            base node, dx*A columns) to shared memory.
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
#include "tma.cuh"

namespace mpm {

#ifndef MPM_P2G_MINBLK
#define MPM_P2G_MINBLK 4
#endif
#ifndef MPM_P2G_TMA
#define MPM_P2G_TMA 0  // 1: the 25 particle streams of the block arrive as one 2-D TMA box in shared memory
#endif
#ifndef MPM_P2G_CAP
#define MPM_P2G_CAP 0  // 0 = runs as long as the warp allows
#endif
#ifndef MPM_P2G_BLOCK
#define MPM_P2G_BLOCK 256
#endif
constexpr int kP2gBlock = MPM_P2G_BLOCK;
constexpr int kRunPosBits = kP2gBlock > 256 ? 9 : 8;  // run list entry = first particle | (length - 1) << kRunPosBits
static_assert(kP2gBlock <= 512, "run list entries are 16 bits");
constexpr uint32_t kInvalidKey = 0xffffffffu;
constexpr int kKeyBias = 4;  // base node >= -3 for particles that are not skipped

struct P2gSmem {
  // per-particle payload record, 4 x float4 used of a 5 x float4 (80 B) stride: consecutive
  // records start 20 banks apart, so the 128-bit reads of 8 different runs are conflict-free
  //   [0] = (fx, fy, fz, mass)  [1] = (q0.xyz, cx.x)  [2] = (cx.y, cx.z, cy.x, cy.y)  [3] = (cy.z, cz.xyz)
  float4 pay[kP2gBlock][5];
  uint32_t key[kP2gBlock];
  uint32_t hist[32];          // bin b = runs of length 32 - b
  uint16_t runs[kP2gBlock];   // first particle | (length - 1) << 8, longest first
};

// what phase 0 hands to phase 1 for one particle
struct P2gPayload {
  float f[3];      // fractional position relative to the base node, in cells: [0.5, 1.5)
  float mass;
  float q0[3];     // m v + A (x_base - x)
  float cx[3], cy[3], cz[3];  // dx * A columns: the change of q per node step along x, y, z
  uint32_t key;    // packed biased base node, kInvalidKey for a particle outside the domain
};

// P2G particle preparation (reference TransferScheme.h:66-86 + MaterialModel.cuh:85-93), folded:
//   A = -Dinv dt vol PF + m C,  PF = 2 mu (F - R) F^T + lambda (Jp - 1) Jp I
//   q(node) = m v + A (x_node - x) = q0 + i cx + j cy + k cz,  x_base - x = -dx f
template <int MODEL, class O, bool EXACT>
__device__ __forceinline__ P2gPayload p2g_prepare(const float x[3], const float v[3], const Mat3& F, const Mat3& C, float Jp,
                                                  const MpmMaterial& m, const KParams& k) {
  P2gPayload o;
  int base[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float g = x[a] * k.dx_inv;
    base[a] = (int)(g - 0.5f);  // C truncation like the reference's cast<int>()
    o.f[a] = g - (float)base[a];
  }
  bool inside = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) inside = inside && !(base[a] + 3 < 0 || base[a] >= k.N);  // src/mpm.cu:31-35
  o.key = inside ? (((uint32_t)(base[0] + kKeyBias) << 20) | ((uint32_t)(base[1] + kKeyBias) << 10) | (uint32_t)(base[2] + kKeyBias))
                 : kInvalidKey;
  Mat3 R;
  if constexpr (EXACT) R = polar_rotation<O>(F); else R = polar_rotation_newton(F);
  float mu = m.mu0, lambda = m.lambda0;
  if (MODEL == MPM_MODEL_SNOW) {
    float e;
    if (EXACT) e = (float)exp((double)m.hardening * (1.0 - (double)Jp));
    else e = (m.hardening == 0.0f) ? 1.0f : __expf(m.hardening * (1.0f - Jp));
    mu *= e;
    lambda *= e;
  }
  float lam_term;
  if (EXACT) lam_term = (float)((double)lambda * (((double)Jp - 1.0) * (double)Jp));
  else lam_term = lambda * ((Jp - 1.0f) * Jp);
  const float kk = (((-k.dinv) * k.dt) * m.particleVolume) * k.dx;  // -Dinv dt vol, times dx for the columns
  const float s_dev = kk * (2.0f * mu), s_vol = kk * lam_term, s_c = m.particleMass * k.dx;
  Mat3 D;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) D.m[i][j] = F.m[i][j] - R.m[i][j];
  const Mat3 M = mul_abt(D, F);  // (F - R) F^T
  float Ad[3][3];                // dx * A
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Ad[i][j] = fmaf(s_dev, M.m[i][j], s_c * C.m[i][j]) + ((i == j) ? s_vol : 0.0f);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o.cx[c] = Ad[c][0];
    o.cy[c] = Ad[c][1];
    o.cz[c] = Ad[c][2];
    o.q0[c] = v[c] * m.particleMass - (Ad[c][0] * o.f[0] + Ad[c][1] * o.f[1] + Ad[c][2] * o.f[2]);
  }
  o.mass = m.particleMass;
  return o;
}

// quadratic B-spline weights from the fractional position (InterpolationKernel.cuh:61-66);
// w2 = w0 + (f - 1) is the same polynomial with two operations fewer
__device__ __forceinline__ void bspline_w(float f, float w[3]) {
  const float a0 = 1.5f - f, a1 = f - 1.0f;
  w[0] = (0.5f * a0) * a0;
  w[1] = fmaf(-a1, a1, 0.75f);
  w[2] = w[0] + a1;
}

struct BlockBarrier {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// phase R: warp-local runs of equal keys, listed block-wide in order of descending length.
// Expects sm.hist zeroed and sm.pay / sm.key of this thread written; two barriers inside.
// other_hist (optional): histogram of the other payload buffer of a double-buffered caller, zeroed
// here between the two barriers (see g2p2g.cuh).  Returns the number of runs.
template <class Bar>
__device__ __forceinline__ int p2g_list_runs(P2gSmem& sm, uint32_t key, int tid, Bar bar, uint32_t* other_hist = nullptr) {
  const int lane = tid & 31;
  const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
  bool head = (lane == 0) || (prev != key);
  uint32_t heads = __ballot_sync(0xffffffffu, head);
#if MPM_P2G_CAP
  {  // long runs are cut into pieces of at most MPM_P2G_CAP particles: shorter dependent chains in phase 1
    const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
    head = head || ((lane - start) % MPM_P2G_CAP == 0);
    heads = __ballot_sync(0xffffffffu, head);
  }
#endif
  const uint32_t rest = (lane == 31) ? 0u : (heads >> (lane + 1));
  const int len = rest ? __ffs(rest) : (32 - lane);  // meaningful for heads
  const bool listed = head && key != kInvalidKey;     // skipped particles / the tail of the last block scatter nothing
  uint32_t slot = 0;
  if (listed) slot = atomicAdd(&sm.hist[32 - len], 1u);
  bar();
  uint32_t incl = sm.hist[lane];
  if (other_hist && tid < 32) other_hist[tid] = 0;
  const uint32_t cnt = incl;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const int n_runs = (int)__shfl_sync(0xffffffffu, incl, 31);
  const uint32_t bin_base = __shfl_sync(0xffffffffu, incl - cnt, listed ? (32 - len) : 0);
  if (listed) sm.runs[bin_base + slot] = (uint16_t)(tid | ((len - 1) << kRunPosBits));
  bar();
  return n_runs;
}

// phase 1: three threads per run, one per stencil z-node (the 9 (x, y) nodes of that z = 36
// accumulators in registers).  The three lanes of a run then reduce into three CONSECUTIVE float4
// nodes (z is the fastest grid index): 48 contiguous bytes per run and reduction instruction.  With
// one thread per x-slab instead, every lane of a reduction hit its own 128-byte line, and the
// reductions alone took a fifth of the L1 data-pipe wavefronts of the kernel
// (profiles/r01_ncu_v4_fused_ws.txt).
// (tid = index of the thread among the `nthreads` that share the tile's run list)
__device__ __forceinline__ void p2g_scatter_runs(const P2gSmem& sm, int n_runs, int tid, float4* __restrict__ grid, const KParams& k,
                                                 int nthreads = kP2gBlock) {
  const long long NN = (long long)k.N * k.N;
  const int gx_lo = max(0, k.x0), gx_hi = min(k.N, k.x0 + k.nxl);
  for (int u = tid; u < 3 * n_runs; u += nthreads) {
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
#if defined(MPM_P2G_EXP) && (MPM_P2G_EXP & 4)  // experiment: no accumulation
    if (n_runs < 0)
#endif
#pragma unroll 1
    for (int s = s0; s < s1; ++s) {
      const float4 r0 = sm.pay[s][0], r1 = sm.pay[s][1], r2 = sm.pay[s][2], r3 = sm.pay[s][3];
      const float dz = r0.z - wc;
      const float wzc = fmaf(wb * dz, dz, wa);
      float wx[3], wy[3];
      bspline_w(r0.x, wx);
      bspline_w(r0.y, wy);
      const float mass = r0.w;
      float q[3] = {fmaf(fc, r3.y, r1.x), fmaf(fc, r3.z, r1.y), fmaf(fc, r3.w, r1.z)};
      const float cx[3] = {r1.w, r2.x, r2.y};
      const float cy[3] = {r2.z, r2.w, r3.x};
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
#if defined(MPM_P2G_EXP) && (MPM_P2G_EXP & 2)  // experiment: no reductions
    if (acc[0][0].x + acc[1][1].y + acc[2][2].z + acc[0][1].w + acc[0][2].x + acc[1][0].x + acc[1][2].x + acc[2][0].x + acc[2][1].x != 1.2345e30f) continue;
#endif
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

template <class Bar>
__device__ __forceinline__ void p2g_list_and_scatter_runs(P2gSmem& sm, uint32_t key, int tid, float4* __restrict__ grid, const KParams& k, Bar bar) {
  const int n_runs = p2g_list_runs(sm, key, tid, bar);
  p2g_scatter_runs(sm, n_runs, tid, grid, k);
}

// KEYS: also emit the cell keys (and the identity permutation) of a re-bin that follows in the same
// substep — the positions are in registers here anyway, which saves the sort its own pass over them
// (cell_key_kernel, sort.cuh; same key: clamped base node, x local to the slab, z fastest).
template <int MODEL, class O, bool EXACT, bool ONE_MAT, bool KEYS>
__global__ void __launch_bounds__(kP2gBlock, MPM_P2G_MINBLK)
p2g_sched_kernel(Soa p, size_t count, const MpmMaterial* __restrict__ mats, const MpmMaterial mat0, float4* __restrict__ grid,
                 KParams k, const __grid_constant__ CUtensorMap tm_streams, uint32_t* __restrict__ sort_keys, uint32_t* __restrict__ sort_vals) {
  __shared__ P2gSmem sm;
  const int tid = threadIdx.x;
  const size_t pi = (size_t)blockIdx.x * kP2gBlock + tid;
#if MPM_P2G_TMA
  // The block's 25 x 256 floats as one tiled TMA load: no registers are held while the data
  // travels, and phase 0 reads its inputs with immediate-offset LDS instead of 25 address
  // computations + LDG.  Columns beyond the stream length arrive as zeros and are not used.
  __shared__ __align__(128) float st[NSTREAM * kP2gBlock];
  __shared__ uint64_t st_bar;
  if (tid == 0) {
    mbar_init(&st_bar, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&st_bar, (uint32_t)sizeof(st));
    tma_load_2d(st, &tm_streams, (int)(blockIdx.x * kP2gBlock), 0, &st_bar);
  }
#endif
  if (tid < 32) sm.hist[tid] = 0;
  __syncthreads();
#if MPM_P2G_TMA
  mbar_wait(&st_bar, 0);
#define MPM_P2G_IN(stream) st[(stream) * kP2gBlock + tid]
#else
#define MPM_P2G_IN(stream) p.s(stream)[pi]
#endif

  // ---------------- phase 0: per-particle payload ----------------
  uint32_t key = kInvalidKey;
  if (pi < count) {
    float x[3], v[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      x[a] = MPM_P2G_IN(SX + a);
      v[a] = MPM_P2G_IN(SV + a);
    }
    Mat3 F, C;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        F.m[r][c] = MPM_P2G_IN(SF + 3 * r + c);
        C.m[r][c] = MPM_P2G_IN(SC + 3 * r + c);
      }
    const float Jp = (MODEL == MPM_MODEL_SNOW) ? MPM_P2G_IN(SJ) : 1.0f;  // fixed-corotated never changes Jp
#undef MPM_P2G_IN
    if (KEYS) {
      int b[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) b[a] = min(max((int)(x[a] * k.dx_inv - 0.5f), 0), k.N - 1);
      const int bx = min(max(b[0] - k.x0, 0), k.nxl - 1);
      sort_keys[pi] = (uint32_t)((bx * k.N + b[1]) * k.N + b[2]);
      sort_vals[pi] = (uint32_t)pi;
    }
    MpmMaterial m;
    if constexpr (ONE_MAT) m = mat0; else m = load_material(mats, p.mat[pi]);  // ONE_MAT: operands straight from the constant bank
    const P2gPayload o = p2g_prepare<MODEL, O, EXACT>(x, v, F, C, Jp, m, k);
    key = o.key;
    sm.pay[tid][0] = make_float4(o.f[0], o.f[1], o.f[2], o.mass);
    sm.pay[tid][1] = make_float4(o.q0[0], o.q0[1], o.q0[2], o.cx[0]);
    sm.pay[tid][2] = make_float4(o.cx[1], o.cx[2], o.cy[0], o.cy[1]);
    sm.pay[tid][3] = make_float4(o.cy[2], o.cz[0], o.cz[1], o.cz[2]);
  }
  sm.key[tid] = key;
  p2g_list_and_scatter_runs(sm, key, tid, grid, k, BlockBarrier());
}

}  // namespace mpm
