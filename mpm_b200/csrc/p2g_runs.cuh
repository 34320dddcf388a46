// Stage (2), production kernel: P2G with in-register pre-reduction of same-cell particles.
//
// The direct scatter (p2g_kernel, kernels.cuh) is bound by reduction lanes: 27 REDG per particle
// at ~1 lane/clk/SM.  Particles are cell-sorted, so consecutive particles mostly share the base
// node and therefore all 27 target nodes.  Per block of kP2gBlock consecutive particles:
//   phase 0  one thread per particle: load the 25 streams (coalesced), stress via polar/svd3,
//            affine matrix; write a 16-float payload (fractional position, m v + A d at the base
//            node, dx*A columns, mass) and the packed base node to shared memory.
//   phase R  run detection: a run = maximal stretch of consecutive particles with equal base
//            node, cut at kRunCap particles (bounded trip count -> balanced warps).
//   phase 1  three threads per run, one per stencil x-slab (9 nodes, 36 accumulators in
//            registers): loop over the run's payloads, accumulate w*(q, m) with FFMAs, then ONE
//            red.global.add.v4.f32 per node per run — ~6x fewer reduction lanes than per particle.
// Correctness does not depend on the order being perfectly sorted (a stale order only shortens
// the runs).  Reference behaviour: src/mpm.cu:14-74, TransferScheme.h:66-100.
#pragma once
#include "common.cuh"

namespace mpm {

#ifndef MPM_P2G_MINBLK
#define MPM_P2G_MINBLK 3
#endif
#ifndef MPM_RUN_CAP
#define MPM_RUN_CAP 16
#endif
constexpr int kP2gBlock = 256;
constexpr int kRunCap = MPM_RUN_CAP;
constexpr uint32_t kInvalidKey = 0xffffffffu;
constexpr int kKeyBias = 4;  // base node >= -3 for particles that are not skipped

struct P2gSmem {
  // per-particle payload record, 4 x float4 used of a 5 x float4 (80 B) stride: consecutive
  // records start 20 banks apart, so the 128-bit reads of 8 different runs are conflict-free
  //   [0] = (fx, fy, fz, mass)  [1] = (q0.xyz, cx.x)  [2] = (cx.y, cx.z, cy.x, cy.y)  [3] = (cy.z, cz.xyz)
  float4 pay[kP2gBlock][5];
  uint32_t key[kP2gBlock];
  uint16_t run_start[kP2gBlock + 1];
  uint32_t warp_tmp[kP2gBlock / 32];
  uint32_t n_runs;
};

template <int MODEL, class O, bool EXACT>
__global__ void __launch_bounds__(kP2gBlock, MPM_P2G_MINBLK)
p2g_runs_kernel(Soa p, size_t count, const MpmMaterial* __restrict__ mats, float4* __restrict__ grid, KParams k) {
  __shared__ P2gSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t pi = (size_t)blockIdx.x * kP2gBlock + tid;

  // ---------------- phase 0: per-particle payload ----------------
  uint32_t key = kInvalidKey;
  if (pi < count) {
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
    float fx[3], w[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) bspline(x[a], k.dx_inv, base[a], fx[a], w);
    const Mat3 PF = compute_PF<MODEL, O, EXACT>(F, Jp, m);
    const float kk = ((-k.dinv) * k.dt) * m.particleVolume;
    Mat3 A;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) A.m[i][j] = kk * PF.m[i][j] + m.particleMass * C.m[i][j];
    bool inside = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) inside = inside && !(base[a] + 3 < 0 || base[a] >= k.N);
    if (inside) key = ((uint32_t)(base[0] + kKeyBias) << 20) | ((uint32_t)(base[1] + kKeyBias) << 10) | (uint32_t)(base[2] + kKeyBias);
    float d0[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) d0[a] = (float)base[a] * k.dx - x[a];
    float q0[3], ccx[3], ccy[3], ccz[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      q0[c] = v[c] * m.particleMass + (A.m[c][0] * d0[0] + A.m[c][1] * d0[1] + A.m[c][2] * d0[2]);
      ccx[c] = A.m[c][0] * k.dx;
      ccy[c] = A.m[c][1] * k.dx;
      ccz[c] = A.m[c][2] * k.dx;
    }
    sm.pay[tid][0] = make_float4(fx[0], fx[1], fx[2], m.particleMass);
    sm.pay[tid][1] = make_float4(q0[0], q0[1], q0[2], ccx[0]);
    sm.pay[tid][2] = make_float4(ccx[1], ccx[2], ccy[0], ccy[1]);
    sm.pay[tid][3] = make_float4(ccy[2], ccz[0], ccz[1], ccz[2]);
  }
  sm.key[tid] = key;
  __syncthreads();

  // ---------------- phase R: runs ----------------
  const bool head = (tid == 0) || (sm.key[tid - 1] != key);
  // start of the natural run containing tid: inclusive max-scan of (head ? tid : 0)
  uint32_t rs = head ? (uint32_t)tid : 0u;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, rs, o);
    if (lane >= o) rs = max(rs, t);
  }
  if (lane == 31) sm.warp_tmp[warp] = rs;
  __syncthreads();
  {
    uint32_t carry = 0;
    for (int w2 = 0; w2 < warp; ++w2) carry = max(carry, sm.warp_tmp[w2]);
    rs = max(rs, carry);
  }
  const bool head2 = head || (((uint32_t)tid - rs) % kRunCap == 0);
  const uint32_t ballot = __ballot_sync(0xffffffffu, head2);
  __syncthreads();  // warp_tmp reuse
  if (lane == 0) sm.warp_tmp[warp] = __popc(ballot);
  __syncthreads();
  {
    uint32_t off = 0;
    for (int w2 = 0; w2 < warp; ++w2) off += sm.warp_tmp[w2];
    const uint32_t idx = off + __popc(ballot & ((1u << lane) - 1u));
    if (head2) sm.run_start[idx] = (uint16_t)tid;
    if (tid == kP2gBlock - 1) {
      const uint32_t n = idx + (head2 ? 1u : 0u);
      sm.n_runs = n;
      sm.run_start[n] = kP2gBlock;
    }
  }
  __syncthreads();
  const int n_runs = (int)sm.n_runs;

  // ---------------- phase 1: 3 threads per run ----------------
  const long long NN = (long long)k.N * k.N;
  for (int u = tid; u < 3 * n_runs; u += kP2gBlock) {
    const int r = u / 3, i = u - 3 * r;
    const int s0 = sm.run_start[r], s1 = sm.run_start[r + 1];
    const uint32_t rk = sm.key[s0];
    if (rk == kInvalidKey) continue;  // skipped particles / tail of the last block
    float4 acc[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int kz = 0; kz < 3; ++kz) acc[j][kz] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float fi = (float)i;
    for (int s = s0; s < s1; ++s) {
      const float4 r0 = sm.pay[s][0], r1 = sm.pay[s][1], r2 = sm.pay[s][2], r3 = sm.pay[s][3];
      const float fx = r0.x, fy = r0.y, fz = r0.z;
      // weight of this thread's x-slab
      const float dxi = (i == 0) ? (1.5f - fx) : ((i == 1) ? (fx - 1.0f) : (fx - 0.5f));
      const float wxi = (i == 1) ? (0.75f - dxi * dxi) : (0.5f * (dxi * dxi));
      float wy[3], wz[3];
      {
        const float a0 = 1.5f - fy, a1 = fy - 1.0f, a2 = fy - 0.5f;
        wy[0] = 0.5f * (a0 * a0); wy[1] = 0.75f - a1 * a1; wy[2] = 0.5f * (a2 * a2);
        const float b0 = 1.5f - fz, b1 = fz - 1.0f, b2 = fz - 0.5f;
        wz[0] = 0.5f * (b0 * b0); wz[1] = 0.75f - b1 * b1; wz[2] = 0.5f * (b2 * b2);
      }
      const float mass = r0.w;
      const float qi[3] = {r1.x + fi * r1.w, r1.y + fi * r2.x, r1.z + fi * r2.y};
      const float cy[3] = {r2.z, r2.w, r3.x};
      const float cz[3] = {r3.y, r3.z, r3.w};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float wij = wxi * wy[j];
        float q[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) q[c] = qi[c] + (float)j * cy[c];
#pragma unroll
        for (int kz = 0; kz < 3; ++kz) {
          const float wt = wij * wz[kz];
          acc[j][kz].x += wt * q[0];
          acc[j][kz].y += wt * q[1];
          acc[j][kz].z += wt * q[2];
          acc[j][kz].w += wt * mass;
#pragma unroll
          for (int c = 0; c < 3; ++c) q[c] += cz[c];
        }
      }
    }
    // flush: one vector reduction per node of this slab
    const int bx = (int)(rk >> 20) - kKeyBias, by = (int)((rk >> 10) & 1023u) - kKeyBias, bz = (int)(rk & 1023u) - kKeyBias;
    const int gx = bx + i;
    if (gx < 0 || gx >= k.N || gx < k.x0 || gx >= k.x0 + k.nxl) continue;
    float4* gp = grid + ((long long)(gx - k.x0) * NN + (long long)by * k.N + bz);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int gy = by + j;
      if (gy < 0 || gy >= k.N) continue;
#pragma unroll
      for (int kz = 0; kz < 3; ++kz) {
        const int gz = bz + kz;
        if (gz < 0 || gz >= k.N) continue;
        atomicAdd(gp + (j * k.N + kz), acc[j][kz]);
      }
    }
  }
}

}  // namespace mpm
