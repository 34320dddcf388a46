// Launchers of the particle stages for one MaterialModel type, and the model registry.
//
// Templates cannot cross a C ABI, so libmpm_b200.so instantiates the substep kernels for the
// shipped plugin tuples — (MLS_APIC_Scheme, QuadraticInterpolationKernel, {MMSnow, MMFixedCorotated,
// MMJelly} x {ExactOps, FastOps}) — and selects one by MpmParams.model / svd_mode, which mirrors the
// reference's compile-time aliases (include/mpm.cuh:24-27).  A user-defined material is compiled in
// by one more instantiation: see include/mpm_b200/plugin.cuh.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "g2p_tile.cuh"
#include "kernels.cuh"
#include "p2g_sched.cuh"

namespace mpm {

// everything a particle-stage launch needs from the handle
struct LaunchCtx {
  cudaStream_t stream;
  int n_sms;
  Soa soa;
  size_t count;
  const void* mats_dev;   // n_mats material objects of the model's type, device memory
  const void* mats_host;  // the same on the host
  int n_mats;
  float4* grid;
  KParams k;
  DeviceDiag* diag;
  int p2g_mode, g2p_mode;
  // P2G
  bool handover_in;           // the C rows hold dx * affine (written by the previous G2P)
  size_t tile_begin, tile_end;  // tiles [begin, end) of the SoA (split launches of slab handles)
  bool stale_order;           // enough particles changed cell since the last re-bin: warps re-order their records
  uint32_t* sort_keys;        // non-null: also emit the cell keys of the re-bin that follows
  uint32_t* sort_vals;
  // G2P
  bool emit;                  // write dx * affine of the next P2G into the C rows
  unsigned long long* moved;  // non-null: count cell crossings (adaptive re-bin)
  unsigned int* tile_counters;
  int tile_parity;
};

struct ModelOps {
  const char* name;
  size_t material_bytes;
  // n records of the C ABI (MpmMaterial = MMSnow's 7 floats) -> n objects of this model's type
  void (*from_abi)(const MpmMaterial* in, int n, void* out);
  bool staged;  // has the staged kernels (P2G_RUNS / G2P_TILE / hand-over); false = generic kernels only
  void (*p2g)(const LaunchCtx&);
  void (*g2p)(const LaunchCtx&);
};

// model id -> ops, per svd_mode (0 exact, 1 fast).  Ids 0..15 are reserved for the shipped models.
int register_model(uint32_t id, const ModelOps* exact, const ModelOps* fast);
const ModelOps* find_model(uint32_t id, uint32_t svd_mode);

inline unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

template <class Material, class Kernel = DefaultKernel, class Scheme = DefaultScheme>
struct ModelImpl {
  static constexpr bool kStaged = std::is_same<Kernel, DefaultKernel>::value && std::is_same<Scheme, DefaultScheme>::value;

  static MatTable<Material> table(const LaunchCtx& c) {
    MatTable<Material> t;
    t.one = *static_cast<const Material*>(c.mats_host);
    t.all = static_cast<const Material*>(c.mats_dev);
    t.n = c.n_mats;
    return t;
  }

  template <bool ONE_MAT, bool HANDOVER, bool SORT>
  static void p2g_sched(const LaunchCtx& c) {
    Soa view = c.soa;
    view.f += c.tile_begin * (size_t)kTileFloats;  // the kernel indexes tiles and ids from its first block
    view.id += c.tile_begin * kTile;
    view.mat += c.tile_begin * kTile;
    const size_t first = c.tile_begin * kTile;
    const size_t n = std::min(c.count, c.tile_end * kTile) - first;
    p2g_sched_kernel<Material, ONE_MAT, HANDOVER, SORT><<<(unsigned)(c.tile_end - c.tile_begin), kP2gBlock, 0, c.stream>>>(
        view, n, table(c), c.grid, c.k, c.sort_keys ? c.sort_keys + first : nullptr, c.sort_vals ? c.sort_vals + first : nullptr,
        (uint32_t)first, c.diag);
  }

  static void p2g(const LaunchCtx& c) {
    if (c.count == 0 || c.tile_begin >= c.tile_end) return;
    if constexpr (kStaged) {
      if (c.p2g_mode == MPM_P2G_RUNS && c.k.N <= kP2gMaxN) {
        const bool one = c.n_mats == 1;
        if (c.stale_order) {
          if (c.handover_in) one ? p2g_sched<true, true, true>(c) : p2g_sched<false, true, true>(c);
          else one ? p2g_sched<true, false, true>(c) : p2g_sched<false, false, true>(c);
        } else {
          if (c.handover_in) one ? p2g_sched<true, true, false>(c) : p2g_sched<false, true, false>(c);
          else one ? p2g_sched<true, false, false>(c) : p2g_sched<false, false, false>(c);
        }
        return;
      }
    }
    p2g_generic_kernel<Material, Kernel, Scheme><<<blocks_for(c.count, kParticleBlock), kParticleBlock, 0, c.stream>>>(
        c.soa, c.count, static_cast<const Material*>(c.mats_dev), c.grid, c.k, Kernel());
  }

  template <bool ONE_MAT, bool EMIT>
  static void g2p_tile(const LaunchCtx& c) {
    const size_t smem = G2pTileLayout::bytes(MaterialTraits<Material>::kMutatesJp);
    const size_t n_tiles = (c.count + kTile - 1) / kTile;
    const unsigned ctas = (unsigned)std::min<size_t>(n_tiles, (size_t)c.n_sms * 4);  // 4 CTAs per SM: __launch_bounds__(256, 4), 40 KB each
    g2p_tile_kernel<Material, ONE_MAT, EMIT><<<ctas, kG2pThreads, smem, c.stream>>>(c.soa, table(c), c.grid, c.k, c.count, c.moved,
                                                                                     c.tile_counters, c.tile_parity, c.diag);
  }

  static void g2p(const LaunchCtx& c) {
    if (c.count == 0) return;
    if constexpr (kStaged) {
      if (c.g2p_mode == MPM_G2P_TILE) {
        const bool one = c.n_mats == 1;
        if (c.emit) one ? g2p_tile<true, true>(c) : g2p_tile<false, true>(c);
        else one ? g2p_tile<true, false>(c) : g2p_tile<false, false>(c);
        return;
      }
    }
    g2p_generic_kernel<Material, Kernel, Scheme><<<blocks_for(c.count, kParticleBlock), kParticleBlock, 0, c.stream>>>(
        c.soa, c.count, static_cast<const Material*>(c.mats_dev), c.grid, c.k, Kernel());
  }

  // default conversion from the C-ABI record: the leading floats of MMSnow's layout
  // (volume, mass, mu0, lambda0, hardening, clamp lo, clamp hi), as many as the type holds
  static void from_abi_prefix(const MpmMaterial* in, int n, void* out) {
    static_assert(sizeof(Material) <= sizeof(MpmMaterial), "materials larger than MpmMaterial need their own from_abi / mpm_create_raw");
    for (int i = 0; i < n; ++i) memcpy(static_cast<char*>(out) + (size_t)i * sizeof(Material), &in[i], sizeof(Material));
  }

  static const ModelOps* ops(const char* name) {
    static const ModelOps o = {name, sizeof(Material), &from_abi_prefix, kStaged, &p2g, &g2p};
    return &o;
  }
};

}  // namespace mpm
