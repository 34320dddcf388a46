// C ABI of the B200-native MLS-MPM substep: handle, buffers, stage launches.
// Replaces the device side of the reference's Simulation class (src/mpm.cu:180-329).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "../../include/mpm_b200.h"
#include "comm.cuh"
#include "kernels.cuh"
#include "g2p_tile.cuh"
#include "g2p2g.cuh"
#include "p2g_sched.cuh"
#include "sort.cuh"

using namespace mpm;

constexpr int kLtSmall = 48, kLtLarge = 96;  // box lengths the staged kernels are instantiated for

static thread_local std::string g_create_error;

struct MpmSim {
  MpmParams par{};
  KParams k{};
  int device = 0;
  int n_sms = 148;
  cudaStream_t stream = nullptr;

  // particles: two SoA buffers (sort permutes from one into the other)
  Soa soa[2]{};
  int cur = 0;
  size_t capacity = 0;
  size_t count = 0;
  uint32_t first_id = 0;

  // grid: nxl * N * N float4
  float4* grid = nullptr;
  size_t grid_nodes = 0;
  // fused mode (g2p2g.cuh): second grid, the two swap roles every substep.  grid_ready = `grid`
  // already holds the updated velocities of the NEXT substep (scattered by the previous fused
  // kernel from the particles' current state); any change to the particles or the grid clears it.
  float4* grid_b = nullptr;
  bool fused = false;
  bool grid_ready = false;

  MpmMaterial* mats = nullptr;
  MpmMaterial mat0{};  // host copy of material 0: single-material handles pass it as a kernel parameter
  int n_mats = 0;

  // sort scratch
  uint32_t* keys[2] = {nullptr, nullptr};
  uint32_t* vals[2] = {nullptr, nullptr};
  uint32_t* table = nullptr;
  size_t table_len = 0;
  uint32_t* scan_sums = nullptr;
  size_t scan_sums_len = 0;
  int key_bits = 0;
  int ghost = 0;
  bool whole_domain = true;
  int sorted_cur = 0;  // which keys[] buffer holds the keys of the current order
  // particle tiles of the staged kernels (tiles.cuh), rebuilt at every re-bin
  TileDesc* tiles = nullptr;
  size_t tiles_cap = 0;
  uint32_t* row_first = nullptr;  // n_rows + 1
  uint32_t* tile_base = nullptr;  // n_rows + 1
  uint32_t* d_n_tiles = nullptr;
  uint32_t n_rows = 0;
  // box length (nodes along z) of the staged kernels, chosen from the measured cell span of the
  // tiles at the last re-bin (read back asynchronously, never waited for)
  int tile_lt = kLtLarge;
  unsigned int* d_span = nullptr;  // [0] tiles whose box fits kLtSmall, [1] tiles
  unsigned int* h_span = nullptr;  // pinned mirror
  cudaEvent_t span_ev = nullptr;
  bool span_pending = false;
  // TMA descriptors: grid as [nxl][N][N][4 floats] with a 5 x 5 x LT box; particle streams of
  // each SoA buffer as [NSTREAM][stride] with kTile-column boxes of 12 / 13 / 25 rows
  CUtensorMap tm_grid[2];         // kLtSmall, kLtLarge
  CUtensorMap tm_grid_b[2];       // the same over grid_b (fused mode)
  CUtensorMap tm_streams[2][3];   // [soa buffer][12, 13, 25 rows]

  MpmParticle* aos_stage = nullptr;  // device AoS staging for upload/download
  size_t aos_stage_cap = 0;
  unsigned long long* d_counter = nullptr;

  // adaptive re-bin (MpmParams.rebin_permille): cell crossings counted by the G2P tile kernel since the
  // last re-bin, read back asynchronously (never waited for)
  unsigned int* d_tile_counters = nullptr;  // [2], used alternately by the dynamic tile scheduler of G2P
  int tile_parity = 0;
  unsigned long long* d_moved = nullptr;
  unsigned long long* h_moved = nullptr;  // pinned
  cudaEvent_t moved_ev = nullptr;
  bool moved_pending = false;
  uint64_t moved_issued_at = 0;
  unsigned long long moved_seen = 0;
  uint64_t rebins = 0;

  uint64_t substeps = 0;
  uint64_t steps_since_sort = 0;
  uint64_t launches = 0;
  double t = 0.0;

  bool timing = false;
  cudaEvent_t ev[2]{};
  float stage_ms[MPM_STAGE_COUNT]{};

  Comm comm;
  std::string err;
};

namespace {

int fail(MpmSim* s, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (s) s->err = buf; else g_create_error = buf;
  return 1;
}

#define CK(call)                                                                                    \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) return fail(sim, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)) ? (int)e_ : (int)e_; \
  } while (0)

inline unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int make_stream_maps(MpmSim* sim, int buf) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail(sim, "cuTensorMapEncodeTiled is not available from this driver");
  const Soa& s = sim->soa[buf];
  const int rows[3] = {12, 13, NSTREAM};
  for (int r = 0; r < 3; ++r) {
    const cuuint64_t dims[2] = {(cuuint64_t)s.stride, (cuuint64_t)NSTREAM};
    const cuuint64_t strides[1] = {(cuuint64_t)s.stride * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kTile, (cuuint32_t)rows[r]};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = enc(&sim->tm_streams[buf][r], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, s.f, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(sim, "cuTensorMapEncodeTiled(streams) failed: %d", (int)rc);
  }
  return 0;
}

int make_grid_maps(MpmSim* sim, float4* grid, CUtensorMap* out) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail(sim, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t N = (cuuint64_t)sim->k.N;
  const int lts[2] = {kLtSmall, kLtLarge};
  for (int i = 0; i < 2; ++i) {
    const cuuint64_t dims[4] = {4, N, N, (cuuint64_t)sim->k.nxl};
    const cuuint64_t strides[3] = {16, 16 * N, 16 * N * N};
    const cuuint32_t box[4] = {4, (cuuint32_t)lts[i], (cuuint32_t)kBoxW, (cuuint32_t)kBoxW};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult rc = enc(&out[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, grid, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(sim, "cuTensorMapEncodeTiled(grid) failed: %d", (int)rc);
  }
  return 0;
}

int alloc_soa(MpmSim* sim, Soa& s, size_t cap) {
  s.stride = (cap + 31) / 32 * 32;
  CK(cudaMalloc(&s.f, sizeof(float) * NSTREAM * s.stride));
  CK(cudaMalloc(&s.id, sizeof(uint32_t) * s.stride));
  CK(cudaMalloc(&s.mat, s.stride));
  return 0;
}
void free_soa(Soa& s) {
  cudaFree(s.f);
  cudaFree(s.id);
  cudaFree(s.mat);
  s = Soa{};
}

int ensure_capacity(MpmSim* sim, size_t cap) {
  if (cap <= sim->capacity) return 0;
  if (sim->capacity != 0) {
    for (int b = 0; b < 2; ++b) free_soa(sim->soa[b]);
    for (int b = 0; b < 2; ++b) {
      cudaFree(sim->keys[b]);
      cudaFree(sim->vals[b]);
    }
    cudaFree(sim->table);
    cudaFree(sim->scan_sums);
    cudaFree(sim->tiles);
  }
  for (int b = 0; b < 2; ++b) {
    if (int rc = alloc_soa(sim, sim->soa[b], cap)) return rc;
    if (int rc = make_stream_maps(sim, b)) return rc;
    CK(cudaMalloc(&sim->keys[b], sizeof(uint32_t) * cap));
    CK(cudaMalloc(&sim->vals[b], sizeof(uint32_t) * cap));
  }
  sim->tiles_cap = cap / kTileMax + sim->n_rows + 1;  // every row adds at most one partial tile
  CK(cudaMalloc(&sim->tiles, sizeof(TileDesc) * sim->tiles_cap));
  const size_t n_tiles = (cap + kSortTile - 1) / kSortTile;
  sim->table_len = n_tiles * kRadix;
  CK(cudaMalloc(&sim->table, sizeof(uint32_t) * sim->table_len));
  sim->scan_sums_len = (sim->table_len + kScanTile - 1) / kScanTile;
  CK(cudaMalloc(&sim->scan_sums, sizeof(uint32_t) * sim->scan_sums_len));
  sim->capacity = cap;
  return 0;
}

int ensure_stage(MpmSim* sim, size_t n) {
  if (n <= sim->aos_stage_cap) return 0;
  cudaFree(sim->aos_stage);
  sim->aos_stage = nullptr;
  sim->aos_stage_cap = 0;
  CK(cudaMalloc(&sim->aos_stage, sizeof(MpmParticle) * n));
  sim->aos_stage_cap = n;
  return 0;
}

struct StageTimer {
  MpmSim* s;
  int stage;
  StageTimer(MpmSim* s_, int st) : s(s_), stage(st) {
    if (s->timing) cudaEventRecord(s->ev[0], s->stream);
  }
  ~StageTimer() {
    if (s->timing) {
      cudaEventRecord(s->ev[1], s->stream);
      cudaEventSynchronize(s->ev[1]);
      float ms = 0;
      cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]);
      s->stage_ms[stage] += ms;
    }
  }
};

// ---- stages -------------------------------------------------------------------------------------
// tile descriptors of the current (freshly sorted) order, from keys[sorted_cur]
int build_tiles(MpmSim* sim) {
  const uint32_t* keys = sim->keys[sim->sorted_cur];
  const uint32_t n_rows = sim->n_rows;
  row_bounds_kernel<<<blocks_for((size_t)n_rows + 1, 256), 256, 0, sim->stream>>>(keys, (uint32_t)sim->count, (uint32_t)sim->k.N, n_rows,
                                                                                  sim->row_first);
  row_tiles_scan_kernel<<<1, 1024, 0, sim->stream>>>(sim->row_first, n_rows, sim->tile_base, sim->d_n_tiles);
  CK(cudaMemsetAsync(sim->d_span, 0, 2 * sizeof(unsigned int), sim->stream));
  tile_fill_kernel<<<blocks_for(n_rows, 256), 256, 0, sim->stream>>>(keys, sim->row_first, sim->tile_base, n_rows, sim->tiles, kLtSmall,
                                                                     sim->d_span);
  sim->launches += 3;
  CK(cudaMemcpyAsync(sim->h_span, sim->d_span, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaEventRecord(sim->span_ev, sim->stream));
  sim->span_pending = true;
  return 0;
}

// partial: called between the grid update and G2P, permutes only what G2P reads (sort.cuh)
int do_sort(MpmSim* sim, bool partial = false, bool keys_ready = false) {
  StageTimer tm(sim, MPM_STAGE_SORT);
  sim->steps_since_sort = 0;
  sim->rebins++;
  sim->moved_seen = 0;
  sim->moved_pending = false;
  if (sim->d_moved) CK(cudaMemsetAsync(sim->d_moved, 0, sizeof(unsigned long long), sim->stream));
  size_t n_dead = 0;
  if (sim->comm.active()) {  // leavers out (tombstoned), arrivals appended, before the re-bin
    if (sim->comm.migrate(sim->soa[sim->cur], &sim->count, sim->capacity, sim->k, sim->stream, &sim->launches, &n_dead))
      return fail(sim, "particle migration failed: %s", sim->comm.error());
  }
  const size_t n = sim->count;
  if (n == 0) {
    CK(cudaMemsetAsync(sim->d_n_tiles, 0, sizeof(uint32_t), sim->stream));
    return 0;
  }
  Soa& src = sim->soa[sim->cur];
  const uint32_t dead_key = 1u << sim->key_bits;
  const int sort_bits = sim->key_bits + (sim->comm.active() ? 1 : 0);
  if (!keys_ready) {  // else the P2G of this substep wrote them (p2g_sched.cuh, KEYS)
    cell_key_kernel<<<blocks_for(n, 256), 256, 0, sim->stream>>>(src, n, sim->k, sim->keys[0], sim->vals[0], dead_key);
    sim->launches++;
  }
  const int n_tiles = (int)((n + kSortTile - 1) / kSortTile);
  const size_t table_len = (size_t)n_tiles * kRadix;
  const unsigned scan_blocks = blocks_for(table_len, kScanTile);
  int in = 0;
  for (int shift = 0; shift < sort_bits; shift += kRadixBits) {
    radix_hist_kernel<<<n_tiles, kSortThreads, 0, sim->stream>>>(sim->keys[in], n, shift, sim->table, n_tiles);
    scan_tile_sums_kernel<<<scan_blocks, kScanThreads, 0, sim->stream>>>(sim->table, table_len, sim->scan_sums);
    scan_sums_kernel<<<1, 1024, 0, sim->stream>>>(sim->scan_sums, scan_blocks);
    scan_downsweep_kernel<<<scan_blocks, kScanThreads, 0, sim->stream>>>(sim->table, table_len, sim->scan_sums);
    radix_scatter_kernel<<<n_tiles, kSortThreads, 0, sim->stream>>>(sim->keys[in], sim->vals[in], sim->keys[in ^ 1],
                                                                     sim->vals[in ^ 1], n, shift, sim->table, n_tiles);
    sim->launches += 5;
    in ^= 1;
  }
  sim->sorted_cur = in;
  Soa& dst = sim->soa[sim->cur ^ 1];
  if (partial) permute_kernel<true><<<blocks_for(n, 256), 256, 0, sim->stream>>>(src, dst, sim->vals[in], n, sim->k);
  else permute_kernel<false><<<blocks_for(n, 256), 256, 0, sim->stream>>>(src, dst, sim->vals[in], n, sim->k);
  sim->launches++;
  sim->cur ^= 1;
  sim->count = n - n_dead;  // tombstones were sorted behind the live particles
  // row tiles are needed by the fused kernel and by the shared-memory node sources of G2P only
  if (sim->fused || !MPM_G2P_FLAT_TILES) {
    if (int rc = build_tiles(sim)) return rc;
  }
  CK(cudaGetLastError());
  return 0;
}

int do_reset(MpmSim* sim, float4* grid = nullptr) {
  StageTimer tm(sim, MPM_STAGE_RESET);
  CK(cudaMemsetAsync(grid ? grid : sim->grid, 0, sizeof(float4) * sim->grid_nodes, sim->stream));
  return 0;
}

template <int MODEL, class O, bool EXACT, bool KEYS>
void launch_p2g_sched_k(MpmSim* sim) {
  const size_t n = sim->count;
  const unsigned nbr = blocks_for(n, kP2gBlock);
  if (sim->n_mats == 1)
    p2g_sched_kernel<MODEL, O, EXACT, true, KEYS><<<nbr, kP2gBlock, 0, sim->stream>>>(sim->soa[sim->cur], n, sim->mats, sim->mat0, sim->grid, sim->k,
                                                                                      sim->tm_streams[sim->cur][2], sim->keys[0], sim->vals[0]);
  else
    p2g_sched_kernel<MODEL, O, EXACT, false, KEYS><<<nbr, kP2gBlock, 0, sim->stream>>>(sim->soa[sim->cur], n, sim->mats, sim->mat0, sim->grid, sim->k,
                                                                                       sim->tm_streams[sim->cur][2], sim->keys[0], sim->vals[0]);
}
template <int MODEL, class O, bool EXACT>
void launch_p2g_sched(MpmSim* sim, bool keys) {
  if (keys) launch_p2g_sched_k<MODEL, O, EXACT, true>(sim); else launch_p2g_sched_k<MODEL, O, EXACT, false>(sim);
}

// keys: also write the cell keys of the re-bin that follows in this substep; *keys_done says whether that happened
template <int MODEL>
int launch_p2g(MpmSim* sim, bool keys, bool* keys_done) {
  const size_t n = sim->count;
  Soa& p = sim->soa[sim->cur];
  *keys_done = false;
  if (sim->par.p2g_mode == MPM_P2G_RUNS && sim->k.N + kKeyBias <= 1023) {
    if (sim->par.svd_mode == MPM_SVD_EXACT) launch_p2g_sched<MODEL, ExactOps, true>(sim, keys); else launch_p2g_sched<MODEL, FastOps, false>(sim, keys);
    *keys_done = keys;
    return 0;
  }
  const unsigned nb = blocks_for(n, kParticleBlock);
  if (sim->par.svd_mode == MPM_SVD_EXACT)
    p2g_kernel<MODEL, ExactOps, true><<<nb, kParticleBlock, 0, sim->stream>>>(p, n, sim->mats, sim->grid, sim->k);
  else
    p2g_kernel<MODEL, FastOps, false><<<nb, kParticleBlock, 0, sim->stream>>>(p, n, sim->mats, sim->grid, sim->k);
  return 0;
}
int do_p2g(MpmSim* sim, bool keys = false, bool* keys_done = nullptr) {
  StageTimer tm(sim, MPM_STAGE_P2G);
  bool done = false;
  if (keys_done) *keys_done = false;
  if (sim->count == 0) return 0;
  if (sim->par.model == MPM_MODEL_SNOW) launch_p2g<MPM_MODEL_SNOW>(sim, keys, &done); else launch_p2g<MPM_MODEL_FIXED_COROTATED>(sim, keys, &done);
  if (keys_done) *keys_done = done;
  sim->launches++;
  CK(cudaGetLastError());
  return 0;
}

int do_grid(MpmSim* sim) {
  StageTimer tm(sim, MPM_STAGE_GRID);
  const size_t nodes = sim->grid_nodes;
  grid_update_kernel<<<blocks_for(nodes, 256), 256, 0, sim->stream>>>(sim->grid, sim->k, 0, sim->k.nxl);
  sim->launches++;
  CK(cudaGetLastError());
  return 0;
}

template <int MODEL, class O, int LT, bool COUNT_MOVED>
void launch_g2p_tile_impl(MpmSim* sim) {
  const size_t smem = G2pTileLayout<MODEL>::bytes(LT);
  static int per_sm = 0;  // per instantiation
  if (!per_sm) {
    cudaFuncSetAttribute(g2p_tile_kernel<MODEL, O, LT, COUNT_MOVED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, g2p_tile_kernel<MODEL, O, LT, COUNT_MOVED>, kG2pThreads, smem);
    per_sm = std::max(per_sm, 1);
  }
  const size_t max_tiles = MPM_G2P_FLAT_TILES ? (sim->count + kTile - 1) / kTile : sim->count / kTileMax + sim->n_rows + 1;
  const unsigned ctas = (unsigned)std::min<size_t>(max_tiles, (size_t)sim->n_sms * per_sm);
  g2p_tile_kernel<MODEL, O, LT, COUNT_MOVED><<<ctas, kG2pThreads, smem, sim->stream>>>(
      sim->soa[sim->cur], sim->mats, sim->mat0, sim->n_mats == 1, sim->grid, sim->k, sim->tiles, sim->d_n_tiles, sim->tm_grid[LT == kLtSmall ? 0 : 1],
      sim->tm_streams[sim->cur][MODEL == MPM_MODEL_SNOW ? 1 : 0], sim->d_moved, sim->count, sim->d_tile_counters, sim->tile_parity);
  sim->tile_parity ^= 1;
}
template <int MODEL, class O, int LT>
void launch_g2p_tile(MpmSim* sim) {  // the cell-crossing count costs a register the default path cannot spare
  if (sim->par.rebin_permille) launch_g2p_tile_impl<MODEL, O, LT, true>(sim); else launch_g2p_tile_impl<MODEL, O, LT, false>(sim);
}

template <int MODEL>
int launch_g2p(MpmSim* sim) {
  const size_t n = sim->count;
  if (sim->par.g2p_mode == MPM_G2P_TILE) {
    if (sim->span_pending && cudaEventQuery(sim->span_ev) == cudaSuccess) {
      sim->span_pending = false;
      // the small box when (nearly) every tile fits it; particles beyond the box take the
      // global-memory gather, which is correct but slow
      sim->tile_lt = (sim->h_span[1] && (double)sim->h_span[0] >= 0.95 * (double)sim->h_span[1]) ? kLtSmall : kLtLarge;
    }
    if (const char* e = getenv("MPM_TILE_LT")) sim->tile_lt = atoi(e) == kLtSmall ? kLtSmall : kLtLarge;  // experiments only
    const bool exact = sim->par.svd_mode == MPM_SVD_EXACT;
    if (sim->tile_lt == kLtSmall) {
      if (exact) launch_g2p_tile<MODEL, ExactOps, kLtSmall>(sim); else launch_g2p_tile<MODEL, FastOps, kLtSmall>(sim);
    } else {
      if (exact) launch_g2p_tile<MODEL, ExactOps, kLtLarge>(sim); else launch_g2p_tile<MODEL, FastOps, kLtLarge>(sim);
    }
    return 0;
  }
  const unsigned nb = blocks_for(n, kG2pBlock);
  Soa& p = sim->soa[sim->cur];
  if (sim->par.svd_mode == MPM_SVD_EXACT)
    g2p_kernel<MODEL, ExactOps><<<nb, kG2pBlock, 0, sim->stream>>>(p, n, sim->mats, sim->grid, sim->k);
  else
    g2p_kernel<MODEL, FastOps><<<nb, kG2pBlock, 0, sim->stream>>>(p, n, sim->mats, sim->grid, sim->k);
  return 0;
}
int do_g2p(MpmSim* sim) {
  StageTimer tm(sim, MPM_STAGE_G2P);
  if (sim->count == 0) return 0;
  if (sim->par.model == MPM_MODEL_SNOW) launch_g2p<MPM_MODEL_SNOW>(sim); else launch_g2p<MPM_MODEL_FIXED_COROTATED>(sim);
  sim->launches++;
  CK(cudaGetLastError());
  return 0;
}

template <int MODEL, class O, bool EXACT>
void launch_g2p2g(MpmSim* sim) {
  const size_t smem = FusedLayout<MODEL>::bytes();
  const bool one = sim->n_mats == 1;
  auto kern = one ? g2p2g_kernel<MODEL, O, EXACT, true> : g2p2g_kernel<MODEL, O, EXACT, false>;
  static int per_sm[2] = {0, 0};  // per instantiation
  if (!per_sm[one]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[one], kern, kFusedThreads, smem);
    per_sm[one] = std::max(per_sm[one], 1);
  }
  const size_t max_tiles = sim->count / kTileMax + sim->n_rows + 1;
  const unsigned ctas = (unsigned)std::min<size_t>(max_tiles, (size_t)sim->n_sms * per_sm[one]);
  kern<<<ctas, kFusedThreads, smem, sim->stream>>>(sim->soa[sim->cur], sim->mats, sim->mat0, sim->grid, sim->grid_b, sim->k, sim->tiles,
                                                 sim->d_n_tiles, sim->tm_streams[sim->cur][MODEL == MPM_MODEL_SNOW ? 1 : 0]);
}

// G2P of this substep from sim->grid + P2G of the next substep into sim->grid_b (zeroed by the caller)
int do_g2p2g(MpmSim* sim) {
  StageTimer tm(sim, MPM_STAGE_G2P2G);
  if (sim->count == 0) return 0;
  const bool exact = sim->par.svd_mode == MPM_SVD_EXACT;
  if (sim->par.model == MPM_MODEL_SNOW) {
    if (exact) launch_g2p2g<MPM_MODEL_SNOW, ExactOps, true>(sim); else launch_g2p2g<MPM_MODEL_SNOW, FastOps, false>(sim);
  } else {
    if (exact) launch_g2p2g<MPM_MODEL_FIXED_COROTATED, ExactOps, true>(sim); else launch_g2p2g<MPM_MODEL_FIXED_COROTATED, FastOps, false>(sim);
  }
  sim->launches++;
  CK(cudaGetLastError());
  return 0;
}

int do_exchange(MpmSim* sim) {
  if (!sim->comm.active()) return 0;
  StageTimer tm(sim, MPM_STAGE_EXCHANGE);
  if (sim->comm.exchange_halo(sim->grid, sim->k, sim->stream, &sim->launches)) return fail(sim, "halo exchange failed: %s", sim->comm.error());
  return 0;
}

int bits_for(size_t n) {
  int b = 1;
  while (((size_t)1 << b) < n) ++b;
  return b;
}

}  // namespace

extern "C" {

int mpm_abi_version(void) { return MPM_B200_ABI_VERSION; }

void mpm_make_material(double volume, double density, double E, double Nu, double hardening, double lo, double hi,
                       MpmMaterial* out) {
  // arguments arrive as doubles from the TOML reader and are narrowed to real at the call, then
  // the constructor arithmetic runs in f32 (int literals promote to float)
  const float vol = (float)volume, rho = (float)density, e = (float)E, nu = (float)Nu;
  out->particleVolume = vol;
  out->particleMass = rho * vol;
  out->mu0 = e / (2 * (1 + nu));
  out->lambda0 = e * nu / ((1 + nu) * (1 - 2 * nu));
  out->hardening = (float)hardening;
  out->plast_clamp_lower = (float)lo;
  out->plast_clamp_higher = (float)hi;
}

int mpm_create(const MpmParams* params, const MpmMaterial* materials, int n_materials, MpmSim** out) {
  MpmSim* sim = nullptr;
  if (!params || !out) return fail(nullptr, "mpm_create: null argument");
  if (params->N < 4) return fail(nullptr, "mpm_create: N must be >= 4");
  if (n_materials < 1 || n_materials > 256 || !materials) return fail(nullptr, "mpm_create: need 1..256 materials");
  if (params->model > MPM_MODEL_FIXED_COROTATED || params->svd_mode > MPM_SVD_FAST || params->p2g_mode > MPM_P2G_DIRECT ||
      params->g2p_mode > MPM_G2P_DIRECT || params->fuse_mode > MPM_FUSE_G2P2G || params->rebin_permille > 1000 || params->reserved_ != 0)
    return fail(nullptr, "mpm_create: bad model / svd_mode / p2g_mode / g2p_mode / fuse_mode");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, "mpm_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  sim = new (std::nothrow) MpmSim();
  if (!sim) return fail(nullptr, "mpm_create: out of host memory");
  sim->par = *params;
  if (params->device >= 0) {
    sim->device = params->device;
  } else {
    cudaGetDevice(&sim->device);
  }
#define CKC(call)                                                                          \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      fail(nullptr, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
      mpm_destroy(sim);                                                                    \
      return (int)e_;                                                                      \
    }                                                                                      \
  } while (0)
  CKC(cudaSetDevice(sim->device));
  cudaDeviceProp prop;
  CKC(cudaGetDeviceProperties(&prop, sim->device));
  sim->n_sms = prop.multiProcessorCount;
  if (prop.major < 10) {
    fail(nullptr, "mpm_create: device %d is sm_%d%d; this library is built for sm_100a only", sim->device, prop.major, prop.minor);
    mpm_destroy(sim);
    return 1;
  }
  const int N = (int)params->N;
  const int xb = (int)params->x_begin;
  const int xe = params->x_end ? (int)params->x_end : N;
  if (xb >= xe || xe > N) {
    fail(nullptr, "mpm_create: bad slab [%d,%d) for N=%d", xb, xe, N);
    mpm_destroy(sim);
    return 1;
  }
  KParams& k = sim->k;
  k.dt = params->dt;
  k.N = N;
  k.dx = (float)(1.0 / (double)N);
  k.dx_inv = (float)(1.0 / (double)k.dx);
  k.dinv = (4.0f * k.dx_inv) * k.dx_inv;
  k.x_own_begin = xb;
  k.x_own_end = xe;
  sim->whole_domain = (xb == 0 && xe == N);
  sim->ghost = sim->whole_domain ? 0 : (params->ghost ? (int)params->ghost : 1);
  k.x0 = std::max(0, xb - sim->ghost);
  k.nxl = std::min(N, xe + 2 + sim->ghost) - k.x0;  // owned + 2 stencil planes above + ghost planes either side
  sim->grid_nodes = (size_t)k.nxl * N * N;
  sim->n_rows = (uint32_t)k.nxl * (uint32_t)N;
  sim->key_bits = bits_for(sim->grid_nodes);
  CKC(cudaStreamCreateWithFlags(&sim->stream, cudaStreamNonBlocking));
  CKC(cudaEventCreate(&sim->ev[0]));
  CKC(cudaEventCreate(&sim->ev[1]));
  CKC(cudaMalloc(&sim->grid, sizeof(float4) * sim->grid_nodes));
  CKC(cudaMemsetAsync(sim->grid, 0, sizeof(float4) * sim->grid_nodes, sim->stream));
  // the fused kernel is built from the run-based P2G and the tile-based G2P
  sim->fused = params->fuse_mode == MPM_FUSE_G2P2G && params->p2g_mode == MPM_P2G_RUNS && params->g2p_mode == MPM_G2P_TILE &&
               N + kKeyBias <= 1023;
  if (sim->fused) CKC(cudaMalloc(&sim->grid_b, sizeof(float4) * sim->grid_nodes));
  sim->n_mats = n_materials;
  sim->mat0 = materials[0];
  CKC(cudaMalloc(&sim->mats, sizeof(MpmMaterial) * n_materials));
  CKC(cudaMemcpy(sim->mats, materials, sizeof(MpmMaterial) * n_materials, cudaMemcpyHostToDevice));
  CKC(cudaMalloc(&sim->d_counter, sizeof(unsigned long long)));
  CKC(cudaMalloc(&sim->row_first, sizeof(uint32_t) * ((size_t)sim->n_rows + 1)));
  CKC(cudaMalloc(&sim->tile_base, sizeof(uint32_t) * ((size_t)sim->n_rows + 1)));
  CKC(cudaMalloc(&sim->d_n_tiles, sizeof(uint32_t)));
  CKC(cudaMemsetAsync(sim->d_n_tiles, 0, sizeof(uint32_t), sim->stream));
  if (make_grid_maps(sim, sim->grid, sim->tm_grid) || (sim->fused && make_grid_maps(sim, sim->grid_b, sim->tm_grid_b))) {
    g_create_error = sim->err;
    mpm_destroy(sim);
    return 1;
  }
  CKC(cudaMalloc(&sim->d_span, 2 * sizeof(unsigned int)));
  CKC(cudaMallocHost(&sim->h_span, 2 * sizeof(unsigned int)));
  CKC(cudaEventCreateWithFlags(&sim->span_ev, cudaEventDisableTiming));
  CKC(cudaMalloc(&sim->d_tile_counters, 2 * sizeof(unsigned int)));
  CKC(cudaMemsetAsync(sim->d_tile_counters, 0, 2 * sizeof(unsigned int), sim->stream));
  CKC(cudaMalloc(&sim->d_moved, sizeof(unsigned long long)));
  CKC(cudaMemsetAsync(sim->d_moved, 0, sizeof(unsigned long long), sim->stream));
  CKC(cudaMallocHost(&sim->h_moved, sizeof(unsigned long long)));
  CKC(cudaEventCreateWithFlags(&sim->moved_ev, cudaEventDisableTiming));
  if (params->capacity) {
    if (int rc = ensure_capacity(sim, (size_t)params->capacity)) {
      g_create_error = sim->err;
      mpm_destroy(sim);
      return rc;
    }
  }
#undef CKC
  *out = sim;
  return 0;
}

void mpm_destroy(MpmSim* sim) {
  if (!sim) return;
  cudaSetDevice(sim->device);
  if (sim->stream) cudaStreamSynchronize(sim->stream);
  sim->comm.destroy();
  for (int b = 0; b < 2; ++b) {
    free_soa(sim->soa[b]);
    cudaFree(sim->keys[b]);
    cudaFree(sim->vals[b]);
  }
  cudaFree(sim->table);
  cudaFree(sim->scan_sums);
  cudaFree(sim->grid);
  cudaFree(sim->grid_b);
  cudaFree(sim->mats);
  cudaFree(sim->aos_stage);
  cudaFree(sim->d_counter);
  cudaFree(sim->d_span);
  cudaFree(sim->tiles);
  cudaFree(sim->row_first);
  cudaFree(sim->tile_base);
  cudaFree(sim->d_n_tiles);
  cudaFree(sim->d_moved);
  cudaFree(sim->d_tile_counters);
  if (sim->h_moved) cudaFreeHost(sim->h_moved);
  if (sim->moved_ev) cudaEventDestroy(sim->moved_ev);
  if (sim->h_span) cudaFreeHost(sim->h_span);
  if (sim->span_ev) cudaEventDestroy(sim->span_ev);
  if (sim->ev[0]) cudaEventDestroy(sim->ev[0]);
  if (sim->ev[1]) cudaEventDestroy(sim->ev[1]);
  if (sim->stream) cudaStreamDestroy(sim->stream);
  delete sim;
}

const char* mpm_last_error(const MpmSim* sim) { return sim ? sim->err.c_str() : g_create_error.c_str(); }

static int upload_impl(MpmSim* sim, const MpmParticle* particles, size_t count, const uint32_t* ids) {
  if (!sim || (!particles && count)) return fail(sim, "mpm_upload_particles_aos: null argument");
  CK(cudaSetDevice(sim->device));
  if (int rc = ensure_capacity(sim, std::max<size_t>(count, 1))) return rc;
  if (int rc = ensure_stage(sim, std::max<size_t>(count, 1))) return rc;
  sim->count = count;
  sim->first_id = 0;
  sim->cur = 0;
  sim->grid_ready = false;
  if (count) {
    CK(cudaMemcpyAsync(sim->aos_stage, particles, sizeof(MpmParticle) * count, cudaMemcpyHostToDevice, sim->stream));
    aos_to_soa_kernel<<<blocks_for(count, 256), 256, 0, sim->stream>>>(sim->aos_stage, sim->soa[0], count, 0);
    sim->launches++;
    CK(cudaGetLastError());
    if (ids) CK(cudaMemcpyAsync(sim->soa[0].id, ids, sizeof(uint32_t) * count, cudaMemcpyHostToDevice, sim->stream));
  }
  // bin immediately: the substep kernels assume cell-sorted order for locality
  return do_sort(sim);
}

int mpm_upload_particles_aos(MpmSim* sim, const MpmParticle* particles, size_t count) {
  return upload_impl(sim, particles, count, nullptr);
}
int mpm_upload_particles_with_ids(MpmSim* sim, const MpmParticle* particles, const uint32_t* ids, size_t count) {
  if (sim && sim->whole_domain && ids) return fail(sim, "mpm_upload_particles_with_ids: only for slab handles");
  return upload_impl(sim, particles, count, ids);
}

int mpm_append_particles_aos(MpmSim* sim, const MpmParticle* particles, size_t count) {
  if (!sim || (!particles && count)) return fail(sim, "mpm_append_particles_aos: null argument");
  CK(cudaSetDevice(sim->device));
  if (!sim->whole_domain) return fail(sim, "mpm_append_particles_aos: not for slab handles");
  if (count == 0) return 0;
  if (sim->count == 0) return upload_impl(sim, particles, count, nullptr);
  if (sim->count + count > sim->capacity)
    return fail(sim, "mpm_append_particles_aos: %zu + %zu particles exceed the capacity %zu (set MpmParams.capacity)", sim->count, count, sim->capacity);
  if (int rc = ensure_stage(sim, count)) return rc;
  CK(cudaMemcpyAsync(sim->aos_stage, particles, sizeof(MpmParticle) * count, cudaMemcpyHostToDevice, sim->stream));
  // ids continue the upload order, so a later download returns old particles first, then these
  aos_to_soa_kernel<<<blocks_for(count, 256), 256, 0, sim->stream>>>(sim->aos_stage, sim->soa[sim->cur], count, sim->first_id, sim->count);
  sim->launches++;
  CK(cudaGetLastError());
  sim->count += count;
  sim->grid_ready = false;
  return do_sort(sim);
}

int mpm_download_particles_aos(MpmSim* sim, MpmParticle* particles, size_t capacity, size_t* count) {
  if (!sim) return 1;
  CK(cudaSetDevice(sim->device));
  if (count) *count = sim->count;
  if (capacity < sim->count) return fail(sim, "mpm_download_particles_aos: capacity %zu < %zu", capacity, sim->count);
  if (sim->count == 0) return 0;
  if (int rc = ensure_stage(sim, sim->count)) return rc;
  soa_to_aos_kernel<<<blocks_for(sim->count, 256), 256, 0, sim->stream>>>(sim->soa[sim->cur], sim->count, sim->aos_stage, sim->first_id, sim->whole_domain);
  sim->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(particles, sim->aos_stage, sizeof(MpmParticle) * sim->count, cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return 0;
}

int mpm_download_positions_async(MpmSim* sim, float* xyz, size_t capacity, size_t* count) {
  if (!sim) return 1;
  CK(cudaSetDevice(sim->device));
  if (count) *count = sim->count;
  if (capacity < sim->count) return fail(sim, "mpm_download_positions: capacity too small");
  if (sim->count == 0) return 0;
  if (int rc = ensure_stage(sim, (sim->count * 12 + sizeof(MpmParticle) - 1) / sizeof(MpmParticle))) return rc;
  float* stage = reinterpret_cast<float*>(sim->aos_stage);
  positions_kernel<<<blocks_for(sim->count, 256), 256, 0, sim->stream>>>(sim->soa[sim->cur], sim->count, stage, sim->first_id, sim->whole_domain);
  sim->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(xyz, stage, sizeof(float) * 3 * sim->count, cudaMemcpyDeviceToHost, sim->stream));
  return 0;
}
int mpm_download_positions(MpmSim* sim, float* xyz, size_t capacity, size_t* count) {
  if (int rc = mpm_download_positions_async(sim, xyz, capacity, count)) return rc;
  CK(cudaStreamSynchronize(sim->stream));
  return 0;
}

int mpm_generate_dense_block(MpmSim* sim, uint64_t first_id, uint64_t count, uint32_t seed, float lo, float hi, uint8_t material) {
  if (!sim) return 1;
  CK(cudaSetDevice(sim->device));
  const bool whole = sim->whole_domain;
  if (whole) {
    if (int rc = ensure_capacity(sim, std::max<uint64_t>(count, 1))) return rc;
  } else if (sim->capacity == 0) {
    return fail(sim, "mpm_generate_dense_block: slab handles need MpmParams.capacity");
  }
  sim->cur = 0;
  sim->grid_ready = false;
  sim->first_id = (uint32_t)first_id;
  CK(cudaMemsetAsync(sim->d_counter, 0, sizeof(unsigned long long), sim->stream));
  if (count) {
    generate_block_kernel<<<blocks_for(count, 256), 256, 0, sim->stream>>>(sim->soa[0], first_id, count, lowbias32(seed), lo, hi,
                                                                            material, sim->k, whole, sim->d_counter, sim->capacity);
    sim->launches++;
    CK(cudaGetLastError());
  }
  if (whole) {
    sim->count = count;
  } else {
    unsigned long long n = 0;
    CK(cudaMemcpyAsync(&n, sim->d_counter, sizeof(n), cudaMemcpyDeviceToHost, sim->stream));
    CK(cudaStreamSynchronize(sim->stream));
    if (n > sim->capacity) return fail(sim, "mpm_generate_dense_block: %llu particles exceed capacity %zu", n, sim->capacity);
    sim->count = (size_t)n;
  }
  return do_sort(sim);
}

size_t mpm_particle_count(const MpmSim* sim) { return sim ? sim->count : 0; }
size_t mpm_grid_nodes(const MpmSim* sim) { return sim ? sim->grid_nodes : 0; }
double mpm_time(const MpmSim* sim) { return sim ? sim->t : 0.0; }
uint64_t mpm_substeps_done(const MpmSim* sim) { return sim ? sim->substeps : 0; }
uint64_t mpm_kernel_launches(const MpmSim* sim) { return sim ? sim->launches : 0; }
uint64_t mpm_rebins_done(const MpmSim* sim) { return sim ? sim->rebins : 0; }
void* mpm_stream(MpmSim* sim) { return sim ? (void*)sim->stream : nullptr; }

int mpm_stage_sort(MpmSim* sim) { if (!sim) return 1; CK(cudaSetDevice(sim->device)); return do_sort(sim); }
// the single stages work on the primary grid with the separate kernels and leave the fused
// pipeline's look-ahead grid invalid (mpm_advance rebuilds it)
int mpm_stage_reset_grid(MpmSim* sim) { if (!sim) return 1; CK(cudaSetDevice(sim->device)); sim->grid_ready = false; return do_reset(sim); }
int mpm_stage_p2g(MpmSim* sim) { if (!sim) return 1; CK(cudaSetDevice(sim->device)); sim->grid_ready = false; return do_p2g(sim); }
int mpm_stage_grid_update(MpmSim* sim) { if (!sim) return 1; CK(cudaSetDevice(sim->device)); sim->grid_ready = false; return do_grid(sim); }
int mpm_stage_g2p(MpmSim* sim) { if (!sim) return 1; CK(cudaSetDevice(sim->device)); sim->grid_ready = false; return do_g2p(sim); }

int mpm_advance(MpmSim* sim, int n_substeps) {
  if (!sim) return 1;
  CK(cudaSetDevice(sim->device));
  for (int s = 0; s < n_substeps; ++s) {
    bool rebin_late = false;
    bool due = sim->par.sort_every && sim->steps_since_sort >= sim->par.sort_every;
    // (not for slab handles: every rank must reach the migration of a re-bin in the same substep)
    const bool adaptive = sim->par.rebin_permille && !sim->fused && sim->par.g2p_mode == MPM_G2P_TILE && !sim->comm.active();
    if (adaptive) {
      // re-bin on measured disorder: the count of cell crossings since the last re-bin arrives a
      // substep or two late (asynchronous read-back), which is early enough for a locality heuristic
      // (the host may enqueue substeps far ahead of the device: a read-back older than 4 substeps is
      // waited for, which also bounds that run-ahead)
      if (sim->moved_pending && (sim->substeps - sim->moved_issued_at >= 4 || cudaEventQuery(sim->moved_ev) == cudaSuccess)) {
        CK(cudaEventSynchronize(sim->moved_ev));
        sim->moved_pending = false;
        sim->moved_seen = *sim->h_moved;
      }
      due = due || (sim->steps_since_sort > 0 && sim->moved_seen * 1000ull >= (unsigned long long)sim->par.rebin_permille * sim->count && sim->count > 0);
    }
    if (due) {
      // (slab handles migrate whole particle records at the re-bin: v and C are still the previous
      // substep's there, so leavers carry a complete state; the fused pipeline has no such gap)
      rebin_late = !sim->fused;
      if (!rebin_late) {
        if (int rc = do_sort(sim)) return rc;
      }
    }
    if (sim->fused) {
      // `grid` = velocities of this substep (built here on the first substep after the particles
      // changed, otherwise left by the previous fused kernel); the fused kernel gathers from it
      // and scatters the next substep into grid_b, which then becomes `grid`.
      if (!sim->grid_ready) {
        if (int rc = do_reset(sim)) return rc;
        if (int rc = do_p2g(sim)) return rc;
        if (int rc = do_exchange(sim)) return rc;
        if (int rc = do_grid(sim)) return rc;
      }
      if (int rc = do_reset(sim, sim->grid_b)) return rc;
      if (int rc = do_g2p2g(sim)) return rc;
      std::swap(sim->grid, sim->grid_b);
      for (int i = 0; i < 2; ++i) std::swap(sim->tm_grid[i], sim->tm_grid_b[i]);
      if (int rc = do_exchange(sim)) return rc;
      if (int rc = do_grid(sim)) return rc;
      sim->grid_ready = true;
    } else {
      bool keys_ready = false;  // slab handles tombstone leavers at the re-bin, which changes their keys
      if (int rc = do_reset(sim)) return rc;
      if (int rc = do_p2g(sim, rebin_late && !sim->comm.active(), &keys_ready)) return rc;
      if (int rc = do_exchange(sim)) return rc;
      if (int rc = do_grid(sim)) return rc;
      // a due re-bin runs here when it can: G2P is about to overwrite v and C, so only x, F, Jp
      // have to move (the keys come from the positions this substep started with)
      if (rebin_late) {
        if (int rc = do_sort(sim, true, keys_ready)) return rc;
      }
      if (int rc = do_g2p(sim)) return rc;
      if (adaptive && !sim->moved_pending) {
        CK(cudaMemcpyAsync(sim->h_moved, sim->d_moved, sizeof(unsigned long long), cudaMemcpyDeviceToHost, sim->stream));
        CK(cudaEventRecord(sim->moved_ev, sim->stream));
        sim->moved_pending = true;
        sim->moved_issued_at = sim->substeps;
      }
    }
    sim->t += (double)sim->par.dt;
    sim->substeps++;
    sim->steps_since_sort++;
  }
  return 0;
}

int mpm_sync(MpmSim* sim) {
  if (!sim) return 1;
  CK(cudaSetDevice(sim->device));
  CK(cudaStreamSynchronize(sim->stream));
  return 0;
}

int mpm_debug_download_grid(MpmSim* sim, float* vec4, size_t n_nodes) {
  if (!sim || !vec4) return 1;
  CK(cudaSetDevice(sim->device));
  if (n_nodes != sim->grid_nodes) return fail(sim, "grid has %zu nodes, caller passed %zu", sim->grid_nodes, n_nodes);
  CK(cudaMemcpyAsync(vec4, sim->grid, sizeof(float4) * n_nodes, cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return 0;
}
int mpm_debug_upload_grid(MpmSim* sim, const float* vec4, size_t n_nodes) {
  if (!sim || !vec4) return 1;
  CK(cudaSetDevice(sim->device));
  if (n_nodes != sim->grid_nodes) return fail(sim, "grid has %zu nodes, caller passed %zu", sim->grid_nodes, n_nodes);
  sim->grid_ready = false;
  CK(cudaMemcpyAsync(sim->grid, vec4, sizeof(float4) * n_nodes, cudaMemcpyHostToDevice, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return 0;
}
int mpm_debug_overwrite_particles_aos(MpmSim* sim, const MpmParticle* particles, size_t count) {
  if (!sim || !particles) return 1;
  CK(cudaSetDevice(sim->device));
  if (count != sim->count || !sim->whole_domain) return fail(sim, "overwrite needs the same particle count on a whole-domain handle");
  if (count == 0) return 0;
  sim->grid_ready = false;
  if (int rc = ensure_stage(sim, count)) return rc;
  CK(cudaMemcpyAsync(sim->aos_stage, particles, sizeof(MpmParticle) * count, cudaMemcpyHostToDevice, sim->stream));
  aos_overwrite_kernel<<<blocks_for(count, 256), 256, 0, sim->stream>>>(sim->aos_stage, sim->soa[sim->cur], count, sim->first_id);
  sim->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(sim->stream));
  return 0;
}
int mpm_debug_download_sort(MpmSim* sim, uint32_t* keys, uint32_t* ids, size_t capacity) {
  if (!sim) return 1;
  CK(cudaSetDevice(sim->device));
  if (capacity < sim->count) return fail(sim, "capacity too small");
  if (sim->count == 0) return 0;
  if (keys) CK(cudaMemcpyAsync(keys, sim->keys[sim->sorted_cur], sizeof(uint32_t) * sim->count, cudaMemcpyDeviceToHost, sim->stream));
  if (ids) CK(cudaMemcpyAsync(ids, sim->soa[sim->cur].id, sizeof(uint32_t) * sim->count, cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return 0;
}

int mpm_get_stage_times(MpmSim* sim, float ms[MPM_STAGE_COUNT]) {
  if (!sim) return 1;
  sim->timing = true;
  for (int i = 0; i < MPM_STAGE_COUNT; ++i) {
    if (ms) ms[i] = sim->stage_ms[i];
    sim->stage_ms[i] = 0.f;
  }
  return 0;
}

int mpm_comm_unique_id(void* id128) { return Comm::unique_id(id128); }
int mpm_attach_comm(MpmSim* sim, const void* id128, int rank, int nranks) {
  if (!sim) return 1;
  CK(cudaSetDevice(sim->device));
  if (sim->capacity == 0) return fail(sim, "mpm_attach_comm: set MpmParams.capacity for slab handles");
  if (sim->comm.init(id128, rank, nranks, sim->k, sim->ghost, sim->capacity, sim->stream)) return fail(sim, "mpm_attach_comm: %s", sim->comm.error());
  return 0;
}

// ---- linalg hooks ---------------------------------------------------------------------------------
static int linalg_run(const float* A, size_t n, int mode, int what, float* o0, float* o1, float* o2) {
  MpmSim* sim = nullptr;
  float *dA = nullptr, *d0 = nullptr, *d1 = nullptr, *d2 = nullptr;
  const size_t sz0 = (what == 2) ? n : 9 * n;
  CK(cudaMalloc(&dA, sizeof(float) * 9 * n));
  CK(cudaMalloc(&d0, sizeof(float) * sz0));
  if (what == 0) {
    CK(cudaMalloc(&d1, sizeof(float) * 3 * n));
    CK(cudaMalloc(&d2, sizeof(float) * 9 * n));
  }
  CK(cudaMemcpy(dA, A, sizeof(float) * 9 * n, cudaMemcpyHostToDevice));
  const unsigned nb = blocks_for(n, 128);
  if (what == 0) {
    if (mode == MPM_SVD_EXACT) svd3_batch_kernel<ExactOps><<<nb, 128>>>(dA, d0, d1, d2, n);
    else svd3_batch_kernel<FastOps><<<nb, 128>>>(dA, d0, d1, d2, n);
  } else if (what == 1) {
    if (mode == MPM_SVD_EXACT) polar_batch_kernel<ExactOps><<<nb, 128>>>(dA, d0, n);
    else polar_batch_kernel<FastOps><<<nb, 128>>>(dA, d0, n);
  } else {
    det_batch_kernel<<<nb, 128>>>(dA, d0, n);
  }
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(o0, d0, sizeof(float) * sz0, cudaMemcpyDeviceToHost));
  if (what == 0) {
    CK(cudaMemcpy(o1, d1, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(o2, d2, sizeof(float) * 9 * n, cudaMemcpyDeviceToHost));
  }
  cudaFree(dA);
  cudaFree(d0);
  cudaFree(d1);
  cudaFree(d2);
  return 0;
}
int mpm_svd3_batch(const float* A, float* U, float* S, float* V, size_t n, int svd_mode) {
  if (!n) return 0;
  return linalg_run(A, n, svd_mode, 0, U, S, V);
}
int mpm_polar_batch(const float* A, float* R, size_t n, int svd_mode) {
  if (!n) return 0;
  return linalg_run(A, n, svd_mode, 1, R, nullptr, nullptr);
}
int mpm_determinant_batch(const float* A, float* det, size_t n) {
  if (!n) return 0;
  return linalg_run(A, n, 0, 2, det, nullptr, nullptr);
}

}  // extern "C"
