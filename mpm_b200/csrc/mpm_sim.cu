// C ABI of the B200-native MLS-MPM substep: handle, buffers, stage launches.
// Replaces the device side of the reference's Simulation class (src/mpm.cu:180-329).
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/mpm_b200.h"
#include "comm.cuh"
#include "handle_kernels.cuh"
#include "sort.cuh"
#include "substep.cuh"

using namespace mpm;

static thread_local std::string g_create_error;

// ---- model registry (substep.cuh) -----------------------------------------------------------------
namespace mpm {
namespace {
constexpr int kMaxModels = 64;
const ModelOps* g_models[kMaxModels][2];
std::mutex g_models_mu;
}  // namespace
int register_model(uint32_t id, const ModelOps* exact, const ModelOps* fast) {
  if (id >= (uint32_t)kMaxModels || !exact || !fast) return 1;
  std::lock_guard<std::mutex> lock(g_models_mu);
  g_models[id][0] = exact;
  g_models[id][1] = fast;
  return 0;
}
const ModelOps* find_model(uint32_t id, uint32_t svd_mode) {
  if (id >= (uint32_t)kMaxModels || svd_mode > 1) return nullptr;
  std::lock_guard<std::mutex> lock(g_models_mu);
  return g_models[id][svd_mode];
}
// the shipped models, one translation unit each (models_*.cu)
const ModelOps* model_snow(int svd_mode);
const ModelOps* model_fixed_corotated(int svd_mode);
const ModelOps* model_jelly(int svd_mode);
namespace {
struct ShippedModels {
  ShippedModels() {
    register_model(MPM_MODEL_SNOW, model_snow(0), model_snow(1));
    register_model(MPM_MODEL_FIXED_COROTATED, model_fixed_corotated(0), model_fixed_corotated(1));
    register_model(MPM_MODEL_JELLY, model_jelly(0), model_jelly(1));
  }
} g_shipped_models;
}  // namespace
}  // namespace mpm

struct MpmSim {
  MpmParams par{};
  KParams k{};
  int device = 0;
  int n_sms = 148;
  cudaStream_t stream = nullptr;
  const ModelOps* ops = nullptr;

  // particles: two SoA buffers (the re-bin permutes from one into the other)
  Soa soa[2]{};
  int cur = 0;
  size_t capacity = 0;
  size_t count = 0;
  uint32_t first_id = 0;
  // what the C rows of the particles hold: the APIC matrix C (always, at the API boundary) or
  // dx * affine of the next P2G (between two substeps of one mpm_advance call, common.cuh)
  bool form_ad = false;
  bool handover = false;  // the pipeline hands over (MPM_PIPE_HANDOVER with the staged kernels)

  // grid: nxl * N * N float4
  float4* grid = nullptr;
  size_t grid_nodes = 0;

  void* mats_dev = nullptr;   // n_mats objects of the model's material type
  void* mats_host = nullptr;
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
  // the sorted cell keys of the CURRENT particle order (valid for the first okeys_count slots) and a 1/64
  // sample of them: the merge re-bin compares them with the new keys (sort.cuh)
  uint32_t* okeys[2] = {nullptr, nullptr};
  int ocur = 0;
  size_t okeys_count = 0;
  uint32_t* ocoarse = nullptr;
  // merge re-bin scratch: ballot words + their prefix counts, per-tile moved counts, the moved pairs
  uint32_t* mmask = nullptr;
  uint32_t* wprefix = nullptr;
  uint32_t* tile_moved = nullptr;
  uint32_t* mk[2] = {nullptr, nullptr};
  uint32_t* mi[2] = {nullptr, nullptr};
  size_t moved_cap = 0;
  uint32_t* d_n_moved = nullptr;
  uint64_t merge_rebins = 0;

  MpmParticle* aos_stage = nullptr;  // device AoS staging for upload/download
  size_t aos_stage_cap = 0;
  // overlapped transfers (mpm_prefetch_particles_aos / mpm_download_particles_aos_async): a second staging
  // buffer filled by its own copy stream while the substeps run, and the read-back of aos_stage on another
  MpmParticle* aos_prefetch = nullptr;
  size_t aos_prefetch_cap = 0;
  const MpmParticle* prefetched_ptr = nullptr;  // host buffer whose copy sits in (or is on its way to) aos_prefetch
  size_t prefetched_count = 0;
  cudaStream_t io_in = nullptr, io_out = nullptr;
  cudaEvent_t ev_prefetched = nullptr, ev_consumed = nullptr, ev_staged = nullptr, ev_downloaded = nullptr;
  bool prefetch_consumed_recorded = false;
  bool download_pending = false;
  unsigned long long* d_counter = nullptr;
  DeviceDiag* d_diag = nullptr;
  DeviceDiag* h_diag = nullptr;  // pinned

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

  // slab handles: the halo exchange runs on its own stream while the interior particles scatter
  // (see do_p2g_and_exchange).  split[0..1] = tiles whose particles can touch the planes shared with
  // the lower neighbour / first tile of those that can touch the planes shared with the upper one.
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_boundary = nullptr, ev_exchanged = nullptr;
  uint32_t* d_split = nullptr;
  uint32_t* h_split = nullptr;  // pinned
  cudaEvent_t split_ev = nullptr;
  bool split_pending = false, split_valid = false;

  // CUDA graphs of whole mpm_advance calls (MpmParams.graph_mode), keyed by everything the launch
  // sequence depends on; dropped whenever anything but mpm_advance touches the handle
  struct GraphEntry {
    int n_substeps, cur, ocur, tile_parity;
    uint64_t steps_since_sort;
    size_t count;
    cudaGraphExec_t exec;
    // host state after the call
    int cur_after, ocur_after, tile_parity_after;
    uint64_t steps_since_sort_after, rebins_delta, launches_delta;
  };
  std::vector<GraphEntry> graphs;
  uint64_t graph_replays = 0;
  bool capturing = false;  // inside a stream capture: no host-side read-backs

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

int alloc_soa(MpmSim* sim, Soa& s, size_t cap) {
  s.capacity = cap;  // multiple of kTile
  CK(cudaMalloc(&s.f, sizeof(float) * NSTREAM * cap));
  CK(cudaMalloc(&s.id, sizeof(uint32_t) * cap));
  CK(cudaMalloc(&s.mat, cap));
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
    for (int b = 0; b < 2; ++b) {
      cudaFree(sim->okeys[b]);
      cudaFree(sim->mk[b]);
      cudaFree(sim->mi[b]);
    }
    cudaFree(sim->ocoarse);
    cudaFree(sim->mmask);
    cudaFree(sim->wprefix);
    cudaFree(sim->tile_moved);
  }
  cap = (cap + kTile - 1) / kTile * kTile;
  for (int b = 0; b < 2; ++b) {
    if (int rc = alloc_soa(sim, sim->soa[b], cap)) return rc;
    CK(cudaMalloc(&sim->keys[b], sizeof(uint32_t) * cap));
    CK(cudaMalloc(&sim->vals[b], sizeof(uint32_t) * cap));
  }
  const size_t n_tiles = (cap + kSortTile - 1) / kSortTile;
  sim->table_len = n_tiles * kMaxRadix;
  CK(cudaMalloc(&sim->table, sizeof(uint32_t) * sim->table_len));
  sim->scan_sums_len = (sim->table_len + kScanTile - 1) / kScanTile;
  CK(cudaMalloc(&sim->scan_sums, sizeof(uint32_t) * sim->scan_sums_len));
  sim->moved_cap = cap / 8 + 4096;  // the merge re-bin is taken up to an eighth of the particles moved
  for (int b = 0; b < 2; ++b) {
    CK(cudaMalloc(&sim->okeys[b], sizeof(uint32_t) * cap));
    CK(cudaMalloc(&sim->mk[b], sizeof(uint32_t) * sim->moved_cap));
    CK(cudaMalloc(&sim->mi[b], sizeof(uint32_t) * sim->moved_cap));
  }
  CK(cudaMalloc(&sim->ocoarse, sizeof(uint32_t) * (cap / kCoarse + 2)));
  CK(cudaMalloc(&sim->mmask, sizeof(uint32_t) * (cap / 32 + 4)));
  CK(cudaMalloc(&sim->wprefix, sizeof(uint32_t) * (cap / 32 + 4)));
  CK(cudaMalloc(&sim->tile_moved, sizeof(uint32_t) * (cap / kMergeTile + 4)));
  sim->okeys_count = 0;
  sim->capacity = cap;
  return 0;
}

// the AoS staging buffer for n records; whatever the caller queues on sim->stream next runs after a pending
// asynchronous read-back of the buffer (mpm_download_particles_aos_async)
int ensure_stage(MpmSim* sim, size_t n) {
  if (sim->download_pending) CK(cudaStreamWaitEvent(sim->stream, sim->ev_downloaded, 0));
  if (n <= sim->aos_stage_cap) return 0;
  if (sim->download_pending) CK(cudaEventSynchronize(sim->ev_downloaded));
  cudaFree(sim->aos_stage);
  sim->aos_stage = nullptr;
  sim->aos_stage_cap = 0;
  CK(cudaMalloc(&sim->aos_stage, sizeof(MpmParticle) * n));
  sim->aos_stage_cap = n;
  return 0;
}

const char* const kStageNames[MPM_STAGE_COUNT] = {"mpm:sort", "mpm:reset", "mpm:p2g", "mpm:grid", "mpm:g2p", "mpm:exchange"};

// NVTX range per stage (visible in nsys / ncu timelines); with timing on also CUDA events + a host
// synchronisation, which serialises the stages — profiling only
bool nvtx_on() {
  static const bool on = [] {
    const char* e = getenv("MPM_B200_NVTX");
    return !(e && e[0] == '0');
  }();
  return on;
}
struct StageTimer {
  MpmSim* s;
  int stage;
  StageTimer(MpmSim* s_, int st) : s(s_), stage(st) {
    if (nvtx_on()) nvtxRangePushA(kStageNames[st]);
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
    if (nvtx_on()) nvtxRangePop();
  }
};

// P2G re-orders its payload records warp by warp once this share of the particles (1 / divisor) has changed
// cell since the last re-bin (G2P counts the crossings, the host reads the count back without waiting)
constexpr unsigned long long kStaleOrderDivisor = 50;

LaunchCtx make_ctx(MpmSim* sim) {
  LaunchCtx c{};
  c.stream = sim->stream;
  c.n_sms = sim->n_sms;
  c.soa = sim->soa[sim->cur];
  c.count = sim->count;
  c.mats_dev = sim->mats_dev;
  c.mats_host = sim->mats_host;
  c.n_mats = sim->n_mats;
  c.grid = sim->grid;
  c.k = sim->k;
  c.diag = sim->d_diag;
  c.p2g_mode = (int)sim->par.p2g_mode;
  c.g2p_mode = (int)sim->par.g2p_mode;
  c.handover_in = sim->form_ad;
  c.stale_order = sim->moved_seen * kStaleOrderDivisor >= (unsigned long long)sim->count && sim->count > 0;
  c.tile_begin = 0;
  c.tile_end = (sim->count + kTile - 1) / kTile;
  return c;
}

// ---- stages -------------------------------------------------------------------------------------
// one pass of the stable LSD radix sort: (kin, vin) -> (kout, vout) by the digit at `shift`
template <int BITS>
void radix_pass(MpmSim* sim, const uint32_t* kin, const uint32_t* vin, uint32_t* kout, uint32_t* vout, size_t n, int shift, int n_tiles) {
  const size_t table_len = (size_t)n_tiles << BITS;
  const unsigned scan_blocks = blocks_for(table_len, kScanTile);
  radix_hist_kernel<BITS><<<n_tiles, kSortThreads, 0, sim->stream>>>(kin, n, shift, sim->table, n_tiles);
  scan_tile_sums_kernel<<<scan_blocks, kScanThreads, 0, sim->stream>>>(sim->table, table_len, sim->scan_sums);
  scan_sums_kernel<<<1, 1024, 0, sim->stream>>>(sim->scan_sums, scan_blocks);
  scan_downsweep_kernel<<<scan_blocks, kScanThreads, 0, sim->stream>>>(sim->table, table_len, sim->scan_sums);
  radix_scatter_kernel<BITS><<<n_tiles, kSortThreads, 0, sim->stream>>>(kin, vin, kout, vout, n, shift, sim->table, n_tiles);
  sim->launches += 5;
}

// Stable LSD radix sort of n (key, value) pairs over `key_bits` bits between the buffer pairs k[2] / v[2],
// starting in k[0] / v[0]; returns the index of the pair that holds the result.  final_keys (optional): the
// last pass writes the sorted keys there instead.
int radix_sort(MpmSim* sim, uint32_t* const k[2], uint32_t* const v[2], size_t n, int key_bits, uint32_t* final_keys = nullptr) {
  const int n_tiles = (int)((n + kSortTile - 1) / kSortTile);
  const int bits = radix_bits(key_bits);
  int in = 0;
  for (int shift = 0; shift < key_bits; shift += bits) {
    uint32_t* kout = (final_keys && shift + bits >= key_bits) ? final_keys : k[in ^ 1];
    if (bits == 8) radix_pass<8>(sim, k[in], v[in], kout, v[in ^ 1], n, shift, n_tiles);
    else radix_pass<9>(sim, k[in], v[in], kout, v[in ^ 1], n, shift, n_tiles);
    in ^= 1;
  }
  return in;
}

// Sorts the pairs in keys[0] / vals[0]: the sorted keys go to okeys[ocur ^ 1] (which becomes current), the
// permutation to the returned buffer.
uint32_t* sort_pairs(MpmSim* sim, size_t n) {
  const int in = radix_sort(sim, sim->keys, sim->vals, n, sim->key_bits, sim->okeys[sim->ocur ^ 1]);
  sim->ocur ^= 1;
  sim->okeys_count = n;
  coarse_keys_kernel<<<blocks_for((n + kCoarse - 1) / kCoarse, 256), 256, 0, sim->stream>>>(sim->okeys[sim->ocur], (uint32_t)n, sim->ocoarse);
  sim->launches++;
  return sim->vals[in];
}

// The same result by merging (sort.cuh, "merge re-bin"): keys[0] holds the new keys of the particles in their
// current order, okeys[ocur] the keys they were sorted by.  Counts the particles whose key changed (one host
// synchronisation: the count sizes the launches that follow); with more than an eighth of them changed it
// returns nullptr and the caller radix-sorts.  Sorted keys to okeys[ocur ^ 1], permutation to vals[1].
uint32_t* merge_pairs(MpmSim* sim, size_t n, size_t n_old, int* rc) {
  *rc = 0;
  const uint32_t* newk = sim->keys[0];
  const uint32_t* oldk = sim->okeys[sim->ocur];
  uint32_t* keys_out = sim->okeys[sim->ocur ^ 1];
  uint32_t* perm = sim->vals[1];
  const unsigned n_tiles = blocks_for(n, kMergeTile);
  rebin_flags_kernel<<<n_tiles, kMergeThreads, 0, sim->stream>>>(newk, oldk, (uint32_t)n, (uint32_t)n_old, sim->mmask, sim->tile_moved);
  exclusive_scan_total_kernel<<<1, 1024, 0, sim->stream>>>(sim->tile_moved, n_tiles, sim->d_n_moved);
  sim->launches += 3;  // (with the read-back kernel below)
  uint32_t* h_n = reinterpret_cast<uint32_t*>(sim->h_moved);  // pinned scratch
  if (readback_words(h_n, sim->d_n_moved, 1, sim->stream) != cudaSuccess ||
      cudaStreamSynchronize(sim->stream) != cudaSuccess) {
    *rc = 1;
    return nullptr;
  }
  const size_t n_moved = *h_n;
  if (n_moved * 8 > n || n_moved > sim->moved_cap) return nullptr;
  rebin_compact_kernel<<<n_tiles, kMergeThreads, 0, sim->stream>>>(newk, (uint32_t)n, sim->mmask, sim->tile_moved, sim->wprefix, sim->mk[0], sim->mi[0],
                                                                  (uint32_t)n_moved, sim->d_n_moved);
  sim->launches++;
  int in = 0;
  if (n_moved) {
    in = radix_sort(sim, sim->mk, sim->mi, n_moved, sim->key_bits);
    rebin_place_moved_kernel<<<blocks_for(n_moved, 256), 256, 0, sim->stream>>>(sim->mk[in], sim->mi[in], (uint32_t)n_moved, oldk, sim->ocoarse,
                                                                             (uint32_t)n_old, sim->mmask, sim->wprefix, perm, keys_out);
    sim->launches++;
  }
  uint32_t* slice_begin = sim->tile_moved;  // (the compaction is done with it)
  rebin_tile_slices_kernel<<<blocks_for(n_tiles + 1, 128), 128, 0, sim->stream>>>(newk, (uint32_t)n, sim->mmask, sim->mk[in], sim->mi[in],
                                                                                 (uint32_t)n_moved, n_tiles, slice_begin);
  rebin_place_stayed_kernel<<<n_tiles, kMergeThreads, 0, sim->stream>>>(newk, (uint32_t)n, sim->mmask, sim->wprefix, sim->mk[in], sim->mi[in],
                                                                       slice_begin, perm, keys_out);
  sim->launches++;
  sim->ocur ^= 1;
  sim->okeys_count = n;
  coarse_keys_kernel<<<blocks_for((n + kCoarse - 1) / kCoarse, 256), 256, 0, sim->stream>>>(sim->okeys[sim->ocur], (uint32_t)n, sim->ocoarse);
  sim->launches += 2;
  sim->merge_rebins++;
  return perm;
}

// partial: called between the grid update and G2P, permutes only what G2P reads (sort.cuh)
int do_sort(MpmSim* sim, bool partial = false, bool keys_ready = false) {
  StageTimer tm(sim, MPM_STAGE_SORT);
  sim->steps_since_sort = 0;
  sim->rebins++;
  const bool merge_allowed = partial && !sim->capturing;
  const unsigned long long crossings_estimate = sim->moved_seen;  // a substep or two old: good enough to skip hopeless attempts
  sim->moved_seen = 0;
  sim->moved_pending = false;
  if (sim->d_moved) CK(cudaMemsetAsync(sim->d_moved, 0, sizeof(unsigned long long), sim->stream));
  size_t n_dead = 0, n_arrived = 0;
  if (sim->comm.active()) {  // leavers out (tombstoned), arrivals appended, before the re-bin
    // particles whose stencil left the planes held here since the last re-bin lost mass on the grid:
    // that is a configuration error (ghost width / re-bin cadence too small for the velocities), and
    // every rank must stop for it, not only the one that saw it
    unsigned int escaped = 0;
    const size_t count_before = sim->count;
    // (with the cell keys P2G wrote for this re-bin: leavers' keys become the tombstone key, arrivals get theirs)
    if (sim->comm.migrate(sim->soa[sim->cur], &sim->count, sim->capacity, sim->k, sim->stream, &sim->launches, &n_dead, &sim->d_diag->escaped,
                          &escaped, keys_ready ? sim->keys[0] : nullptr, keys_ready ? sim->vals[0] : nullptr, (uint32_t)sim->grid_nodes))
      return fail(sim, "particle migration failed: %s", sim->comm.error());
    CK(cudaMemsetAsync(&sim->d_diag->escaped, 0, sizeof(unsigned int), sim->stream));
    n_arrived = sim->count - count_before;
    if (escaped)
      return fail(sim, "%u particle-substeps (all ranks) scattered outside the %d ghost plane(s) of their slab since the last re-bin; "
                       "this rank holds [%d,%d): raise MpmParams.ghost or lower sort_every", escaped, sim->ghost, sim->k.x_own_begin, sim->k.x_own_end);
  }
  const size_t n = sim->count;
  if (n == 0) return 0;
  Soa& src = sim->soa[sim->cur];
  const uint32_t dead_key = (uint32_t)sim->grid_nodes;  // behind every live key
  if (!keys_ready) {  // else the P2G of this substep wrote them (p2g_sched.cuh)
    cell_key_kernel<<<blocks_for(n, 256), 256, 0, sim->stream>>>(src, n, sim->k, sim->keys[0], sim->vals[0], dead_key);
    sim->launches++;
  }
  // Merge when few particles changed cell (the last read-back of G2P's crossing count says whether the attempt
  // is worth its counting pass; the decision itself rests on the exact number of changed keys).  Otherwise —
  // upload, DIRECT kernels, fast flows, while a graph is being captured — the radix sort.
  const uint32_t* perm = nullptr;
  const size_t n_old = n - n_arrived;
  if (keys_ready && merge_allowed && sim->okeys_count >= n_old && n >= (size_t)kMergeTile && crossings_estimate * 6 <= n) {
    int rc = 0;
    perm = merge_pairs(sim, n, n_old, &rc);
    if (rc) return fail(sim, "merge re-bin: reading the moved count back failed");
  }
  if (!perm) perm = sort_pairs(sim, n);
  Soa& dst = sim->soa[sim->cur ^ 1];
  if (partial) permute_kernel<true><<<blocks_for(n, 256), 256, 0, sim->stream>>>(src, dst, perm, n, sim->k);
  else permute_kernel<false><<<blocks_for(n, 256), 256, 0, sim->stream>>>(src, dst, perm, n, sim->k);
  sim->launches++;
  sim->cur ^= 1;
  sim->count = n - n_dead;  // tombstones were sorted behind the live particles
  if (sim->comm.active() && sim->count) {
    // tiles whose particles can reach the planes shared with a neighbour (drift <= ghost cells included)
    const int g = sim->ghost;
    const uint32_t NN = (uint32_t)sim->k.N * (uint32_t)sim->k.N;
    const uint32_t key_lo = sim->comm.has_lo ? (uint32_t)(sim->k.x_own_begin + 2 + 2 * g - sim->k.x0) * NN : 0u;
    const uint32_t key_hi = sim->comm.has_hi ? (uint32_t)std::max(0, sim->k.x_own_end - 2 - 2 * g - sim->k.x0) * NN : 0xffffffffu;
    split_bounds_kernel<<<1, 32, 0, sim->stream>>>(sim->okeys[sim->ocur], (uint32_t)sim->count, key_lo, key_hi, sim->d_split);
    sim->launches++;
    CK(readback_words(sim->h_split, sim->d_split, 2, sim->stream));
    sim->launches++;
    CK(cudaEventRecord(sim->split_ev, sim->stream));
    sim->split_pending = true;
    sim->split_valid = false;
  }
  CK(cudaGetLastError());
  return 0;
}

int do_reset(MpmSim* sim) {
  StageTimer tm(sim, MPM_STAGE_RESET);
  CK(cudaMemsetAsync(sim->grid, 0, sizeof(float4) * sim->grid_nodes, sim->stream));
  return 0;
}

// keys: also write the cell keys of the re-bin that follows in this substep (staged kernel only);
// *keys_done says whether that happened
int do_p2g(MpmSim* sim, bool keys, bool* keys_done, size_t tile_begin, size_t tile_end) {
  if (keys_done) *keys_done = false;
  if (sim->count == 0) return 0;
  LaunchCtx c = make_ctx(sim);
  c.tile_begin = tile_begin;
  c.tile_end = std::min(tile_end, c.tile_end);
  if (c.tile_begin >= c.tile_end) return 0;
  const bool staged = sim->ops->staged && sim->par.p2g_mode == MPM_P2G_RUNS && sim->k.N <= kP2gMaxN;
  if (keys && staged) {
    c.sort_keys = sim->keys[0];
    c.sort_vals = sim->vals[0];
    if (keys_done) *keys_done = true;
  }
  sim->ops->p2g(c);
  sim->launches++;
  CK(cudaGetLastError());
  return 0;
}

int do_exchange_on(MpmSim* sim, cudaStream_t stream) {
  if (sim->comm.exchange_halo(sim->grid, sim->k, stream, &sim->launches)) return fail(sim, "halo exchange failed: %s", sim->comm.error());
  return 0;
}

// P2G and, on slab handles, the halo exchange (partial sums of the planes shared with the
// neighbours, comm.cuh).  The particles are sorted x-major, so those that can reach a shared plane
// are the first and the last tiles of the SoA: they scatter first, then the exchange runs on its
// own stream while the interior particles — which cannot touch the shared planes — scatter.
int do_p2g_and_exchange(MpmSim* sim, bool keys, bool* keys_done) {
  const size_t n_tiles = (sim->count + kTile - 1) / kTile;
  if (keys_done) *keys_done = false;
  if (!sim->comm.active()) {
    StageTimer tm(sim, MPM_STAGE_P2G);
    return do_p2g(sim, keys, keys_done, 0, n_tiles);
  }
  if (sim->split_pending && !sim->timing) {  // written at the last re-bin, long done by now
    CK(cudaEventSynchronize(sim->split_ev));
    sim->split_pending = false;
    sim->split_valid = true;
  }
  const bool overlap = sim->split_valid && !sim->timing;
  if (!overlap) {
    {
      StageTimer tm(sim, MPM_STAGE_P2G);
      if (int rc = do_p2g(sim, keys, keys_done, 0, n_tiles)) return rc;
    }
    StageTimer tm(sim, MPM_STAGE_EXCHANGE);
    return do_exchange_on(sim, sim->stream);
  }
  // tiles [0, a) and [b, n_tiles) hold the boundary particles
  const size_t a = std::min<size_t>(n_tiles, ((size_t)sim->h_split[0] + kTile - 1) / kTile);
  const size_t b = std::max<size_t>(a, std::min<size_t>(n_tiles, (size_t)sim->h_split[1] / kTile));
  bool kd = false, kd2 = false, kd3 = false;
  if (int rc = do_p2g(sim, keys, &kd, 0, a)) return rc;
  if (int rc = do_p2g(sim, keys, &kd2, b, n_tiles)) return rc;
  CK(cudaEventRecord(sim->ev_boundary, sim->stream));
  CK(cudaStreamWaitEvent(sim->comm_stream, sim->ev_boundary, 0));
  if (int rc = do_exchange_on(sim, sim->comm_stream)) return rc;
  CK(cudaEventRecord(sim->ev_exchanged, sim->comm_stream));
  if (int rc = do_p2g(sim, keys, &kd3, a, b)) return rc;
  CK(cudaStreamWaitEvent(sim->stream, sim->ev_exchanged, 0));
  if (keys_done) *keys_done = keys && (kd || a == 0) && (kd2 || b == n_tiles) && (kd3 || a == b);
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

int do_g2p(MpmSim* sim, bool emit) {
  StageTimer tm(sim, MPM_STAGE_G2P);
  sim->form_ad = emit;
  if (sim->count == 0) return 0;
  LaunchCtx c = make_ctx(sim);
  c.emit = emit;
  c.moved = sim->d_moved;
  c.tile_counters = sim->d_tile_counters;
  c.tile_parity = sim->tile_parity;
  sim->ops->g2p(c);
  if (sim->ops->staged && sim->par.g2p_mode == MPM_G2P_TILE) sim->tile_parity ^= 1;
  sim->launches++;
  CK(cudaGetLastError());
  return 0;
}

int bits_for(size_t n) {
  int b = 1;
  while (((size_t)1 << b) < n) ++b;
  return b;
}

}  // namespace

static void drop_graphs(MpmSim* sim) {
  for (auto& g : sim->graphs) cudaGraphExecDestroy(g.exec);
  sim->graphs.clear();
}

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

int mpm_create_raw(const MpmParams* params, const void* materials, size_t material_bytes, int n_materials, MpmSim** out) {
  MpmSim* sim = nullptr;
  if (!params || !out) return fail(nullptr, "mpm_create: null argument");
  if (params->N < 4) return fail(nullptr, "mpm_create: N must be >= 4");
  if (n_materials < 1 || n_materials > 256 || !materials) return fail(nullptr, "mpm_create: need 1..256 materials");
  if (params->svd_mode > MPM_SVD_FAST || params->p2g_mode > MPM_P2G_DIRECT || params->g2p_mode > MPM_G2P_DIRECT ||
      params->pipeline > MPM_PIPE_CLASSIC || params->rebin_permille > 1000 || params->graph_mode > MPM_GRAPH_ON)
    return fail(nullptr, "mpm_create: bad svd_mode / p2g_mode / g2p_mode / pipeline");
  const ModelOps* ops = find_model(params->model, params->svd_mode);
  if (!ops) return fail(nullptr, "mpm_create: no material model registered under id %u", params->model);
  if (material_bytes != ops->material_bytes)
    return fail(nullptr, "mpm_create: model %s takes materials of %zu bytes, caller passed %zu", ops->name, ops->material_bytes, material_bytes);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, "mpm_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  sim = new (std::nothrow) MpmSim();
  if (!sim) return fail(nullptr, "mpm_create: out of host memory");
  sim->par = *params;
  sim->ops = ops;
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
  if (sim->grid_nodes >= 0xffffffffull) {
    fail(nullptr, "mpm_create: %zu grid nodes do not fit 32-bit cell keys", sim->grid_nodes);
    mpm_destroy(sim);
    return 1;
  }
  sim->key_bits = bits_for(sim->grid_nodes + (sim->whole_domain ? 0 : 1));  // slab handles: + the tombstone key
  sim->handover = params->pipeline == MPM_PIPE_HANDOVER && ops->staged && params->p2g_mode == MPM_P2G_RUNS &&
                  params->g2p_mode == MPM_G2P_TILE && N <= kP2gMaxN;
  CKC(cudaStreamCreateWithFlags(&sim->stream, cudaStreamNonBlocking));
  CKC(cudaEventCreate(&sim->ev[0]));
  CKC(cudaEventCreate(&sim->ev[1]));
  CKC(cudaMalloc(&sim->grid, sizeof(float4) * sim->grid_nodes));
  CKC(cudaMemsetAsync(sim->grid, 0, sizeof(float4) * sim->grid_nodes, sim->stream));
  sim->n_mats = n_materials;
  sim->mats_host = malloc(material_bytes * (size_t)n_materials);
  if (!sim->mats_host) {
    fail(nullptr, "mpm_create: out of host memory");
    mpm_destroy(sim);
    return 1;
  }
  memcpy(sim->mats_host, materials, material_bytes * (size_t)n_materials);
  CKC(cudaMalloc(&sim->mats_dev, material_bytes * (size_t)n_materials));
  CKC(cudaMemcpy(sim->mats_dev, materials, material_bytes * (size_t)n_materials, cudaMemcpyHostToDevice));
  CKC(cudaMalloc(&sim->d_counter, sizeof(unsigned long long)));
  CKC(cudaMalloc(&sim->d_n_moved, sizeof(uint32_t)));
  CKC(cudaMalloc(&sim->d_diag, sizeof(DeviceDiag)));
  CKC(cudaMemsetAsync(sim->d_diag, 0, sizeof(DeviceDiag), sim->stream));
  CKC(cudaMallocHost(&sim->h_diag, sizeof(DeviceDiag)));
  memset(sim->h_diag, 0, sizeof(DeviceDiag));
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

int mpm_create(const MpmParams* params, const MpmMaterial* materials, int n_materials, MpmSim** out) {
  if (!params || !out) return fail(nullptr, "mpm_create: null argument");
  if (n_materials < 1 || n_materials > 256 || !materials) return fail(nullptr, "mpm_create: need 1..256 materials");
  const ModelOps* ops = find_model(params->model, std::min<uint32_t>(params->svd_mode, 1u));
  if (!ops) return fail(nullptr, "mpm_create: no material model registered under id %u", params->model);
  std::string buf(ops->material_bytes * (size_t)n_materials, '\0');
  ops->from_abi(materials, n_materials, &buf[0]);
  return mpm_create_raw(params, buf.data(), ops->material_bytes, n_materials, out);
}

void mpm_destroy(MpmSim* sim) {
  if (!sim) return;
  cudaSetDevice(sim->device);
  if (sim->stream) cudaStreamSynchronize(sim->stream);
  if (sim->comm_stream) cudaStreamSynchronize(sim->comm_stream);
  drop_graphs(sim);
  sim->comm.destroy();
  for (int b = 0; b < 2; ++b) {
    free_soa(sim->soa[b]);
    cudaFree(sim->keys[b]);
    cudaFree(sim->vals[b]);
  }
  cudaFree(sim->table);
  cudaFree(sim->scan_sums);
  for (int b = 0; b < 2; ++b) {
    cudaFree(sim->okeys[b]);
    cudaFree(sim->mk[b]);
    cudaFree(sim->mi[b]);
  }
  cudaFree(sim->ocoarse);
  cudaFree(sim->mmask);
  cudaFree(sim->wprefix);
  cudaFree(sim->tile_moved);
  cudaFree(sim->d_n_moved);
  cudaFree(sim->grid);
  cudaFree(sim->mats_dev);
  free(sim->mats_host);
  if (sim->io_in) cudaStreamSynchronize(sim->io_in);
  if (sim->io_out) cudaStreamSynchronize(sim->io_out);
  cudaFree(sim->aos_stage);
  cudaFree(sim->aos_prefetch);
  for (cudaEvent_t e : {sim->ev_prefetched, sim->ev_consumed, sim->ev_staged, sim->ev_downloaded})
    if (e) cudaEventDestroy(e);
  if (sim->io_in) cudaStreamDestroy(sim->io_in);
  if (sim->io_out) cudaStreamDestroy(sim->io_out);
  cudaFree(sim->d_counter);
  cudaFree(sim->d_diag);
  cudaFree(sim->d_moved);
  cudaFree(sim->d_tile_counters);
  cudaFree(sim->d_split);
  if (sim->h_split) cudaFreeHost(sim->h_split);
  if (sim->h_diag) cudaFreeHost(sim->h_diag);
  if (sim->h_moved) cudaFreeHost(sim->h_moved);
  if (sim->moved_ev) cudaEventDestroy(sim->moved_ev);
  if (sim->split_ev) cudaEventDestroy(sim->split_ev);
  if (sim->ev_boundary) cudaEventDestroy(sim->ev_boundary);
  if (sim->ev_exchanged) cudaEventDestroy(sim->ev_exchanged);
  if (sim->ev[0]) cudaEventDestroy(sim->ev[0]);
  if (sim->ev[1]) cudaEventDestroy(sim->ev[1]);
  if (sim->comm_stream) cudaStreamDestroy(sim->comm_stream);
  if (sim->stream) cudaStreamDestroy(sim->stream);
  delete sim;
}

const char* mpm_last_error(const MpmSim* sim) { return sim ? sim->err.c_str() : g_create_error.c_str(); }

static int upload_impl(MpmSim* sim, const MpmParticle* particles, size_t count, const uint32_t* ids) {
  if (!sim || (!particles && count)) return fail(sim, "mpm_upload_particles_aos: null argument");
  drop_graphs(sim);
  CK(cudaSetDevice(sim->device));
  if (int rc = ensure_capacity(sim, std::max<size_t>(count, 1))) return rc;
  // the records are already on the device (or on their way) if this buffer was announced with mpm_prefetch_particles_aos
  const bool prefetched = count && particles == sim->prefetched_ptr && count == sim->prefetched_count && !ids;
  sim->prefetched_ptr = nullptr;
  if (!prefetched)  // (the prefetched path does not touch the staging buffer, which a read-back may still be using)
    if (int rc = ensure_stage(sim, std::max<size_t>(count, 1))) return rc;
  sim->count = count;
  sim->first_id = 0;
  sim->cur = 0;
  sim->form_ad = false;
  CK(cudaMemsetAsync(&sim->d_diag->jp_not_one, 0, sizeof(unsigned int), sim->stream));
  const MpmParticle* records = sim->aos_stage;
  if (prefetched) {
    CK(cudaStreamWaitEvent(sim->stream, sim->ev_prefetched, 0));
    records = sim->aos_prefetch;
  } else if (count) {
    CK(cudaMemcpyAsync(sim->aos_stage, particles, sizeof(MpmParticle) * count, cudaMemcpyHostToDevice, sim->stream));
  }
  if (count && !sim->comm.active() && !ids) {
    // the common path: keys from the records, sort the (key, index) pairs, then ONE pass that gathers the
    // records in cell order and writes the SoA — the particles never travel through the SoA unsorted
    StageTimer tm(sim, MPM_STAGE_SORT);
    sim->steps_since_sort = 0;
    sim->rebins++;
    sim->moved_seen = 0;
    sim->moved_pending = false;
    CK(cudaMemsetAsync(sim->d_moved, 0, sizeof(unsigned long long), sim->stream));
    aos_keys_kernel<<<blocks_for(count, 256), 256, 0, sim->stream>>>(records, count, sim->k, sim->keys[0], sim->vals[0], sim->d_diag);
    const uint32_t* perm = sort_pairs(sim, count);
    aos_gather_to_soa_kernel<<<blocks_for(count, kTile), kTile, 0, sim->stream>>>(records, perm, sim->soa[0], count, 0);
    sim->launches += 2;
    CK(cudaGetLastError());
    if (prefetched) {  // the next prefetch may overwrite the buffer from here on
      CK(cudaEventRecord(sim->ev_consumed, sim->stream));
      sim->prefetch_consumed_recorded = true;
    }
    return 0;
  }
  if (count) {
    aos_to_soa_kernel<<<blocks_for(count, 256), 256, 0, sim->stream>>>(records, sim->soa[0], count, 0, 0, sim->d_diag);
    sim->launches++;
    CK(cudaGetLastError());
    if (prefetched) {
      CK(cudaEventRecord(sim->ev_consumed, sim->stream));
      sim->prefetch_consumed_recorded = true;
    }
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
  drop_graphs(sim);
  CK(cudaSetDevice(sim->device));
  if (!sim->whole_domain) return fail(sim, "mpm_append_particles_aos: not for slab handles");
  if (count == 0) return 0;
  if (sim->count == 0) return upload_impl(sim, particles, count, nullptr);
  if (sim->count + count > sim->capacity)
    return fail(sim, "mpm_append_particles_aos: %zu + %zu particles exceed the capacity %zu (set MpmParams.capacity)", sim->count, count, sim->capacity);
  if (int rc = ensure_stage(sim, count)) return rc;
  CK(cudaMemcpyAsync(sim->aos_stage, particles, sizeof(MpmParticle) * count, cudaMemcpyHostToDevice, sim->stream));
  // ids continue the upload order, so a later download returns old particles first, then these
  aos_to_soa_kernel<<<blocks_for(count, 256), 256, 0, sim->stream>>>(sim->aos_stage, sim->soa[sim->cur], count, sim->first_id, sim->count, sim->d_diag);
  sim->launches++;
  CK(cudaGetLastError());
  sim->count += count;
  return do_sort(sim);
}

int mpm_remove_particles(MpmSim* sim, size_t first, size_t count) {
  if (!sim) return 1;
  CK(cudaSetDevice(sim->device));
  if (!sim->whole_domain) return fail(sim, "mpm_remove_particles: not for slab handles");
  if (first + count > sim->count) return fail(sim, "mpm_remove_particles: [%zu, %zu) is not a range of the %zu particles", first, first + count, sim->count);
  if (count == 0) return 0;
  drop_graphs(sim);
  const size_t n = sim->count;
  if (count == n) {
    sim->count = 0;
    sim->okeys_count = 0;
    return 0;
  }
  const uint32_t id0 = sim->first_id + (uint32_t)first, id1 = id0 + (uint32_t)count;
  const unsigned n_tiles = blocks_for(n, kMergeTile);
  Soa& src = sim->soa[sim->cur];
  Soa& dst = sim->soa[sim->cur ^ 1];
  uint32_t* perm = sim->vals[1];
  remove_flags_kernel<<<n_tiles, kMergeThreads, 0, sim->stream>>>(src.id, (uint32_t)n, id0, id1, sim->mmask, sim->tile_moved);
  exclusive_scan_total_kernel<<<1, 1024, 0, sim->stream>>>(sim->tile_moved, n_tiles, sim->d_n_moved);
  remove_perm_kernel<<<n_tiles, kMergeThreads, 0, sim->stream>>>((uint32_t)n, sim->mmask, sim->tile_moved, perm);
  const size_t left = n - count;
  permute_kernel<false><<<blocks_for(left, 256), 256, 0, sim->stream>>>(src, dst, perm, left, sim->k);
  renumber_ids_kernel<<<blocks_for(left, 256), 256, 0, sim->stream>>>(dst.id, (uint32_t)left, id1, (uint32_t)count);
  sim->launches += 5;
  CK(cudaGetLastError());
  sim->cur ^= 1;
  sim->count = left;
  sim->okeys_count = 0;  // the keys of the last re-bin no longer line up with the slots: the next re-bin sorts in full
  return 0;
}

int mpm_download_particles_aos(MpmSim* sim, MpmParticle* particles, size_t capacity, size_t* count) {
  if (!sim) return 1;
  CK(cudaSetDevice(sim->device));
  if (count) *count = sim->count;
  if (capacity < sim->count) return fail(sim, "mpm_download_particles_aos: capacity %zu < %zu", capacity, sim->count);
  if (sim->count == 0) return 0;
  if (int rc = ensure_stage(sim, sim->count)) return rc;
  soa_to_aos_kernel<<<blocks_for(sim->count, kTile), kTile, 0, sim->stream>>>(sim->soa[sim->cur], sim->count, sim->aos_stage, sim->first_id, sim->whole_domain);
  sim->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(particles, sim->aos_stage, sizeof(MpmParticle) * sim->count, cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return 0;
}

static int io_objects(MpmSim* sim) {
  if (sim->io_in) return 0;
  CK(cudaStreamCreateWithFlags(&sim->io_in, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&sim->io_out, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&sim->ev_prefetched, &sim->ev_consumed, &sim->ev_staged, &sim->ev_downloaded})
    CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  return 0;
}

int mpm_prefetch_particles_aos(MpmSim* sim, const MpmParticle* particles, size_t count) {
  if (!sim || !particles || count == 0) return fail(sim, "mpm_prefetch_particles_aos: null argument");
  CK(cudaSetDevice(sim->device));
  if (int rc = io_objects(sim)) return rc;
  if (count > sim->aos_prefetch_cap) {
    CK(cudaStreamSynchronize(sim->io_in));
    if (sim->prefetch_consumed_recorded) CK(cudaEventSynchronize(sim->ev_consumed));
    cudaFree(sim->aos_prefetch);
    sim->aos_prefetch = nullptr;
    sim->aos_prefetch_cap = 0;
    CK(cudaMalloc(&sim->aos_prefetch, sizeof(MpmParticle) * count));
    sim->aos_prefetch_cap = count;
  }
  // an upload that still reads the previous contents goes first
  if (sim->prefetch_consumed_recorded) CK(cudaStreamWaitEvent(sim->io_in, sim->ev_consumed, 0));
  CK(cudaMemcpyAsync(sim->aos_prefetch, particles, sizeof(MpmParticle) * count, cudaMemcpyHostToDevice, sim->io_in));
  CK(cudaEventRecord(sim->ev_prefetched, sim->io_in));
  sim->prefetched_ptr = particles;
  sim->prefetched_count = count;
  return 0;
}

int mpm_download_particles_aos_async(MpmSim* sim, MpmParticle* particles, size_t capacity, size_t* count) {
  if (!sim) return 1;
  CK(cudaSetDevice(sim->device));
  if (count) *count = sim->count;
  if (capacity < sim->count) return fail(sim, "mpm_download_particles_aos_async: capacity %zu < %zu", capacity, sim->count);
  if (sim->count == 0) return 0;
  if (int rc = io_objects(sim)) return rc;
  if (int rc = ensure_stage(sim, sim->count)) return rc;  // (also orders this conversion behind a read-back still in flight)
  soa_to_aos_kernel<<<blocks_for(sim->count, kTile), kTile, 0, sim->stream>>>(sim->soa[sim->cur], sim->count, sim->aos_stage, sim->first_id, sim->whole_domain);
  sim->launches++;
  CK(cudaGetLastError());
  CK(cudaEventRecord(sim->ev_staged, sim->stream));
  CK(cudaStreamWaitEvent(sim->io_out, sim->ev_staged, 0));
  CK(cudaMemcpyAsync(particles, sim->aos_stage, sizeof(MpmParticle) * sim->count, cudaMemcpyDeviceToHost, sim->io_out));
  CK(cudaEventRecord(sim->ev_downloaded, sim->io_out));
  sim->download_pending = true;
  return 0;
}

int mpm_download_wait(MpmSim* sim) {
  if (!sim) return 1;
  if (!sim->download_pending) return 0;
  CK(cudaSetDevice(sim->device));
  CK(cudaEventSynchronize(sim->ev_downloaded));
  sim->download_pending = false;
  return 0;
}

int mpm_download_positions_async(MpmSim* sim, float* xyz, size_t capacity, size_t* count) {
  if (!sim) return 1;
  CK(cudaSetDevice(sim->device));
  if (count) *count = sim->count;
  if (capacity < sim->count) return fail(sim, "mpm_download_positions: capacity too small");
  if (sim->count == 0) return 0;
  if (int rc = io_objects(sim)) return rc;
  if (int rc = ensure_stage(sim, (sim->count * 12 + sizeof(MpmParticle) - 1) / sizeof(MpmParticle))) return rc;
  float* stage = reinterpret_cast<float*>(sim->aos_stage);
  positions_kernel<<<blocks_for(sim->count, 256), 256, 0, sim->stream>>>(sim->soa[sim->cur], sim->count, stage, sim->first_id, sim->whole_domain);
  sim->launches++;
  CK(cudaGetLastError());
  // the copy runs on the read-back stream: substeps issued after this call do not wait for it
  CK(cudaEventRecord(sim->ev_staged, sim->stream));
  CK(cudaStreamWaitEvent(sim->io_out, sim->ev_staged, 0));
  CK(cudaMemcpyAsync(xyz, stage, sizeof(float) * 3 * sim->count, cudaMemcpyDeviceToHost, sim->io_out));
  CK(cudaEventRecord(sim->ev_downloaded, sim->io_out));
  sim->download_pending = true;
  return 0;
}
int mpm_download_positions(MpmSim* sim, float* xyz, size_t capacity, size_t* count) {
  if (int rc = mpm_download_positions_async(sim, xyz, capacity, count)) return rc;
  return mpm_download_wait(sim);
}

int mpm_generate_dense_block_stressed(MpmSim* sim, uint64_t first_id, uint64_t count, uint32_t seed, float lo, float hi, uint8_t material,
                                      float shear, float f_noise) {
  if (!sim) return 1;
  drop_graphs(sim);
  CK(cudaSetDevice(sim->device));
  const bool whole = sim->whole_domain;
  if (whole) {
    if (int rc = ensure_capacity(sim, std::max<uint64_t>(count, 1))) return rc;
  } else if (sim->capacity == 0) {
    return fail(sim, "mpm_generate_dense_block: slab handles need MpmParams.capacity");
  }
  sim->cur = 0;
  sim->form_ad = false;
  sim->first_id = (uint32_t)first_id;
  CK(cudaMemsetAsync(sim->d_counter, 0, sizeof(unsigned long long), sim->stream));
  CK(cudaMemsetAsync(&sim->d_diag->jp_not_one, 0, sizeof(unsigned int), sim->stream));
  if (count) {
    generate_block_kernel<<<blocks_for(count, 256), 256, 0, sim->stream>>>(sim->soa[0], first_id, count, lowbias32(seed), lo, hi, material,
                                                                            sim->k, whole, sim->d_counter, sim->capacity, BlockStress{shear, f_noise});
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
int mpm_generate_dense_block(MpmSim* sim, uint64_t first_id, uint64_t count, uint32_t seed, float lo, float hi, uint8_t material) {
  return mpm_generate_dense_block_stressed(sim, first_id, count, seed, lo, hi, material, 0.0f, 0.0f);
}

size_t mpm_particle_count(const MpmSim* sim) { return sim ? sim->count : 0; }
size_t mpm_grid_nodes(const MpmSim* sim) { return sim ? sim->grid_nodes : 0; }
double mpm_time(const MpmSim* sim) { return sim ? sim->t : 0.0; }
uint64_t mpm_substeps_done(const MpmSim* sim) { return sim ? sim->substeps : 0; }
uint64_t mpm_kernel_launches(const MpmSim* sim) { return sim ? sim->launches : 0; }
uint64_t mpm_rebins_done(const MpmSim* sim) { return sim ? sim->rebins : 0; }
uint64_t mpm_graph_replays(const MpmSim* sim) { return sim ? sim->graph_replays : 0; }
uint64_t mpm_merge_rebins(const MpmSim* sim) { return sim ? sim->merge_rebins : 0; }
void* mpm_stream(MpmSim* sim) { return sim ? (void*)sim->stream : nullptr; }

// single stages, for parity tests and profiling: the same kernels mpm_advance runs, on particles in
// the reference's form (C, not the handed-over affine matrix)
int mpm_stage_sort(MpmSim* sim) { if (!sim) return 1; drop_graphs(sim); CK(cudaSetDevice(sim->device)); return do_sort(sim); }
int mpm_stage_reset_grid(MpmSim* sim) { if (!sim) return 1; CK(cudaSetDevice(sim->device)); return do_reset(sim); }
int mpm_stage_p2g(MpmSim* sim) { if (!sim) return 1; CK(cudaSetDevice(sim->device)); return do_p2g_and_exchange(sim, false, nullptr); }
int mpm_stage_grid_update(MpmSim* sim) { if (!sim) return 1; CK(cudaSetDevice(sim->device)); return do_grid(sim); }
int mpm_stage_g2p(MpmSim* sim) { if (!sim) return 1; drop_graphs(sim); CK(cudaSetDevice(sim->device)); return do_g2p(sim, false); }

static int advance_impl(MpmSim* sim, int n_substeps) {
  for (int s = 0; s < n_substeps; ++s) {
    bool due = sim->par.sort_every && sim->steps_since_sort >= sim->par.sort_every;
    // Cell crossings since the last re-bin, counted by the staged G2P and read back asynchronously (it
    // arrives a substep or two late, early enough for heuristics; the host may enqueue substeps far
    // ahead of the device: a read-back older than 4 substeps is waited for, which also bounds that
    // run-ahead).  Drives the stale-order variant of P2G and, with rebin_permille, the re-bin itself
    // (not on slab handles: every rank must reach the migration of a re-bin in the same substep).
    const bool track = sim->ops->staged && sim->par.g2p_mode == MPM_G2P_TILE && !sim->capturing;
    const bool adaptive = track && sim->par.rebin_permille && !sim->comm.active();
    if (track && sim->moved_pending && (sim->substeps - sim->moved_issued_at >= 4 || cudaEventQuery(sim->moved_ev) == cudaSuccess)) {
      CK(cudaEventSynchronize(sim->moved_ev));
      sim->moved_pending = false;
      sim->moved_seen = *sim->h_moved;
    }
    if (adaptive)
      due = due || (sim->steps_since_sort > 0 && sim->moved_seen * 1000ull >= (unsigned long long)sim->par.rebin_permille * sim->count && sim->count > 0);
    if (due) {  // the keys P2G writes come with fresh out-of-domain / non-finite counts
      CK(cudaMemsetAsync(&sim->d_diag->nonfinite, 0, 2 * sizeof(unsigned int), sim->stream));
    }
    bool keys_ready = false;
    if (int rc = do_reset(sim)) return rc;
    if (int rc = do_p2g_and_exchange(sim, due, &keys_ready)) return rc;
    if (int rc = do_grid(sim)) return rc;
    // a due re-bin runs here: G2P is about to overwrite v and C, so only x, F, Jp have to move (the
    // keys come from the positions this substep started with).  Slab handles migrate whole particle
    // records at this point; the new owner's G2P needs x, F, Jp only.
    if (due) {
      if (int rc = do_sort(sim, true, keys_ready)) return rc;
    }
    // hand-over: every G2P but the last of this call leaves the next P2G's affine matrix in the C rows
    if (int rc = do_g2p(sim, sim->handover && s + 1 < n_substeps)) return rc;
    if (track && !sim->moved_pending) {
      CK(readback_words(sim->h_moved, sim->d_moved, 2, sim->stream));
      sim->launches++;
      CK(cudaEventRecord(sim->moved_ev, sim->stream));
      sim->moved_pending = true;
      sim->moved_issued_at = sim->substeps;
    }
    sim->t += (double)sim->par.dt;
    sim->substeps++;
    sim->steps_since_sort++;
  }
  return 0;
}

int mpm_advance(MpmSim* sim, int n_substeps) {
  if (!sim) return 1;
  CK(cudaSetDevice(sim->device));
  if (n_substeps <= 0) return 0;
  const bool want = sim->par.graph_mode == MPM_GRAPH_ON || (sim->par.graph_mode == MPM_GRAPH_AUTO && sim->count <= (size_t)4 << 20);
  const bool graphable = want && sim->count > 0 && n_substeps <= 256 && !sim->par.rebin_permille && !sim->comm.active() && !sim->timing;
  if (!graphable) return advance_impl(sim, n_substeps);
  for (auto& g : sim->graphs) {
    if (g.n_substeps == n_substeps && g.cur == sim->cur && g.ocur == sim->ocur && g.tile_parity == sim->tile_parity &&
        g.steps_since_sort == sim->steps_since_sort && g.count == sim->count) {
      CK(cudaGraphLaunch(g.exec, sim->stream));
      sim->cur = g.cur_after;
      sim->ocur = g.ocur_after;
      sim->tile_parity = g.tile_parity_after;
      sim->steps_since_sort = g.steps_since_sort_after;
      sim->rebins += g.rebins_delta;
      sim->launches += g.launches_delta;
      sim->substeps += (uint64_t)n_substeps;
      for (int s = 0; s < n_substeps; ++s) sim->t += (double)sim->par.dt;  // the same sum as the loop
      sim->graph_replays++;
      return 0;
    }
  }
  // first call in this state: capture the launches of the ordinary code path, then run the graph
  MpmSim::GraphEntry e{};
  e.n_substeps = n_substeps;
  e.cur = sim->cur;
  e.ocur = sim->ocur;
  e.tile_parity = sim->tile_parity;
  e.steps_since_sort = sim->steps_since_sort;
  e.count = sim->count;
  const uint64_t rebins0 = sim->rebins, launches0 = sim->launches, substeps0 = sim->substeps;
  const double t0 = sim->t;
  CK(cudaStreamBeginCapture(sim->stream, cudaStreamCaptureModeThreadLocal));
  sim->capturing = true;
  const int rc = advance_impl(sim, n_substeps);
  sim->capturing = false;
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(sim->stream, &graph);
  cudaGraphExec_t exec = nullptr;
  if (rc == 0 && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
    cudaGraphDestroy(graph);
    e.exec = exec;
    e.cur_after = sim->cur;
    e.ocur_after = sim->ocur;
    e.tile_parity_after = sim->tile_parity;
    e.steps_since_sort_after = sim->steps_since_sort;
    e.rebins_delta = sim->rebins - rebins0;
    e.launches_delta = sim->launches - launches0;
    if (sim->graphs.size() >= 32) {
      cudaGraphExecDestroy(sim->graphs.front().exec);
      sim->graphs.erase(sim->graphs.begin());
    }
    sim->graphs.push_back(e);
    CK(cudaGraphLaunch(exec, sim->stream));
    return 0;
  }
  // capture failed: nothing was enqueued; restore the host state and take the ordinary path
  if (graph) cudaGraphDestroy(graph);
  cudaGetLastError();
  sim->cur = e.cur;
  sim->ocur = e.ocur;
  sim->tile_parity = e.tile_parity;
  sim->steps_since_sort = e.steps_since_sort;
  sim->rebins = rebins0;
  sim->launches = launches0;
  sim->substeps = substeps0;
  sim->t = t0;
  sim->form_ad = false;
  sim->par.graph_mode = MPM_GRAPH_OFF;
  return advance_impl(sim, n_substeps);
}

int mpm_sync(MpmSim* sim) {
  if (!sim) return 1;
  CK(cudaSetDevice(sim->device));
  CK(cudaStreamSynchronize(sim->stream));
  if (sim->download_pending) {  // read-backs queued by the *_async calls
    CK(cudaEventSynchronize(sim->ev_downloaded));
    sim->download_pending = false;
  }
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
  CK(cudaMemcpyAsync(sim->grid, vec4, sizeof(float4) * n_nodes, cudaMemcpyHostToDevice, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return 0;
}
int mpm_debug_overwrite_particles_aos(MpmSim* sim, const MpmParticle* particles, size_t count) {
  if (!sim || !particles) return 1;
  CK(cudaSetDevice(sim->device));
  if (count != sim->count || !sim->whole_domain) return fail(sim, "overwrite needs the same particle count on a whole-domain handle");
  drop_graphs(sim);
  if (count == 0) return 0;
  if (int rc = ensure_stage(sim, count)) return rc;
  CK(cudaMemcpyAsync(sim->aos_stage, particles, sizeof(MpmParticle) * count, cudaMemcpyHostToDevice, sim->stream));
  aos_overwrite_kernel<<<blocks_for(count, 256), 256, 0, sim->stream>>>(sim->aos_stage, sim->soa[sim->cur], count, sim->first_id, sim->d_diag);
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
  if (keys) CK(cudaMemcpyAsync(keys, sim->okeys[sim->ocur], sizeof(uint32_t) * sim->count, cudaMemcpyDeviceToHost, sim->stream));
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
int mpm_set_stage_timing(MpmSim* sim, int on) {
  if (!sim) return 1;
  sim->timing = on != 0;
  return 0;
}
int mpm_get_diagnostics(MpmSim* sim, MpmDiagnostics* out) {
  if (!sim || !out) return 1;
  CK(cudaSetDevice(sim->device));
  static_assert(sizeof(MpmDiagnostics) == sizeof(DeviceDiag), "same four counters");
  CK(cudaMemcpyAsync(sim->h_diag, sim->d_diag, sizeof(DeviceDiag), cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  memcpy(out, sim->h_diag, sizeof(DeviceDiag));
  return 0;
}

int mpm_comm_unique_id(void* id128) { return Comm::unique_id(id128); }
int mpm_attach_comm(MpmSim* sim, const void* id128, int rank, int nranks) {
  if (!sim) return 1;
  CK(cudaSetDevice(sim->device));
  if (sim->capacity == 0) return fail(sim, "mpm_attach_comm: set MpmParams.capacity for slab handles");
  if (sim->par.sort_every == 0)
    return fail(sim, "mpm_attach_comm: slab handles migrate particles at the re-bin, sort_every = 0 would never migrate");
  if (sim->comm.init(id128, rank, nranks, sim->k, sim->ghost, sim->capacity, sim->stream)) return fail(sim, "mpm_attach_comm: %s", sim->comm.error());
  int lo = 0, hi = 0;
  CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CK(cudaStreamCreateWithPriority(&sim->comm_stream, cudaStreamNonBlocking, hi));
  CK(cudaEventCreateWithFlags(&sim->ev_boundary, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&sim->ev_exchanged, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&sim->split_ev, cudaEventDisableTiming));
  CK(cudaMalloc(&sim->d_split, 2 * sizeof(uint32_t)));
  CK(cudaMallocHost(&sim->h_split, 2 * sizeof(uint32_t)));
  return 0;
}

// ---- linalg hooks ---------------------------------------------------------------------------------
static int linalg_run(const float* A, size_t n, int mode, int what, float* o0, float* o1, float* o2, float aux0 = 0.f, float aux1 = 0.f) {
  MpmSim* sim = nullptr;
  float *dA = nullptr, *d0 = nullptr, *d1 = nullptr, *d2 = nullptr;
  const size_t in_sz = (what == 3) ? 3 * n : 9 * n;
  const size_t sz0 = (what == 2) ? n : 9 * n;
  CK(cudaMalloc(&dA, sizeof(float) * in_sz));
  CK(cudaMalloc(&d0, sizeof(float) * sz0));
  if (what == 0) {
    CK(cudaMalloc(&d1, sizeof(float) * 3 * n));
    CK(cudaMalloc(&d2, sizeof(float) * 9 * n));
  }
  CK(cudaMemcpy(dA, A, sizeof(float) * in_sz, cudaMemcpyHostToDevice));
  const unsigned nb = blocks_for(n, 128);
  if (what == 0) {
    if (mode == MPM_SVD_EXACT) svd3_batch_kernel<ExactOps><<<nb, 128>>>(dA, d0, d1, d2, n);
    else svd3_batch_kernel<FastOps><<<nb, 128>>>(dA, d0, d1, d2, n);
  } else if (what == 1) {
    if (mode == MPM_SVD_EXACT) polar_batch_kernel<ExactOps><<<nb, 128>>>(dA, d0, n);
    else polar_batch_kernel<FastOps><<<nb, 128>>>(dA, d0, n);
  } else if (what == 2) {
    det_batch_kernel<<<nb, 128>>>(dA, d0, n);
  } else {
    dinv_batch_kernel<<<nb, 128>>>(dA, d0, n, aux0, aux1);
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
int mpm_dinv_batch(const float* xyz, float* Dinv9, size_t n, uint32_t N) {
  if (!n) return 0;
  const float dx = (float)(1.0 / (double)N);
  const float dx_inv = (float)(1.0 / (double)dx);
  return linalg_run(xyz, n, 0, 3, Dinv9, nullptr, nullptr, dx, dx_inv);
}

}  // extern "C"
