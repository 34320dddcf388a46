// Stage (1): particle cell binning — cell keys, stable LSD radix sort of (key, index) pairs, SoA
// permute.  No reference counterpart (the reference never sorts); the contract is
// the north star's: keys and permutation bit-exact against a stable sort on the key.
//
// Per 8-bit pass: (a) per-tile digit histogram, (b) exclusive scan of the digit-major table,
// (c) scatter with a deterministic in-tile rank.  The rank is computed warp-synchronously with
// match.any (peers holding the same digit) — no atomics decide an output position, so the
// result is the unique stable permutation.
#pragma once
#include "common.cuh"

namespace mpm {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 8;                                     // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;              // keys per block
// digit width per pass: 8 bits unless 9 bits save a whole pass (slab handles at N = 512 have 25-bit keys
// plus the tombstone key: 3 passes of 9 instead of 4 of 8)
constexpr int kMaxRadixBits = 9;
constexpr int kMaxRadix = 1 << kMaxRadixBits;
inline int radix_passes(int key_bits) { return (key_bits + kMaxRadixBits - 1) / kMaxRadixBits; }
inline int radix_bits(int key_bits) { return ((key_bits + 7) / 8 > radix_passes(key_bits)) ? 9 : 8; }

// key = N^2*bi + N*bj + bk of the clamped base node, bi local to the slab
// tombstoned particles (migrated away) get dead_key, which sorts behind every live key
__global__ void cell_key_kernel(Soa p, size_t count, KParams k, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                uint32_t dead_key) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  vals[i] = (uint32_t)i;
  if (p.id[i] == kDeadId) {
    keys[i] = dead_key;
    return;
  }
  const float* c = p.col(i);
  int b[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float fx, w[3];
    bspline(c[(SX + a) * kTile], k.dx_inv, b[a], fx, w);
    b[a] = min(max(b[a], 0), k.N - 1);
  }
  const int bx = min(max(b[0] - k.x0, 0), k.nxl - 1);
  keys[i] = (uint32_t)((bx * k.N + b[1]) * k.N + b[2]);
}

// (a) table[d * n_tiles + tile] = number of keys of this tile whose digit is d
template <int BITS>
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint32_t* __restrict__ keys, size_t count, int shift, uint32_t* __restrict__ table, int n_tiles) {
  constexpr int kRadix = 1 << BITS;
  __shared__ uint32_t h[kRadix];
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) h[d] = 0;
  __syncthreads();
  const size_t tile0 = (size_t)blockIdx.x * kSortTile;
  static_assert(kSortItems == 8, "two 16-byte loads per thread");
  // the histogram does not care which thread counts which key of the tile: thread t takes the 8
  // consecutive keys at tile0 + 8 t with two independent 16-byte loads issued before any is used
  // (one 4-byte load feeding one atomic at a time left the kernel at 2 TB/s, latency-bound)
  const size_t i0 = tile0 + (size_t)threadIdx.x * kSortItems;
  if (i0 + kSortItems <= count) {
    const uint4 a = *reinterpret_cast<const uint4*>(keys + i0);
    const uint4 b = *reinterpret_cast<const uint4*>(keys + i0 + 4);
    const uint32_t kk[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int it = 0; it < kSortItems; ++it) atomicAdd(&h[(kk[it] >> shift) & (kRadix - 1)], 1u);
  } else {
    for (size_t i = i0; i < count && i < i0 + kSortItems; ++i) atomicAdd(&h[(keys[i] >> shift) & (kRadix - 1)], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) table[(size_t)d * n_tiles + blockIdx.x] = h[d];
}

// (b) device-wide exclusive scan of uint32, three launches: tile sums, scan of sums, downsweep
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem /*>= 32*/, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t s = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    smem[lane] = s;
  }
  __syncthreads();
  const uint32_t warp_off = warp ? smem[warp - 1] : 0;
  total = smem[(blockDim.x >> 5) - 1];
  __syncthreads();
  return warp_off + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ sums) {
  __shared__ uint32_t sm[32];
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
    if (base + j < n) s += in[base + j];
  uint32_t total;
  block_exclusive_scan(s, sm, total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
// single block: exclusive scan of `sums` in place (n up to a few 10^5)
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t* sums, size_t n) {
  __shared__ uint32_t sm[32];
  uint32_t carry = 0;
  for (size_t base = 0; base < n; base += 1024) {
    const size_t i = base + threadIdx.x;
    const uint32_t v = (i < n) ? sums[i] : 0;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, sm, total);
    if (i < n) sums[i] = carry + ex;
    carry += total;
  }
}
__global__ void __launch_bounds__(kScanThreads) scan_downsweep_kernel(uint32_t* __restrict__ data, size_t n, const uint32_t* __restrict__ sums) {
  __shared__ uint32_t sm[32];
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    v[j] = (base + j < n) ? data[base + j] : 0;
    s += v[j];
  }
  uint32_t total;
  uint32_t run = block_exclusive_scan(s, sm, total) + sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    if (base + j < n) data[base + j] = run;
    run += v[j];
  }
}

// (c) stable scatter.  Warp w owns the contiguous chunk [tile0 + w*256, +256) of its tile and
// walks it 32 keys at a time, so (warp, iteration, lane) order is the input order.
constexpr int kScatterMinBlocks = 5;
template <int BITS>
__global__ void __launch_bounds__(kSortThreads, BITS == 8 ? kScatterMinBlocks : 4)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                     uint32_t* __restrict__ vals_out, size_t count, int shift, const uint32_t* __restrict__ table, int n_tiles) {
  constexpr int kRadix = 1 << BITS;
  __shared__ uint32_t cnt[kSortWarps][kRadix];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < kSortWarps * kRadix; d += kSortThreads) (&cnt[0][0])[d] = 0;
  __syncthreads();
  const size_t chunk0 = (size_t)blockIdx.x * kSortTile + (size_t)warp * (kSortItems * 32);
  // Only the ranks stay in registers between the two loops; keys and values are read again for the
  // write-out (they are in L1/L2 by then).  Holding them cost 16 registers and a third of the
  // resident warps of a kernel that does little but wait for memory.
  uint32_t rank[kSortItems];
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const size_t i = chunk0 + (size_t)it * 32 + lane;
    const bool valid = i < count;
    const uint32_t key = valid ? keys_in[i] : 0u;
    const uint32_t digit = valid ? ((key >> shift) & (kRadix - 1)) : (uint32_t)(kRadix + lane);
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const int leader = __ffs(peers) - 1;
    uint32_t before = 0;
    if (valid && lane == leader) {
      before = cnt[warp][digit];
      cnt[warp][digit] = before + __popc(peers);
    }
    before = __shfl_sync(0xffffffffu, before, leader);
    rank[it] = before + __popc(peers & lt_mask);
    __syncwarp();
  }
  __syncthreads();
  // turn per-warp counts into output bases: global tile offset + counts of earlier warps
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) {
    uint32_t run = table[(size_t)d * n_tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = cnt[w][d];
      cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kSortItems; ++it) {
    const size_t i = chunk0 + (size_t)it * 32 + lane;
    if (i < count) {
      const uint32_t key = keys_in[i];
      const uint32_t digit = (key >> shift) & (kRadix - 1);
      const uint32_t pos = cnt[warp][digit] + rank[it];
      keys_out[pos] = key;
      vals_out[pos] = vals_in[i];
    }
  }
}

// histogram tiles must match the scatter's tiles: tile t = keys [t*kSortTile, (t+1)*kSortTile)
static_assert(kSortTile == kSortWarps * kSortItems * 32, "tile layout");

// SoA permute: dst[r] = src[perm[r]] for every stream (+ id, material).
// PARTIAL: only what G2P reads (x, F, Jp = stream rows SX..SJ, + id, material) — for a re-bin placed
// between the grid update and G2P, which overwrites v and C of every particle it processes.  The
// particles G2P leaves untouched (whole stencil outside the domain, src/mpm.cu:128-132) keep all
// their streams.  58 % of the full permute's bytes.
template <bool PARTIAL>
__global__ void __launch_bounds__(256) permute_kernel(Soa src, Soa dst, const uint32_t* __restrict__ perm, size_t count, KParams k) {
  static_assert(SX == 12 && SJ == NSTREAM - 1, "rows SX.. = x, F, Jp");
  constexpr int R0 = PARTIAL ? SX : 0;
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= count) return;
  const uint32_t s = perm[r];
  const float* __restrict__ in = src.col(s);
  float* __restrict__ out = dst.tile(blockIdx.x) + threadIdx.x;
  float t[NSTREAM - R0];
#pragma unroll
  for (int q = R0; q < NSTREAM; ++q) t[q - R0] = in[q * kTile];
  const uint32_t id = src.id[s];
  const uint8_t mt = src.mat[s];
#pragma unroll
  for (int q = R0; q < NSTREAM; ++q) out[q * kTile] = t[q - R0];
  dst.id[r] = id;
  dst.mat[r] = mt;
  if (PARTIAL) {
    int base[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) base[a] = (int)(t[SX - R0 + a] * k.dx_inv - 0.5f);
    if (stencil_outside(base, k.N)) {
      for (int q = 0; q < R0; ++q) out[q * kTile] = in[q * kTile];
    }
  }
}

// ---- merge re-bin -------------------------------------------------------------------------------------
// Between two re-bins most particles stay in their cell.  Those form a subsequence of the current order
// that is ALREADY sorted by the new keys (their new key is their old one), so the stable sort of the whole
// sequence is the merge of that subsequence with the (few) particles whose key changed, sorted among
// themselves by (key, current index):
//   1. flags: moved[i] = new key != old key (or no old key: arrivals of a slab handle); one ballot word per
//      32 particles + its prefix count, so that "moved before position p" is an O(1) lookup;
//   2. the host reads the moved count back (it sizes the launches that follow; above an eighth of the
//      particles the radix sort takes over); the moved (key, index) pairs are compacted and radix-sorted;
//   3. a moved pair's final position = its rank among the moved + the number of stayed particles that
//      precede it, found by binary search in the OLD sorted keys (two levels: a 1/64 sample first);
//   4. a stayed particle's final position = its rank among the stayed + the number of moved pairs that
//      precede it, found in the slice of the sorted moved list that its 2048-particle tile can see
//      (one search per tile bounds the slice, which is then staged in shared memory).
// The result (keys and permutation) is identical to the LSD radix sort's, for a fraction of the traffic when
// few particles moved.
constexpr int kMergeTile = 2048;
constexpr int kMergeThreads = 256;
constexpr int kCoarse = 64;                 // sampling stride of the old keys for the two-level search
constexpr uint32_t kMergeSentinel = 0xffffffffu;

// moved flags of tile `blockIdx.x` as ballot words (mask[i >> 5]) and the tile's moved count
__global__ void __launch_bounds__(kMergeThreads)
rebin_flags_kernel(const uint32_t* __restrict__ newk, const uint32_t* __restrict__ oldk, uint32_t n, uint32_t n_old, uint32_t* __restrict__ mask,
                   uint32_t* __restrict__ tile_moved) {
  __shared__ uint32_t warp_cnt[kMergeThreads / 32];
  const uint32_t tile0 = blockIdx.x * (uint32_t)kMergeTile;
  uint32_t cnt = 0;
#pragma unroll
  for (int r = 0; r < kMergeTile / kMergeThreads; ++r) {
    const uint32_t i = tile0 + r * kMergeThreads + threadIdx.x;
    const bool moved = i < n && !(i < n_old && newk[i] == oldk[i]);
    const uint32_t m = __ballot_sync(0xffffffffu, moved);
    if ((threadIdx.x & 31) == 0) {
      if (i < n + 32) mask[i >> 5] = m;  // (i is a multiple of 32 here; one word beyond the end stays in bounds: see the allocation)
      cnt += __popc(m);
    }
  }
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kMergeThreads / 32; ++w) t += warp_cnt[w];
    tile_moved[blockIdx.x] = t;
  }
}

// single block: exclusive scan of data[0..n) in place, total to *total (and to data[n])
__global__ void __launch_bounds__(1024) exclusive_scan_total_kernel(uint32_t* data, uint32_t n, uint32_t* total) {
  __shared__ uint32_t sm[32];
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = (i < n) ? data[i] : 0;
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan(v, sm, tot);
    if (i < n) data[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) {
    data[n] = carry;
    *total = carry;
  }
}

// per 32-particle word: moved count before it (wprefix); the moved pairs compacted in order (m_up = their
// number as the host read it back; a larger m_up pads the pair buffers with sentinels that sort last)
__global__ void __launch_bounds__(kMergeThreads)
rebin_compact_kernel(const uint32_t* __restrict__ newk, uint32_t n, const uint32_t* __restrict__ mask, const uint32_t* __restrict__ tile_base,
                     uint32_t* __restrict__ wprefix, uint32_t* __restrict__ mk, uint32_t* __restrict__ mi, uint32_t m_up,
                     const uint32_t* __restrict__ n_moved_ptr) {
  constexpr int kWords = kMergeTile / 32;  // 64 ballot words per tile
  static_assert(kWords == 64, "two warps scan the tile's ballot words");
  __shared__ uint32_t wpre[kWords];
  __shared__ uint32_t half_total;
  const uint32_t tile0 = blockIdx.x * (uint32_t)kMergeTile;
  const uint32_t n_words = (n + 31) / 32;
  const uint32_t word0 = tile0 / 32;
  uint32_t c = 0, inc = 0;
  if (threadIdx.x < kWords) {  // exclusive prefix of the 64 popcounts: one scan per warp, then the first half's total
    const uint32_t w = word0 + threadIdx.x;
    c = (w < n_words) ? __popc(mask[w]) : 0u;
    inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if (threadIdx.x == 31) half_total = inc;
  }
  __syncthreads();
  if (threadIdx.x < kWords) wpre[threadIdx.x] = inc - c + (threadIdx.x >= 32 ? half_total : 0u);
  __syncthreads();
  const uint32_t base = tile_base[blockIdx.x];
  if (threadIdx.x < kWords && word0 + threadIdx.x < n_words) wprefix[word0 + threadIdx.x] = base + wpre[threadIdx.x];
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < kMergeTile / kMergeThreads; ++r) {
    const uint32_t i = tile0 + r * kMergeThreads + threadIdx.x;
    if (i >= n) break;
    const uint32_t m = mask[i >> 5];
    if ((m >> lane) & 1u) {
      const uint32_t rank = base + wpre[(i - tile0) >> 5] + __popc(m & ((1u << lane) - 1u));
      mk[rank] = newk[i];
      mi[rank] = i;
    }
  }
  const uint32_t n_moved = *n_moved_ptr;
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // the end marker of the O(1) lookups
    wprefix[n_words] = n_moved;
  }
  for (uint32_t j = n_moved + blockIdx.x * kMergeThreads + threadIdx.x; j < m_up; j += gridDim.x * kMergeThreads) {
    mk[j] = kMergeSentinel;
    mi[j] = kMergeSentinel;
  }
}

// coarse[q] = sorted_keys[q * kCoarse]
__global__ void __launch_bounds__(256) coarse_keys_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ coarse) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if ((size_t)q * kCoarse < n) coarse[q] = keys[(size_t)q * kCoarse];
}

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* __restrict__ a, uint32_t lo, uint32_t hi, uint32_t v) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// first position of the sorted old keys whose key >= v, through the 1/kCoarse sample
__device__ __forceinline__ uint32_t lower_bound_two_level(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ coarse, uint32_t n,
                                                          uint32_t v) {
  const uint32_t nq = (n + kCoarse - 1) / kCoarse;
  const uint32_t q = lower_bound_u32(coarse, 0, nq, v);  // first sample >= v: the answer lies in ((q-1) * kCoarse, q * kCoarse]
  const uint32_t lo = q ? (q - 1) * kCoarse + 1 : 0;
  const uint32_t hi = min(n, q * (uint32_t)kCoarse);
  return lower_bound_u32(keys, min(lo, hi), hi, v);
}
// moved particles before position p (p in [0, n]) from the ballot words and their prefix counts
__device__ __forceinline__ uint32_t moved_before(const uint32_t* __restrict__ mask, const uint32_t* __restrict__ wprefix, uint32_t p) {
  return wprefix[p >> 5] + __popc(mask[p >> 5] & ((1u << (p & 31)) - 1u));
}

// step 3: final position of every moved pair (sorted by (key, index) in mk / mi)
__global__ void __launch_bounds__(256)
rebin_place_moved_kernel(const uint32_t* __restrict__ mk, const uint32_t* __restrict__ mi, uint32_t m_up, const uint32_t* __restrict__ oldk,
                         const uint32_t* __restrict__ coarse, uint32_t n_old, const uint32_t* __restrict__ mask,
                         const uint32_t* __restrict__ wprefix, uint32_t* __restrict__ perm, uint32_t* __restrict__ keys_out) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m_up) return;
  const uint32_t K = mk[j];
  const uint32_t idx = mi[j];
  if (K == kMergeSentinel && idx == kMergeSentinel) return;
  const uint32_t lb = lower_bound_two_level(oldk, coarse, n_old, K);
  const uint32_t ub = (K == 0xffffffffu) ? n_old : lower_bound_two_level(oldk, coarse, n_old, K + 1u);
  // stayed particles with this key sit in [lb, ub) of the current order; those before `idx` precede the pair
  const uint32_t p = idx >= n_old ? ub : min(max(idx, lb), ub);
  const uint32_t stayed_before = p - moved_before(mask, wprefix, p);
  const uint32_t pos = j + stayed_before;
  perm[pos] = idx;
  keys_out[pos] = K;
}

// moved pairs below (key, index): pairs compare by key, then by index
__device__ __forceinline__ uint32_t pairs_below(const uint32_t* __restrict__ k, const uint32_t* __restrict__ ix, uint32_t lo, uint32_t hi, uint32_t key,
                                                uint32_t index) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    const uint32_t km = k[mid];
    if (km < key || (km == key && ix[mid] < index)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// step 4a, one thread per tile: slice_begin[t] = moved pairs below the first stayed particle at or after the
// start of tile t (n_moved when there is none; also the entry behind the last tile).  Stayed particles have
// ascending keys, so the pairs a tile's stayed particles can see as their boundary lie in
// [slice_begin[t], slice_begin[t + 1]].
__global__ void __launch_bounds__(128)
rebin_tile_slices_kernel(const uint32_t* __restrict__ newk, uint32_t n, const uint32_t* __restrict__ mask, const uint32_t* __restrict__ mk,
                         const uint32_t* __restrict__ mi, uint32_t n_moved, uint32_t n_tiles, uint32_t* __restrict__ slice_begin) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  uint32_t first = 0xffffffffu;
  const uint32_t n_words = (n + 31) / 32;
  for (uint32_t w = t * (uint32_t)(kMergeTile / 32); w < n_words; ++w) {
    uint32_t stay = ~mask[w];
    if (w * 32 + 32 > n) stay &= (1u << (n - w * 32)) - 1u;
    if (stay) {
      first = w * 32 + (uint32_t)__ffs(stay) - 1u;
      break;
    }
  }
  slice_begin[t] = (first == 0xffffffffu) ? n_moved : pairs_below(mk, mi, 0, n_moved, newk[first], first);
}

// step 4b: final position of every stayed particle of tile `blockIdx.x`
__global__ void __launch_bounds__(kMergeThreads)
rebin_place_stayed_kernel(const uint32_t* __restrict__ newk, uint32_t n, const uint32_t* __restrict__ mask, const uint32_t* __restrict__ wprefix,
                          const uint32_t* __restrict__ mk, const uint32_t* __restrict__ mi, const uint32_t* __restrict__ slice_begin,
                          uint32_t* __restrict__ perm, uint32_t* __restrict__ keys_out) {
  constexpr int kWindow = 2048;
  __shared__ uint32_t wk[kWindow], wi[kWindow];
  const uint32_t tile0 = blockIdx.x * (uint32_t)kMergeTile;
  const uint32_t a = slice_begin[blockIdx.x], b = slice_begin[blockIdx.x + 1];
  const bool staged = (b - a) <= (uint32_t)kWindow;
  if (staged) {
    for (uint32_t e = threadIdx.x; e < b - a; e += kMergeThreads) {
      wk[e] = mk[a + e];
      wi[e] = mi[a + e];
    }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  constexpr int kRounds = kMergeTile / kMergeThreads;
  uint32_t key[kRounds], m[kRounds], wp[kRounds];
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {  // all loads first: the searches below then overlap
    const uint32_t i = tile0 + r * kMergeThreads + threadIdx.x;
    const bool in = i < n;
    key[r] = in ? newk[i] : 0u;
    m[r] = in ? mask[i >> 5] : 0xffffffffu;
    wp[r] = in ? wprefix[i >> 5] : 0u;
  }
  if (staged) {
    const uint32_t len = b - a;
    uint32_t lo[kRounds], hi[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) lo[r] = 0, hi[r] = len;
    // the eight searches of a thread advance together (branch-free steps; len <= 2048: at most 12 of them)
    for (uint32_t span = len; span; span >>= 1) {
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const uint32_t i = tile0 + r * kMergeThreads + threadIdx.x;
        if (lo[r] < hi[r]) {
          const uint32_t mid = lo[r] + ((hi[r] - lo[r]) >> 1);
          const uint32_t km = wk[mid];
          const bool below = km < key[r] || (km == key[r] && wi[mid] < i);
          lo[r] = below ? mid + 1 : lo[r];
          hi[r] = below ? hi[r] : mid;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      const uint32_t i = tile0 + r * kMergeThreads + threadIdx.x;
      if ((m[r] >> lane) & 1u) continue;  // moved (or beyond the end): placed by rebin_place_moved_kernel
      const uint32_t pos = (i - (wp[r] + __popc(m[r] & ((1u << lane) - 1u)))) + a + lo[r];
      perm[pos] = i;
      keys_out[pos] = key[r];
    }
  } else {
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      const uint32_t i = tile0 + r * kMergeThreads + threadIdx.x;
      if ((m[r] >> lane) & 1u) continue;
      const uint32_t pos = (i - (wp[r] + __popc(m[r] & ((1u << lane) - 1u)))) + pairs_below(mk, mi, a, b, key[r], i);
      perm[pos] = i;
      keys_out[pos] = key[r];
    }
  }
}

// ---- removal of an id range (an object whose lifetime ends, include/mpm.cuh:36-40) ------------------------
// Stable compaction of the survivors: the order within the survivors — the cell order — is kept, so no
// re-bin is needed afterwards.  Same building blocks as the merge re-bin: ballot words of the REMOVED
// particles, their per-tile counts, an exclusive scan, then perm[rank among survivors] = slot.
__global__ void __launch_bounds__(kMergeThreads)
remove_flags_kernel(const uint32_t* __restrict__ ids, uint32_t n, uint32_t id_begin, uint32_t id_end, uint32_t* __restrict__ mask,
                    uint32_t* __restrict__ tile_removed) {
  __shared__ uint32_t warp_cnt[kMergeThreads / 32];
  const uint32_t tile0 = blockIdx.x * (uint32_t)kMergeTile;
  uint32_t cnt = 0;
#pragma unroll
  for (int r = 0; r < kMergeTile / kMergeThreads; ++r) {
    const uint32_t i = tile0 + r * kMergeThreads + threadIdx.x;
    const bool removed = i < n && ids[i] >= id_begin && ids[i] < id_end;
    const uint32_t m = __ballot_sync(0xffffffffu, removed);
    if ((threadIdx.x & 31) == 0) {
      if (i < n + 32) mask[i >> 5] = m;
      cnt += __popc(m);
    }
  }
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kMergeThreads / 32; ++w) t += warp_cnt[w];
    tile_removed[blockIdx.x] = t;
  }
}
// perm[rank of slot i among the survivors] = i (tile_base = exclusive scan of the removed counts per tile)
__global__ void __launch_bounds__(kMergeThreads)
remove_perm_kernel(uint32_t n, const uint32_t* __restrict__ mask, const uint32_t* __restrict__ tile_base, uint32_t* __restrict__ perm) {
  constexpr int kWords = kMergeTile / 32;
  __shared__ uint32_t wpre[kWords];
  const uint32_t tile0 = blockIdx.x * (uint32_t)kMergeTile;
  const uint32_t n_words = (n + 31) / 32, word0 = tile0 / 32;
  if (threadIdx.x == 0) {  // 64 popcounts: a serial prefix is as fast as anything here
    uint32_t run = 0;
    for (int w = 0; w < kWords; ++w) {
      wpre[w] = run;
      if (word0 + w < n_words) run += __popc(mask[word0 + w]);
    }
  }
  __syncthreads();
  const uint32_t base = tile_base[blockIdx.x];
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < kMergeTile / kMergeThreads; ++r) {
    const uint32_t i = tile0 + r * kMergeThreads + threadIdx.x;
    if (i >= n) break;
    const uint32_t m = mask[i >> 5];
    if ((m >> lane) & 1u) continue;
    const uint32_t removed_before = base + wpre[(i - tile0) >> 5] + __popc(m & ((1u << lane) - 1u));
    perm[i - removed_before] = i;
  }
}
// ids stay the positions in upload order: those behind the removed range move down
__global__ void __launch_bounds__(256) renumber_ids_kernel(uint32_t* __restrict__ ids, uint32_t n, uint32_t id_end, uint32_t removed) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && ids[i] >= id_end) ids[i] -= removed;
}

// out[0] = first sorted position whose key >= key_lo, out[1] = first whose key >= key_hi (slab handles:
// the particles before out[0] / from out[1] on can reach the planes shared with a neighbour)
__global__ void split_bounds_kernel(const uint32_t* __restrict__ keys, uint32_t count, uint32_t key_lo, uint32_t key_hi,
                                    uint32_t* __restrict__ out) {
  if (threadIdx.x >= 2) return;
  const uint32_t target = threadIdx.x == 0 ? key_lo : key_hi;
  uint32_t lo = 0, hi = count;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  out[threadIdx.x] = lo;
}

}  // namespace mpm
