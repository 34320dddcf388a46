// Particle tiles: the unit of work of the staged (TMA-fed) P2G and G2P kernels.
//
// After a re-bin the particles are sorted by cell key, z fastest, so the particles of one (x, y)
// grid row are contiguous.  A tile is a run of at most kTileMax consecutive particles OF ONE ROW:
// tiles never straddle rows, so the grid nodes a tile touches form one box
// (x-1..x+3) x (y-1..y+3) x (z_first-1 .. z_last+3) that a single tiled TMA load can fetch.
// Built from row-level arrays only (binary search per row, one small scan), no per-particle pass.
#pragma once
#include "common.cuh"

namespace mpm {

constexpr int kTile = 256;     // threads per tile and columns of the TMA stream box
constexpr int kTileMax = 252;  // particles per tile: the stream box must start at a 16-byte aligned
                               // column (start & ~3), so up to 3 leading columns belong to the tile before

struct TileDesc {
  uint32_t start;   // first particle (slot) of the tile
  uint32_t n;       // particles in the tile, 1..kTileMax
  uint32_t kfirst;  // cell key of the first / last particle at the re-bin
  uint32_t klast;
};

// row_first[r] = first sorted position whose key >= r*N  (r = 0..n_rows)
__global__ void __launch_bounds__(256) row_bounds_kernel(const uint32_t* __restrict__ keys, uint32_t count, uint32_t N, uint32_t n_rows,
                                                         uint32_t* __restrict__ row_first) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  const unsigned long long target = (unsigned long long)r * N;
  uint32_t lo = 0, hi = count;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if ((unsigned long long)keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  row_first[r] = lo;
}

// single block: tile_base[r] = exclusive scan of ceil(count_r / kTileMax); tile_base[n_rows] = *n_tiles = total
__global__ void __launch_bounds__(1024) row_tiles_scan_kernel(const uint32_t* __restrict__ row_first, uint32_t n_rows,
                                                              uint32_t* __restrict__ tile_base, uint32_t* __restrict__ n_tiles) {
  __shared__ uint32_t warp_sum[32];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < n_rows; base += 1024) {
    const uint32_t r = base + threadIdx.x;
    const uint32_t v = (r < n_rows) ? (row_first[r + 1] - row_first[r] + kTileMax - 1) / kTileMax : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t s = warp_sum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
      }
      warp_sum[lane] = s;
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    const uint32_t ex = carry + (warp ? warp_sum[warp - 1] : 0u) + inc - v;
    if (r < n_rows) tile_base[r] = ex;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = ex + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    tile_base[n_rows] = carry_s;
    *n_tiles = carry_s;
  }
}

// one thread per row: writes the row's tile descriptors; stats[0] counts tiles whose cell span
// fits the small box (span + 5 <= lt_small), stats[1] all tiles
__global__ void __launch_bounds__(256) tile_fill_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ row_first,
                                                        const uint32_t* __restrict__ tile_base, uint32_t n_rows,
                                                        TileDesc* __restrict__ tiles, int lt_small, unsigned int* __restrict__ stats) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const uint32_t b = row_first[r], e = row_first[r + 1];
  uint32_t t = tile_base[r];
  unsigned int fit = 0, all = 0;
  for (uint32_t s = b; s < e; s += kTileMax, ++t) {
    TileDesc d;
    d.start = s;
    d.n = min((uint32_t)kTileMax, e - s);
    d.kfirst = keys[s];
    d.klast = keys[s + d.n - 1];
    tiles[t] = d;
    ++all;
    if ((int)(d.klast - d.kfirst) + 5 <= lt_small) ++fit;
  }
  if (all) {
    atomicAdd(stats + 0, fit);
    atomicAdd(stats + 1, all);
  }
}

}  // namespace mpm
