// Multi-GPU plumbing for the slab decomposition along x (one handle = one rank = one GPU).
// No reference counterpart (the reference is single-GPU).
//
// Geometry.  Rank r owns x-planes [xb, xe) and the particles whose base node lies there at the
// last migration.  Its local grid holds planes [xb - g, xe + 2 + g) (clipped to the domain): 2
// planes above because the stencil reaches base+2, plus g ghost planes either side so particles
// may drift g cells out of their slab between migrations.  Neighbouring ranks therefore both
// hold the 2 + 2g planes around their common boundary.
//
// Halo exchange (every substep, between P2G and the grid update): each side sends its partial
// sums for the shared planes to the other and adds what it receives.  IEEE addition is
// commutative, so both ranks end up with bit-identical totals and can run the grid update
// redundantly on the shared planes: ONE exchange per substep instead of reduce + broadcast.
//
// Migration (every re-bin): leavers are packed into per-neighbour send buffers and tombstoned
// (id = kDeadId), arrivals are appended to the SoA tail, and the cell sort that follows moves the
// tombstones behind the live particles.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>

#include <string>

#include "common.cuh"

namespace mpm {

__global__ void __launch_bounds__(256) halo_add_kernel(float4* __restrict__ grid, const float4* __restrict__ recv, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = grid[i];
  const float4 b = recv[i];
  a.x += b.x;
  a.y += b.y;
  a.z += b.z;
  a.w += b.w;
  grid[i] = a;
}

// dest 0 = lower neighbour, 1 = upper neighbour.  Packed layout: stream s of destination d at
// buf + (d * kMigRows + s) * cap; rows NSTREAM / NSTREAM+1 carry id and material as bits.
constexpr int kMigRows = NSTREAM + 2;
// keys (optional): the cell keys the P2G of this substep wrote for the re-bin; a leaver's becomes dead_key
__global__ void __launch_bounds__(256) migrate_pack_kernel(Soa p, size_t count, KParams k, float* __restrict__ buf, size_t cap,
                                                           unsigned int* __restrict__ counters, uint32_t* __restrict__ keys, uint32_t dead_key) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  if (p.id[i] == kDeadId) return;
  const float* __restrict__ c = p.col(i);
  int b;
  float fx, w[3];
  bspline(c[SX * kTile], k.dx_inv, b, fx, w);
  b = min(max(b, 0), k.N - 1);
  int d;
  if (b < k.x_own_begin) d = 0;
  else if (b >= k.x_own_end) d = 1;
  else return;
  const unsigned int slot = atomicAdd(&counters[d], 1u);
  if (slot < cap) {
    float* base = buf + (size_t)d * kMigRows * cap + slot;
#pragma unroll
    for (int s = 0; s < NSTREAM; ++s) base[(size_t)s * cap] = c[s * kTile];
    base[(size_t)NSTREAM * cap] = __uint_as_float(p.id[i]);
    base[(size_t)(NSTREAM + 1) * cap] = __uint_as_float((uint32_t)p.mat[i]);
  }
  p.id[i] = kDeadId;  // tombstone: sorted behind the live particles by the re-bin that follows
  if (keys) keys[i] = dead_key;
}

// arrivals from one neighbour (stream-major in `buf`, row stride cap) into slots [first, first + n);
// keys / vals (optional): their entries of the (key, index) pairs the re-bin sorts
__global__ void __launch_bounds__(256) migrate_unpack_kernel(Soa p, size_t first, size_t n, const float* __restrict__ buf, size_t cap, KParams k,
                                                             uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* __restrict__ c = p.col(first + i);
#pragma unroll
  for (int s = 0; s < NSTREAM; ++s) c[s * kTile] = buf[(size_t)s * cap + i];
  p.id[first + i] = __float_as_uint(buf[(size_t)NSTREAM * cap + i]);
  p.mat[first + i] = (uint8_t)__float_as_uint(buf[(size_t)(NSTREAM + 1) * cap + i]);
  if (keys) {
    int b[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float fx, w[3];
      bspline(buf[(size_t)(SX + a) * cap + i], k.dx_inv, b[a], fx, w);
      b[a] = min(max(b[a], 0), k.N - 1);
    }
    const int bx = min(max(b[0] - k.x0, 0), k.nxl - 1);
    keys[first + i] = (uint32_t)((bx * k.N + b[1]) * k.N + b[2]);
    vals[first + i] = (uint32_t)(first + i);
  }
}

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  bool on = false;
  bool has_lo = false, has_hi = false;
  // shared plane ranges, in local plane indices
  int lo_begin = 0, lo_end = 0, hi_begin = 0, hi_end = 0;
  float4* recv_lo = nullptr;
  float4* recv_hi = nullptr;
  // migration
  size_t mig_cap = 0;
  float* send_buf = nullptr;  // 2 * kMigRows * mig_cap
  float* recv_buf = nullptr;  // 2 * kMigRows * mig_cap
  unsigned int* d_counts = nullptr;  // [0..1] out, [2..3] in, [4] escape count summed over all ranks
  unsigned int* h_counts = nullptr;  // pinned mirror
  std::string err;

  bool active() const { return on; }
  const char* error() const { return err.c_str(); }

  static int unique_id(void* id128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    return ncclGetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)) == ncclSuccess ? 0 : 1;
  }

  int fail(const char* what, ncclResult_t r) {
    err = std::string(what) + ": " + ncclGetErrorString(r);
    return 1;
  }
  int failc(const char* what, cudaError_t e) {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return 1;
  }
  // a peer that died or a transport error shows up here, not in the enqueue calls
  int check_async() {
    ncclResult_t st = ncclSuccess;
    const ncclResult_t r = ncclCommGetAsyncError(comm, &st);
    if (r != ncclSuccess) return fail("ncclCommGetAsyncError", r);
    if (st != ncclSuccess && st != ncclInProgress) return fail("NCCL asynchronous error", st);
    return 0;
  }

  int init(const void* id128, int rank_, int nranks_, const KParams& k, int ghost, size_t capacity, cudaStream_t) {
    rank = rank_;
    nranks = nranks_;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t r = ncclCommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) return fail("ncclCommInitRank", r);
    has_lo = k.x_own_begin > 0;
    has_hi = k.x_own_end < k.N;
    if ((has_lo && rank == 0) || (has_hi && rank == nranks - 1)) {
      err = "slab order must follow rank order (rank 0 owns x = 0)";
      return 1;
    }
    const int own = k.x_own_end - k.x_own_begin;
    if (own < 2 + 2 * ghost) {
      err = "slab thinner than 2 + 2*ghost planes";
      return 1;
    }
    const size_t plane = (size_t)k.N * k.N;
    if (has_lo) {
      lo_begin = (k.x_own_begin - ghost) - k.x0;
      lo_end = min(k.N, k.x_own_begin + 2 + ghost) - k.x0;
      if (lo_begin < 0) { err = "ghost planes below the domain"; return 1; }
      cudaError_t e = cudaMalloc(&recv_lo, sizeof(float4) * plane * (lo_end - lo_begin));
      if (e != cudaSuccess) return failc("cudaMalloc(recv_lo)", e);
    }
    if (has_hi) {
      hi_begin = (k.x_own_end - ghost) - k.x0;
      hi_end = min(k.N, k.x_own_end + 2 + ghost) - k.x0;
      cudaError_t e = cudaMalloc(&recv_hi, sizeof(float4) * plane * (hi_end - hi_begin));
      if (e != cudaSuccess) return failc("cudaMalloc(recv_hi)", e);
    }
    mig_cap = capacity / 16 + 4096;  // a re-bin may move at most this many particles to one neighbour
    cudaError_t e = cudaMalloc(&send_buf, sizeof(float) * 2 * kMigRows * mig_cap);
    if (e != cudaSuccess) return failc("cudaMalloc(send_buf)", e);
    e = cudaMalloc(&recv_buf, sizeof(float) * 2 * kMigRows * mig_cap);
    if (e != cudaSuccess) return failc("cudaMalloc(recv_buf)", e);
    e = cudaMalloc(&d_counts, sizeof(unsigned int) * 8);
    if (e != cudaSuccess) return failc("cudaMalloc(counts)", e);
    e = cudaMallocHost(&h_counts, sizeof(unsigned int) * 8);
    if (e != cudaSuccess) return failc("cudaMallocHost(counts)", e);
    on = true;
    return 0;
  }

  // one grouped send/recv pair per neighbour, then the add
  int exchange_halo(float4* grid, const KParams& k, cudaStream_t stream, uint64_t* launches) {
    if (!on) return 0;
    const size_t plane = (size_t)k.N * k.N;
    const size_t n_lo = plane * (lo_end - lo_begin), n_hi = plane * (hi_end - hi_begin);
    ncclResult_t r = ncclGroupStart();
    if (r != ncclSuccess) return fail("ncclGroupStart", r);
    if (has_lo) {
      ncclSend(grid + plane * lo_begin, n_lo * 4, ncclFloat, rank - 1, comm, stream);
      ncclRecv(recv_lo, n_lo * 4, ncclFloat, rank - 1, comm, stream);
    }
    if (has_hi) {
      ncclSend(grid + plane * hi_begin, n_hi * 4, ncclFloat, rank + 1, comm, stream);
      ncclRecv(recv_hi, n_hi * 4, ncclFloat, rank + 1, comm, stream);
    }
    r = ncclGroupEnd();
    if (r != ncclSuccess) return fail("ncclGroupEnd(halo)", r);
    if (int rc = check_async()) return rc;
    if (has_lo) {
      halo_add_kernel<<<(unsigned)((n_lo + 255) / 256), 256, 0, stream>>>(grid + plane * lo_begin, recv_lo, n_lo);
      ++*launches;
    }
    if (has_hi) {
      halo_add_kernel<<<(unsigned)((n_hi + 255) / 256), 256, 0, stream>>>(grid + plane * hi_begin, recv_hi, n_hi);
      ++*launches;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return failc("halo_add_kernel", e);
    return 0;
  }

  // Moves leavers to the neighbours and appends arrivals behind `count`.  On return *count
  // includes the arrivals and still includes the tombstoned leavers; *n_dead is their number.
  // d_escaped: this rank's count of particle-substeps that scattered outside its planes since the last
  // re-bin; *escaped_all = the sum over ALL ranks, so that every rank takes the same decision about it.
  int migrate(Soa& p, size_t* count, size_t capacity, const KParams& k, cudaStream_t stream, uint64_t* launches, size_t* n_dead,
              const unsigned int* d_escaped, unsigned int* escaped_all, uint32_t* keys = nullptr, uint32_t* vals = nullptr, uint32_t dead_key = 0) {
    *n_dead = 0;
    *escaped_all = 0;
    if (!on) return 0;
    cudaError_t e = cudaMemsetAsync(d_counts, 0, sizeof(unsigned int) * 8, stream);
    if (e != cudaSuccess) return failc("memset(counts)", e);
    if (*count) {
      migrate_pack_kernel<<<(unsigned)((*count + 255) / 256), 256, 0, stream>>>(p, *count, k, send_buf, mig_cap, d_counts, keys, dead_key);
      ++*launches;
    }
    // counts: mine out -> neighbours' in
    ncclResult_t r = ncclGroupStart();
    if (r != ncclSuccess) return fail("ncclGroupStart", r);
    if (has_lo) {
      ncclSend(d_counts + 0, 1, ncclUint32, rank - 1, comm, stream);
      ncclRecv(d_counts + 2, 1, ncclUint32, rank - 1, comm, stream);
    }
    if (has_hi) {
      ncclSend(d_counts + 1, 1, ncclUint32, rank + 1, comm, stream);
      ncclRecv(d_counts + 3, 1, ncclUint32, rank + 1, comm, stream);
    }
    r = ncclGroupEnd();
    if (r != ncclSuccess) return fail("ncclGroupEnd(counts)", r);
    r = ncclAllReduce(d_escaped, d_counts + 4, 1, ncclUint32, ncclSum, comm, stream);
    if (r != ncclSuccess) return fail("ncclAllReduce(escaped)", r);
    e = readback_words(h_counts, d_counts, 8, stream);
    if (e != cudaSuccess) return failc("memcpy(counts)", e);
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return failc("sync(counts)", e);
    *escaped_all = h_counts[4];
    const size_t out_lo = has_lo ? h_counts[0] : 0, out_hi = has_hi ? h_counts[1] : 0;
    const size_t in_lo = has_lo ? h_counts[2] : 0, in_hi = has_hi ? h_counts[3] : 0;
    if (out_lo > mig_cap || out_hi > mig_cap || in_lo > mig_cap || in_hi > mig_cap) {
      err = "migration buffer overflow (sort more often or raise capacity)";
      return 1;
    }
    if (*count + in_lo + in_hi > capacity) {
      err = "particle capacity exceeded by migration";
      return 1;
    }
    const size_t first_lo = *count, first_hi = *count + in_lo;
    r = ncclGroupStart();
    if (r != ncclSuccess) return fail("ncclGroupStart", r);
    // the packed rows of one direction are one contiguous block only up to the row stride: send row by row
    for (int s = 0; s < kMigRows; ++s) {
      if (has_lo) {
        if (out_lo) ncclSend(send_buf + (size_t)s * mig_cap, out_lo, ncclFloat, rank - 1, comm, stream);
        if (in_lo) ncclRecv(recv_buf + (size_t)s * mig_cap, in_lo, ncclFloat, rank - 1, comm, stream);
      }
      if (has_hi) {
        if (out_hi) ncclSend(send_buf + (size_t)(kMigRows + s) * mig_cap, out_hi, ncclFloat, rank + 1, comm, stream);
        if (in_hi) ncclRecv(recv_buf + (size_t)(kMigRows + s) * mig_cap, in_hi, ncclFloat, rank + 1, comm, stream);
      }
    }
    r = ncclGroupEnd();
    if (r != ncclSuccess) return fail("ncclGroupEnd(payload)", r);
    if (int rc = check_async()) return rc;
    if (in_lo) {
      migrate_unpack_kernel<<<(unsigned)((in_lo + 255) / 256), 256, 0, stream>>>(p, first_lo, in_lo, recv_buf, mig_cap, k, keys, vals);
      ++*launches;
    }
    if (in_hi) {
      migrate_unpack_kernel<<<(unsigned)((in_hi + 255) / 256), 256, 0, stream>>>(p, first_hi, in_hi, recv_buf + (size_t)kMigRows * mig_cap, mig_cap, k, keys,
                                                                                 vals);
      ++*launches;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return failc("migrate kernels", e);
    *count += in_lo + in_hi;
    *n_dead = out_lo + out_hi;
    return 0;
  }

  void destroy() {
    if (comm) ncclCommDestroy(comm);
    comm = nullptr;
    cudaFree(recv_lo);
    cudaFree(recv_hi);
    cudaFree(send_buf);
    cudaFree(recv_buf);
    cudaFree(d_counts);
    if (h_counts) cudaFreeHost(h_counts);
    recv_lo = recv_hi = nullptr;
    send_buf = recv_buf = nullptr;
    d_counts = h_counts = nullptr;
    on = false;
  }
};

}  // namespace mpm
