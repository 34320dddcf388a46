// Multi-GPU plumbing for slab decomposition along x (one handle per rank): halo-plane exchange
// over NCCL.  No reference counterpart (the reference is single-GPU).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "common.cuh"

namespace mpm {

struct Comm {
  bool active() const { return false; }
  const char* error() const { return err.c_str(); }
  static int unique_id(void*) { return 1; }
  int init(const void*, int, int, const KParams&, cudaStream_t) {
    err = "multi-GPU support not built yet";
    return 1;
  }
  int exchange_halo(float4*, const KParams&, cudaStream_t, uint64_t*) { return 0; }
  void destroy() {}
  std::string err;
};

}  // namespace mpm
