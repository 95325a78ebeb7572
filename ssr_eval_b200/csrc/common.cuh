// common.cuh -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/ssr_b200.h"

namespace ssr {

std::string& last_error_ref();
std::atomic<uint64_t>& launch_counter();

inline int fail(int code, const std::string& msg) {
  last_error_ref() = msg;
  return code;
}

#define SSR_CUDA_TRY(expr)                                                               \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      return ::ssr::fail(SSR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    }                                                                                    \
  } while (0)

#define SSR_LAUNCH_CHECK(name)                                                           \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      return ::ssr::fail(SSR_ERR_CUDA, std::string(name) + " launch: " + cudaGetErrorString(_e)); \
    }                                                                                    \
    ::ssr::launch_counter().fetch_add(1, std::memory_order_relaxed);                     \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// every batched entry point indexes device buffers (and its workspace) with the ABSOLUTE offsets: they have to start
// at 0 and be non-decreasing (a caller with a sub-batch passes pointers to its first utterance and rebased offsets)
inline int check_offsets(const int64_t* offsets_host, int n, const char* who) {
  if (offsets_host[0] != 0) return fail(SSR_ERR_INVALID, std::string(who) + ": offsets must start at 0");
  for (int u = 0; u < n; ++u)
    if (offsets_host[u + 1] < offsets_host[u]) return fail(SSR_ERR_INVALID, std::string(who) + ": offsets must be non-decreasing");
  return SSR_OK;
}

// multiprocessor count of the current device (queried once per process; the persistent kernels size their grids and
// their work-item granularity with it)
inline int sm_count() {
  static int n = 0;
  if (n <= 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      n = v;
    else
      return 148;  // B200; not cached, so a later call can still succeed
  }
  return n;
}

}  // namespace ssr
