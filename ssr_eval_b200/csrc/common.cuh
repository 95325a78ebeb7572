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

}  // namespace ssr
