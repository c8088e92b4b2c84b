// Shared helpers for libdmvs_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/dmvs_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libdmvs_b200 is written for sm_100a (B200) only"
#endif

namespace dmvs {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized against this

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return DMVS_ERR_CUDA;
  }
  return DMVS_OK;
}

#define DMVS_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) {                    \
      ::dmvs::set_error(__VA_ARGS__); \
      return (code);                  \
    }                                 \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

constexpr int kMaxDevices = 64;
// Per-device one-time state of a kernel template instance (the > 48 KB shared-memory opt-in is a per-device function attribute, so
// is the occupancy it allows).  Index with the CURRENT device; idempotent, racing threads write the same values.
struct PerDevice {
  bool configured[kMaxDevices] = {};
  int value[kMaxDevices] = {};
};
inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  return dev;
}

}  // namespace dmvs
