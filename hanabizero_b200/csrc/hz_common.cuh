// Shared host/device helpers for libhzb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/hzb200.h"

#define HZ_WARP 32
#define HZ_FULL 0xffffffffu

namespace hz {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int fail_cuda(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return HZ_ERR_CUDA;
}

#define HZ_CUDA(call)                                        \
  do {                                                       \
    cudaError_t _e = (call);                                 \
    if (_e != cudaSuccess) return hz::fail_cuda(_e, #call);  \
  } while (0)

#define HZ_LAUNCH_CHECK(name)                                \
  do {                                                       \
    hz::g_launches.fetch_add(1, std::memory_order_relaxed);  \
    cudaError_t _e = cudaGetLastError();                     \
    if (_e != cudaSuccess) return hz::fail_cuda(_e, name);   \
  } while (0)

// RAII device guard: the handle's device is made current for the call and restored after.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    target = dev;
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != target) cudaSetDevice(prev);
  }
  int target = -1;
};

}  // namespace hz
