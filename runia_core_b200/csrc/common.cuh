// Shared helpers for libruniab200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>

#include <atomic>

#include <nvtx3/nvToolsExt.h>

#include "../../include/runia_b200.h"

namespace runia {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int finish_launch(const char *what);  // cudaGetLastError -> return code (+ message)

#define RUNIA_REQUIRE(cond, code, ...)            \
  do {                                            \
    if (!(cond)) {                                \
      ::runia::set_error(__VA_ARGS__);            \
      return (code);                              \
    }                                             \
  } while (0)

#define RUNIA_CUDA(call)                                                          \
  do {                                                                            \
    cudaError_t _e = (call);                                                      \
    if (_e != cudaSuccess) {                                                      \
      ::runia::set_error("%s failed: %s", #call, cudaGetErrorString(_e));         \
      return (int)_e;                                                             \
    }                                                                             \
  } while (0)

// "Done once" flag for per-device state (cudaFuncSetAttribute is per device / context): reads as false until it
// has been set on the CURRENT device.  A racing second thread at worst repeats an idempotent attribute call.
struct PerDeviceFlag {
  std::atomic<uint64_t> mask{0};
  static int cur() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
  }
  bool operator!() const {
    const int dev = cur();
    return dev >= 64 || !((mask.load(std::memory_order_acquire) >> dev) & 1ull);
  }
  PerDeviceFlag &operator=(bool v) {
    const int dev = cur();
    if (v && dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release);
    return *this;
  }
};

// NVTX range around every C-ABI entry point (header-only nvtx3: a no-op unless a profiler is attached), so that
// nsys / ncu --nvtx timelines show the library calls by name.
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define RUNIA_NVTX() ::runia::NvtxRange runia_nvtx_range_(__func__)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float warp_sum16(float v) {  // over aligned groups of 16 lanes
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ float warp_max16(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return v;
}
__device__ __forceinline__ float warp_sum32(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return warp_sum16(v);
}
__device__ __forceinline__ float warp_max32(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
  return warp_max16(v);
}
// Fixed-order float64 sum over the 32 lanes (xor butterfly: 16, 8, 4, 2, 1).  IEEE addition is
// commutative, so every lane ends with the same bits; the oracle replays the same tree.
__device__ __forceinline__ double warp_tree_sum_f64(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

}  // namespace runia
