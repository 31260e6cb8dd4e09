// Shared host/device helpers for libalaya_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/alaya_b200.h"

namespace ab {

constexpr uint32_t kEmptyId = 0xFFFFFFFFu;  // reference: Graph::kEmptyId == -1 (include/index/graph/graph.hpp:49)
constexpr uint32_t kIdMask = 0x7FFFFFFFu;   // reference: LinearPool::kMask (include/utils/query_utils.hpp:302)
constexpr uint32_t kCheckedBit = 0x80000000u;
constexpr int kRowAlign = 64;  // reference: SequentialStorage rows are padded to 64 B (sequential_storage.hpp:53-57)

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

extern std::atomic<uint64_t> g_kernel_launches;
inline void count_launch(uint64_t n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

#define AB_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      throw ::ab::Error(AB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +   \
                                         __FILE__ + ":" + std::to_string(__LINE__) + ")");        \
    }                                                                                              \
  } while (0)

#define AB_CHECK(cond, code, msg)                 \
  do {                                            \
    if (!(cond)) throw ::ab::Error((code), (msg)); \
  } while (0)

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
#ifdef __CUDACC__
__host__ __device__
#endif
inline size_t round_up16(size_t v) { return (v + 15) & ~size_t(15); }

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  // grow-only allocation (contents undefined after growth)
  void reserve(size_t count) {
    if (count <= n) return;
    release();
    AB_CUDA(cudaMalloc(reinterpret_cast<void **>(&p), count * sizeof(T)));
    n = count;
  }
  void alloc_zero(size_t count, cudaStream_t s) {
    reserve(count);
    AB_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), s));
  }
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint4 ldg128(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// Ask the TMA unit to pull `bytes` (a multiple of 16, 16-byte aligned) towards L2: one instruction per row, no
// register or shared-memory landing zone, nothing to wait for.
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
#endif

}  // namespace ab
