// Shared helpers for the sm_100a kernels of libmgcn.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "mgcn.h"

namespace mgcn {

extern std::atomic<long long> g_launch_count;

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// carve a sub-buffer out of a workspace; with base == nullptr only the running size is advanced
struct WorkspaceCarver {
  char* base;
  size_t off = 0;
  explicit WorkspaceCarver(void* b) : base(static_cast<char*>(b)) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
  size_t bytes() const { return align_up(off, 256); }
};

}  // namespace mgcn

// Launch + count + error check.  Used inside functions returning int.
#define MGCN_LAUNCH(kernel, grid, block, smem, stream, ...)                          \
  do {                                                                               \
    kernel<<<(grid), (block), (smem), static_cast<cudaStream_t>(stream)>>>(__VA_ARGS__); \
    ::mgcn::g_launch_count.fetch_add(1, std::memory_order_relaxed);                  \
    cudaError_t mgcn_err__ = cudaGetLastError();                                     \
    if (mgcn_err__ != cudaSuccess) return static_cast<int>(mgcn_err__);              \
  } while (0)

#define MGCN_CHECK_CUDA(expr)                                       \
  do {                                                              \
    cudaError_t mgcn_err__ = (expr);                                \
    if (mgcn_err__ != cudaSuccess) return static_cast<int>(mgcn_err__); \
  } while (0)

#define MGCN_REQUIRE(cond, code) \
  do {                           \
    if (!(cond)) return (code);  \
  } while (0)
