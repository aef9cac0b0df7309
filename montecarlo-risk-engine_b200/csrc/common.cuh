// Shared host-side helpers of libmcre_b200: error reporting, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <vector>
#include "../../include/mcre.h"

namespace mcre {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char *fmt, const char *a = "", long long b = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}
inline int cuda_fail(cudaError_t e, const char *what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}
#define MCRE_CUDA(call)                                      \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return mcre::cuda_fail(e__, #call); \
  } while (0)
#define MCRE_LAUNCHED()                                           \
  do {                                                            \
    mcre::g_launches.fetch_add(1, std::memory_order_relaxed);     \
    cudaError_t e__ = cudaGetLastError();                         \
    if (e__ != cudaSuccess) return mcre::cuda_fail(e__, "kernel launch"); \
  } while (0)

// Device buffer filled from a host array (plan tables are tiny; cudaMalloc is fine).
template <typename T>
struct DevArray {
  T *p = nullptr;
  size_t n = 0;
  int upload(const T *host, size_t count) {
    n = count;
    if (count == 0 || host == nullptr) { p = nullptr; n = 0; return 0; }
    MCRE_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
    MCRE_CUDA(cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

int sm_count();

}  // namespace mcre
