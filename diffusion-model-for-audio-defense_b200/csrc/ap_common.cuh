// Shared internals of libaudiopure_b200: error plumbing, launch accounting, small device-buffer helper.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include <string>
#include <vector>

#include "../../include/audiopure.h"

namespace ap {

extern thread_local std::string g_last_error;
extern std::atomic<unsigned long long> g_launches;
// bumped by every device (re)allocation or release of the library: a captured CUDA graph that replays the library's kernels holds
// raw pointers into its workspaces and must be re-captured when this changes (ap_alloc_generation, certify.py)
extern std::atomic<unsigned long long> g_alloc_generation;

inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap_;
  va_start(ap_, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap_);
  va_end(ap_);
  g_last_error = buf;
  return code;
}

#define AP_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return ::ap::fail(AP_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
  } while (0)

#define AP_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) return ::ap::fail(AP_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Every kernel launch of the library goes through this (bench.py reports the count as gpu_launches).
#define AP_LAUNCH_CHECK()                                                                                  \
  do {                                                                                                     \
    ::ap::g_launches.fetch_add(1, std::memory_order_relaxed);                                              \
    cudaError_t e__ = cudaGetLastError();                                                                  \
    if (e__ != cudaSuccess)                                                                                \
      return ::ap::fail(AP_ERR_CUDA, "%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// Owning device allocation (freed with the handle).
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) {
      cudaFree(p);
      g_alloc_generation.fetch_add(1, std::memory_order_relaxed);
    }
    p = nullptr;
    bytes = 0;
  }
  cudaError_t alloc(size_t n) {
    release();
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n;
    g_alloc_generation.fetch_add(1, std::memory_order_relaxed);
    return e;
  }
  cudaError_t upload(const void* host, size_t n) {
    cudaError_t e = alloc(n);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(p, host, n, cudaMemcpyHostToDevice);
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

inline uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

int select_device(int device);  // cudaSetDevice + capability check (sm_100 required); AP_OK or error

}  // namespace ap
