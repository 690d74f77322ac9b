// common.cuh -- error handling, launch accounting and small device helpers shared by all TUs.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string>

namespace sdb {

void set_error(const std::string& msg);
extern long long g_launch_count;
// one-time per-device runtime setup (keeps the stream-ordered memory pool's blocks cached instead of
// returning them to the OS at every synchronisation) + a small persistent pinned scratch buffer
void ensure_runtime_init();
unsigned int* pinned_scratch();     // 64 uint32, per process

// optional live profiling of named kernels with CUDA events on the launching stream (bench.py's
// roofline numbers): enabled through sdb_profile_enable(); record_begin/end bracket one launch
struct ProfSpan { cudaEvent_t a, b; };
bool profile_enabled();
void profile_begin(const char* name, cudaStream_t st, ProfSpan* sp);
void profile_end(const char* name, cudaStream_t st, ProfSpan* sp);
void profile_add_units(const char* name, double units);     // e.g. pairs processed

#define SDB_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      char _b[512];                                                                         \
      snprintf(_b, sizeof(_b), "%s:%d: %s failed: %s", __FILE__, __LINE__, #call,           \
               cudaGetErrorString(_e));                                                     \
      sdb::set_error(_b);                                                                   \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

// launch + count + check the launch itself (not completion)
#define SDB_LAUNCH(kernel, grid, block, smem, stream, ...)                                  \
  do {                                                                                      \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                             \
    sdb::g_launch_count++;                                                                  \
    SDB_CUDA(cudaGetLastError());                                                           \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// RAII for stream-ordered scratch allocations
struct DevBuf {
  void* p = nullptr;
  cudaStream_t s = 0;
  size_t bytes = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  cudaError_t alloc(size_t n, cudaStream_t st) {
    ensure_runtime_init();
    release();
    s = st; bytes = n;
    if (n == 0) n = 16;
    return cudaMallocAsync(&p, n, st);
  }
  void release() { if (p) { cudaFreeAsync(p, s); p = nullptr; } }
  ~DevBuf() { release(); }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// order-preserving float -> uint32 (ascending)
__host__ __device__ inline uint32_t float_order_key(float f) {
  uint32_t u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

}  // namespace sdb
