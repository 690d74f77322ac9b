// runtime.cu -- error string, launch counter, device query.
#include "common.cuh"
#include "../../include/stardist_b200.h"
#include <mutex>

namespace sdb {
static std::mutex g_err_mu;
static std::string g_err;
long long g_launch_count = 0;
void set_error(const std::string& msg) { std::lock_guard<std::mutex> l(g_err_mu); g_err = msg; }

static std::mutex g_init_mu;
static bool g_init_done[64] = {false};
static unsigned int* g_pinned = nullptr;
void ensure_runtime_init() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  if (g_init_done[dev]) return;
  std::lock_guard<std::mutex> l(g_init_mu);
  if (g_init_done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  g_init_done[dev] = true;
}
unsigned int* pinned_scratch() {
  std::lock_guard<std::mutex> l(g_init_mu);
  if (!g_pinned) { if (cudaMallocHost(&g_pinned, 64 * sizeof(unsigned int)) != cudaSuccess) g_pinned = nullptr; }
  return g_pinned;
}
}  // namespace sdb

extern "C" const char* sdb_last_error(void) {
  static thread_local std::string copy;
  std::lock_guard<std::mutex> l(sdb::g_err_mu);
  copy = sdb::g_err;
  return copy.c_str();
}

extern "C" long long sdb_launch_count(int reset) {
  long long v = sdb::g_launch_count;
  if (reset) sdb::g_launch_count = 0;
  return v;
}

extern "C" int sdb_device_info(int* n_devices, int* sm_count, int* cc_major, int* cc_minor) {
  int n = 0;
  SDB_CUDA(cudaGetDeviceCount(&n));
  if (n_devices) *n_devices = n;
  if (n == 0) { sdb::set_error("no CUDA device"); return 1; }
  int dev = 0;
  SDB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  SDB_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}
