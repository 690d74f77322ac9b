// runtime.cu -- error string, launch counter, device query.
#include "common.cuh"
#include "../../include/stardist_b200.h"
#include <mutex>
#include <map>
#include <vector>

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

struct ProfEntry { std::vector<ProfSpan> spans; double units = 0; double ms = 0; long long launches = 0; };
static bool g_prof_on = false;
static std::map<std::string, ProfEntry> g_prof;
bool profile_enabled() { return g_prof_on; }
void profile_begin(const char* name, cudaStream_t st, ProfSpan* sp) {
  if (!g_prof_on) return;
  cudaEventCreate(&sp->a); cudaEventCreate(&sp->b);
  cudaEventRecord(sp->a, st);
}
void profile_end(const char* name, cudaStream_t st, ProfSpan* sp) {
  if (!g_prof_on) return;
  cudaEventRecord(sp->b, st);
  g_prof[name].spans.push_back(*sp);
}
void profile_add_units(const char* name, double units) { if (g_prof_on) g_prof[name].units += units; }
static void profile_collect() {
  for (auto& kv : g_prof) {
    for (auto& sp : kv.second.spans) {
      float ms = 0;
      if (cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) { kv.second.ms += ms; kv.second.launches++; }
      cudaEventDestroy(sp.a); cudaEventDestroy(sp.b);
    }
    kv.second.spans.clear();
  }
}
}  // namespace sdb

extern "C" int sdb_profile_enable(int on) { sdb::profile_collect(); sdb::g_prof.clear(); sdb::g_prof_on = on != 0; return 0; }
// accumulated device time (ms), launch count and unit count of a named kernel since sdb_profile_enable(1)
extern "C" int sdb_profile_get(const char* name, double* ms, long long* launches, double* units) {
  sdb::profile_collect();
  auto it = sdb::g_prof.find(name);
  if (it == sdb::g_prof.end()) { if (ms) *ms = 0; if (launches) *launches = 0; if (units) *units = 0; return 1; }
  if (ms) *ms = it->second.ms; if (launches) *launches = it->second.launches; if (units) *units = it->second.units;
  return 0;
}

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
