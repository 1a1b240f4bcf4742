// Library-wide state of the C ABI: last error (thread-local), launch counter, the
// device-memory pool behind DevBuf and the pinned host staging pool.
#include <atomic>
#include <map>
#include <mutex>
#include "common.cuh"
#include "../../include/ursm_b200.h"

namespace ursm {
static thread_local std::string g_error;
static std::atomic<long long> g_launches{0};
void set_error(const std::string& msg) { g_error = msg; }
void count_launch(int n) { g_launches += n; }

int device_sm_count() {
  int dev = 0, v = 0;
  URSM_CUDA(cudaGetDevice(&dev));
  URSM_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
  return v;
}

size_t device_smem_optin() {
  int dev = 0, v = 0;
  URSM_CUDA(cudaGetDevice(&dev));
  URSM_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return (size_t)v;
}

// ---- device pool ---------------------------------------------------------------
static std::mutex g_pool_mu;
static unsigned long long g_pool_ready = 0;  // bit d: default pool of device d configured

static void pool_configure() {
  int dev = 0;
  URSM_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_pool_mu);
  if (dev < 64 && (g_pool_ready >> dev) & 1ull) return;
  cudaMemPool_t pool;
  URSM_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
  unsigned long long keep = ~0ull;  // never trim: freed blocks stay cached for the next shard
  URSM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  if (dev < 64) g_pool_ready |= 1ull << dev;
}

void* pool_alloc(size_t bytes) {
  pool_configure();
  void* p = nullptr;
  URSM_CUDA(cudaMallocAsync(&p, bytes, (cudaStream_t)0));
  return p;
}

void pool_free(void* p) {
  if (p) cudaFreeAsync(p, (cudaStream_t)0);
}

// ---- pinned host pool ------------------------------------------------------------
// Page-locked blocks for results that are read back whole (E[S], L x N doubles): the
// device writes them at PCIe speed and the host never touches a fresh page.  Blocks are
// cached by exact size; the cache is capped so that a long-lived process does not pin
// memory without bound.
static std::mutex g_host_mu;
static std::multimap<size_t, void*> g_host_free;
static std::map<void*, size_t> g_host_live;
static size_t g_host_cached = 0;
static const size_t kHostCacheCap = (size_t)12 << 30;

void* host_pool_alloc(size_t bytes) {
  {
    std::lock_guard<std::mutex> lock(g_host_mu);
    auto it = g_host_free.find(bytes);
    if (it != g_host_free.end()) {
      void* p = it->second;
      g_host_free.erase(it);
      g_host_cached -= bytes;
      g_host_live[p] = bytes;
      return p;
    }
  }
  void* p = nullptr;
  URSM_CUDA(cudaHostAlloc(&p, bytes, cudaHostAllocPortable));
  std::lock_guard<std::mutex> lock(g_host_mu);
  g_host_live[p] = bytes;
  return p;
}

void host_pool_free(void* p) {
  if (!p) return;
  size_t bytes = 0;
  {
    std::lock_guard<std::mutex> lock(g_host_mu);
    auto it = g_host_live.find(p);
    if (it == g_host_live.end()) return;  // not ours
    bytes = it->second;
    g_host_live.erase(it);
    if (g_host_cached + bytes <= kHostCacheCap) {
      g_host_free.emplace(bytes, p);
      g_host_cached += bytes;
      return;
    }
  }
  cudaFreeHost(p);
}

void host_pool_trim() {
  std::lock_guard<std::mutex> lock(g_host_mu);
  for (auto& kv : g_host_free) cudaFreeHost(kv.second);
  g_host_free.clear();
  g_host_cached = 0;
}
}  // namespace ursm

extern "C" {
const char* ursm_last_error(void) { return ursm::g_error.c_str(); }
int ursm_version(void) { return 101; }
int64_t ursm_launch_count(void) { return ursm::g_launches.load(); }
void ursm_launch_count_reset(void) { ursm::g_launches = 0; }

int ursm_host_alloc(void** out, int64_t bytes) {
  URSM_API_BEGIN
  URSM_REQUIRE(out && bytes > 0, "ursm_host_alloc: bad argument");
  *out = ursm::host_pool_alloc((size_t)bytes);
  URSM_API_END
}

int ursm_host_free(void* p) {
  URSM_API_BEGIN
  ursm::host_pool_free(p);
  URSM_API_END
}

int ursm_host_trim(void) {
  URSM_API_BEGIN
  ursm::host_pool_trim();
  URSM_API_END
}
}
