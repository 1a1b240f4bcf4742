// Shared plumbing for the C-ABI translation units: error capture, launch
// counting, warp/block reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

namespace ursm {

void set_error(const std::string& msg);
void count_launch(int n = 1);

struct Failure {
  std::string msg;
};

#define URSM_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      char buf__[512];                                                                   \
      snprintf(buf__, sizeof buf__, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call,   \
               cudaGetErrorString(e__));                                                 \
      throw ::ursm::Failure{buf__};                                                      \
    }                                                                                    \
  } while (0)

#define URSM_REQUIRE(cond, text)                                                         \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      char buf__[512];                                                                   \
      snprintf(buf__, sizeof buf__, "%s:%d: %s", __FILE__, __LINE__, text);              \
      throw ::ursm::Failure{buf__};                                                      \
    }                                                                                    \
  } while (0)

#define URSM_LAUNCH_CHECK()                \
  do {                                     \
    ::ursm::count_launch();                \
    URSM_CUDA(cudaGetLastError());         \
  } while (0)

// wraps a C-ABI body: exceptions -> status code + ursm_last_error()
#define URSM_API_BEGIN try {
#define URSM_API_END                          \
  return 0;                                   \
  }                                           \
  catch (const ::ursm::Failure& f) {          \
    ::ursm::set_error(f.msg);                 \
    return 1;                                 \
  }                                           \
  catch (const std::exception& e) {           \
    ::ursm::set_error(e.what());              \
    return 2;                                 \
  }

// Device memory comes from the CUDA stream-ordered pool of the current device with its
// release threshold raised to "never" (common.cu): a shard that is destroyed and rebuilt
// -- one per fit in the reference's API -- reuses its blocks instead of paying
// cudaMalloc/cudaFree (milliseconds each, and a device synchronize) every time.
// Allocation and release are ordered on the legacy default stream, which every blocking
// stream synchronizes with; the *_destroy entry points synchronize the device first.
// attributes of the current device (cudaDeviceGetAttribute: microseconds, where
// cudaGetDeviceProperties takes milliseconds per call)
int device_sm_count();
size_t device_smem_optin();

void* pool_alloc(size_t bytes);
void pool_free(void* p);

template <class T>
struct DevBuf {  // owning device buffer
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  void alloc(size_t count) {
    release();
    n = count;
    if (count) p = static_cast<T*>(pool_alloc(count * sizeof(T)));
  }
  void zero(cudaStream_t s = 0) {
    if (n) URSM_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
  void release() {
    if (p) pool_free(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

// host fp64 matrix -> device fp32 (staged through the device in chunks, cast there; sc_gibbs.cu)
void upload_f64_as_f32(const double* h, float* d, size_t n);
// row sums of an fp32 [L][N] matrix; and out[type][i] += sum_{rows of the type} Y[row][i] / R[row]
// (type-sorted `order` and `G`, or both null: one type, natural order) -- the per-type means of
// the row-normalised counts that parameter initialisation needs (gem.py:225-255); sc_gibbs.cu
void launch_rowsum(const float* Y, long long L, int N, double* R, cudaStream_t st);
void launch_normalised_colsum(const float* Y, const int* order, const int* G, const double* R,
                              long long L, int N, int num_sms, double* out, cudaStream_t st);

#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum; `red` is shared scratch of >= 32 doubles; result valid in all threads
__device__ __forceinline__ double block_sum(double v, double* red) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double x = (lane < nw) ? red[lane] : 0.0;
  return warp_sum(x);
}
#endif

}  // namespace ursm
