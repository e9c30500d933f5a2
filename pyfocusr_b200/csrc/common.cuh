// Shared plumbing of libfocusr_b200.so: error reporting, launch checks, small device helpers.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <stdint.h>
#include <stdio.h>

#include "../../include/focusr_b200.h"

namespace fb {

enum {
  FB_OK = 0,
  FB_ERR_CUDA = 100,
  FB_ERR_ARG = 101,
  FB_ERR_WORKSPACE = 102,
  FB_ERR_UNSUPPORTED = 103
};

void set_error(const char* fmt, ...);

#define FB_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      fb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return fb::FB_ERR_CUDA;                                                              \
    }                                                                                      \
  } while (0)

#define FB_LAUNCH_CHECK()                                                                         \
  do {                                                                                            \
    cudaError_t e__ = cudaGetLastError();                                                         \
    if (e__ != cudaSuccess) {                                                                     \
      fb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__));   \
      return fb::FB_ERR_CUDA;                                                                     \
    }                                                                                             \
  } while (0)

#define FB_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      fb::set_error(__VA_ARGS__);  \
      return fb::FB_ERR_ARG;       \
    }                              \
  } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Carves typed arrays out of a caller-provided workspace.
struct Carver {
  char* base;
  size_t used;
  size_t cap;
  Carver(void* p, size_t bytes) : base((char*)p), used(0), cap(bytes) {}
  template <class T>
  T* take(size_t n) {
    used = align_up(used);
    T* p = (T*)(base + used);
    used += n * sizeof(T);
    return p;
  }
  bool ok() const { return used <= cap && (base != nullptr || used == 0); }
};

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
// SMs of the current device (148 on B200): asked per call, never cached process-wide (a process may use several GPUs)
inline int sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
    return 148;
  return sms;
}

// number of launches issued by this library since load (bench.py reports it as gpu_launches); atomic because
// independent calls may come from several host threads (one stream each), as the batched CPD does
extern std::atomic<unsigned long long> g_launch_count;
#define FB_COUNT_LAUNCH(n) (fb::g_launch_count += (unsigned long long)(n))

// exclusive prefix sum of int32 (out[n] = total); tmp needs scan_tmp_ints(n) ints
size_t scan_tmp_ints(int n);
int exclusive_scan_i32(const int* in, int* out, int n, int* tmp, cudaStream_t stream);

}  // namespace fb
