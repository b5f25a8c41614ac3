// Shared helpers for the hsg_b200 sm_100a kernels (error plumbing, warp
// primitives, small device utilities).  Internal to the library; the public
// boundary is include/hsg_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/hsg_b200.h"

namespace hsg {

// ----------------------------------------------------------------- errors
void set_error(const char* fmt, ...);

#define HSG_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      ::hsg::set_error(__VA_ARGS__);            \
      return (code);                            \
    }                                           \
  } while (0)

#define HSG_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      ::hsg::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                       __FILE__, __LINE__);                                   \
      return HSG_E_CUDA;                                                      \
    }                                                                         \
  } while (0)

#define HSG_LAUNCH_CHECK()                                                    \
  do {                                                                        \
    ::hsg::count_launch();                                                    \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) {                                                 \
      ::hsg::set_error("kernel launch failed: %s (%s:%d)",                    \
                       cudaGetErrorString(e__), __FILE__, __LINE__);          \
      return HSG_E_CUDA;                                                      \
    }                                                                         \
  } while (0)

int num_sms();   // SM count of the current device (cached per device)
void count_launch();

// optional per-phase device timing (hsg_profile_* in the C ABI): records a pair
// of CUDA events on the launching stream around a phase when enabled
enum ProfPhase { PROF_PREP = 0, PROF_MSTEP_SORT, PROF_MSTEP_GATHER, PROF_MSTEP_COMBINE,
                 PROF_ESTEP, PROF_ESTEP_FIXUP, PROF_RELABEL, PROF_POOL, PROF_NCE_FWD,
                 PROF_NCE_BWD, PROF_CONVERT, PROF_KMEANS /* one whole hsg_kmeans_* call */, PROF_NUM };
struct ProfSuppress {   // inner ranges are skipped while one of these is alive (per thread)
  ProfSuppress();
  ~ProfSuppress();
};
struct ProfRange {
  int slot;
  cudaStream_t st;
  ProfRange(int phase, cudaStream_t stream);
  ~ProfRange();
};

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// carve a workspace pointer into aligned sub-buffers
struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* p) : base(static_cast<char*>(p)), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
  size_t used() const { return align_up(off, 256); }
};

// ----------------------------------------------------------------- device
#ifdef __CUDACC__

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// inclusive warp scan
__device__ __forceinline__ int warp_scan_incl(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// reference rule for normalize_embedding (hsg/utils/general/common.py:116-120):
// divide by the norm, or by eps when the norm is below eps.
__device__ __forceinline__ float safe_norm(float sumsq) {
  float n = sqrtf(sumsq);
  return n >= 1e-12f ? n : 1e-12f;
}

// streaming (read-once) global loads
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ld_stream2(const float* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

#endif  // __CUDACC__

}  // namespace hsg
