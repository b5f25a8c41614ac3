// Shared by the tensor-core attention kernels (attention_tc.cu: forward, attention_bwd_tc.cu: backward).
#pragma once

#include "tc_common.cuh"

namespace hsg {

constexpr int AC_BM = 128;                // query rows (forward, dq) or keys (dk/dv) per CTA
constexpr int AC_HD = 64;                 // head dim of the tensor-core path
constexpr int AC_SLAB = AC_BM * 64 * 2;   // [128 rows x 64 fp16] = 16 KiB
constexpr int AC_TSLAB = AC_HD * 64 * 2;  // [64 rows x 64 fp16] = 8 KiB: a slab of a transposed operand

#ifdef __CUDACC__
__device__ __forceinline__ bool attn_dropout_keep(uint64_t seed, uint32_t bh, uint32_t row, uint32_t col, float p) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (((uint64_t)bh << 40) ^ ((uint64_t)row << 20) ^ col);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f) >= p;
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// multiplier of an operand split: mul, or mul / *amax (0 when *amax == 0) when a device-side normaliser is given
__device__ __forceinline__ float attn_mul(float mul, const float* amax) {
  if (!amax) return mul;
  const float m = *amax;
  return m > 0.f ? mul / m : 0.f;
}

// byte offset of the 16-byte chunk holding columns [8*chunk, 8*chunk+8) of row r in a 128B-swizzled
// K-major slab of 64 fp16 columns (rows of 128 bytes, 8-row groups 1024 bytes apart)
__device__ __forceinline__ uint32_t swz128_off(int r, int chunk) {
  return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
}
#endif

static inline int attn_pad64(int n) { return (n + 63) / 64 * 64; }

// operand preparation (attention_tc.cu); `amax` (device, optional): the multiplier becomes mul / *amax (0 if *amax == 0)
int attn_split_rows(const float* src, int64_t R, float mul, const float* amax, __half* dst, cudaStream_t st);
int attn_split_transposed(const float* src, int64_t BH, int n, int np, float mul, const float* amax, __half* dst,
                          cudaStream_t st);

}  // namespace hsg
