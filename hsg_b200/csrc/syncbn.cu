// Synchronised BatchNorm for one process per GPU (lib/nn/sync_batchnorm/batchnorm.py:55-118): the per-layer
// work either side of the ONE all-reduce of [2C] sums.  The reference computes sum and square-sum with two
// full-tensor reductions, rendezvous over Python queues, ReduceAddCoalesced + Broadcast, then ~6 elementwise
// launches; here: one statistics kernel, (host: all-reduce), one apply kernel -- same again for the backward.
//   x viewed as [B, C, L] (L = 1 for 2-D inputs); statistics per channel over B*L.
#include "common.cuh"

namespace hsg {

constexpr int BN_THREADS = 256;

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < BN_THREADS / 32 ? sh[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) t = warp_sum(t);
  __syncthreads();
  return t;                                        // valid in thread 0
}

// stats[c] = sum_x, stats[C + c] = sum_x^2   (or, with g: sum_g and sum_g * xhat)
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                              const float* __restrict__ mean, const float* __restrict__ inv,
                                                              int B, int C, int L, float* __restrict__ stats) {
  __shared__ float sh[BN_THREADS / 32];
  const int c = blockIdx.x;
  float s0 = 0.f, s1 = 0.f;
  const float mu = g ? mean[c] : 0.f, iv = g ? inv[c] : 0.f;
  const int64_t per = (int64_t)B * L;
  for (int64_t i = threadIdx.x; i < per; i += BN_THREADS) {
    const int64_t b = i / L, l = i % L;
    const int64_t at = (b * C + c) * L + l;
    const float v = x[at];
    if (g) {
      const float gv = g[at];
      s0 += gv;
      s1 = fmaf(gv, (v - mu) * iv, s1);
    } else {
      s0 += v;
      s1 = fmaf(v, v, s1);
    }
  }
  s0 = block_sum_256(s0, sh);
  s1 = block_sum_256(s1, sh);
  if (threadIdx.x == 0) { stats[c] = s0; stats[C + c] = s1; }
}

// forward: y = (x - mean) * inv * w + b ; backward: gx = (g - mean_g - xhat * mean_gx) * w * inv
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                              const float* __restrict__ mean, const float* __restrict__ inv,
                                                              const float* __restrict__ weight, const float* __restrict__ bias,
                                                              const float* __restrict__ gstats, float inv_count,
                                                              int64_t total, int C, int L, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * BN_THREADS + threadIdx.x;
  if (i >= total) return;
  const int c = (int)((i / L) % C);
  const float w = weight ? weight[c] : 1.f;
  const float xhat = (x[i] - mean[c]) * inv[c];
  if (g) {
    const float mg = gstats[c] * inv_count, mgx = gstats[C + c] * inv_count;
    out[i] = (g[i] - mg - xhat * mgx) * (w * inv[c]);
  } else {
    out[i] = xhat * w + (bias ? bias[c] : 0.f);
  }
}

}  // namespace hsg

using namespace hsg;

extern "C" {

int hsg_bn_stats_f32(const float* x, const float* grad_or_null, const float* mean, const float* inv_std, int B, int C,
                     int L, float* stats_out, void* stream) {
  HSG_REQUIRE(B > 0 && C > 0 && L > 0, HSG_E_INVALID, "bn_stats: bad shape");
  HSG_REQUIRE(x && stats_out && (!grad_or_null || (mean && inv_std)), HSG_E_INVALID, "bn_stats: null pointer");
  bn_stats_kernel<<<C, BN_THREADS, 0, (cudaStream_t)stream>>>(x, grad_or_null, mean, inv_std, B, C, L, stats_out);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int hsg_bn_apply_f32(const float* x, const float* grad_or_null, const float* mean, const float* inv_std,
                     const float* weight, const float* bias, const float* grad_stats, float inv_count, int B, int C,
                     int L, float* out, void* stream) {
  HSG_REQUIRE(B > 0 && C > 0 && L > 0, HSG_E_INVALID, "bn_apply: bad shape");
  HSG_REQUIRE(x && mean && inv_std && out && (!grad_or_null || grad_stats), HSG_E_INVALID, "bn_apply: null pointer");
  const int64_t total = (int64_t)B * C * L;
  bn_apply_kernel<<<(unsigned)ceil_div64(total, BN_THREADS), BN_THREADS, 0, (cudaStream_t)stream>>>(
      x, grad_or_null, mean, inv_std, weight, bias, grad_stats, inv_count, total, C, L, out);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

}  // extern "C"
