// K5: fused attention core of the clustering transformer.
//
// Reference: nn.MultiheadAttention on its need_weights=True slow path
// (hsg/models/heads/transformer.py:235,300,304; torch/nn/functional.py
// multi_head_attention_forward): q scaled by 1/sqrt(hd), baddbmm with the
// additive -inf key-padding mask, softmax over keys, dropout on the
// probabilities, bmm with v -- five kernels and two [B*h, L, S] temporaries per
// call.  Here: one kernel, exact fp32, online softmax, probabilities never
// stored in the forward pass.
//
// Sequence lengths on this path are tiny (S <= 256 padded prototypes, L in
// {256, 64, 16, 8, 4}; head dim 32 or 64), so round 1 keeps this on CUDA cores:
// the arithmetic of one call is ~1 GFLOP and the win is launch count and
// traffic.  (tcgen05 version: next round)
//
// Layout: q [BH, L, hd], k/v [BH, S, hd], key_padding_mask [B, S] (1 = ignore),
// BH = B * heads with the head index fastest (b = bh / heads), like torch.
#include "common.cuh"

#include <float.h>

namespace hsg {

constexpr int AT_WARPS = 8;           // query rows per CTA (one warp each)
constexpr int AT_TS = 64;             // keys per shared-memory tile
constexpr int AT_HD_MAX = 128;

// counter-based keep mask for dropout on the probabilities (not torch's Philox
// stream: train-mode parity is defined with p = 0)
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint32_t bh, uint32_t row, uint32_t col, float p) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (((uint64_t)bh << 40) ^ ((uint64_t)row << 20) ^ col);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f) >= p;
}

struct AttnArgs {
  const float* q;
  const float* k;
  const float* v;
  const unsigned char* mask;   // [B,S] or NULL
  int BH, heads, L, S, hd;
  float scale;
  float drop_p;
  uint64_t seed;
};

// forward: out [BH,L,hd], lse [BH,L] (log-sum-exp of the scaled, masked scores)
template <int NV>   // NV = ceil(hd / 32): output dims per lane
__global__ void __launch_bounds__(AT_WARPS * 32) attn_fwd_kernel(const AttnArgs a, float* __restrict__ out,
                                                                 float* __restrict__ lse) {
  extern __shared__ float sm[];
  const int hdp = a.hd + 1;
  float* Ks = sm;                              // [AT_TS][hd+1]
  float* Vs = Ks + AT_TS * hdp;                // [AT_TS][hd+1]
  float* Qs = Vs + AT_TS * hdp;                // [AT_WARPS][hd]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int row = blockIdx.x * AT_WARPS + warp;
  const bool active = row < a.L;
  const int b = bh / a.heads;
  const float* kb = a.k + (int64_t)bh * a.S * a.hd;
  const float* vb = a.v + (int64_t)bh * a.S * a.hd;
  if (active)
    for (int d = lane; d < a.hd; d += 32) Qs[warp * a.hd + d] = a.q[((int64_t)bh * a.L + row) * a.hd + d] * a.scale;
  float m = -INFINITY, l = 0.f, o[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) o[i] = 0.f;
  const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;

  for (int s0 = 0; s0 < a.S; s0 += AT_TS) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < AT_TS * a.hd; idx += blockDim.x) {
      const int j = idx / a.hd, d = idx % a.hd;
      const bool ok = s0 + j < a.S;
      Ks[j * hdp + d] = ok ? kb[(int64_t)(s0 + j) * a.hd + d] : 0.f;
      Vs[j * hdp + d] = ok ? vb[(int64_t)(s0 + j) * a.hd + d] : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    // scores of this lane's two keys
    float sc[AT_TS / 32];
#pragma unroll
    for (int t = 0; t < AT_TS / 32; ++t) {
      const int j = lane + 32 * t;
      float acc = 0.f;
      for (int d = 0; d < a.hd; ++d) acc = fmaf(Qs[warp * a.hd + d], Ks[j * hdp + d], acc);
      const bool valid = s0 + j < a.S && !(a.mask && a.mask[(int64_t)b * a.S + s0 + j]);
      sc[t] = valid ? acc : -INFINITY;
    }
    float tmax = sc[0];
#pragma unroll
    for (int t = 1; t < AT_TS / 32; ++t) tmax = fmaxf(tmax, sc[t]);
    tmax = warp_max(tmax);
    const float m_new = fmaxf(m, tmax);
    if (m_new == -INFINITY) continue;            // everything masked so far
    const float corr = __expf(m - m_new);        // exp(-inf) = 0 on the first live tile
    float p[AT_TS / 32], psum = 0.f;
#pragma unroll
    for (int t = 0; t < AT_TS / 32; ++t) {
      p[t] = expf(sc[t] - m_new);
      psum += p[t];
    }
    l = l * corr + warp_sum(psum);
#pragma unroll
    for (int i = 0; i < NV; ++i) o[i] *= corr;
    m = m_new;
#pragma unroll
    for (int t = 0; t < AT_TS / 32; ++t) {
      for (int jj = 0; jj < 32; ++jj) {
        float pj = __shfl_sync(FULL, p[t], jj);
        const int j = jj + 32 * t;
        if (a.drop_p > 0.f) pj = dropout_keep(a.seed, bh, row, s0 + j, a.drop_p) ? pj * keep_scale : 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int d = lane + 32 * i;
          if (d < a.hd) o[i] = fmaf(pj, Vs[j * hdp + d], o[i]);
        }
      }
    }
  }
  if (active) {
    const float inv = 1.f / l;                   // l == 0 (all keys masked) -> NaN, like the reference
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int d = lane + 32 * i;
      if (d < a.hd) out[((int64_t)bh * a.L + row) * a.hd + d] = o[i] * inv;
    }
    if (lane == 0) lse[(int64_t)bh * a.L + row] = m + logf(l);
  }
}

// backward, pass 1 (one warp per query row): recompute p, write
//   pd[bh,row,j] = dropped/scaled probability, ds[bh,row,j] = p * (dp - delta)
// and dq = scale * ds K.
template <int NV>
__global__ void __launch_bounds__(AT_WARPS * 32) attn_bwd_rows_kernel(const AttnArgs a, const float* __restrict__ out,
                                                                      const float* __restrict__ lse,
                                                                      const float* __restrict__ dout,
                                                                      float* __restrict__ dq, float* __restrict__ pd,
                                                                      float* __restrict__ ds) {
  extern __shared__ float sm[];
  const int hdp = a.hd + 1;
  float* Ks = sm;
  float* Vs = Ks + AT_TS * hdp;
  float* Qs = Vs + AT_TS * hdp;                  // [AT_WARPS][hd] scaled q
  float* Gs = Qs + AT_WARPS * a.hd;              // [AT_WARPS][hd] dout
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int row = blockIdx.x * AT_WARPS + warp;
  const bool active = row < a.L;
  const int b = bh / a.heads;
  const float* kb = a.k + (int64_t)bh * a.S * a.hd;
  const float* vb = a.v + (int64_t)bh * a.S * a.hd;
  float delta = 0.f, row_lse = 0.f;
  if (active) {
    const int64_t off = ((int64_t)bh * a.L + row) * a.hd;
    for (int d = lane; d < a.hd; d += 32) {
      Qs[warp * a.hd + d] = a.q[off + d] * a.scale;
      const float g = dout[off + d];
      Gs[warp * a.hd + d] = g;
      delta = fmaf(g, out[off + d], delta);
    }
    delta = warp_sum(delta);
    row_lse = lse[(int64_t)bh * a.L + row];
  }
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;

  for (int s0 = 0; s0 < a.S; s0 += AT_TS) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < AT_TS * a.hd; idx += blockDim.x) {
      const int j = idx / a.hd, d = idx % a.hd;
      const bool ok = s0 + j < a.S;
      Ks[j * hdp + d] = ok ? kb[(int64_t)(s0 + j) * a.hd + d] : 0.f;
      Vs[j * hdp + d] = ok ? vb[(int64_t)(s0 + j) * a.hd + d] : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    float dsv[AT_TS / 32];
#pragma unroll
    for (int t = 0; t < AT_TS / 32; ++t) {
      const int j = lane + 32 * t;
      float sc = 0.f, dp = 0.f;
      for (int d = 0; d < a.hd; ++d) {
        sc = fmaf(Qs[warp * a.hd + d], Ks[j * hdp + d], sc);
        dp = fmaf(Gs[warp * a.hd + d], Vs[j * hdp + d], dp);
      }
      const bool valid = s0 + j < a.S && !(a.mask && a.mask[(int64_t)b * a.S + s0 + j]);
      const float p = valid ? expf(sc - row_lse) : 0.f;
      float keep = 1.f;
      if (a.drop_p > 0.f) keep = dropout_keep(a.seed, bh, row, s0 + j, a.drop_p) ? keep_scale : 0.f;
      // out = sum_j p_j keep_j v_j ; d(out)/d(p_j) = keep_j v_j ; softmax backward with delta = <dout, out>
      dsv[t] = p * (dp * keep - delta);
      if (s0 + j < a.S) {
        const int64_t o2 = ((int64_t)bh * a.L + row) * a.S + s0 + j;
        pd[o2] = p * keep;
        ds[o2] = dsv[t];
      }
    }
#pragma unroll
    for (int t = 0; t < AT_TS / 32; ++t) {
      for (int jj = 0; jj < 32; ++jj) {
        const float dj = __shfl_sync(FULL, dsv[t], jj);
        const int j = jj + 32 * t;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int d = lane + 32 * i;
          if (d < a.hd) acc[i] = fmaf(dj, Ks[j * hdp + d], acc[i]);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int d = lane + 32 * i;
      if (d < a.hd) dq[((int64_t)bh * a.L + row) * a.hd + d] = acc[i] * a.scale;
    }
  }
}

// backward, pass 2 (one warp per key): dk_j = scale * sum_i ds_ij q_i ; dv_j = sum_i pd_ij dout_i
template <int NV>
__global__ void __launch_bounds__(AT_WARPS * 32) attn_bwd_keys_kernel(const AttnArgs a, const float* __restrict__ dout,
                                                                      const float* __restrict__ pd,
                                                                      const float* __restrict__ ds,
                                                                      float* __restrict__ dk, float* __restrict__ dv) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int j = blockIdx.x * AT_WARPS + warp;
  if (j >= a.S) return;
  float ak[NV], av[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { ak[i] = 0.f; av[i] = 0.f; }
  for (int r = 0; r < a.L; ++r) {
    const int64_t o2 = ((int64_t)bh * a.L + r) * a.S + j;
    const float dsv = ds[o2], pv = pd[o2];
    const int64_t off = ((int64_t)bh * a.L + r) * a.hd;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int d = lane + 32 * i;
      if (d < a.hd) {
        ak[i] = fmaf(dsv, a.q[off + d], ak[i]);
        av[i] = fmaf(pv, dout[off + d], av[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int d = lane + 32 * i;
    if (d < a.hd) {
      dk[((int64_t)bh * a.S + j) * a.hd + d] = ak[i] * a.scale;
      dv[((int64_t)bh * a.S + j) * a.hd + d] = av[i];
    }
  }
}

static int fill(AttnArgs& a, const float* q, const float* k, const float* v, const unsigned char* mask, int B,
                int heads, int L, int S, int hd, float scale, float drop_p, uint64_t seed) {
  HSG_REQUIRE(B > 0 && heads > 0 && L > 0 && S > 0 && hd > 0, HSG_E_INVALID, "mha: bad shape");
  HSG_REQUIRE(hd <= AT_HD_MAX, HSG_E_UNSUPPORTED, "mha: head dim %d (max %d)", hd, AT_HD_MAX);
  HSG_REQUIRE((int64_t)B * heads <= 65535, HSG_E_UNSUPPORTED, "mha: batch*heads %lld (max 65535)", (long long)B * heads);
  HSG_REQUIRE(drop_p >= 0.f && drop_p < 1.f, HSG_E_INVALID, "mha: dropout %f", drop_p);
  HSG_REQUIRE(q && k && v, HSG_E_INVALID, "mha: null pointer");
  a.q = q; a.k = k; a.v = v; a.mask = mask; a.BH = B * heads; a.heads = heads; a.L = L; a.S = S; a.hd = hd;
  a.scale = scale; a.drop_p = drop_p; a.seed = seed;
  return HSG_OK;
}

}  // namespace hsg

using namespace hsg;

extern "C" {

size_t hsg_mha_workspace_bytes(int B, int heads, int L, int S) {
  return (size_t)2 * B * heads * L * S * sizeof(float) + 256;
}

int hsg_mha_fwd_f32(const float* q, const float* k, const float* v, const unsigned char* key_padding_mask,
                    int B, int heads, int L, int S, int hd, float scale, float dropout_p,
                    unsigned long long seed, float* out, float* lse, void* stream) {
  AttnArgs a;
  int rc = fill(a, q, k, v, key_padding_mask, B, heads, L, S, hd, scale, dropout_p, seed);
  if (rc) return rc;
  HSG_REQUIRE(out && lse, HSG_E_INVALID, "mha_fwd: null output");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)(2 * AT_TS * (hd + 1) + AT_WARPS * hd) * sizeof(float);
  dim3 grid((L + AT_WARPS - 1) / AT_WARPS, a.BH);
  const int nv = (hd + 31) / 32;
#define LAUNCH_FWD(NV)                                                                                      \
  do {                                                                                                      \
    if (smem > 48 * 1024) HSG_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    attn_fwd_kernel<NV><<<grid, AT_WARPS * 32, smem, st>>>(a, out, lse);                                    \
  } while (0)
  if (nv <= 1) LAUNCH_FWD(1); else if (nv <= 2) LAUNCH_FWD(2); else LAUNCH_FWD(4);
#undef LAUNCH_FWD
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int hsg_mha_bwd_f32(const float* q, const float* k, const float* v, const unsigned char* key_padding_mask,
                    int B, int heads, int L, int S, int hd, float scale, float dropout_p,
                    unsigned long long seed, const float* out, const float* lse, const float* dout,
                    float* dq, float* dk, float* dv, void* workspace, size_t workspace_bytes, void* stream) {
  AttnArgs a;
  int rc = fill(a, q, k, v, key_padding_mask, B, heads, L, S, hd, scale, dropout_p, seed);
  if (rc) return rc;
  HSG_REQUIRE(out && lse && dout && dq && dk && dv, HSG_E_INVALID, "mha_bwd: null pointer");
  HSG_REQUIRE(workspace && workspace_bytes >= hsg_mha_workspace_bytes(B, heads, L, S), HSG_E_WORKSPACE,
              "mha_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* pd = (float*)workspace;
  float* ds = pd + (size_t)a.BH * L * S;
  const size_t smem = (size_t)(2 * AT_TS * (hd + 1) + 2 * AT_WARPS * hd) * sizeof(float);
  dim3 g1((L + AT_WARPS - 1) / AT_WARPS, a.BH), g2((S + AT_WARPS - 1) / AT_WARPS, a.BH);
  const int nv = (hd + 31) / 32;
#define LAUNCH_BWD(NV)                                                                                      \
  do {                                                                                                      \
    if (smem > 48 * 1024) HSG_CUDA(cudaFuncSetAttribute(attn_bwd_rows_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    attn_bwd_rows_kernel<NV><<<g1, AT_WARPS * 32, smem, st>>>(a, out, lse, dout, dq, pd, ds);               \
    HSG_LAUNCH_CHECK();                                                                                     \
    attn_bwd_keys_kernel<NV><<<g2, AT_WARPS * 32, 0, st>>>(a, dout, pd, ds, dk, dv);                        \
  } while (0)
  if (nv <= 1) LAUNCH_BWD(1); else if (nv <= 2) LAUNCH_BWD(2); else LAUNCH_BWD(4);
#undef LAUNCH_BWD
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

}  // extern "C"
