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

#include <algorithm>

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

// Work split.  A warp owns R query rows (R = 4, or 1 when L is tiny); a lane owns two keys of the
// 64-key tile in the score phase and NV output dims in the value phase.  Every K/V element read
// from shared memory (128-bit loads; rows padded to hd+4 floats) is used for R rows, and the
// probabilities go through a per-warp shared tile so the value phase reads them as broadcast
// float4s: ~0.3 shared-memory instructions per FMA instead of the 1.5 of a row-per-warp layout.
constexpr int AT_PAD = 4;

__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}

// stage one 64-key tile of K and V (rows padded to hd + AT_PAD floats; hd % 4 == 0)
__device__ __forceinline__ void load_kv_tile(const AttnArgs& a, const float* __restrict__ kb, const float* __restrict__ vb,
                                             int s0, float* Ks, float* Vs) {
  const int hdp = a.hd + AT_PAD, q4 = a.hd >> 2;
  for (int idx = threadIdx.x; idx < AT_TS * q4; idx += blockDim.x) {
    const int j = idx / q4, d = (idx - j * q4) << 2;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
    if (s0 + j < a.S) {
      kv = *reinterpret_cast<const float4*>(kb + (int64_t)(s0 + j) * a.hd + d);
      vv = *reinterpret_cast<const float4*>(vb + (int64_t)(s0 + j) * a.hd + d);
    }
    *reinterpret_cast<float4*>(Ks + j * hdp + d) = kv;
    *reinterpret_cast<float4*>(Vs + j * hdp + d) = vv;
  }
}

// forward: out [BH,L,hd], lse [BH,L] (log-sum-exp of the scaled, masked scores)
template <int NV, int R>   // NV = ceil(hd / 32): output dims per lane; R query rows per warp
__global__ void __launch_bounds__(AT_WARPS * 32) attn_fwd_kernel(const AttnArgs a, float* __restrict__ out,
                                                                 float* __restrict__ lse) {
  extern __shared__ __align__(16) float sm[];
  const int hdp = a.hd + AT_PAD;
  float* Ks = sm;                              // [AT_TS][hdp]
  float* Vs = Ks + AT_TS * hdp;                // [AT_TS][hdp]
  float* Qs = Vs + AT_TS * hdp;                // [AT_WARPS*R][hd]   scaled q
  float* Ps = Qs + AT_WARPS * R * a.hd;        // [AT_WARPS*R][AT_TS] probabilities of the current tile
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int row0 = (blockIdx.x * AT_WARPS + warp) * R;
  const int b = bh / a.heads;
  const float* kb = a.k + (int64_t)bh * a.S * a.hd;
  const float* vb = a.v + (int64_t)bh * a.S * a.hd;
#pragma unroll
  for (int r = 0; r < R; ++r)
    for (int d = lane; d < a.hd; d += 32)
      Qs[(warp * R + r) * a.hd + d] = row0 + r < a.L ? a.q[((int64_t)bh * a.L + row0 + r) * a.hd + d] * a.scale : 0.f;
  float m[R], l[R], o[R][NV];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) o[r][i] = 0.f;
  }
  const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  const bool active = row0 < a.L;

  for (int s0 = 0; s0 < a.S; s0 += AT_TS) {
    __syncthreads();
    load_kv_tile(a, kb, vb, s0, Ks, Vs);
    __syncthreads();
    if (!active) continue;
    // scores of this lane's two keys, for the R rows of the warp
    float sc[R][AT_TS / 32];
#pragma unroll
    for (int t = 0; t < AT_TS / 32; ++t) {
      const int j = lane + 32 * t;
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = 0.f;
      for (int d = 0; d < a.hd; d += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(Ks + j * hdp + d);
#pragma unroll
        for (int r = 0; r < R; ++r)
          acc[r] = dot4(*reinterpret_cast<const float4*>(Qs + (warp * R + r) * a.hd + d), k4, acc[r]);
      }
      const bool valid = s0 + j < a.S && !(a.mask && a.mask[(int64_t)b * a.S + s0 + j]);
#pragma unroll
      for (int r = 0; r < R; ++r) sc[r][t] = valid ? acc[r] : -INFINITY;
    }
    __syncwarp();                                // the previous tile's Ps reads are done
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float tmax = sc[r][0];
#pragma unroll
      for (int t = 1; t < AT_TS / 32; ++t) tmax = fmaxf(tmax, sc[r][t]);
      tmax = warp_max(tmax);
      const float m_new = fmaxf(m[r], tmax);
      float* prow = Ps + (warp * R + r) * AT_TS;
      if (m_new == -INFINITY) {                  // everything masked so far (warp-uniform)
#pragma unroll
        for (int t = 0; t < AT_TS / 32; ++t) prow[lane + 32 * t] = 0.f;
        continue;
      }
      const float corr = __expf(m[r] - m_new);   // exp(-inf) = 0 on the first live tile
      float psum = 0.f;
#pragma unroll
      for (int t = 0; t < AT_TS / 32; ++t) {
        float pj = expf(sc[r][t] - m_new);
        psum += pj;
        if (a.drop_p > 0.f) pj = dropout_keep(a.seed, bh, row0 + r, s0 + lane + 32 * t, a.drop_p) ? pj * keep_scale : 0.f;
        prow[lane + 32 * t] = pj;
      }
      l[r] = l[r] * corr + warp_sum(psum);
#pragma unroll
      for (int i = 0; i < NV; ++i) o[r][i] *= corr;
      m[r] = m_new;
    }
    __syncwarp();
    for (int j0 = 0; j0 < AT_TS; j0 += 4) {
      float vv[4][NV];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int d = lane + 32 * i;
          vv[k][i] = d < a.hd ? Vs[(j0 + k) * hdp + d] : 0.f;
        }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 p4 = *reinterpret_cast<const float4*>(Ps + (warp * R + r) * AT_TS + j0);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          o[r][i] = fmaf(p4.x, vv[0][i], o[r][i]);
          o[r][i] = fmaf(p4.y, vv[1][i], o[r][i]);
          o[r][i] = fmaf(p4.z, vv[2][i], o[r][i]);
          o[r][i] = fmaf(p4.w, vv[3][i], o[r][i]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (row0 + r >= a.L) continue;
    const float inv = 1.f / l[r];                // l == 0 (all keys masked) -> NaN, like the reference
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int d = lane + 32 * i;
      if (d < a.hd) out[((int64_t)bh * a.L + row0 + r) * a.hd + d] = o[r][i] * inv;
    }
    if (lane == 0) lse[(int64_t)bh * a.L + row0 + r] = m[r] + logf(l[r]);
  }
}

// backward, pass 1 (R query rows per warp): recompute p, write
//   pd[bh,row,j] = dropped/scaled probability, ds[bh,row,j] = p * (dp - delta)
// and dq = scale * ds K.
template <int NV, int R>
__global__ void __launch_bounds__(AT_WARPS * 32) attn_bwd_rows_kernel(const AttnArgs a, const float* __restrict__ out,
                                                                      const float* __restrict__ lse,
                                                                      const float* __restrict__ dout,
                                                                      float* __restrict__ dq, float* __restrict__ pd,
                                                                      float* __restrict__ ds) {
  extern __shared__ __align__(16) float sm[];
  const int hdp = a.hd + AT_PAD;
  float* Ks = sm;
  float* Vs = Ks + AT_TS * hdp;
  float* Qs = Vs + AT_TS * hdp;                  // [AT_WARPS*R][hd] scaled q
  float* Gs = Qs + AT_WARPS * R * a.hd;          // [AT_WARPS*R][hd] dout
  float* Ds = Gs + AT_WARPS * R * a.hd;          // [AT_WARPS*R][AT_TS] ds of the current tile
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int row0 = (blockIdx.x * AT_WARPS + warp) * R;
  const bool active = row0 < a.L;
  const int b = bh / a.heads;
  const float* kb = a.k + (int64_t)bh * a.S * a.hd;
  const float* vb = a.v + (int64_t)bh * a.S * a.hd;
  float delta[R], row_lse[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const bool in = row0 + r < a.L;
    const int64_t off = ((int64_t)bh * a.L + row0 + r) * a.hd;
    float dl = 0.f;
    for (int d = lane; d < a.hd; d += 32) {
      const float g = in ? dout[off + d] : 0.f;
      Qs[(warp * R + r) * a.hd + d] = in ? a.q[off + d] * a.scale : 0.f;
      Gs[(warp * R + r) * a.hd + d] = g;
      if (in) dl = fmaf(g, out[off + d], dl);
    }
    delta[r] = warp_sum(dl);
    row_lse[r] = in ? lse[(int64_t)bh * a.L + row0 + r] : 0.f;
  }
  float acc[R][NV];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[r][i] = 0.f;
  const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;

  for (int s0 = 0; s0 < a.S; s0 += AT_TS) {
    __syncthreads();
    load_kv_tile(a, kb, vb, s0, Ks, Vs);
    __syncthreads();
    if (!active) continue;
    __syncwarp();
#pragma unroll
    for (int t = 0; t < AT_TS / 32; ++t) {
      const int j = lane + 32 * t;
      float sc[R], dp[R];
#pragma unroll
      for (int r = 0; r < R; ++r) { sc[r] = 0.f; dp[r] = 0.f; }
      for (int d = 0; d < a.hd; d += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(Ks + j * hdp + d);
        const float4 v4 = *reinterpret_cast<const float4*>(Vs + j * hdp + d);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          sc[r] = dot4(*reinterpret_cast<const float4*>(Qs + (warp * R + r) * a.hd + d), k4, sc[r]);
          dp[r] = dot4(*reinterpret_cast<const float4*>(Gs + (warp * R + r) * a.hd + d), v4, dp[r]);
        }
      }
      const bool valid = s0 + j < a.S && !(a.mask && a.mask[(int64_t)b * a.S + s0 + j]);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float p = valid ? expf(sc[r] - row_lse[r]) : 0.f;
        float keep = 1.f;
        if (a.drop_p > 0.f) keep = dropout_keep(a.seed, bh, row0 + r, s0 + j, a.drop_p) ? keep_scale : 0.f;
        // out = sum_j p_j keep_j v_j ; d(out)/d(p_j) = keep_j v_j ; softmax backward with delta = <dout, out>
        const float dsv = p * (dp[r] * keep - delta[r]);
        Ds[(warp * R + r) * AT_TS + j] = dsv;
        if (s0 + j < a.S && row0 + r < a.L) {
          const int64_t o2 = ((int64_t)bh * a.L + row0 + r) * a.S + s0 + j;
          pd[o2] = p * keep;
          ds[o2] = dsv;
        }
      }
    }
    __syncwarp();
    for (int j0 = 0; j0 < AT_TS; j0 += 4) {
      float kk[4][NV];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int d = lane + 32 * i;
          kk[k][i] = d < a.hd ? Ks[(j0 + k) * hdp + d] : 0.f;
        }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 d4 = *reinterpret_cast<const float4*>(Ds + (warp * R + r) * AT_TS + j0);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          acc[r][i] = fmaf(d4.x, kk[0][i], acc[r][i]);
          acc[r][i] = fmaf(d4.y, kk[1][i], acc[r][i]);
          acc[r][i] = fmaf(d4.z, kk[2][i], acc[r][i]);
          acc[r][i] = fmaf(d4.w, kk[3][i], acc[r][i]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (row0 + r >= a.L) continue;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int d = lane + 32 * i;
      if (d < a.hd) dq[((int64_t)bh * a.L + row0 + r) * a.hd + d] = acc[r][i] * a.scale;
    }
  }
}

// backward, pass 2 (four keys per warp): dk_j = scale * sum_i ds_ij q_i ; dv_j = sum_i pd_ij dout_i.
// The query rows go by in shared-memory tiles; a row's ds / pd of the warp's four keys is one
// 128-bit load each (S % 4 == 0 on this path; otherwise scalar loads).
template <int NV>
__global__ void __launch_bounds__(AT_WARPS * 32) attn_bwd_keys_kernel(const AttnArgs a, const float* __restrict__ dout,
                                                                      const float* __restrict__ pd,
                                                                      const float* __restrict__ ds,
                                                                      float* __restrict__ dk, float* __restrict__ dv) {
  extern __shared__ __align__(16) float sm[];
  float* Qt = sm;                                // [AT_TS][hd]
  float* Gt = Qt + AT_TS * a.hd;                 // [AT_TS][hd]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int j0 = (blockIdx.x * AT_WARPS + warp) * 4;
  const bool vec = (a.S & 3) == 0;
  float ak[4][NV], av[4][NV];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < NV; ++i) { ak[k][i] = 0.f; av[k][i] = 0.f; }
  for (int r0 = 0; r0 < a.L; r0 += AT_TS) {
    __syncthreads();
    const int q4 = a.hd >> 2;
    for (int idx = threadIdx.x; idx < AT_TS * q4; idx += blockDim.x) {
      const int r = idx / q4, d = (idx - r * q4) << 2;
      float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), gv = qv;
      if (r0 + r < a.L) {
        const int64_t off = ((int64_t)bh * a.L + r0 + r) * a.hd + d;
        qv = *reinterpret_cast<const float4*>(a.q + off);
        gv = *reinterpret_cast<const float4*>(dout + off);
      }
      *reinterpret_cast<float4*>(Qt + r * a.hd + d) = qv;
      *reinterpret_cast<float4*>(Gt + r * a.hd + d) = gv;
    }
    __syncthreads();
    if (j0 >= a.S) continue;
    const int nr = min(AT_TS, a.L - r0);
    for (int r = 0; r < nr; ++r) {
      const int64_t o2 = ((int64_t)bh * a.L + r0 + r) * a.S + j0;
      float dsv[4], pv[4];
      if (vec) {
        const float4 d4 = *reinterpret_cast<const float4*>(ds + o2), p4 = *reinterpret_cast<const float4*>(pd + o2);
        dsv[0] = d4.x; dsv[1] = d4.y; dsv[2] = d4.z; dsv[3] = d4.w;
        pv[0] = p4.x; pv[1] = p4.y; pv[2] = p4.z; pv[3] = p4.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          dsv[k] = j0 + k < a.S ? ds[o2 + k] : 0.f;
          pv[k] = j0 + k < a.S ? pd[o2 + k] : 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int d = lane + 32 * i;
        const float qd = d < a.hd ? Qt[r * a.hd + d] : 0.f;
        const float gd = d < a.hd ? Gt[r * a.hd + d] : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ak[k][i] = fmaf(dsv[k], qd, ak[k][i]);
          av[k][i] = fmaf(pv[k], gd, av[k][i]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (j0 + k >= a.S) continue;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int d = lane + 32 * i;
      if (d < a.hd) {
        dk[((int64_t)bh * a.S + j0 + k) * a.hd + d] = ak[k][i] * a.scale;
        dv[((int64_t)bh * a.S + j0 + k) * a.hd + d] = av[k][i];
      }
    }
  }
}

static int fill(AttnArgs& a, const float* q, const float* k, const float* v, const unsigned char* mask, int B,
                int heads, int L, int S, int hd, float scale, float drop_p, uint64_t seed) {
  HSG_REQUIRE(B > 0 && heads > 0 && L > 0 && S > 0 && hd > 0, HSG_E_INVALID, "mha: bad shape");
  HSG_REQUIRE(hd <= AT_HD_MAX && hd % 4 == 0, HSG_E_UNSUPPORTED, "mha: head dim %d (multiple of 4, max %d)", hd, AT_HD_MAX);
  HSG_REQUIRE((int64_t)B * heads <= 65535, HSG_E_UNSUPPORTED, "mha: batch*heads %lld (max 65535)", (long long)B * heads);
  HSG_REQUIRE(drop_p >= 0.f && drop_p < 1.f, HSG_E_INVALID, "mha: dropout %f", drop_p);
  HSG_REQUIRE(q && k && v, HSG_E_INVALID, "mha: null pointer");
  a.q = q; a.k = k; a.v = v; a.mask = mask; a.BH = B * heads; a.heads = heads; a.L = L; a.S = S; a.hd = hd;
  a.scale = scale; a.drop_p = drop_p; a.seed = seed;
  return HSG_OK;
}

// tensor-core forward (attention_tc.cu)
bool attn_tc_supported(int B, int heads, int L, int S, int hd);
bool attn_tc_profitable(int B, int heads, int L, int S, int hd);
size_t attn_tc_workspace_bytes(int B, int heads, int L, int S);
int attn_fwd_tc(const float* q, const float* k, const float* v, const unsigned char* mask, int B, int heads, int L,
                int S, float scale, float drop_p, uint64_t seed, float* out, float* lse, void* workspace,
                cudaStream_t st);
// tensor-core backward (attention_bwd_tc.cu)
bool attn_bwd_tc_supported(int B, int heads, int L, int S, int hd);
bool attn_bwd_tc_profitable(int B, int heads, int L, int S, int hd);
size_t attn_bwd_tc_workspace_bytes(int B, int heads, int L, int S);
int attn_bwd_tc(const float* q, const float* k, const float* v, const unsigned char* mask, int B, int heads, int L,
                int S, float scale, float drop_p, uint64_t seed, const float* out, const float* lse, const float* dout,
                float* dq, float* dk, float* dv, void* workspace, cudaStream_t st);
extern int g_debug_flags;      // nce.cu; tests: bit 4 (16) keeps the attention forward on the CUDA-core kernel,
                               // bit 5 (32) takes the tensor-core kernel for every shape it supports;
                               // bits 6 (64) / 7 (128): the same two switches for the backward

}  // namespace hsg

using namespace hsg;

extern "C" {

static size_t mha_bwd_simt_bytes(int B, int heads, int L, int S) {
  return (size_t)2 * B * heads * L * S * sizeof(float) + 256;
}

// enough for either backward (the entry point has no head dim: the tensor-core layout is counted whenever L, S fit it)
size_t hsg_mha_workspace_bytes(int B, int heads, int L, int S) {
  size_t need = mha_bwd_simt_bytes(B, heads, L, S);
  if (attn_bwd_tc_supported(B, heads, L, S, 64)) need = std::max(need, attn_bwd_tc_workspace_bytes(B, heads, L, S));
  return need;
}

size_t hsg_mha_fwd_workspace_bytes(int B, int heads, int L, int S, int hd) {
  const bool use = (g_debug_flags & 32) ? attn_tc_supported(B, heads, L, S, hd) : attn_tc_profitable(B, heads, L, S, hd);
  return use && !(g_debug_flags & 16) ? attn_tc_workspace_bytes(B, heads, L, S) : 0;
}

int hsg_mha_fwd_f32(const float* q, const float* k, const float* v, const unsigned char* key_padding_mask,
                    int B, int heads, int L, int S, int hd, float scale, float dropout_p,
                    unsigned long long seed, float* out, float* lse, void* workspace, size_t workspace_bytes,
                    void* stream) {
  AttnArgs a;
  int rc = fill(a, q, k, v, key_padding_mask, B, heads, L, S, hd, scale, dropout_p, seed);
  if (rc) return rc;
  HSG_REQUIRE(out && lse, HSG_E_INVALID, "mha_fwd: null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (!(g_debug_flags & 16) && attn_tc_supported(B, heads, L, S, hd) && workspace &&
      workspace_bytes >= attn_tc_workspace_bytes(B, heads, L, S))
    return attn_fwd_tc(q, k, v, key_padding_mask, B, heads, L, S, scale, dropout_p, seed, out, lse, workspace, st);
  const int R = L >= 64 ? 4 : 1;                   // rows per warp (tiny L: keep the CTAs many)
  const size_t smem = (size_t)(2 * AT_TS * (hd + AT_PAD) + AT_WARPS * R * (hd + AT_TS)) * sizeof(float);
  dim3 grid((L + AT_WARPS * R - 1) / (AT_WARPS * R), a.BH);
  const int nv = (hd + 31) / 32;
#define LAUNCH_FWD(NV, RR)                                                                                  \
  do {                                                                                                      \
    if (smem > 48 * 1024) HSG_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<NV, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    attn_fwd_kernel<NV, RR><<<grid, AT_WARPS * 32, smem, st>>>(a, out, lse);                                \
  } while (0)
  if (R == 4) { if (nv <= 1) LAUNCH_FWD(1, 4); else if (nv <= 2) LAUNCH_FWD(2, 4); else LAUNCH_FWD(4, 4); }
  else { if (nv <= 1) LAUNCH_FWD(1, 1); else if (nv <= 2) LAUNCH_FWD(2, 1); else LAUNCH_FWD(4, 1); }
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
  HSG_REQUIRE(workspace && workspace_bytes >= mha_bwd_simt_bytes(B, heads, L, S), HSG_E_WORKSPACE,
              "mha_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc = (g_debug_flags & 128) ? attn_bwd_tc_supported(B, heads, L, S, hd) : attn_bwd_tc_profitable(B, heads, L, S, hd);
  if (tc && !(g_debug_flags & 64) && workspace_bytes >= attn_bwd_tc_workspace_bytes(B, heads, L, S))
    return attn_bwd_tc(q, k, v, key_padding_mask, B, heads, L, S, scale, dropout_p, seed, out, lse, dout, dq, dk, dv,
                       workspace, st);
  float* pd = (float*)workspace;
  float* ds = pd + (size_t)a.BH * L * S;
  const int R = L >= 64 ? 4 : 1;
  const size_t smem = (size_t)(2 * AT_TS * (hd + AT_PAD) + AT_WARPS * R * (2 * hd + AT_TS)) * sizeof(float);
  const size_t smem2 = (size_t)2 * AT_TS * hd * sizeof(float);
  dim3 g1((L + AT_WARPS * R - 1) / (AT_WARPS * R), a.BH), g2((S + AT_WARPS * 4 - 1) / (AT_WARPS * 4), a.BH);
  const int nv = (hd + 31) / 32;
#define LAUNCH_BWD(NV, RR)                                                                                  \
  do {                                                                                                      \
    if (smem > 48 * 1024) HSG_CUDA(cudaFuncSetAttribute(attn_bwd_rows_kernel<NV, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    attn_bwd_rows_kernel<NV, RR><<<g1, AT_WARPS * 32, smem, st>>>(a, out, lse, dout, dq, pd, ds);           \
    HSG_LAUNCH_CHECK();                                                                                     \
    if (smem2 > 48 * 1024) HSG_CUDA(cudaFuncSetAttribute(attn_bwd_keys_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
    attn_bwd_keys_kernel<NV><<<g2, AT_WARPS * 32, smem2, st>>>(a, dout, pd, ds, dk, dv);                    \
  } while (0)
  if (R == 4) { if (nv <= 1) LAUNCH_BWD(1, 4); else if (nv <= 2) LAUNCH_BWD(2, 4); else LAUNCH_BWD(4, 4); }
  else { if (nv <= 1) LAUNCH_BWD(1, 1); else if (nv <= 2) LAUNCH_BWD(2, 1); else LAUNCH_BWD(4, 1); }
#undef LAUNCH_BWD
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

}  // extern "C"
