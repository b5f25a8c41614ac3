// K5 on tensor cores: forward of the clustering transformer's attention core (hd = 64, S <= 256).
//
// Reference: nn.MultiheadAttention's need_weights=True slow path (hsg/models/heads/transformer.py:235,
// 300,304): q/sqrt(hd), baddbmm with the -inf key-padding mask, softmax, dropout, bmm with v.
// One CTA per (batch*head, 128 query rows):
//   TMA      : Q tile [128 x 64], all keys K [S x 64], then V^T [64 x S]; every operand pre-split into
//              fp16 (hi | lo) halves (three tcgen05 passes per contraction: fp32-grade, as in nce_tc.cu)
//   tcgen05  : S = Q K^T  -> TMEM [128 x S] (q pre-scaled by log2(e)/sqrt(hd): scores in the base-2 domain)
//   softmax  : a thread per query row sweeps its TMEM lane twice (max, then exp2 / sum), applies the
//              key-padding mask and dropout, and writes the probabilities as fp16 (hi, lo) straight into
//              shared memory in the 128B-swizzled K-major layout the tensor core reads (over the dead Q/K tiles)
//   tcgen05  : O = P V     -> TMEM [128 x 64]
//   epilogue : O / row sum -> out, log-sum-exp -> lse (the CUDA-core backward reuses both)
// The [L,S] score matrix never exists in memory.
#include "attention_tc.cuh"

#include <float.h>

#include <algorithm>

namespace hsg {

constexpr int AC_THREADS = 192;           // warp0: TMA + MMA issue, warp1: TMEM allocation, warps 2-5: softmax / epilogue

struct AttnTcParams {
  int BH, heads, L, S, Sp;        // Sp = S rounded up to 64
  const unsigned char* mask;      // [B,S] or NULL
  float drop_p;
  uint64_t seed;
  float* out;                     // [BH,L,64]
  float* lse;                     // [BH,L]
};

__global__ void __launch_bounds__(AC_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_vt, const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nks = p.Sp / 64;                                   // key slabs
  // phase 1: Q hi, Q lo (2 x 16 KiB), K hi, K lo (2 x Sp*128 B).  phase 2: P hi / P lo slabs reuse that space.
  const uint32_t sQ = base;
  const uint32_t sK = sQ + 2 * AC_SLAB;
  const uint32_t k_bytes = (uint32_t)p.Sp * 128u;
  const uint32_t sP = base;                                    // [2][nks] slabs of 16 KiB: P hi then P lo
  const uint32_t p_bytes = 2u * nks * AC_SLAB;
  const uint32_t region1 = max(2u * AC_SLAB + 2u * k_bytes, p_bytes);
  const uint32_t sV = base + region1;                          // V^T hi, lo: [2][nks] slabs of [64 x 64] = 8 KiB
  const uint32_t v_slab = AC_HD * 64 * 2;
  uint8_t* misc = smem_raw + (sV + 2u * nks * v_slab - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  uint32_t* live_bits = tmem_slot + 2;                         // [8] bit k of word w: key 32w+k takes part
  const uint32_t bar_qk = smem_u32(bars), bar_v = bar_qk + 8, bar_s = bar_qk + 16, bar_p = bar_qk + 24, bar_o = bar_qk + 32;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int row0 = blockIdx.x * AC_BM;
  if (warp == 0 && lane == 0) {
    mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 4); mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // TMEM: Sp columns of scores + 64 of output, as a power of two (short key sets leave room for more CTAs per SM)
  const uint32_t ncols = p.Sp + AC_HD <= 128 ? 128u : p.Sp + AC_HD <= 256 ? 256u : 512u;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_s = tmem_base, t_o = tmem_base + (uint32_t)p.Sp;

  if (warp == 0) {
    if (lane == 0) {
      // ---- loads
      mbar_expect_tx(bar_qk, 2 * AC_SLAB + 2 * k_bytes);
      tma_load_2d(sQ, &tmap_q, 0, bh * p.L + row0, bar_qk);
      tma_load_2d(sQ + AC_SLAB, &tmap_q, AC_HD, bh * p.L + row0, bar_qk);
      tma_load_2d(sK, &tmap_k, 0, bh * p.S, bar_qk);
      tma_load_2d(sK + k_bytes, &tmap_k, AC_HD, bh * p.S, bar_qk);
      mbar_expect_tx(bar_v, 2 * nks * v_slab);
      for (int ks = 0; ks < nks; ++ks) {
        tma_load_2d(sV + ks * v_slab, &tmap_vt, ks * 64, bh * AC_HD, bar_v);
        tma_load_2d(sV + (nks + ks) * v_slab, &tmap_vt, p.Sp + ks * 64, bh * AC_HD, bar_v);
      }
      // ---- S = Q K^T (three passes)
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Sp >> 3) << 17) | ((uint32_t)(AC_BM >> 4) << 24);
        const uint64_t qh = umma_desc(sQ, 1024, 2), ql = umma_desc(sQ + AC_SLAB, 1024, 2);
        const uint64_t kh = umma_desc(sK, 1024, 2), kl = umma_desc(sK + k_bytes, 1024, 2);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16(t_s, qh + 2 * k4, kh + 2 * k4, idesc, k4 ? 1u : 0u);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16(t_s, ql + 2 * k4, kh + 2 * k4, idesc, 1u);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16(t_s, qh + 2 * k4, kl + 2 * k4, idesc, 1u);
        tc_commit(bar_s);
      }
      // ---- O = P V (three passes) once the probabilities are in shared memory
      mbar_wait(bar_p, 0);
      mbar_wait(bar_v, 0);
      tc_fence_after();
      {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(AC_HD >> 3) << 17) | ((uint32_t)(AC_BM >> 4) << 24);
        uint32_t first = 1;
        for (int ks = 0; ks < nks; ++ks) {
          const uint64_t ph = umma_desc(sP + ks * AC_SLAB, 1024, 2), pl = umma_desc(sP + (nks + ks) * AC_SLAB, 1024, 2);
          const uint64_t vh = umma_desc(sV + ks * v_slab, 1024, 2), vl = umma_desc(sV + (nks + ks) * v_slab, 1024, 2);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) { tc_mma_f16(t_o, ph + 2 * k4, vh + 2 * k4, idesc, first ? 0u : 1u); first = 0; }
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16(t_o, pl + 2 * k4, vh + 2 * k4, idesc, 1u);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16(t_o, ph + 2 * k4, vl + 2 * k4, idesc, 1u);
        }
        tc_commit(bar_o);
      }
    }
  } else if (warp >= 2) {
    // ===================== softmax / epilogue: one query row per thread =====================
    const int q = warp & 3;
    const int r = 32 * q + lane;
    const int row = row0 + r;
    const bool inb = row < p.L;
    const int b = bh / p.heads;
    const unsigned char* mrow = p.mask ? p.mask + (int64_t)b * p.S : nullptr;
    const uint32_t trow = ((uint32_t)(32 * q) << 16);
    const int nchunk = p.Sp / 16;
    // which keys are live (inside S and not padded): 256 bits, built once per CTA
    {
      const int t = threadIdx.x - 64;                            // 0..127
#pragma unroll
      for (int rep = 0; rep < 2; ++rep) {
        const int key = t + 128 * rep;
        const bool lv = key < p.S && !(mrow && mrow[key]);
        const unsigned bal = __ballot_sync(FULL, lv);
        if (lane == 0) live_bits[key >> 5] = bal;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    uint32_t lw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) lw[i] = live_bits[i];
    mbar_wait(bar_s, 0);
    tc_fence_after();
    // sweep 1: row maximum over the live keys
    float mx = -INFINITY;
    for (int c = 0; c < nchunk; ++c) {
      uint32_t v[16];
      tc_ld16(t_s + trow + c * 16, v);
      tc_ld_wait();
#pragma unroll
      const uint32_t bits = lw[c >> 1] >> ((c & 1) * 16);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if ((bits >> j) & 1u) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    // sweep 2: probabilities -> shared memory (fp16 hi / lo, K-major, 128B swizzle), row sum
    const float keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    float sum = 0.f;
    uint8_t* sp = smem_raw + (sP - smem_u32(smem_raw));
    for (int c = 0; c < nchunk; ++c) {
      uint32_t v[16];
      tc_ld16(t_s + trow + c * 16, v);
      tc_ld_wait();
      __align__(16) __half hi[16], lo[16];
#pragma unroll
      const uint32_t bits = lw[c >> 1] >> ((c & 1) * 16);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int key = c * 16 + j;
        const bool live = (bits >> j) & 1u;
        float pj = live ? ex2f(__uint_as_float(v[j]) - mx) : 0.f;      // mx = -inf (nothing live): NaN, like the reference
        sum += pj;
        if (p.drop_p > 0.f) pj = attn_dropout_keep(p.seed, bh, row, key, p.drop_p) ? pj * keep_scale : 0.f;
        hi[j] = __float2half_rn(pj);
        lo[j] = __float2half_rn(pj - __half2float(hi[j]));
      }
      // element (r, key) of slab key/64: byte (r/8)*1024 + (r%8)*128 + ((chunk ^ (r%8)) * 16), chunk = (key%64)/8
      const int ks = c >> 2;
      const uint32_t rbase = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
#pragma unroll
      for (int half8 = 0; half8 < 2; ++half8) {
        const int chunk = ((c & 3) << 1) + half8;
        const uint32_t off = rbase + (uint32_t)((chunk ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(sp + ks * AC_SLAB + off) = *reinterpret_cast<const uint4*>(hi + 8 * half8);
        *reinterpret_cast<uint4*>(sp + (nks + ks) * AC_SLAB + off) = *reinterpret_cast<const uint4*>(lo + 8 * half8);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the tensor core
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    // epilogue
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = 1.f / sum;
#pragma unroll 1
    for (int c = 0; c < AC_HD / 16; ++c) {
      uint32_t v[16];
      tc_ld16(t_o + trow + c * 16, v);
      tc_ld_wait();
      if (inb) {
        float4* dst = reinterpret_cast<float4*>(p.out + ((int64_t)bh * p.L + row) * AC_HD + c * 16);
#pragma unroll
        for (int w = 0; w < 4; ++w)
          dst[w] = make_float4(__uint_as_float(v[4 * w]) * inv, __uint_as_float(v[4 * w + 1]) * inv,
                               __uint_as_float(v[4 * w + 2]) * inv, __uint_as_float(v[4 * w + 3]) * inv);
      }
    }
    if (inb) p.lse[(int64_t)bh * p.L + row] = (mx + log2f(sum)) * 0.6931471805599453f;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols));
  }
}

// ---------------------------------------------------------------- operand preparation
// rows [R, 64] fp32 -> [R, 128] fp16 (hi | lo) of mul * x
__global__ void __launch_bounds__(256) attn_split_rows_kernel(const float* __restrict__ src, int64_t R, float mul,
                                                              const float* __restrict__ amax, __half* __restrict__ dst) {
  mul = attn_mul(mul, amax);
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= R * AC_HD) return;
  const int64_t r = i / AC_HD;
  const int d = (int)(i - r * AC_HD);
  const float4 x = *reinterpret_cast<const float4*>(src + i);
  const float v[4] = {x.x * mul, x.y * mul, x.z * mul, x.w * mul};
  __align__(8) __half hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    hi[j] = __float2half_rn(v[j]);
    lo[j] = __float2half_rn(v[j] - __half2float(hi[j]));
  }
  *reinterpret_cast<uint2*>(dst + r * 2 * AC_HD + d) = *reinterpret_cast<const uint2*>(hi);
  *reinterpret_cast<uint2*>(dst + r * 2 * AC_HD + AC_HD + d) = *reinterpret_cast<const uint2*>(lo);
}

// x [BH, n, 64] -> xt [BH, 64, 2*np]: xt[bh, d, s] = hi(mul x[bh, s, d]), xt[bh, d, np + s] = lo; zero for s >= n
__global__ void __launch_bounds__(256) attn_split_transposed_kernel(const float* __restrict__ x, int n, int np, float mul,
                                                                    const float* __restrict__ amax,
                                                                    __half* __restrict__ xt) {
  __shared__ float tile[32][33];
  mul = attn_mul(mul, amax);
  const int bh = blockIdx.z;
  const int s0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = x + (int64_t)bh * n * AC_HD;
  for (int k = ty; k < 32; k += 8) {
    const int s = s0 + k;
    tile[k][tx] = s < n ? src[(int64_t)s * AC_HD + d0 + tx] * mul : 0.f;
  }
  __syncthreads();
  __half* dst = xt + (int64_t)bh * AC_HD * 2 * np;
  for (int k = ty; k < 32; k += 8) {
    const int d = d0 + k, s = s0 + tx;
    if (s < np) {
      const float v = tile[tx][k];
      const __half hi = __float2half_rn(v);
      dst[(int64_t)d * 2 * np + s] = hi;
      dst[(int64_t)d * 2 * np + np + s] = __float2half_rn(v - __half2float(hi));
    }
  }
}

int attn_split_rows(const float* src, int64_t R, float mul, const float* amax, __half* dst, cudaStream_t st) {
  attn_split_rows_kernel<<<(unsigned)ceil_div64(R * AC_HD, 1024), 256, 0, st>>>(src, R, mul, amax, dst);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int attn_split_transposed(const float* src, int64_t BH, int n, int np, float mul, const float* amax, __half* dst,
                          cudaStream_t st) {
  dim3 grid((unsigned)(np / 32), AC_HD / 32, (unsigned)BH);
  attn_split_transposed_kernel<<<grid, 256, 0, st>>>(src, n, np, mul, amax, dst);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

// ---------------------------------------------------------------- host side
bool attn_tc_supported(int B, int heads, int L, int S, int hd) {
  return hd == AC_HD && S >= 1 && S <= 256 && L >= 1 && (int64_t)B * heads <= 65535 &&
         (int64_t)B * heads * (int64_t)(L > S ? L : S) < (1ll << 31);
}

// Where it pays (measured, profiles/r1_attention_core.txt): three operand-preparation launches put a floor of
// ~50 us under the tensor-core path; the CUDA-core kernel costs ~20 us per 2^20 (row, key) pairs.
bool attn_tc_profitable(int B, int heads, int L, int S, int hd) {
  return attn_tc_supported(B, heads, L, S, hd) && (int64_t)B * heads * L * S >= 3 * (1ll << 20);
}

static int attn_sp(int S) { return attn_pad64(S); }

size_t attn_tc_workspace_bytes(int B, int heads, int L, int S) {
  const int64_t bh = (int64_t)B * heads;
  Carver c(nullptr);
  c.take<__half>((size_t)(bh * L + AC_BM) * 2 * AC_HD);
  c.take<__half>((size_t)(bh * S + 256) * 2 * AC_HD);
  c.take<__half>((size_t)bh * AC_HD * 2 * attn_sp(S));
  return c.used() + 256;
}

int attn_fwd_tc(const float* q, const float* k, const float* v, const unsigned char* mask, int B, int heads, int L,
                int S, float scale, float drop_p, uint64_t seed, float* out, float* lse, void* workspace,
                cudaStream_t st) {
  const int64_t bh = (int64_t)B * heads;
  const int Sp = attn_sp(S);
  Carver c(workspace);
  __half* q2 = c.take<__half>((size_t)(bh * L + AC_BM) * 2 * AC_HD);
  __half* k2 = c.take<__half>((size_t)(bh * S + 256) * 2 * AC_HD);
  __half* vt2 = c.take<__half>((size_t)bh * AC_HD * 2 * Sp);
  int rc;
  if ((rc = attn_split_rows(q, bh * L, scale * 1.4426950408889634f, nullptr, q2, st))) return rc;
  if ((rc = attn_split_rows(k, bh * S, 1.f, nullptr, k2, st))) return rc;
  if ((rc = attn_split_transposed(v, bh, S, Sp, 1.f, nullptr, vt2, st))) return rc;

  AttnTcParams p;
  p.BH = (int)bh; p.heads = heads; p.L = L; p.S = S; p.Sp = Sp; p.mask = mask; p.drop_p = drop_p; p.seed = seed;
  p.out = out; p.lse = lse;
  CUtensorMap mq, mk, mv;
  if ((rc = encode_2d_f16(&mq, q2, (uint64_t)(bh * L), 2 * AC_HD, 64, AC_BM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = encode_2d_f16(&mk, k2, (uint64_t)(bh * S), 2 * AC_HD, 64, (uint32_t)Sp, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = encode_2d_f16(&mv, vt2, (uint64_t)(bh * AC_HD), (uint64_t)2 * Sp, 64, AC_HD, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  const int nks = Sp / 64;
  const size_t region1 = std::max((size_t)2 * AC_SLAB + (size_t)2 * Sp * 128, (size_t)2 * nks * AC_SLAB);
  const size_t smem = 1024 + region1 + (size_t)2 * nks * AC_HD * 128 + 192;
  HSG_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((L + AC_BM - 1) / AC_BM), (unsigned)bh);
  attn_fwd_tc_kernel<<<grid, AC_THREADS, smem, st>>>(mq, mk, mv, p);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

}  // namespace hsg
