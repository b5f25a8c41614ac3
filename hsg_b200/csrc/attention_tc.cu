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

constexpr int AC_THREADS = 320;           // warp0: TMA + MMA issue, warp1: TMEM allocation, warps 2-9: softmax / epilogue

struct AttnTcParams {
  int BH, heads, L, S, Sp;        // Sp = S rounded up to 64
  const unsigned char* mask;      // [B,S] or NULL
  float drop_p;
  uint64_t seed;
  float* out;                     // [BH,L,64]
  float* lse;                     // [BH,L]
  const float* q;                 // [BH,L,64] fp32: split into fp16 (hi | lo) inside the kernel
  const float* k;                 // [BH,S,64]
  float q_mul;                    // scale * log2(e)
};

// r2 restructuring: two CTAs per SM and two threads per query row, so that one tile's softmax runs under the other
// tile's products (r1: one CTA per SM, load / product / sweep / product back to back: tensor pipe 15 %).
//   * shared memory 96 KB instead of 192: the probabilities and V^T move through a ring of two 64-key slabs
//     (P hi|lo 32 KB + V^T hi|lo 16 KB each) laid over the Q / K tiles, which are dead once S is in TMEM;
//   * TMEM 256 columns instead of 512: O accumulates over columns 0..63 of S, which the second sweep has consumed
//     by the time the first slab of probabilities is handed to the tensor core;
//   * O = P V starts after the FIRST slab of probabilities, not after the whole row;
//   * eight softmax warps: warps w and w + 4 own the same 32 TMEM lanes and split every slab's columns
//     (row maximum and row sum exchanged through shared memory);
//   * Q and K are read as fp32 and split into fp16 (hi | lo) by the CTA itself, straight into the swizzled
//     K-major tiles (r1: two operand-preparation launches and a round trip of both tensors through HBM).
__global__ void __launch_bounds__(AC_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_vt, const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nks = p.Sp / 64;                                   // key slabs
  const int nring = nks < 2 ? nks : 2;
  // phase 1: Q hi, Q lo (2 x 16 KiB), K hi, K lo (2 x Sp*128 B).  phase 2, same bytes: P ring, then V^T ring.
  const uint32_t sQ = base;
  const uint32_t sK = sQ + 2 * AC_SLAB;
  const uint32_t k_bytes = (uint32_t)p.Sp * 128u;
  const uint32_t sP = base;                                    // slot s: P hi slab, P lo slab (2 x 16 KiB)
  const uint32_t sV = base + (uint32_t)nring * 2u * AC_SLAB;   // slot s: V^T hi slab, lo slab (2 x 8 KiB)
  const uint32_t v_slab = AC_HD * 64 * 2;
  const uint32_t region = max(2u * AC_SLAB + 2u * k_bytes, (uint32_t)nring * (2u * AC_SLAB + 2u * v_slab));
  uint8_t* misc = smem_raw + (base + region - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  uint32_t* live_bits = tmem_slot + 2;                         // [8] bit k of word w: key 32w+k takes part
  float* xch = reinterpret_cast<float*>(live_bits + 8);        // [2][128] row maximum / row sum of the other column half
  const uint32_t bar_qk = smem_u32(bars), bar_s = bar_qk + 8, bar_o = bar_qk + 16;
  const uint32_t bar_p = bar_qk + 24, bar_v = bar_qk + 40, bar_free = bar_qk + 56;     // two each

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int row0 = blockIdx.x * AC_BM;
  if (warp == 0 && lane == 0) {
    mbar_init(bar_qk, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_p + 8 * i, 8); mbar_init(bar_v + 8 * i, 1); mbar_init(bar_free + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // TMEM: Sp columns of scores; the 64 output columns lie over the first 64 of them
  const uint32_t ncols = p.Sp <= 64 ? 64u : p.Sp <= 128 ? 128u : 256u;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // ---- Q (pre-scaled: scores in the base-2 domain) and K: fp32 rows -> fp16 hi / lo tiles, eight features per piece
  {
    uint8_t* sq = smem_raw + (sQ - smem_u32(smem_raw));
    uint8_t* sk = smem_raw + (sK - smem_u32(smem_raw));
    const float* qsrc = p.q + ((int64_t)bh * p.L + row0) * AC_HD;
    const float* ksrc = p.k + (int64_t)bh * p.S * AC_HD;
    const int q_rows = min(AC_BM, p.L - row0);
    const int n_pieces = (AC_BM + p.Sp) * 8;
    for (int i0 = threadIdx.x; i0 < n_pieces; i0 += 4 * AC_THREADS) {
      float4 a[4][2];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * AC_THREADS;
        const bool isq = i < AC_BM * 8;
        const int rr = (isq ? i : i - AC_BM * 8) >> 3, chunk = i & 7;
        const bool on = i < n_pieces && rr < (isq ? q_rows : p.S);
        const float4* src = reinterpret_cast<const float4*>((isq ? qsrc : ksrc) + (int64_t)rr * AC_HD + chunk * 8);
        a[u][0] = on ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f);
        a[u][1] = on ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * AC_THREADS;
        if (i >= n_pieces) break;
        const bool isq = i < AC_BM * 8;
        const int rr = (isq ? i : i - AC_BM * 8) >> 3, chunk = i & 7;
        const float mul = isq ? p.q_mul : 1.f;
        const float x[8] = {a[u][0].x * mul, a[u][0].y * mul, a[u][0].z * mul, a[u][0].w * mul,
                            a[u][1].x * mul, a[u][1].y * mul, a[u][1].z * mul, a[u][1].w * mul};
        __align__(16) __half hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          hi[j] = __float2half_rn(x[j]);
          lo[j] = __float2half_rn(x[j] - __half2float(hi[j]));
        }
        uint8_t* dst = (isq ? sq : sk) + swz128_off(rr, chunk);
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(dst + (isq ? (uint32_t)AC_SLAB : k_bytes)) = *reinterpret_cast<const uint4*>(lo);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the tensor core
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_s = tmem_base, t_o = tmem_base;

  if (warp == 0) {
    {   // the whole warp walks this role, an elected lane issues (tc_common.cuh: tc_mma_f16_elect)
      // ---- S = Q K^T (three passes)
      {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Sp >> 3) << 17) | ((uint32_t)(AC_BM >> 4) << 24);
        const uint64_t qh = umma_desc(sQ, 1024, 2), ql = umma_desc(sQ + AC_SLAB, 1024, 2);
        const uint64_t kh = umma_desc(sK, 1024, 2), kl = umma_desc(sK + k_bytes, 1024, 2);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16_elect(t_s, qh + 2 * k4, kh + 2 * k4, idesc, k4 ? 1u : 0u);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16_elect(t_s, ql + 2 * k4, kh + 2 * k4, idesc, 1u);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16_elect(t_s, qh + 2 * k4, kl + 2 * k4, idesc, 1u);
        tc_commit_elect(bar_s);
      }
      // ---- the Q / K bytes are free once S is complete: V^T slabs of the first ring round
      mbar_wait(bar_s, 0);
      auto load_v = [&](int ks) {
        const uint32_t slot = (uint32_t)(ks & 1);
        mbar_expect_tx_elect(bar_v + 8 * slot, 2 * v_slab);
        tma_load_2d_elect(sV + slot * 2 * v_slab, &tmap_vt, ks * 64, bh * AC_HD, bar_v + 8 * slot);
        tma_load_2d_elect(sV + slot * 2 * v_slab + v_slab, &tmap_vt, p.Sp + ks * 64, bh * AC_HD, bar_v + 8 * slot);
      };
      for (int ks = 0; ks < nring; ++ks) load_v(ks);
      // ---- O += P_ks V_ks (three passes) as the slabs of probabilities arrive
      const uint32_t idesc = (1u << 4) | ((uint32_t)(AC_HD >> 3) << 17) | ((uint32_t)(AC_BM >> 4) << 24);
      for (int ks = 0; ks < nks; ++ks) {
        const uint32_t slot = (uint32_t)(ks & 1), ph = (uint32_t)(ks >> 1) & 1u;
        mbar_wait(bar_p + 8 * slot, ph);
        mbar_wait(bar_v + 8 * slot, ph);
        tc_fence_after();
        const uint64_t pdh = umma_desc(sP + slot * 2 * AC_SLAB, 1024, 2), pdl = umma_desc(sP + slot * 2 * AC_SLAB + AC_SLAB, 1024, 2);
        const uint64_t vh = umma_desc(sV + slot * 2 * v_slab, 1024, 2), vl = umma_desc(sV + slot * 2 * v_slab + v_slab, 1024, 2);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16_elect(t_o, pdh + 2 * k4, vh + 2 * k4, idesc, (ks || k4) ? 1u : 0u);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16_elect(t_o, pdl + 2 * k4, vh + 2 * k4, idesc, 1u);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16_elect(t_o, pdh + 2 * k4, vl + 2 * k4, idesc, 1u);
        tc_commit_elect(bar_free + 8 * slot);
        if (ks == nks - 1) tc_commit_elect(bar_o);
        if (ks >= 1 && ks + 1 < nks) {            // the slot of slab ks-1 is free once its products are done: V^T of slab ks+1
          mbar_wait(bar_free + 8 * (slot ^ 1u), (uint32_t)((ks - 1) >> 1) & 1u);
          load_v(ks + 1);
        }
      }
    }
  } else if (warp >= 2) {
    // ===================== softmax / epilogue: a query row per pair of threads (column halves) =====================
    const int q = warp & 3;                                      // TMEM lane quarter of this warp
    const int g = (warp - 2) >> 2;                               // column half: chunks 0,1 (g = 0) or 2,3 (g = 1) of every slab
    const int r = 32 * q + lane;
    const int row = row0 + r;
    const bool inb = row < p.L;
    const int b = bh / p.heads;
    const unsigned char* mrow = p.mask ? p.mask + (int64_t)b * p.S : nullptr;
    const uint32_t trow = ((uint32_t)(32 * q) << 16);
    // which keys are live (inside S and not padded): 256 bits, built once per CTA
    {
      const int key = threadIdx.x - 64;                          // 0..255
      const bool lv = key < p.S && !(mrow && mrow[key]);
      const unsigned bal = __ballot_sync(FULL, lv);
      if (lane == 0) live_bits[key >> 5] = bal;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    uint32_t lw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) lw[i] = live_bits[i];
    mbar_wait(bar_s, 0);
    tc_fence_after();
    // sweep 1: row maximum over the live keys of this thread's columns, then of the row
    float mx = -INFINITY;
    for (int ks = 0; ks < nks; ++ks) {
      // both 16-column pieces of this thread's half slab in one round trip to TMEM
      uint32_t v[2][16];
      const int c0 = 4 * ks + 2 * g;
      tc_ld16(t_s + trow + c0 * 16, v[0]);
      tc_ld16(t_s + trow + c0 * 16 + 16, v[1]);
      tc_ld_wait();
      const uint32_t bits = lw[c0 >> 1];
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if ((bits >> j) & 1u) mx = fmaxf(mx, __uint_as_float(v[j >> 4][j & 15]));
    }
    xch[g * 128 + r] = mx;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    mx = fmaxf(mx, xch[(g ^ 1) * 128 + r]);
    // sweep 2: probabilities -> shared memory (fp16 hi / lo, K-major, 128B swizzle) slab by slab, row sum
    const float keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    float sum = 0.f;
    uint8_t* sp = smem_raw + (sP - smem_u32(smem_raw));
    const uint32_t rbase = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
    for (int ks = 0; ks < nks; ++ks) {
      const uint32_t slot = (uint32_t)(ks & 1);
      uint32_t vv[2][16];
      tc_ld16(t_s + trow + (4 * ks + 2 * g) * 16, vv[0]);                // both pieces of the half slab in one round trip
      tc_ld16(t_s + trow + (4 * ks + 2 * g) * 16 + 16, vv[1]);
      if (ks >= 2) mbar_wait(bar_free + 8 * slot, (uint32_t)((ks - 2) >> 1) & 1u);      // the products of slab ks-2 are done
      tc_ld_wait();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 4 * ks + 2 * g + cc;
        const uint32_t (&v)[16] = vv[cc];
        __align__(16) __half hi[16], lo[16];
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
        // element (r, key) of a slab: byte (r/8)*1024 + (r%8)*128 + ((chunk ^ (r%8)) * 16), chunk = (key%64)/8
#pragma unroll
        for (int half8 = 0; half8 < 2; ++half8) {
          const int chunk = ((c & 3) << 1) + half8;
          const uint32_t off = rbase + (uint32_t)((chunk ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(sp + slot * 2 * AC_SLAB + off) = *reinterpret_cast<const uint4*>(hi + 8 * half8);
          *reinterpret_cast<uint4*>(sp + slot * 2 * AC_SLAB + AC_SLAB + off) = *reinterpret_cast<const uint4*>(lo + 8 * half8);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the tensor core
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p + 8 * slot);
    }
    // row sum of both column halves
    xch[256 + g * 128 + r] = sum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    sum = g == 0 ? sum + xch[256 + 128 + r] : xch[256 + r] + sum;          // the same order in both threads of a row
    // epilogue: 32 of the 64 output columns per thread
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = 1.f / sum;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c = 2 * g + cc;
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
    if (inb && g == 0) p.lse[(int64_t)bh * p.L + row] = (mx + log2f(sum)) * 0.6931471805599453f;
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
  c.take<__half>((size_t)(bh * L + AC_BM) * 2 * AC_HD);        // (room the backward's own operand copies use)
  c.take<__half>((size_t)(bh * S + 256) * 2 * AC_HD);
  __half* vt2 = c.take<__half>((size_t)bh * AC_HD * 2 * Sp);
  int rc;
  if ((rc = attn_split_transposed(v, bh, S, Sp, 1.f, nullptr, vt2, st))) return rc;

  AttnTcParams p;
  p.BH = (int)bh; p.heads = heads; p.L = L; p.S = S; p.Sp = Sp; p.mask = mask; p.drop_p = drop_p; p.seed = seed;
  p.out = out; p.lse = lse; p.q = q; p.k = k; p.q_mul = scale * 1.4426950408889634f;
  CUtensorMap mv;
  if ((rc = encode_2d_f16(&mv, vt2, (uint64_t)(bh * AC_HD), (uint64_t)2 * Sp, 64, AC_HD, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  const int nks = Sp / 64, nring = nks < 2 ? nks : 2;
  const size_t region = std::max((size_t)2 * AC_SLAB + (size_t)2 * Sp * 128, (size_t)nring * (2 * AC_SLAB + 2 * AC_HD * 128));
  const size_t smem = 1024 + region + 12 * 8 + 8 + 32 + 4 * 128 * 4 + 64;
  HSG_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((L + AC_BM - 1) / AC_BM), (unsigned)bh);
  attn_fwd_tc_kernel<<<grid, AC_THREADS, smem, st>>>(mv, p);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

}  // namespace hsg
