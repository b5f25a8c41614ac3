// K5 on tensor cores: backward of the clustering transformer's attention core (hd = 64, L, S <= 256).
//
// Reference: autograd through nn.MultiheadAttention's need_weights=True slow path
// (hsg/models/heads/transformer.py:235,300,304): bmm / dropout / softmax / baddbmm backward, which keeps
// two [B*h, L, S] fp32 temporaries in memory.  Here, with p = softmax row, kf = dropout factor,
// delta = <dout, out>:
//     dP = dO V^T          dS = p * (dP * kf - delta)          P~ = p * kf
//     dQ = scale dS K      dK = scale dS^T Q                   dV = P~^T dO
// as two kernels, every contraction on tcgen05 with fp16 (hi | lo) operands (three passes, fp32-grade as in
// the forward); the score matrix is recomputed from q, k and the saved log-sum-exp and never leaves the SM.
//   dq kernel    : one CTA per (batch*head, 128 query rows).  S and dP land in TMEM ([128 x S] each, all 512
//                  columns); a thread per (row, column half) forms dS and writes it as fp16 (hi, lo) in the
//                  128B-swizzled K-major layout; dQ = dS K with K^T streamed in after the first products.
//   dk/dv kernel : one CTA per (batch*head, 128 keys).  The same products transposed (S^T = K Q^T, dP^T = V dO^T:
//                  lanes = keys, columns = query rows), so dS^T and P~^T come out K-major over the query rows
//                  without any transposition in shared memory; dK = dS^T Q, dV = P~^T dO.
// dout is normalised by its largest magnitude (device scalar) before the fp16 split so that small upstream
// gradients do not fall into the fp16 subnormals; the epilogues multiply it back.
#include "attention_tc.cuh"

#include <algorithm>

namespace hsg {

constexpr int AB_THREADS = 320;           // warp0: TMA + MMA issue, warp1: TMEM allocation, warps 2-9: sweep / epilogue
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct AttnBwdTcParams {
  int BH, heads, L, S, Lp, Sp;    // Lp, Sp: rounded up to 64
  const unsigned char* mask;      // [B,S] or NULL
  float drop_p;
  uint64_t seed;
  float scale;
  const float* lse;               // [BH,L] natural-log log-sum-exp of the scaled scores (forward)
  const float* delta;             // [BH,L] <dout, out> / amax
  const float* amax;              // max |dout|
  float* dq;                      // [BH,L,64]
  float* dk;                      // [BH,S,64]
  float* dv;                      // [BH,S,64]
};

// three fp16 passes of one [128 x 64k] x [N x 64k]^T product: ah*bh + al*bh + ah*bl
__device__ __forceinline__ void mma3(uint32_t tmem, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl, uint32_t idesc,
                                     bool fresh) {
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16_elect(tmem, ah + 2 * k4, bh + 2 * k4, idesc, (fresh && k4 == 0) ? 0u : 1u);
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16_elect(tmem, al + 2 * k4, bh + 2 * k4, idesc, 1u);
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16_elect(tmem, ah + 2 * k4, bl + 2 * k4, idesc, 1u);
}

// hi / lo fp16 halves of 16 values -> the two slabs of a (hi, lo) operand, row r, columns [16*c16, 16*c16 + 16) of a slab
__device__ __forceinline__ void store16_split(uint8_t* slab_hi, uint8_t* slab_lo, int r, int c16, const float (&x)[16]) {
  __align__(16) __half hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    hi[j] = __float2half_rn(x[j]);
    lo[j] = __float2half_rn(x[j] - __half2float(hi[j]));
  }
#pragma unroll
  for (int half8 = 0; half8 < 2; ++half8) {
    const uint32_t off = swz128_off(r, (c16 << 1) + half8);
    *reinterpret_cast<uint4*>(slab_hi + off) = *reinterpret_cast<const uint4*>(hi + 8 * half8);
    *reinterpret_cast<uint4*>(slab_lo + off) = *reinterpret_cast<const uint4*>(lo + 8 * half8);
  }
}

// ===================================================================== dq
// shared memory, phase 1: Q hi|lo, dO hi|lo (4 x 16 KiB), K hi|lo, V hi|lo (4 x Sp*128 B)
//                phase 2: dS hi, dS lo (2*nks slabs of 16 KiB), then K^T hi, lo (2*nks slabs of 8 KiB)
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                      const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                      const __grid_constant__ CUtensorMap tmap_kt, const AttnBwdTcParams p, const uint32_t region) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - smem_u32(smem_raw));
  const int nks = p.Sp / 64;
  const uint32_t k_bytes = (uint32_t)p.Sp * 128u;
  const uint32_t sQ = base, sDO = sQ + 2 * AC_SLAB, sK = sDO + 2 * AC_SLAB, sV = sK + 2 * k_bytes;
  const uint32_t oDS = 0, oKT = 2u * nks * AC_SLAB;             // phase-2 offsets from base
  uint64_t* bars = reinterpret_cast<uint64_t*>(gbase + region);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  uint32_t* live_bits = tmem_slot + 2;                          // [8]
  const uint32_t bar_in = smem_u32(bars), bar_s = bar_in + 8, bar_t = bar_in + 16, bar_p = bar_in + 24, bar_o = bar_in + 32;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int row0 = blockIdx.x * AC_BM;
  if (warp == 0 && lane == 0) {
    mbar_init(bar_in, 1); mbar_init(bar_s, 1); mbar_init(bar_t, 1); mbar_init(bar_p, 8); mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_s = tmem_base, t_dp = tmem_base + 256, t_dq = tmem_base;

  if (warp == 0) {
    {   // the whole warp walks this role, an elected lane issues (tc_common.cuh: tc_mma_f16_elect)
      mbar_expect_tx_elect(bar_in, 4 * AC_SLAB + 4 * k_bytes);
      tma_load_2d_elect(sQ, &tmap_q, 0, bh * p.L + row0, bar_in);
      tma_load_2d_elect(sQ + AC_SLAB, &tmap_q, AC_HD, bh * p.L + row0, bar_in);
      tma_load_2d_elect(sDO, &tmap_do, 0, bh * p.L + row0, bar_in);
      tma_load_2d_elect(sDO + AC_SLAB, &tmap_do, AC_HD, bh * p.L + row0, bar_in);
      tma_load_2d_elect(sK, &tmap_k, 0, bh * p.S, bar_in);
      tma_load_2d_elect(sK + k_bytes, &tmap_k, AC_HD, bh * p.S, bar_in);
      tma_load_2d_elect(sV, &tmap_v, 0, bh * p.S, bar_in);
      tma_load_2d_elect(sV + k_bytes, &tmap_v, AC_HD, bh * p.S, bar_in);
      mbar_wait(bar_in, 0);
      tc_fence_after();
      {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Sp >> 3) << 17) | ((uint32_t)(AC_BM >> 4) << 24);
        mma3(t_s, umma_desc(sQ, 1024, 2), umma_desc(sQ + AC_SLAB, 1024, 2), umma_desc(sK, 1024, 2),
             umma_desc(sK + k_bytes, 1024, 2), idesc, true);
        mma3(t_dp, umma_desc(sDO, 1024, 2), umma_desc(sDO + AC_SLAB, 1024, 2), umma_desc(sV, 1024, 2),
             umma_desc(sV + k_bytes, 1024, 2), idesc, true);
        tc_commit_elect(bar_s);
      }
      // the phase-1 operands are dead once the products have completed: bring K^T into their place
      mbar_wait(bar_s, 0);
      mbar_expect_tx_elect(bar_t, 2 * nks * AC_TSLAB);
      for (int ks = 0; ks < nks; ++ks) {
        tma_load_2d_elect(base + oKT + ks * AC_TSLAB, &tmap_kt, ks * 64, bh * AC_HD, bar_t);
        tma_load_2d_elect(base + oKT + (nks + ks) * AC_TSLAB, &tmap_kt, p.Sp + ks * 64, bh * AC_HD, bar_t);
      }
      mbar_wait(bar_p, 0);
      mbar_wait(bar_t, 0);
      tc_fence_after();
      {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(AC_HD >> 3) << 17) | ((uint32_t)(AC_BM >> 4) << 24);
        for (int ks = 0; ks < nks; ++ks)
          mma3(t_dq, umma_desc(base + oDS + ks * AC_SLAB, 1024, 2), umma_desc(base + oDS + (nks + ks) * AC_SLAB, 1024, 2),
               umma_desc(base + oKT + ks * AC_TSLAB, 1024, 2), umma_desc(base + oKT + (nks + ks) * AC_TSLAB, 1024, 2),
               idesc, ks == 0);
        tc_commit_elect(bar_o);
      }
    }
  } else if (warp >= 2) {
    const int q = warp & 3;                                      // TMEM lane quarter this warp may read
    const int h = (warp - 2) >> 2;                               // column half
    const int r = 32 * q + lane;
    const int row = row0 + r;
    const bool inb = row < p.L;
    const int b = bh / p.heads;
    const unsigned char* mrow = p.mask ? p.mask + (int64_t)b * p.S : nullptr;
    const uint32_t trow = ((uint32_t)(32 * q) << 16);
    {
      const int key = threadIdx.x - 64;                          // 0..255
      const bool lv = key < p.S && !(mrow && mrow[key]);
      const unsigned bal = __ballot_sync(FULL, lv);
      if (lane == 0) live_bits[key >> 5] = bal;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    const float l2 = inb ? p.lse[(int64_t)bh * p.L + row] * LOG2E : 0.f;
    const float dl = inb ? p.delta[(int64_t)bh * p.L + row] : 0.f;
    const float keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    const int nchunk = p.Sp / 16;
    const int c_lo = h * (nchunk >> 1), c_hi = c_lo + (nchunk >> 1);
    mbar_wait(bar_s, 0);
    tc_fence_after();
    for (int c = c_lo; c < c_hi; ++c) {
      uint32_t v[16], w[16];
      tc_ld16(t_s + trow + c * 16, v);
      tc_ld16(t_dp + trow + c * 16, w);
      tc_ld_wait();
      const uint32_t bits = live_bits[c >> 1] >> ((c & 1) * 16);
      float x[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const bool live = (bits >> j) & 1u;
        const float pj = ex2f(__uint_as_float(v[j]) - l2);
        float kf = 1.f;
        if (p.drop_p > 0.f) kf = attn_dropout_keep(p.seed, bh, row, c * 16 + j, p.drop_p) ? keep_scale : 0.f;
        x[j] = live ? pj * (__uint_as_float(w[j]) * kf - dl) : 0.f;
      }
      const int ks = c >> 2;
      store16_split(gbase + oDS + ks * AC_SLAB, gbase + oDS + (nks + ks) * AC_SLAB, r, c & 3, x);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    // epilogue: this thread stores columns [32h, 32h + 32) of its row
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float am = *p.amax;
    const float mul = p.scale * am;
#pragma unroll 1
    for (int c = 2 * h; c < 2 * h + 2; ++c) {
      uint32_t v[16];
      tc_ld16(t_dq + trow + c * 16, v);
      tc_ld_wait();
      if (inb) {
        float4* dst = reinterpret_cast<float4*>(p.dq + ((int64_t)bh * p.L + row) * AC_HD + c * 16);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          dst[u] = make_float4(__uint_as_float(v[4 * u]) * mul, __uint_as_float(v[4 * u + 1]) * mul,
                               __uint_as_float(v[4 * u + 2]) * mul, __uint_as_float(v[4 * u + 3]) * mul);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ===================================================================== dk, dv
// One CTA per (batch*head, 128 keys); the query rows go by in halves of 128 (shared memory holds the operands of
// one half), dk / dv accumulate in TMEM across the halves.
// shared memory, phase 1: K hi|lo, V hi|lo tiles (4 x 16 KiB), Q hi|lo, dO hi|lo of the row half (4 x qbox*128 B)
//                phase 2: dS^T hi, lo, P~^T hi, lo (4*nls slabs of 16 KiB), then Q^T hi, lo, dO^T hi, lo (4*nls x 8 KiB)
// Every barrier completes one phase per half (parity = half & 1).
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                       const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                       const __grid_constant__ CUtensorMap tmap_qt, const __grid_constant__ CUtensorMap tmap_dot,
                       const AttnBwdTcParams p, const uint32_t region, const int qbox) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - smem_u32(smem_raw));
  const int halves = (p.Lp + AC_BM - 1) / AC_BM;
  const uint32_t q_bytes = (uint32_t)qbox * 128u;               // the TMA box of Q / dO: qbox = min(128, Lp) rows
  const uint32_t sK = base, sV = sK + 2 * AC_SLAB, sQ = sV + 2 * AC_SLAB, sDO = sQ + 2 * q_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gbase + region);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* s_l2 = reinterpret_cast<float*>(tmem_slot + 4);        // [128] log2-domain log-sum-exp of the query rows
  float* s_dl = s_l2 + AC_BM;                                   // [128] delta
  const uint32_t bar_in = smem_u32(bars), bar_s = bar_in + 8, bar_t = bar_in + 16, bar_p = bar_in + 24, bar_o = bar_in + 32;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int key0 = blockIdx.x * AC_BM;
  if (warp == 0 && lane == 0) {
    mbar_init(bar_in, 1); mbar_init(bar_s, 1); mbar_init(bar_t, 1); mbar_init(bar_p, 8); mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_s = tmem_base, t_dp = tmem_base + 128, t_dk = tmem_base + 256, t_dv = tmem_base + 320;

  if (warp == 0) {
    {   // the whole warp walks this role, an elected lane issues (tc_common.cuh: tc_mma_f16_elect)
      for (int half = 0; half < halves; ++half) {
        const uint32_t ph = half & 1;
        const int l0 = half * AC_BM;
        const int width = min(AC_BM, p.Lp - l0);                // 64 or 128 (padded) query rows
        const int nls = width / 64;
        const uint32_t oDS = 0, oPD = 2u * nls * AC_SLAB, oQT = 4u * nls * AC_SLAB, oDOT = oQT + 2u * nls * AC_TSLAB;
        // the previous half's second products still read the shared memory the loads below overwrite
        if (half > 0) mbar_wait(bar_o, ph ^ 1u);
        mbar_expect_tx_elect(bar_in, 4 * AC_SLAB + 4 * q_bytes);
        tma_load_2d_elect(sK, &tmap_k, 0, bh * p.S + key0, bar_in);
        tma_load_2d_elect(sK + AC_SLAB, &tmap_k, AC_HD, bh * p.S + key0, bar_in);
        tma_load_2d_elect(sV, &tmap_v, 0, bh * p.S + key0, bar_in);
        tma_load_2d_elect(sV + AC_SLAB, &tmap_v, AC_HD, bh * p.S + key0, bar_in);
        tma_load_2d_elect(sQ, &tmap_q, 0, bh * p.L + l0, bar_in);
        tma_load_2d_elect(sQ + q_bytes, &tmap_q, AC_HD, bh * p.L + l0, bar_in);
        tma_load_2d_elect(sDO, &tmap_do, 0, bh * p.L + l0, bar_in);
        tma_load_2d_elect(sDO + q_bytes, &tmap_do, AC_HD, bh * p.L + l0, bar_in);
        mbar_wait(bar_in, ph);
        tc_fence_after();
        {
          const uint32_t idesc = (1u << 4) | ((uint32_t)(width >> 3) << 17) | ((uint32_t)(AC_BM >> 4) << 24);
          mma3(t_s, umma_desc(sK, 1024, 2), umma_desc(sK + AC_SLAB, 1024, 2), umma_desc(sQ, 1024, 2),
               umma_desc(sQ + q_bytes, 1024, 2), idesc, true);
          mma3(t_dp, umma_desc(sV, 1024, 2), umma_desc(sV + AC_SLAB, 1024, 2), umma_desc(sDO, 1024, 2),
               umma_desc(sDO + q_bytes, 1024, 2), idesc, true);
          tc_commit_elect(bar_s);
        }
        // the phase-1 operands are dead once the products have completed: bring Q^T, dO^T into their place
        mbar_wait(bar_s, ph);
        mbar_expect_tx_elect(bar_t, 4 * nls * AC_TSLAB);
        for (int ls = 0; ls < nls; ++ls) {
          tma_load_2d_elect(base + oQT + ls * AC_TSLAB, &tmap_qt, l0 + ls * 64, bh * AC_HD, bar_t);
          tma_load_2d_elect(base + oQT + (nls + ls) * AC_TSLAB, &tmap_qt, p.Lp + l0 + ls * 64, bh * AC_HD, bar_t);
          tma_load_2d_elect(base + oDOT + ls * AC_TSLAB, &tmap_dot, l0 + ls * 64, bh * AC_HD, bar_t);
          tma_load_2d_elect(base + oDOT + (nls + ls) * AC_TSLAB, &tmap_dot, p.Lp + l0 + ls * 64, bh * AC_HD, bar_t);
        }
        mbar_wait(bar_p, ph);
        mbar_wait(bar_t, ph);
        tc_fence_after();
        {
          const uint32_t idesc = (1u << 4) | ((uint32_t)(AC_HD >> 3) << 17) | ((uint32_t)(AC_BM >> 4) << 24);
          for (int ls = 0; ls < nls; ++ls) {
            const bool fresh = half == 0 && ls == 0;
            mma3(t_dk, umma_desc(base + oDS + ls * AC_SLAB, 1024, 2), umma_desc(base + oDS + (nls + ls) * AC_SLAB, 1024, 2),
                 umma_desc(base + oQT + ls * AC_TSLAB, 1024, 2), umma_desc(base + oQT + (nls + ls) * AC_TSLAB, 1024, 2),
                 idesc, fresh);
            mma3(t_dv, umma_desc(base + oPD + ls * AC_SLAB, 1024, 2), umma_desc(base + oPD + (nls + ls) * AC_SLAB, 1024, 2),
                 umma_desc(base + oDOT + ls * AC_TSLAB, 1024, 2), umma_desc(base + oDOT + (nls + ls) * AC_TSLAB, 1024, 2),
                 idesc, fresh);
          }
          tc_commit_elect(bar_o);
        }
      }
    }
  } else if (warp >= 2) {
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int r = 32 * q + lane;
    const int key = key0 + r;
    const int b = bh / p.heads;
    const bool in_s = key < p.S;
    const bool live = in_s && !(p.mask && p.mask[(int64_t)b * p.S + key]);
    const uint32_t trow = ((uint32_t)(32 * q) << 16);
    const float keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    for (int half = 0; half < halves; ++half) {
      const uint32_t ph = half & 1;
      const int l0 = half * AC_BM;
      const int width = min(AC_BM, p.Lp - l0);
      const int nls = width / 64;
      const uint32_t oDS = 0, oPD = 2u * nls * AC_SLAB;
      asm volatile("bar.sync 1, 256;" ::: "memory");             // every sweep thread is done with the previous half's rows
      {
        const int t = threadIdx.x - 64;                          // 0..255; the first 128 fetch one query row each
        if (t < AC_BM) {
          const bool in_l = l0 + t < p.L;
          s_l2[t] = in_l ? p.lse[(int64_t)bh * p.L + l0 + t] * LOG2E : 0.f;
          s_dl[t] = in_l ? p.delta[(int64_t)bh * p.L + l0 + t] : 0.f;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int nchunk = width / 16;                             // 4 or 8
      const int c_lo = h * (nchunk >> 1), c_hi = c_lo + (nchunk >> 1);
      mbar_wait(bar_s, ph);
      tc_fence_after();
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t v[16], w[16];
        tc_ld16(t_s + trow + c * 16, v);
        tc_ld16(t_dp + trow + c * 16, w);
        tc_ld_wait();
        float l2[16], dl[16];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          *reinterpret_cast<float4*>(l2 + 4 * u) = *reinterpret_cast<const float4*>(s_l2 + c * 16 + 4 * u);
          *reinterpret_cast<float4*>(dl + 4 * u) = *reinterpret_cast<const float4*>(s_dl + c * 16 + 4 * u);
        }
        float xs[16], xp[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int row = l0 + c * 16 + j;
          const bool ok = live && row < p.L;
          const float pj = ex2f(__uint_as_float(v[j]) - l2[j]);
          float kf = 1.f;
          if (p.drop_p > 0.f) kf = attn_dropout_keep(p.seed, bh, row, key, p.drop_p) ? keep_scale : 0.f;
          xp[j] = ok ? pj * kf : 0.f;
          xs[j] = ok ? pj * (__uint_as_float(w[j]) * kf - dl[j]) : 0.f;
        }
        const int ls = c >> 2;
        store16_split(gbase + oDS + ls * AC_SLAB, gbase + oDS + (nls + ls) * AC_SLAB, r, c & 3, xs);
        store16_split(gbase + oPD + ls * AC_SLAB, gbase + oPD + (nls + ls) * AC_SLAB, r, c & 3, xp);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    mbar_wait(bar_o, (uint32_t)(halves - 1) & 1u);
    tc_fence_after();
    const float am = *p.amax;
    const float mk = am * LN2;                                   // q was pre-scaled by scale * log2(e)
#pragma unroll 1
    for (int c = 2 * h; c < 2 * h + 2; ++c) {
      uint32_t v[16], w[16];
      tc_ld16(t_dk + trow + c * 16, v);
      tc_ld16(t_dv + trow + c * 16, w);
      tc_ld_wait();
      if (in_s) {
        float4* dk = reinterpret_cast<float4*>(p.dk + ((int64_t)bh * p.S + key) * AC_HD + c * 16);
        float4* dv = reinterpret_cast<float4*>(p.dv + ((int64_t)bh * p.S + key) * AC_HD + c * 16);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          dk[u] = make_float4(__uint_as_float(v[4 * u]) * mk, __uint_as_float(v[4 * u + 1]) * mk,
                              __uint_as_float(v[4 * u + 2]) * mk, __uint_as_float(v[4 * u + 3]) * mk);
          dv[u] = make_float4(__uint_as_float(w[4 * u]) * am, __uint_as_float(w[4 * u + 1]) * am,
                              __uint_as_float(w[4 * u + 2]) * am, __uint_as_float(w[4 * u + 3]) * am);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------- operand preparation
// One pass over x [BH, n, 64]: rows [(BH*n), 128] = fp16 (hi | lo) of mul * x, and the transposed copy
// xt [BH, 64, 2*np] (xt[bh, d, s] = hi, xt[bh, d, np + s] = lo, zero for n <= s < np).  With DELTA also
// delta[bh*n + s] = mul * <x[s], o[s]> (x = dout, o = out, mul = 1 / max|dout|).  One CTA per 32 rows.
template <bool DELTA>
__global__ void __launch_bounds__(256) attn_split_both_kernel(const float* __restrict__ x, const float* __restrict__ o,
                                                              int n, int np, float mul, const float* __restrict__ amax,
                                                              __half* __restrict__ rows, __half* __restrict__ xt,
                                                              float* __restrict__ delta) {
  __shared__ float tile[32][AC_HD + 1];
  mul = attn_mul(mul, amax);
  const int bh = blockIdx.y;
  const int s0 = blockIdx.x * 32;
#pragma unroll
  for (int rep = 0; rep < 2; ++rep) {
    const int idx = threadIdx.x + 256 * rep;                     // 512 float4 of the [32 x 64] tile
    const int r = idx >> 4, c = (idx & 15) << 2;
    const int s = s0 + r;
    const bool in = s < n;
    const int64_t off = ((int64_t)bh * n + s) * AC_HD + c;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in) g = *reinterpret_cast<const float4*>(x + off);
    const float v[4] = {g.x * mul, g.y * mul, g.z * mul, g.w * mul};
#pragma unroll
    for (int j = 0; j < 4; ++j) tile[r][c + j] = v[j];
    if (in) {
      __align__(8) __half hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hi[j] = __float2half_rn(v[j]);
        lo[j] = __float2half_rn(v[j] - __half2float(hi[j]));
      }
      __half* dst = rows + ((int64_t)bh * n + s) * 2 * AC_HD + c;
      *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(dst + AC_HD) = *reinterpret_cast<const uint2*>(lo);
    }
    if (DELTA) {                                                 // the 16 threads of a row are 16 consecutive lanes
      float4 ov = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in) ov = *reinterpret_cast<const float4*>(o + off);
      float dot = v[0] * ov.x + v[1] * ov.y + v[2] * ov.z + v[3] * ov.w;
#pragma unroll
      for (int sh = 8; sh > 0; sh >>= 1) dot += __shfl_xor_sync(FULL, dot, sh);
      if (in && (idx & 15) == 0) delta[(int64_t)bh * n + s] = dot;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = s0 + lane;
  if (s < np) {
    __half* dst = xt + (int64_t)bh * AC_HD * 2 * np;
#pragma unroll
    for (int k = 0; k < AC_HD / 8; ++k) {
      const int d = warp + 8 * k;
      const float v = tile[lane][d];
      const __half hi = __float2half_rn(v);
      dst[(int64_t)d * 2 * np + s] = hi;
      dst[(int64_t)d * 2 * np + np + s] = __float2half_rn(v - __half2float(hi));
    }
  }
}

// ---------------------------------------------------------------- host side
int absmax(const float* x, int64_t n, float* out, cudaStream_t st);      // gemm_tc.cu

bool attn_bwd_tc_supported(int B, int heads, int L, int S, int hd) {
  return hd == AC_HD && S >= 1 && S <= 256 && L >= 1 && L <= 256 && (int64_t)B * heads <= 65535 &&
         (int64_t)B * heads * 256 < (1ll << 31);
}

// Where it pays (measured, profiles/r1_attention_core.txt): the two kernels are preceded by five small
// operand-preparation launches, and a tile of mostly padding (L = 16, S = 64) is slower than the CUDA-core kernels.
bool attn_bwd_tc_profitable(int B, int heads, int L, int S, int hd) {
  return attn_bwd_tc_supported(B, heads, L, S, hd) && L >= 64 && S >= 96 &&
         (int64_t)B * heads * L * S >= (1ll << 21);
}

namespace {
struct BwdCarve {
  __half *q2, *k2, *v2, *do2, *kt2, *qt2, *dot2;
  float *delta, *amax;
  size_t bytes;
};
BwdCarve bwd_carve(void* ws, int64_t bh, int L, int S) {
  const int Lp = attn_pad64(L), Sp = attn_pad64(S);
  Carver c(ws);
  BwdCarve b;
  b.q2 = c.take<__half>((size_t)(bh * L) * 2 * AC_HD);
  b.k2 = c.take<__half>((size_t)(bh * S) * 2 * AC_HD);
  b.v2 = c.take<__half>((size_t)(bh * S) * 2 * AC_HD);
  b.do2 = c.take<__half>((size_t)(bh * L) * 2 * AC_HD);
  b.kt2 = c.take<__half>((size_t)bh * AC_HD * 2 * Sp);
  b.qt2 = c.take<__half>((size_t)bh * AC_HD * 2 * Lp);
  b.dot2 = c.take<__half>((size_t)bh * AC_HD * 2 * Lp);
  b.delta = c.take<float>((size_t)(bh * L));
  b.amax = c.take<float>(64);
  b.bytes = c.used() + 256;
  return b;
}
}  // namespace

size_t attn_bwd_tc_workspace_bytes(int B, int heads, int L, int S) {
  return bwd_carve(nullptr, (int64_t)B * heads, L, S).bytes;
}

int attn_bwd_tc(const float* q, const float* k, const float* v, const unsigned char* mask, int B, int heads, int L,
                int S, float scale, float drop_p, uint64_t seed, const float* out, const float* lse, const float* dout,
                float* dq, float* dk, float* dv, void* workspace, cudaStream_t st) {
  const int64_t bh = (int64_t)B * heads;
  const int Lp = attn_pad64(L), Sp = attn_pad64(S);
  const BwdCarve w = bwd_carve(workspace, bh, L, S);
  int rc;
  if ((rc = absmax(dout, bh * L * AC_HD, w.amax, st))) return rc;
  {
    dim3 gq((unsigned)(Lp / 32), (unsigned)bh), gk((unsigned)(Sp / 32), (unsigned)bh);
    attn_split_both_kernel<false><<<gq, 256, 0, st>>>(q, nullptr, L, Lp, scale * LOG2E, nullptr, w.q2, w.qt2, nullptr);
    HSG_LAUNCH_CHECK();
    attn_split_both_kernel<false><<<gk, 256, 0, st>>>(k, nullptr, S, Sp, 1.f, nullptr, w.k2, w.kt2, nullptr);
    HSG_LAUNCH_CHECK();
    if ((rc = attn_split_rows(v, bh * S, 1.f, nullptr, w.v2, st))) return rc;
    attn_split_both_kernel<true><<<gq, 256, 0, st>>>(dout, out, L, Lp, 1.f, w.amax, w.do2, w.dot2, w.delta);
    HSG_LAUNCH_CHECK();
  }

  AttnBwdTcParams p;
  p.BH = (int)bh; p.heads = heads; p.L = L; p.S = S; p.Lp = Lp; p.Sp = Sp; p.mask = mask; p.drop_p = drop_p; p.seed = seed;
  p.scale = scale; p.lse = lse; p.delta = w.delta; p.amax = w.amax; p.dq = dq; p.dk = dk; p.dv = dv;
  CUtensorMap mq, mdo, mk, mv, mkt, mqt, mdot;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  const int nks = Sp / 64;
  // dq: 128-row tiles of q / dout against every key
  if ((rc = encode_2d_f16(&mq, w.q2, (uint64_t)(bh * L), 2 * AC_HD, 64, AC_BM, sw))) return rc;
  if ((rc = encode_2d_f16(&mdo, w.do2, (uint64_t)(bh * L), 2 * AC_HD, 64, AC_BM, sw))) return rc;
  if ((rc = encode_2d_f16(&mk, w.k2, (uint64_t)(bh * S), 2 * AC_HD, 64, (uint32_t)Sp, sw))) return rc;
  if ((rc = encode_2d_f16(&mv, w.v2, (uint64_t)(bh * S), 2 * AC_HD, 64, (uint32_t)Sp, sw))) return rc;
  if ((rc = encode_2d_f16(&mkt, w.kt2, (uint64_t)(bh * AC_HD), (uint64_t)2 * Sp, 64, AC_HD, sw))) return rc;
  {
    const size_t p1 = (size_t)4 * AC_SLAB + (size_t)4 * Sp * 128, p2 = (size_t)nks * (2 * AC_SLAB + 2 * AC_TSLAB);
    const uint32_t region = (uint32_t)std::max(p1, p2);
    const size_t smem = 1024 + region + 256;
    HSG_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((L + AC_BM - 1) / AC_BM), (unsigned)bh);
    attn_bwd_dq_tc_kernel<<<grid, AB_THREADS, smem, st>>>(mq, mdo, mk, mv, mkt, p, region);
    HSG_LAUNCH_CHECK();
  }
  // dk / dv: 128-key tiles of k / v against the queries, 128 rows at a time
  const int qbox = std::min(AC_BM, Lp);
  if ((rc = encode_2d_f16(&mk, w.k2, (uint64_t)(bh * S), 2 * AC_HD, 64, AC_BM, sw))) return rc;
  if ((rc = encode_2d_f16(&mv, w.v2, (uint64_t)(bh * S), 2 * AC_HD, 64, AC_BM, sw))) return rc;
  if ((rc = encode_2d_f16(&mq, w.q2, (uint64_t)(bh * L), 2 * AC_HD, 64, (uint32_t)qbox, sw))) return rc;
  if ((rc = encode_2d_f16(&mdo, w.do2, (uint64_t)(bh * L), 2 * AC_HD, 64, (uint32_t)qbox, sw))) return rc;
  if ((rc = encode_2d_f16(&mqt, w.qt2, (uint64_t)(bh * AC_HD), (uint64_t)2 * Lp, 64, AC_HD, sw))) return rc;
  if ((rc = encode_2d_f16(&mdot, w.dot2, (uint64_t)(bh * AC_HD), (uint64_t)2 * Lp, 64, AC_HD, sw))) return rc;
  {
    const int nls = qbox / 64;                                   // widest row half
    const size_t p1 = (size_t)4 * AC_SLAB + (size_t)4 * qbox * 128, p2 = (size_t)nls * (4 * AC_SLAB + 4 * AC_TSLAB);
    const uint32_t region = (uint32_t)std::max(p1, p2);
    const size_t smem = 1024 + region + 64 + 32 + 2 * AC_BM * sizeof(float) + 64;
    HSG_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((S + AC_BM - 1) / AC_BM), (unsigned)bh);
    attn_bwd_dkv_tc_kernel<<<grid, AB_THREADS, smem, st>>>(mk, mv, mq, mdo, mqt, mdot, p, region, qbox);
    HSG_LAUNCH_CHECK();
  }
  return HSG_OK;
}

}  // namespace hsg
