// fp32-grade GEMM on the 5th-generation tensor cores:  C[M,N] (+)= alpha * A[M,K] . B[N,K]^T
//
// Used by the NCE backward (hsg/utils/segsort/loss.py:15-82 under autograd: dE = G P,
// dP = G^T E with G the [pixels x prototypes] gradient of the loss w.r.t. the similarities).
// A single fp16/bf16 pass cannot give gradients to 1e-5, so -- as in the forward (nce_tc.cu)
// -- both operands are pre-split into fp16 (hi, lo) pairs stored as rows [hi(K) | lo(K)] and
// three passes accumulate into one TMEM tile:  a.b ~ ah.bh + al.bh + ah.bl.
//
// Kernel: persistent, one CTA per SM, 128-row output tiles, N <= 256 columns, K streamed in
// 64-wide slabs through a TMA/mbarrier ring (one stage = A_hi, A_lo [128x64] and B_hi, B_lo
// [Nx64], 128B-swizzled); tcgen05.mma kind::f16 issued by one thread; two TMEM accumulators so
// the epilogue (tcgen05.ld -> scale -> global) of a tile overlaps the MMAs of the next.
//
// Long K.  The tensor core's fp32 accumulator does not round to nearest: every accumulation step
// loses a little in the same direction, so the error of a TMEM sum grows linearly with the number
// of steps (measured 1.8e-8 relative per step: 1.5e-5 at K = 4160).  K is therefore cut into
// segments of GT_SEG slabs (K = 1024: 192 steps, ~3e-6); each segment is a fresh accumulator and
// the epilogue adds the segments into C in ordinary round-to-nearest fp32.
#include "tc_common.cuh"

namespace hsg {

constexpr int GT_BM = 128;
constexpr int GT_BK = 64;
constexpr int GT_THREADS = 192;          // warp0 TMA, warp1 MMA + TMEM alloc, warps 2-5 epilogue
constexpr int GT_SEG = 16;               // K slabs per accumulator segment

struct GemmTcParams {
  int64_t M;
  int N, K;                  // N in {16..256, multiple of 16}; K multiple of 64
  float* C;
  int64_t ldc;
  const float* inv_scale;    // device scalar multiplied into every output (1 / operand scales), or NULL
  float alpha;
  int accumulate;
  int nst;
};

__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = GT_BM * GT_BK * 2;                  // 16 KiB
  const uint32_t b_bytes = (uint32_t)p.N * GT_BK * 2;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * ((b_bytes + 1023u) & ~1023u);
  const uint32_t b_off = 2 * a_bytes, b_stride = (b_bytes + 1023u) & ~1023u;
  uint8_t* misc = smem_raw + (base - smem_u32(smem_raw)) + (size_t)p.nst * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  const uint32_t bar_full = smem_u32(bars);            // [8]
  const uint32_t bar_empty = bar_full + 64;            // [8]
  const uint32_t bar_tfull = bar_empty + 64;           // [2]
  const uint32_t bar_tempty = bar_tfull + 16;          // [2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nst; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t n_tiles = (p.M + GT_BM - 1) / GT_BM;
  const int n_k = p.K / GT_BK;

  if (warp == 0) {
    {   // the whole warp walks the loop, an elected lane issues the copies
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int row0 = (int)(t * GT_BM);
        for (int ks = 0; ks < n_k; ++ks) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t s0 = base + stage * stage_bytes;
          mbar_expect_tx_elect(bar_full + 8 * stage, 2 * a_bytes + 2 * b_bytes);
          tma_load_2d_elect(s0, &tmap_a, ks * GT_BK, row0, bar_full + 8 * stage);                       // A hi
          tma_load_2d_elect(s0 + a_bytes, &tmap_a, p.K + ks * GT_BK, row0, bar_full + 8 * stage);       // A lo
          tma_load_2d_elect(s0 + b_off, &tmap_b, ks * GT_BK, 0, bar_full + 8 * stage);                  // B hi
          tma_load_2d_elect(s0 + b_off + b_stride, &tmap_b, p.K + ks * GT_BK, 0, bar_full + 8 * stage); // B lo
          if (++stage == p.nst) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {   // the whole warp walks the loop, an elected lane issues (tc_common.cuh: tc_mma_f16_elect)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(GT_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, seq = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
       for (int k0 = 0; k0 < n_k; k0 += GT_SEG, ++seq) {
        const uint32_t acc = seq & 1;
        mbar_wait(bar_tempty + 8 * acc, ((seq >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t first = 1;
        const int k1 = min(n_k, k0 + GT_SEG);
        for (int ks = k0; ks < k1; ++ks) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t s0 = base + stage * stage_bytes;
          const uint64_t ah = umma_desc(s0, 1024, 2), al = umma_desc(s0 + a_bytes, 1024, 2);
          const uint64_t bh = umma_desc(s0 + b_off, 1024, 2), bl = umma_desc(s0 + b_off + b_stride, 1024, 2);
#pragma unroll
          for (int k4 = 0; k4 < GT_BK / 16; ++k4) { tc_mma_f16_elect(d_tmem, ah + 2 * k4, bh + 2 * k4, idesc, first ? 0u : 1u); first = 0; }
#pragma unroll
          for (int k4 = 0; k4 < GT_BK / 16; ++k4) tc_mma_f16_elect(d_tmem, al + 2 * k4, bh + 2 * k4, idesc, 1u);
#pragma unroll
          for (int k4 = 0; k4 < GT_BK / 16; ++k4) tc_mma_f16_elect(d_tmem, ah + 2 * k4, bl + 2 * k4, idesc, 1u);
          tc_commit_elect(bar_empty + 8 * stage);
          if (++stage == p.nst) { stage = 0; phase ^= 1; }
        }
        tc_commit_elect(bar_tfull + 8 * acc);
       }
      }
    }
  } else {
    const int q = warp & 3;
    const int r = 32 * q + lane;
    const float mul = p.alpha * (p.inv_scale ? *p.inv_scale : 1.f);
    uint32_t seq = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
     for (int k0 = 0; k0 < n_k; k0 += GT_SEG, ++seq) {
      const uint32_t acc = seq & 1;
      const int64_t row = t * GT_BM + r;
      const bool add = p.accumulate || k0 > 0;
      mbar_wait(bar_tfull + 8 * acc, (seq >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + acc * 256 + ((uint32_t)(32 * q) << 16);
      for (int c = 0; c < p.N; c += 16) {
        uint32_t v[16];
        tc_ld16(trow + c, v);
        tc_ld_wait();
        if (row < p.M) {
          float4* dst = reinterpret_cast<float4*>(p.C + row * p.ldc + c);
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            float4 o = make_float4(__uint_as_float(v[4 * w]) * mul, __uint_as_float(v[4 * w + 1]) * mul,
                                   __uint_as_float(v[4 * w + 2]) * mul, __uint_as_float(v[4 * w + 3]) * mul);
            if (add) {
              const float4 old = dst[w];
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            dst[w] = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
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

// ---------------------------------------------------------------- C[M,N] += A^T B with both operands stored K-outer
// dP = G^T E (NCE backward): A = G chunk [K pixels, M prototypes], B = E chunk [K pixels, N dims], both row-major
// with the reduction index as the ROW -- "MN-major" operands for tcgen05 (instruction-descriptor bits 15/16).  A
// 128B-swizzled TMA box of {64 elements of a row} x {64 rows} is then exactly one canonical MN-major atom column:
// rows (K) 128 bytes apart, 8-row groups 1024 bytes apart (SBO), the next 64 MN elements one box (8 KiB) further
// (LBO).  So the SAME fp16 (hi|lo) copy of G that feeds dE = G P K-major feeds this product with no transposed
// copy (r1: split_transpose of every G chunk, 8 bytes of HBM traffic per element of G).
// Work items = (128-row tile of C, K split): P = 12288 has only 96 tiles for 148 SMs, so the reduction is cut into
// `ksplit` ranges, each accumulating into its own partial C (fixed owner per launch: deterministic); the caller adds
// the partials.  Same three fp16 passes, TMEM segments of GT_SEG slabs and fp32 read-modify-write epilogue as above.
struct GemmTnParams {
  int64_t M;                 // rows of C
  int N;                     // columns of C: 64, 128, 192 or 256
  int64_t n_k;               // K slabs of 64 (rows beyond the tensor maps read as zero)
  int a_lo_col, b_lo_col;    // column of the lo halves inside a row of A / B
  int64_t b_row0;            // first row of B (A starts at row 0)
  float* C;                  // [ksplit][M][ldc]
  int64_t ldc, split_stride;
  int ksplit;
  const float* inv_scale;
  float alpha;
  int nst;
};

__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;            // between 64-element blocks along M / N
  d |= (uint64_t)(sbo_bytes >> 4) << 32;            // between 8-row groups along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const GemmTnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t BOX = 64 * 128;                           // {64 elements} x {64 rows} of fp16
  const int nb = p.N / 64;                                     // boxes per B half
  const uint32_t a_bytes = 2 * BOX, b_bytes = (uint32_t)nb * BOX;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  uint8_t* misc = smem_raw + (base - smem_u32(smem_raw)) + (size_t)p.nst * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  const uint32_t bar_full = smem_u32(bars);            // [8]
  const uint32_t bar_empty = bar_full + 64;            // [8]
  const uint32_t bar_tfull = bar_empty + 64;           // [2]
  const uint32_t bar_tempty = bar_tfull + 16;          // [2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nst; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t n_tiles = (p.M + GT_BM - 1) / GT_BM;
  const int64_t items = n_tiles * p.ksplit;

  if (warp == 0) {
    {   // the whole warp walks the loop, an elected lane issues the copies
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const int64_t t = it / p.ksplit;
        const int ks = (int)(it - t * p.ksplit);
        const int64_t k_begin = p.n_k * ks / p.ksplit, k_end = p.n_k * (ks + 1) / p.ksplit;
        const int col0 = (int)(t * GT_BM);
        for (int64_t kk = k_begin; kk < k_end; ++kk) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t s0 = base + stage * stage_bytes;
          const uint32_t bar = bar_full + 8 * stage;
          mbar_expect_tx_elect(bar, stage_bytes);
          const int row = (int)(kk * GT_BK);
          for (int h = 0; h < 2; ++h) {
            tma_load_2d_elect(s0 + h * BOX, &tmap_a, col0 + 64 * h, row, bar);                          // A hi
            tma_load_2d_elect(s0 + a_bytes + h * BOX, &tmap_a, p.a_lo_col + col0 + 64 * h, row, bar);   // A lo
          }
          for (int h = 0; h < nb; ++h) {
            tma_load_2d_elect(s0 + 2 * a_bytes + h * BOX, &tmap_b, 64 * h, (int)p.b_row0 + row, bar);                        // B hi
            tma_load_2d_elect(s0 + 2 * a_bytes + b_bytes + h * BOX, &tmap_b, p.b_lo_col + 64 * h, (int)p.b_row0 + row, bar); // B lo
          }
          if (++stage == p.nst) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {   // the whole warp walks the loop, an elected lane issues (tc_common.cuh: tc_mma_f16_elect)
      // D = f32, A = B = f16, BOTH MN-major (bits 15, 16)
      const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(GT_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, seq = 0;
      for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const int64_t t = it / p.ksplit;
        const int ks = (int)(it - t * p.ksplit);
        const int64_t k_begin = p.n_k * ks / p.ksplit, k_end = p.n_k * (ks + 1) / p.ksplit;
        for (int64_t k0 = k_begin; k0 < k_end; k0 += GT_SEG, ++seq) {
          const uint32_t acc = seq & 1;
          mbar_wait(bar_tempty + 8 * acc, ((seq >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 256;
          uint32_t first = 1;
          const int64_t k1 = min(k_end, k0 + GT_SEG);
          for (int64_t kk = k0; kk < k1; ++kk) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t s0 = base + stage * stage_bytes;
            const uint64_t ah = umma_desc_mn(s0, BOX, 1024), al = umma_desc_mn(s0 + a_bytes, BOX, 1024);
            const uint64_t bh = umma_desc_mn(s0 + 2 * a_bytes, BOX, 1024), bl = umma_desc_mn(s0 + 2 * a_bytes + b_bytes, BOX, 1024);
            // one K = 16 step = two 8-row groups = 2048 bytes = 128 descriptor units
#pragma unroll
            for (int k4 = 0; k4 < GT_BK / 16; ++k4) { tc_mma_f16_elect(d_tmem, ah + 128 * k4, bh + 128 * k4, idesc, first ? 0u : 1u); first = 0; }
#pragma unroll
            for (int k4 = 0; k4 < GT_BK / 16; ++k4) tc_mma_f16_elect(d_tmem, al + 128 * k4, bh + 128 * k4, idesc, 1u);
#pragma unroll
            for (int k4 = 0; k4 < GT_BK / 16; ++k4) tc_mma_f16_elect(d_tmem, ah + 128 * k4, bl + 128 * k4, idesc, 1u);
            tc_commit_elect(bar_empty + 8 * stage);
            if (++stage == p.nst) { stage = 0; phase ^= 1; }
          }
          tc_commit_elect(bar_tfull + 8 * acc);
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int r = 32 * q + lane;
    const float mul = p.alpha * (p.inv_scale ? *p.inv_scale : 1.f);
    uint32_t seq = 0;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
      const int64_t t = it / p.ksplit;
      const int ks = (int)(it - t * p.ksplit);
      const int64_t k_begin = p.n_k * ks / p.ksplit, k_end = p.n_k * (ks + 1) / p.ksplit;
      const int64_t row = t * GT_BM + r;
      float* crow = p.C + ks * p.split_stride + row * p.ldc;
      for (int64_t k0 = k_begin; k0 < k_end; k0 += GT_SEG, ++seq) {
        const uint32_t acc = seq & 1;
        mbar_wait(bar_tfull + 8 * acc, (seq >> 1) & 1);
        tc_fence_after();
        const uint32_t trow = tmem_base + acc * 256 + ((uint32_t)(32 * q) << 16);
        for (int c = 0; c < p.N; c += 16) {
          uint32_t v[16];
          tc_ld16(trow + c, v);
          tc_ld_wait();
          if (row < p.M) {
            float4* dst = reinterpret_cast<float4*>(crow + c);
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const float4 old = dst[w];
              dst[w] = make_float4(old.x + __uint_as_float(v[4 * w]) * mul, old.y + __uint_as_float(v[4 * w + 1]) * mul,
                                   old.z + __uint_as_float(v[4 * w + 2]) * mul, old.w + __uint_as_float(v[4 * w + 3]) * mul);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
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

// out[i] (+)= sum over the K splits
__global__ void sum_splits_kernel(const float* __restrict__ part, int64_t n, int ksplit, int64_t stride, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
  for (int k = 0; k < ksplit; ++k) a += part[k * stride + i];
  out[i] = a;
}

// ---------------------------------------------------------------- operand preparation
// dst[r, c] = fp16(s x), dst[r, Kp + c] = fp16(s x - hi) for c < C, zero for C <= c < Kp;  s = mul * (*dev_mul)
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ src, int64_t ld, int64_t R, int C, int Kp,
                                                         float mul, const float* __restrict__ dev_mul,
                                                         __half* __restrict__ dst) {
  const float s = mul * (dev_mul ? *dev_mul : 1.f);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * Kp) return;
  const int64_t r = i / Kp;
  const int c = (int)(i - r * Kp);
  const float v = c < C ? src[r * ld + c] * s : 0.f;
  const __half hi = __float2half_rn(v);
  dst[r * 2 * Kp + c] = hi;
  dst[r * 2 * Kp + Kp + c] = __float2half_rn(v - __half2float(hi));
}

// transpose + split: dst[c, r] = hi(s x[r, c]), dst[c, Rp + r] = lo, for src [R, C]; rows r >= R are zero
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ src, int64_t ld, int64_t R, int C, int64_t Rp,
                                                              float mul, const float* __restrict__ dev_mul,
                                                              __half* __restrict__ dst) {
  __shared__ float tile[32][33];
  const float s = mul * (dev_mul ? *dev_mul : 1.f);
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 8 rows of 32 threads
  for (int k = ty; k < 32; k += 8) {
    const int64_t r = r0 + k;
    const int c = c0 + tx;
    tile[k][tx] = (r < R && c < C) ? src[r * ld + c] * s : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k;
    const int64_t r = r0 + tx;
    if (c < C && r < Rp) {
      const float v = tile[tx][k];
      const __half hi = __float2half_rn(v);
      dst[(int64_t)c * 2 * Rp + r] = hi;
      dst[(int64_t)c * 2 * Rp + Rp + r] = __float2half_rn(v - __half2float(hi));
    }
  }
}

__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));   // non-negative floats order like ints
}

// ---------------------------------------------------------------- host side
bool gemm_tc_supported(int N, int K) { return N >= 16 && N <= 256 && N % 16 == 0 && K > 0 && K % GT_BK == 0; }

int gemm_tc_split(const __half* a2, const __half* b2, int64_t M, int N, int K, float* C, int64_t ldc,
                  const float* inv_scale, float alpha, bool accumulate, cudaStream_t st) {
  HSG_REQUIRE(gemm_tc_supported(N, K), HSG_E_UNSUPPORTED, "gemm_tc: N=%d K=%d", N, K);
  HSG_REQUIRE(M > 0 && M < (1ll << 31) && ldc % 4 == 0, HSG_E_INVALID, "gemm_tc: M=%lld ldc=%lld", (long long)M, (long long)ldc);
  GemmTcParams p;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.inv_scale = inv_scale; p.alpha = alpha; p.accumulate = accumulate ? 1 : 0;
  const size_t b_bytes = ((size_t)N * GT_BK * 2 + 1023) / 1024 * 1024;
  const size_t stage = 2 * (size_t)GT_BM * GT_BK * 2 + 2 * b_bytes;
  int nst = (int)((227 * 1024 - 2048 - 512) / stage);
  if (nst > 8) nst = 8;
  HSG_REQUIRE(nst >= 2, HSG_E_UNSUPPORTED, "gemm_tc: shared memory budget");
  p.nst = nst;
  const size_t smem = 1024 + (size_t)nst * stage + 512;
  CUtensorMap ma, mb;
  int rc;
  if ((rc = encode_2d_f16(&ma, a2, (uint64_t)M, (uint64_t)2 * K, GT_BK, GT_BM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = encode_2d_f16(&mb, b2, (uint64_t)N, (uint64_t)2 * K, GT_BK, (uint32_t)N, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  HSG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = num_sms();
  const int64_t tiles = ceil_div64(M, GT_BM);
  if (grid > tiles) grid = tiles;
  gemm_tc_kernel<<<(unsigned)grid, GT_THREADS, smem, st>>>(ma, mb, p);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

bool gemm_tn_tc_supported(int N) { return N == 64 || N == 128 || N == 192 || N == 256; }

int gemm_tn_ksplit(int64_t M, int64_t n_k) {
  const int64_t tiles = ceil_div64(M, GT_BM);
  int64_t ks = ceil_div64(2 * (int64_t)num_sms(), tiles);
  if (ks > 8) ks = 8;
  if (ks > n_k) ks = n_k;
  return ks < 1 ? 1 : (int)ks;
}

// part[ks][M][N] += (A[0:K, a cols]^T . B[b_row0 : b_row0 + K, :]) / scales, for the K range of split ks.
// a2: [a_rows, 2 * a_half] fp16 (hi | lo), b2: [b_rows, 2 * b_half]; K = 64 * n_k rows (missing rows read as zero).
int gemm_tn_tc_split(const __half* a2, int64_t a_rows, int a_half, const __half* b2, int64_t b_rows, int b_half,
                     int64_t b_row0, int64_t M, int N, int64_t n_k, int ksplit, float* part, const float* inv_scale,
                     cudaStream_t st) {
  HSG_REQUIRE(gemm_tn_tc_supported(N) && N <= b_half, HSG_E_UNSUPPORTED, "gemm_tn_tc: N=%d", N);
  HSG_REQUIRE(M > 0 && n_k > 0 && ksplit >= 1 && ksplit <= 8 && a_rows < (1ll << 31) && b_rows < (1ll << 31), HSG_E_INVALID,
              "gemm_tn_tc: bad shape");
  GemmTnParams p;
  p.M = M; p.N = N; p.n_k = n_k; p.a_lo_col = a_half; p.b_lo_col = b_half; p.b_row0 = b_row0;
  p.C = part; p.ldc = N; p.split_stride = M * N; p.ksplit = ksplit; p.inv_scale = inv_scale; p.alpha = 1.f;
  const size_t stage = 2 * (size_t)2 * 8192 + 2 * (size_t)(N / 64) * 8192;
  int nst = (int)((227 * 1024 - 2048 - 512) / stage);
  if (nst > 8) nst = 8;
  HSG_REQUIRE(nst >= 2, HSG_E_UNSUPPORTED, "gemm_tn_tc: shared memory budget");
  p.nst = nst;
  const size_t smem = 1024 + (size_t)nst * stage + 512;
  CUtensorMap ma, mb;
  int rc;
  if ((rc = encode_2d_f16(&ma, a2, (uint64_t)a_rows, (uint64_t)2 * a_half, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = encode_2d_f16(&mb, b2, (uint64_t)b_rows, (uint64_t)2 * b_half, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  HSG_CUDA(cudaFuncSetAttribute(gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = num_sms();
  const int64_t items = ceil_div64(M, GT_BM) * ksplit;
  if (grid > items) grid = items;
  gemm_tn_tc_kernel<<<(unsigned)grid, GT_THREADS, smem, st>>>(ma, mb, p);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int sum_splits(const float* part, int64_t n, int ksplit, float* out, cudaStream_t st) {
  sum_splits_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(part, n, ksplit, n, out);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int split_rows(const float* src, int64_t ld, int64_t R, int C, int Kp, float mul, const float* dev_mul, __half* dst,
               cudaStream_t st) {
  split_rows_kernel<<<(unsigned)ceil_div64(R * Kp, 256), 256, 0, st>>>(src, ld, R, C, Kp, mul, dev_mul, dst);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int split_transpose(const float* src, int64_t ld, int64_t R, int C, int64_t Rp, float mul, const float* dev_mul,
                    __half* dst, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div64(Rp, 32), (unsigned)ceil_div64(C, 32));
  split_transpose_kernel<<<grid, 256, 0, st>>>(src, ld, R, C, Rp, mul, dev_mul, dst);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int absmax(const float* x, int64_t n, float* out, cudaStream_t st) {
  HSG_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  absmax_kernel<<<296, 256, 0, st>>>(x, n, out);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

}  // namespace hsg
