// K4 on tensor cores: forward of the pixel-to-prototype NCE loss
// (_calculate_log_likelihood, hsg/utils/segsort/loss.py:15-82) as a flash-style
// tcgen05 kernel: S = exp(c E P^T) is produced tile by tile in TMEM and reduced
// per pixel and per label set in the epilogue; nothing of size [N,P] exists.
//
// Precision.  The loss needs fp32-grade similarities (1e-5 relative on the loss
// => ~1e-6 absolute on <e,p> at concentration 16), which a single fp16/bf16 pass
// cannot give.  Both operands are split into fp16 (hi, lo) pairs after scaling
// each by t = sqrt(c log2 e) -- the accumulator is then directly the base-2
// exponent, and the lo parts stay (mostly) normal numbers:
//     t^2 <e,p> ~ eh.ph + el.ph + eh.pl      (el.pl ~ 2^-22 is dropped)
// i.e. three K=D passes accumulated into the same TMEM tile; error ~3*2^-22.
// Rows are stored as [hi | lo] (2D fp16 per row) so a pixel tile [128 x 2D] stays
// resident in shared memory for a whole sweep over the prototypes, which stream
// through a ring of [128 x 64] slabs; each ph slab is used twice (eh and el).
//
// Per pixel and label set the epilogue accumulates  pos = sum_{same class} S
// (own prototype included) and neg = sum_{other class} S, and captures own =
// S[i, inst_i]; the finish is literally the reference's: num = pos - own if that
// is > 0 else own, den = neg + num, l = -log(num/den).
#include "tc_common.cuh"

#include <stdlib.h>

#include <float.h>
#include <limits.h>

namespace hsg {

constexpr int NT_BM = 128;                 // pixels per tile
constexpr int NT_BN = 128;                 // prototype rows per slab (one CTA's half of a tile)
constexpr int NT_BK = 64;
constexpr int NT_SLAB = NT_BM * NT_BK * 2; // 16 KiB
constexpr int NT_THREADS = 384;
constexpr int NT_MAX_SETS = 4;

struct NceTcParams {
  int64_t N, P;
  const int64_t* P_dev;     // optional: the number of valid prototypes lives on the device (<= P, the capacity of the
                            // arrays); rows and labels beyond it are never counted -- no host read before the launch
  int D;                    // 64, 128 or 256
  int n_sets;
  int plus[NT_MAX_SETS];
  const int32_t* inst;      // [N]
  const int32_t* sem;       // [n_sets,N]
  const int32_t* psem;      // [n_sets,Ppad]  (Ppad = P rounded up to 128; padding never read as valid)
  int64_t Ppad;
  float* per_pixel;         // [n_sets,N]
  float* stats;             // [n_sets,N,4] or NULL
  int nstb;                 // stages of the prototype ring
};

__device__ __forceinline__ float ex2_approx(float x) {   // MUFU.EX2, relative error <= 2^-22
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(256) nce_split_kernel(const float* __restrict__ x, int64_t rows, int D, float mul,
                                                        __half* __restrict__ out) {
  // out[r, d] = fp16(mul x), out[r, D + d] = fp16(mul x - hi); eight elements per thread (D % 8 == 0)
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= rows * D) return;
  const int64_t r = i / D;
  const int d = (int)(i - r * D);
  const float4 a = ld_stream4(x + i), b = ld_stream4(x + i + 4);
  const float v[8] = {a.x * mul, a.y * mul, a.z * mul, a.w * mul, b.x * mul, b.y * mul, b.z * mul, b.w * mul};
  __align__(16) __half hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    hi[j] = __float2half_rn(v[j]);
    lo[j] = __float2half_rn(v[j] - __half2float(hi[j]));
  }
  __half* o = out + r * 2 * D + d;
  *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(o + D) = *reinterpret_cast<const uint4*>(lo);
}

__global__ void nce_labels32_kernel(const int64_t* __restrict__ src, int64_t n_src, int64_t n_dst, int rows,
                                    int32_t* __restrict__ dst, int32_t pad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_dst * rows) return;
  const int64_t r = i / n_dst, c = i % n_dst;
  dst[i] = c < n_src ? (int32_t)src[r * n_src + c] : pad;
}

// One accumulator tile [128 lanes x ncols] of this thread's TMEM row: exp2, label masks, sums.
// TMEM loads (and, for up to two label sets, the label loads) of chunk c+1 are in flight while
// chunk c is reduced.
template <int NS>
struct EpiChunk {
  static constexpr int NL = NS <= 2 ? NS : 1;     // labels ride with the chunk only for <= 2 sets (registers)
  uint32_t v[16];
  int4 lab[NL][4];
};

template <int NS>
__device__ __forceinline__ void epi_labels(int4 (&lab)[4], const int32_t* lab_ptr, int64_t Ppad, int s) {
  const int4* lp = reinterpret_cast<const int4*>(lab_ptr + (int64_t)s * Ppad);
#pragma unroll
  for (int w = 0; w < 4; ++w) lab[w] = __ldg(lp + w);
}

template <int NS>
__device__ __forceinline__ void epi_issue(EpiChunk<NS>& ch, uint32_t taddr, const int32_t* lab, int64_t Ppad) {
  tc_ld16(taddr, ch.v);
  if constexpr (NS <= 2) {
#pragma unroll
    for (int s = 0; s < NS; ++s) epi_labels<NS>(ch.lab[s], lab, Ppad, s);
  }
}

template <int NS>
__device__ __forceinline__ void epi_reduce(EpiChunk<NS>& ch, const int32_t* lab_ptr, int64_t Ppad, int col0,
                                           int n_valid, int own_rel, const int (&my_sem)[NS], float (&pos)[NS], float (&neg)[NS], float& own) {
  float sv[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) sv[j] = ex2_approx(__uint_as_float(ch.v[j]));
  if (col0 + 16 > n_valid) {
#pragma unroll
    for (int j = 0; j < 16; ++j) if (col0 + j >= n_valid) sv[j] = 0.f;
  }
  const int rel = own_rel - col0;
  if ((unsigned)rel < 16u) {
#pragma unroll
    for (int j = 0; j < 16; ++j) if (j == rel) own = sv[j];
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    if constexpr (NS > 2) epi_labels<NS>(ch.lab[0], lab_ptr, Ppad, s);
    const int4 (&lab)[4] = ch.lab[NS <= 2 ? s : 0];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      if (lab[w].x == my_sem[s]) pos[s] += sv[4 * w + 0]; else neg[s] += sv[4 * w + 0];
      if (lab[w].y == my_sem[s]) pos[s] += sv[4 * w + 1]; else neg[s] += sv[4 * w + 1];
      if (lab[w].z == my_sem[s]) pos[s] += sv[4 * w + 2]; else neg[s] += sv[4 * w + 2];
      if (lab[w].w == my_sem[s]) pos[s] += sv[4 * w + 3]; else neg[s] += sv[4 * w + 3];
    }
  }
}

// Two CTAs of a cluster form one tcgen05 cta_group::2 tile of 256 pixels x 256 prototypes:
// each CTA keeps its own 128 pixel rows [128 x 2D] resident and streams only ITS half (128
// rows) of every prototype tile; the tensor cores read the other half from the peer's shared
// memory, so L2->SM traffic per flop is half that of a 128x128 single-CTA tile (which sat at
// the chip's ~6300 B/clk L2 budget).  The leader (cluster rank 0) issues every MMA; both CTAs
// run a TMA producer (crediting the leader's "full" barriers) and an epilogue over their own
// 128 TMEM lanes.
constexpr int N2_BN = 256;                  // prototypes per accumulator tile (128 per CTA of the pair)
constexpr int N2_ACC = 2;                   // TMEM accumulator buffers (2 x 256 columns)

template <int NS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT_THREADS, 1)
nce_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const NceTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nslab = p.D / NT_BK;                       // slabs per half (hi or lo)
  const int64_t P_eff = p.P_dev ? max((int64_t)0, min(p.P, *p.P_dev)) : p.P;
  const uint32_t sA = base;                            // [2*nslab] slabs: eh_0.., el_0..
  const uint32_t sB = sA + 2 * nslab * NT_SLAB;        // ring of nstb slabs
  const uint32_t sMisc = sB + p.nstb * NT_SLAB;
  uint8_t* misc = smem_raw + (sMisc - smem_u32(smem_raw));
  float* ex = reinterpret_cast<float*>(misc);                               // [NT_BM][2*NT_MAX_SETS+1]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ex + NT_BM * (2 * NT_MAX_SETS + 1) + 1);
  bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bars) + 7) & ~uintptr_t(7));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
  const uint32_t bar_bfull = smem_u32(bars);            // [8]  (waited on in the leader only)
  const uint32_t bar_bempty = bar_bfull + 64;           // [8]
  const uint32_t bar_afull = bar_bempty + 64;           // [1]  (leader only)
  const uint32_t bar_aempty = bar_afull + 8;            // [1]
  const uint32_t bar_tfull = bar_aempty + 8;            // [2]
  const uint32_t bar_tempty = bar_tfull + 16;           // [2]  (leader only; 8 arrivals: 4 warps x 2 CTAs)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nstb; ++i) { mbar_init(bar_bfull + 8 * i, 1); mbar_init(bar_bempty + 8 * i, 1); }
    mbar_init(bar_afull, 1);
    mbar_init(bar_aempty, 1);
    for (int i = 0; i < N2_ACC; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // the peer's barriers exist before anything remote touches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t n_ptiles = (p.N + 2 * NT_BM - 1) / (2 * NT_BM);     // pair tiles of 256 pixels
  // prototype tiles: with a device-side count only the tiles that hold valid prototypes are visited (every role
  // reads the same count, so producer, issuer and epilogue agree); at least one, so the pipeline protocol is the same
  const int n_ntiles_cap = (int)(p.Ppad / N2_BN);
  const int n_ntiles = p.P_dev ? max(1, min(n_ntiles_cap, (int)((P_eff + N2_BN - 1) / N2_BN))) : n_ntiles_cap;
  const int64_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    {   // the whole warp walks the loop, an elected lane issues the copies
      const uint32_t l_afull = mapa_cluster(bar_afull, 0);
      const uint32_t l_bfull = mapa_cluster(bar_bfull, 0);
      int stage = 0;
      uint32_t phase = 0, a_round = 0;
      for (int64_t pt = pair; pt < n_ptiles; pt += n_pairs, ++a_round) {
        mbar_wait(bar_aempty, (a_round & 1) ^ 1);               // previous pixel tile fully consumed
        if (leader) mbar_expect_tx_elect(bar_afull, 2 * 2 * nslab * NT_SLAB);     // both CTAs' rows
        const int row0 = (int)(pt * 2 * NT_BM + rank * NT_BM);
        for (int j = 0; j < 2 * nslab; ++j)
          tma_load_2d_pair_elect(sA + j * NT_SLAB, &tmap_a, j * NT_BK, row0, l_afull);
        for (int nt = 0; nt < n_ntiles; ++nt) {
          const int prow0 = nt * N2_BN + (int)rank * NT_BN;
          for (int j = 0; j < 2 * nslab; ++j) {                  // ph_0.., then pl_0..
            mbar_wait(bar_bempty + 8 * stage, phase ^ 1);
            if (leader) mbar_expect_tx_elect(bar_bfull + 8 * stage, 2 * NT_SLAB);
            tma_load_2d_pair_elect(sB + stage * NT_SLAB, &tmap_b, j * NT_BK, prow0, l_bfull + 8 * stage);
            if (++stage == p.nstb) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    // the whole warp walks the loop, an elected lane issues (tc_common.cuh: tc_mma_f16_elect)
    if (uniform_i32(leader ? 1 : 0)) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N2_BN >> 3) << 17) | ((uint32_t)((2 * NT_BM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, a_round = 0, seq = 0;
      for (int64_t pt = pair; pt < n_ptiles; pt += n_pairs, ++a_round) {
        mbar_wait(bar_afull, a_round & 1);
        tc_fence_after();
        for (int nt = 0; nt < n_ntiles; ++nt, ++seq) {
          const uint32_t acc = seq & 1;
          mbar_wait(bar_tempty + 8 * acc, ((seq >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * N2_BN;
          uint32_t first = 1;
          for (int j = 0; j < 2 * nslab; ++j) {
            mbar_wait(bar_bfull + 8 * stage, phase);
            tc_fence_after();
            const uint32_t b0 = sB + stage * NT_SLAB;
            const int js = j < nslab ? j : j - nslab;             // matching slab of eh
            const int passes = j < nslab ? 2 : 1;                 // ph: eh.ph and el.ph ; pl: eh.pl
            // descriptors differ only in the 14-bit start-address field: build once, then add
            const uint64_t bd = umma_desc(b0, 1024, 2);
            for (int ps = 0; ps < passes; ++ps) {
              const uint64_t ad = umma_desc(sA + (ps * nslab + js) * NT_SLAB, 1024, 2);
#pragma unroll
              for (int k4 = 0; k4 < NT_BK / 16; ++k4) {
                tc_mma_f16_pair_elect(d_tmem, ad + 2 * k4, bd + 2 * k4, idesc, first ? 0u : 1u);   // +32 bytes per K=16 step
                first = 0;
              }
            }
            tc_commit_pair_elect(bar_bempty + 8 * stage);
            if (++stage == p.nstb) { stage = 0; phase ^= 1; }
          }
          tc_commit_pair_elect(bar_tfull + 8 * acc);
        }
        tc_commit_pair_elect(bar_aempty);                                // arrives when every MMA of this pixel tile retired
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 TMEM lanes) =====================
    // Two groups of four warps; group g owns accumulator buffer g, i.e. every second prototype
    // tile.  A thread owns one pixel row and all 256 columns of its tiles; prototype labels are
    // read straight from global memory (warp-uniform addresses, L1 resident).
    const int q = warp & 3, g = (warp - 4) >> 2;
    const int r = 32 * q + lane;
    const uint32_t l_tempty = mapa_cluster(bar_tempty, 0);
    uint32_t seq = 0;
    for (int64_t pt = pair; pt < n_ptiles; pt += n_pairs) {
      const int64_t pix = pt * 2 * NT_BM + rank * NT_BM + r;
      const bool inb = pix < p.N;
      int my_sem[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) my_sem[s] = inb ? p.sem[(int64_t)s * p.N + pix] : INT_MIN;
      const int my_inst = inb ? p.inst[pix] : -1;
      float pos[NS], neg[NS], own = 0.f;
#pragma unroll
      for (int s = 0; s < NS; ++s) { pos[s] = 0.f; neg[s] = 0.f; }

      for (int nt = 0; nt < n_ntiles; ++nt, ++seq) {
        if ((int)(seq & 1) != g) continue;
        const int n_valid = (int)min((int64_t)N2_BN, P_eff - (int64_t)nt * N2_BN);
        const int nchunk = (max(n_valid, 1) + 15) >> 4;                  // a tile beyond a device-side count still drains once
        const int own_rel = my_inst - nt * N2_BN;                        // column of the own prototype in this tile
        const int32_t* lab_tile = p.psem + (int64_t)nt * N2_BN;

        mbar_wait(bar_tfull + 8 * g, (seq >> 1) & 1);
        tc_fence_after();
        const uint32_t trow = tmem_base + g * N2_BN + ((uint32_t)(32 * q) << 16);
        EpiChunk<NS> ca, cb;
        epi_issue<NS>(ca, trow, lab_tile, p.Ppad);
#pragma unroll 1
        for (int c = 0; c < nchunk; c += 2) {
          tc_ld_wait();
          if (c + 1 < nchunk) epi_issue<NS>(cb, trow + (c + 1) * 16, lab_tile + (c + 1) * 16, p.Ppad);
          epi_reduce<NS>(ca, lab_tile + c * 16, p.Ppad, c * 16, n_valid, own_rel, my_sem, pos, neg, own);
          if (c + 1 < nchunk) {
            tc_ld_wait();
            if (c + 2 < nchunk) epi_issue<NS>(ca, trow + (c + 2) * 16, lab_tile + (c + 2) * 16, p.Ppad);
            epi_reduce<NS>(cb, lab_tile + (c + 1) * 16, p.Ppad, (c + 1) * 16, n_valid, own_rel, my_sem, pos, neg, own);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(l_tempty + 8 * g);
      }

      // combine the two groups' partial sums and finish like the reference
      float* row = ex + r * (2 * NT_MAX_SETS + 1);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (g == 1) {
#pragma unroll
        for (int s = 0; s < NS; ++s) { row[2 * s] = pos[s]; row[2 * s + 1] = neg[s]; }
        row[2 * NT_MAX_SETS] = own;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (g == 0 && inb) {
        own += row[2 * NT_MAX_SETS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const float pp = pos[s] + row[2 * s], nn = neg[s] + row[2 * s + 1];
          const bool own_same = my_inst >= 0 && my_inst < P_eff && p.psem[(int64_t)s * p.Ppad + my_inst] == my_sem[s];
          float num = own, flags = own_same ? 2.f : 0.f;
          if (p.plus[s]) {
            const float ps2 = __fsub_rn(pp, own);            // loss.py:64-66: sum over the class, then subtract own
            if (ps2 > 0.f) { num = ps2; flags += 1.f; }
          }
          const float den = nn + num;
          p.per_pixel[(int64_t)s * p.N + pix] = -logf(num / den);
          if (p.stats) {
            float* st = p.stats + ((int64_t)s * p.N + pix) * 4;
            st[0] = num; st[1] = den; st[2] = own; st[3] = flags;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // neither CTA leaves (or frees TMEM) while the pair still works
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------- backward: G chunk on tensor cores
// G[i, j] = c * S_ij * sum_s w_si * dl_si/dS_ij for the pixels [i_begin, i_end) (the
// fp32 CUDA-core version is nce_grad_kernel in nce.cu).  Same CTA-pair tiles and three fp16 passes as
// the forward; the work items are (pixel pair-tile, prototype tile) pairs so that a chunk of a few
// thousand pixels still fills the machine, and the epilogue writes the tile instead of reducing it.
struct NceGradTcParams {
  NceTcParams f;
  int64_t i_begin, i_end;
  const float* stats;        // [n_sets,N,4] num, den, own, flags (forward)
  const float* w;            // [n_sets,N] upstream gradient per pixel
  float conc;
  float* G;                  // [i_end - i_begin, ldg] fp32, or NULL:
  int64_t ldg;               // multiple of 64, >= P; columns >= P are written as zeros
  __half* G2;                // [i_end - i_begin, 2 * ldg] = fp16 (hi | lo) of G * (*gscale): the operand layout of the
  const float* gscale;       //   two backward GEMMs, written straight from the accumulator (no fp32 round trip)
  int exp_flags;             // timing experiments (HSG_NCE_EXP), 0 in production: 1 = no stores, 2 = direct (uncoalesced) stores
  uint32_t stg_off;          // staging area of the coalesced G2 stores, bytes from the aligned base (0 = none)
};
constexpr int NG_STG_ROW = 80;                       // bytes per staged row: 32 halves + 16 bytes of padding (conflict-free)
constexpr int NG_STG_WARP = 2 * 32 * NG_STG_ROW;     // hi and lo blocks of one warp: 32 rows x 32 columns each

template <int NS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT_THREADS, 1)
nce_grad_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const NceGradTcParams g) {
  const NceTcParams& p = g.f;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nslab = p.D / NT_BK;
  const int64_t P_eff = p.P_dev ? max((int64_t)0, min(p.P, *p.P_dev)) : p.P;
  const uint32_t sA = base;
  const uint32_t sB = sA + 2 * nslab * NT_SLAB;
  const uint32_t sMisc = sB + p.nstb * NT_SLAB;
  uint8_t* misc = smem_raw + (sMisc - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
  const uint32_t bar_bfull = smem_u32(bars);            // [8]  (waited on in the leader only)
  const uint32_t bar_bempty = bar_bfull + 64;           // [8]
  const uint32_t bar_afull = bar_bempty + 64;           // [1]  (leader only)
  const uint32_t bar_aempty = bar_afull + 8;            // [1]
  const uint32_t bar_tfull = bar_aempty + 8;            // [2]
  const uint32_t bar_tempty = bar_tfull + 16;           // [2]  (leader only; 8 arrivals)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nstb; ++i) { mbar_init(bar_bfull + 8 * i, 1); mbar_init(bar_bempty + 8 * i, 1); }
    mbar_init(bar_afull, 1);
    mbar_init(bar_aempty, 1);
    for (int i = 0; i < N2_ACC; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t rows = g.i_end - g.i_begin;
  const int64_t n_ptiles = (rows + 2 * NT_BM - 1) / (2 * NT_BM);
  const int n_ntiles = (int)(p.Ppad / N2_BN);
  const int64_t items = n_ptiles * n_ntiles;
  const int64_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int64_t per = (items + n_pairs - 1) / n_pairs;
  const int64_t it0 = min(items, pair * per), it1 = min(items, it0 + per);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    {   // the whole warp walks the loop, an elected lane issues the copies
      const uint32_t l_afull = mapa_cluster(bar_afull, 0);
      const uint32_t l_bfull = mapa_cluster(bar_bfull, 0);
      int stage = 0;
      uint32_t phase = 0, a_round = 0;
      int64_t cur_pt = -1;
      for (int64_t it = it0; it < it1; ++it) {
        const int64_t pt = it / n_ntiles;
        const int nt = (int)(it - pt * n_ntiles);
        if (pt != cur_pt) {
          mbar_wait(bar_aempty, (a_round & 1) ^ 1);
          if (leader) mbar_expect_tx_elect(bar_afull, 2 * 2 * nslab * NT_SLAB);
          const int row0 = (int)(g.i_begin + pt * 2 * NT_BM + rank * NT_BM);
          for (int j = 0; j < 2 * nslab; ++j)
            tma_load_2d_pair_elect(sA + j * NT_SLAB, &tmap_a, j * NT_BK, row0, l_afull);
          cur_pt = pt;
          ++a_round;
        }
        const int prow0 = nt * N2_BN + (int)rank * NT_BN;
        for (int j = 0; j < 2 * nslab; ++j) {
          mbar_wait(bar_bempty + 8 * stage, phase ^ 1);
          if (leader) mbar_expect_tx_elect(bar_bfull + 8 * stage, 2 * NT_SLAB);
          tma_load_2d_pair_elect(sB + stage * NT_SLAB, &tmap_b, j * NT_BK, prow0, l_bfull + 8 * stage);
          if (++stage == p.nstb) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    // the whole warp walks the loop, an elected lane issues (tc_common.cuh: tc_mma_f16_elect)
    if (uniform_i32(leader ? 1 : 0)) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N2_BN >> 3) << 17) | ((uint32_t)((2 * NT_BM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, a_round = 0, seq = 0;
      int64_t cur_pt = -1;
      for (int64_t it = it0; it < it1; ++it, ++seq) {
        const int64_t pt = it / n_ntiles;
        if (pt != cur_pt) {
          mbar_wait(bar_afull, a_round & 1);
          tc_fence_after();
          cur_pt = pt;
          ++a_round;
        }
        const uint32_t acc = seq & 1;
        mbar_wait(bar_tempty + 8 * acc, ((seq >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * N2_BN;
        uint32_t first = 1;
        for (int j = 0; j < 2 * nslab; ++j) {
          mbar_wait(bar_bfull + 8 * stage, phase);
          tc_fence_after();
          const uint32_t b0 = sB + stage * NT_SLAB;
          const int js = j < nslab ? j : j - nslab;
          const int passes = j < nslab ? 2 : 1;
          const uint64_t bd = umma_desc(b0, 1024, 2);
          for (int ps = 0; ps < passes; ++ps) {
            const uint64_t ad = umma_desc(sA + (ps * nslab + js) * NT_SLAB, 1024, 2);
#pragma unroll
            for (int k4 = 0; k4 < NT_BK / 16; ++k4) {
              tc_mma_f16_pair_elect(d_tmem, ad + 2 * k4, bd + 2 * k4, idesc, first ? 0u : 1u);
              first = 0;
            }
          }
          tc_commit_pair_elect(bar_bempty + 8 * stage);
          if (++stage == p.nstb) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair_elect(bar_tfull + 8 * acc);
        // last item of this pixel tile (in this pair's slice): its rows may be replaced once these MMAs retire
        if (it + 1 == it1 || (it + 1) / n_ntiles != pt) tc_commit_pair_elect(bar_aempty);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 TMEM lanes) =====================
    const int q = warp & 3, grp = (warp - 4) >> 2;
    const int r = 32 * q + lane;
    const uint32_t l_tempty = mapa_cluster(bar_tempty, 0);
    uint32_t seq = 0;
    int64_t cur_pt = -1;
    int my_sem[NS];
    float ca[NS], cb[NS];        // coefficient of a negative / of another positive, per label set (c * w folded in)
    float c_own = 0.f;           // coefficient of the pixel's own prototype (all sets)
    int my_inst = -1;
    int64_t pix = 0;
    bool inb = false;
    // the operand scale of the fp16 copy rides in the per-pixel coefficients (one multiply per pixel and label set
    // instead of one per element); the fp32 G and the fp16 G2 are never requested together
    const float gsc0 = g.G2 ? *g.gscale : 1.f;
    for (int64_t it = it0; it < it1; ++it, ++seq) {
      const int64_t pt = it / n_ntiles;
      const int nt = (int)(it - pt * n_ntiles);
      if (pt != cur_pt) {
        cur_pt = pt;
        pix = g.i_begin + pt * 2 * NT_BM + rank * NT_BM + r;
        inb = pix < g.i_end;
        my_inst = inb ? p.inst[pix] : -1;
        c_own = 0.f;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          my_sem[s] = inb ? p.sem[(int64_t)s * p.N + pix] : INT_MIN;
          ca[s] = 0.f; cb[s] = 0.f;
          if (inb) {
            const float* st = g.stats + ((int64_t)s * p.N + pix) * 4;
            const float num = st[0], den = st[1];
            const int fl = (int)st[3];
            const bool use = fl & 1, own_same = fl & 2;
            const float ws = gsc0 * g.conc * g.w[(int64_t)s * p.N + pix];
            const float inv_den = 1.f / den, dlt = inv_den - 1.f / num;
            ca[s] = ws * inv_den;
            cb[s] = use ? ws * dlt : 0.f;
            const float w_own = use ? (own_same ? 0.f : -1.f) : 1.f;
            c_own += ws * ((own_same ? 0.f : inv_den) + w_own * dlt);
          }
        }
      }
      if ((int)(seq & 1) != grp) continue;
      const int n_valid = (int)min((int64_t)N2_BN, P_eff - (int64_t)nt * N2_BN);
      const int n_store = (int)min((int64_t)N2_BN, g.ldg - (int64_t)nt * N2_BN);      // columns of G that exist
      const int nchunk = (max(n_store, 0) + 15) >> 4;
      const int own_rel = my_inst - nt * N2_BN;
      const int32_t* lab_tile = p.psem + (int64_t)nt * N2_BN;
      float* grow = g.G ? g.G + (pix - g.i_begin) * g.ldg + (int64_t)nt * N2_BN : nullptr;
      __half* g2row = g.G2 ? g.G2 + (pix - g.i_begin) * 2 * g.ldg + (int64_t)nt * N2_BN : nullptr;
      // Coalesced G2 stores.  A thread owns one pixel row, so a direct 16-byte store of a warp touches 32 different
      // lines: the load/store unit, not the tensor pipe, paced this kernel (r2: 1.15 ms per chunk, 0.4 ms with the
      // stores removed).  Each warp stages 32 rows x 32 columns of hi and of lo in shared memory (row pitch 80 bytes:
      // conflict-free both ways) and writes them out 8 rows x 64 contiguous bytes per instruction.
      const bool staged = g2row && g.stg_off && !(g.exp_flags & 2);
      uint8_t* stg = smem_raw + (base - smem_u32(smem_raw)) + g.stg_off + (warp - 4) * NG_STG_WARP;
      const int64_t row0_local = pix - lane - g.i_begin;          // first row of this warp inside the chunk
      const int64_t rows_left = g.i_end - (pix - lane);           // rows of this warp that exist
      auto flush = [&](int col_first, int ncols) {                // ncols = 16 or 32 staged columns starting at col_first
        __syncwarp();
        const int sub = lane & 3, rsel = lane >> 2;
        if (sub * 8 < ncols) {
#pragma unroll
          for (int it4 = 0; it4 < 4; ++it4) {
            const int rr = it4 * 8 + rsel;
            if (rr < rows_left) {
              __half* dst = g.G2 + (row0_local + rr) * 2 * g.ldg + (int64_t)nt * N2_BN + col_first + sub * 8;
              const uint4 h = *reinterpret_cast<const uint4*>(stg + rr * NG_STG_ROW + sub * 16);
              const uint4 l = *reinterpret_cast<const uint4*>(stg + 32 * NG_STG_ROW + rr * NG_STG_ROW + sub * 16);
              *reinterpret_cast<uint4*>(dst) = h;
              *reinterpret_cast<uint4*>(dst + g.ldg) = l;
            }
          }
        }
        __syncwarp();
      };

      mbar_wait(bar_tfull + 8 * grp, (seq >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + grp * N2_BN + ((uint32_t)(32 * q) << 16);
      // chunk c + 1's TMEM load (and label loads) are in flight while chunk c is turned into G (r2: without the
      // overlap this kernel ran its tensor pipe at 28 %, the slowest of the three backward kernels)
      auto process = [&](EpiChunk<NS>& cur, EpiChunk<NS>& nxt, int c) {
        const int col0 = c * 16;
        tc_ld_wait();
        if (c + 1 < nchunk) epi_issue<NS>(nxt, trow + col0 + 16, lab_tile + col0 + 16, p.Ppad);
        float gv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) gv[j] = 0.f;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          if constexpr (NS > 2) epi_labels<NS>(cur.lab[0], lab_tile + col0, p.Ppad, s);
          const int4 (&lab)[4] = cur.lab[NS <= 2 ? s : 0];
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) {
            gv[4 * w4 + 0] += lab[w4].x == my_sem[s] ? cb[s] : ca[s];
            gv[4 * w4 + 1] += lab[w4].y == my_sem[s] ? cb[s] : ca[s];
            gv[4 * w4 + 2] += lab[w4].z == my_sem[s] ? cb[s] : ca[s];
            gv[4 * w4 + 3] += lab[w4].w == my_sem[s] ? cb[s] : ca[s];
          }
        }
        // the epilogue is ALU-bound (r2: tensor pipe 28 %): the own-prototype column and the padding columns are
        // handled off the common path, and the operand scale of G2 is folded into the coefficients
        const int rel = own_rel - col0;
        if ((unsigned)rel < 16u) {
#pragma unroll
          for (int j = 0; j < 16; ++j) if (j == rel) gv[j] = c_own;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) gv[j] *= ex2_approx(__uint_as_float(cur.v[j]));
        if (col0 + 16 > n_valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) if (col0 + j >= n_valid) gv[j] = 0.f;
        }
        if (inb && grow) {
          float4* dst = reinterpret_cast<float4*>(grow + col0);
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) dst[w4] = make_float4(gv[4 * w4], gv[4 * w4 + 1], gv[4 * w4 + 2], gv[4 * w4 + 3]);
        }
        if (inb && g2row && !(g.exp_flags & 1)) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float a = gv[2 * j], b = gv[2 * j + 1];
            const __half2 h = __floats2half2_rn(a, b);
            const __half2 l = __floats2half2_rn(a - __low2float(h), b - __high2float(h));
            hi[j] = *reinterpret_cast<const uint32_t*>(&h);
            lo[j] = *reinterpret_cast<const uint32_t*>(&l);
          }
          uint4* dh = reinterpret_cast<uint4*>(g2row + col0);
          uint4* dl = reinterpret_cast<uint4*>(g2row + g.ldg + col0);
          if (staged) {
            dh = reinterpret_cast<uint4*>(stg + lane * NG_STG_ROW + (c & 1) * 32);
            dl = reinterpret_cast<uint4*>(stg + 32 * NG_STG_ROW + lane * NG_STG_ROW + (c & 1) * 32);
          }
          dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]); dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
        if (staged && !(g.exp_flags & 1) && ((c & 1) || c + 1 == nchunk)) flush((c & ~1) * 16, (c & 1) ? 32 : 16);
      };
      EpiChunk<NS> cha, chb;
      if (nchunk > 0) epi_issue<NS>(cha, trow, lab_tile, p.Ppad);
#pragma unroll 1
      for (int c = 0; c < nchunk; c += 2) {
        process(cha, chb, c);
        if (c + 1 < nchunk) process(chb, cha, c + 1);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(l_tempty + 8 * grp);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------- host side
bool nce_tc_supported(int64_t N, int64_t P, int dim, int n_sets) {
  return (dim == 64 || dim == 128 || dim == 256) && N >= 1 && P >= 1 && P < (1ll << 31) - 256 && N < (1ll << 31) &&
         n_sets >= 1 && n_sets <= NT_MAX_SETS;
}

static int64_t nce_ppad(int64_t P) { return (P + N2_BN - 1) / N2_BN * N2_BN; }

size_t nce_tc_workspace_bytes(int64_t N, int64_t P, int dim, int n_sets) {
  Carver c(nullptr);
  c.take<__half>((size_t)N * 2 * dim);
  c.take<__half>((size_t)nce_ppad(P) * 2 * dim);
  c.take<int32_t>(N);
  c.take<int32_t>((size_t)n_sets * N);
  c.take<int32_t>((size_t)n_sets * nce_ppad(P));
  return c.used() + 256;
}

struct NceTcSetup {
  NceTcParams p;
  CUtensorMap ma, mb;
  size_t smem;
  size_t fixed;              // shared memory next to the prototype ring
  const __half* e2;          // [N, 2 D] fp16 (hi | lo) of t * e
  float t;                   // operand scale sqrt(|c| log2 e)
};

// fp16 (hi|lo) copies of both operands, int32 labels, tensor maps: everything the pair kernels read
static int nce_tc_setup(NceTcSetup& u, const int64_t* P_dev, const float* e, const float* prototypes, int64_t N, int64_t P, int dim,
                        const int64_t* inst, const int64_t* sem, const int64_t* psem, int n_sets, const int32_t* plus,
                        float conc, void* workspace, cudaStream_t st) {
  const int64_t Ppad = nce_ppad(P);
  Carver c(workspace);
  __half* ah = c.take<__half>((size_t)N * 2 * dim);
  __half* bh = c.take<__half>((size_t)Ppad * 2 * dim);
  int32_t* inst32 = c.take<int32_t>(N);
  int32_t* sem32 = c.take<int32_t>((size_t)n_sets * N);
  int32_t* psem32 = c.take<int32_t>((size_t)n_sets * Ppad);

  HSG_CUDA(cudaMemsetAsync(bh, 0, sizeof(__half) * Ppad * 2 * dim, st));
  const float t = sqrtf(fabsf(conc) * 1.4426950408889634f);      // both operands scaled by t: accumulator = c log2(e) <e,p>
  nce_split_kernel<<<(unsigned)ceil_div64(N * dim, 2048), 256, 0, st>>>(e, N, dim, t, ah);
  HSG_LAUNCH_CHECK();
  nce_split_kernel<<<(unsigned)ceil_div64(P * dim, 2048), 256, 0, st>>>(prototypes, P, dim, conc < 0.f ? -t : t, bh);
  HSG_LAUNCH_CHECK();
  nce_labels32_kernel<<<(unsigned)ceil_div64(N, 256), 256, 0, st>>>(inst, N, N, 1, inst32, -1);
  HSG_LAUNCH_CHECK();
  nce_labels32_kernel<<<(unsigned)ceil_div64(N * n_sets, 256), 256, 0, st>>>(sem, N, N, n_sets, sem32, 0);
  HSG_LAUNCH_CHECK();
  nce_labels32_kernel<<<(unsigned)ceil_div64(Ppad * n_sets, 256), 256, 0, st>>>(psem, P, Ppad, n_sets, psem32, INT_MIN + 1);
  HSG_LAUNCH_CHECK();

  u.e2 = ah;
  u.t = t;
  NceTcParams& p = u.p;
  p.N = N; p.P = P; p.P_dev = P_dev; p.D = dim; p.n_sets = n_sets; p.inst = inst32; p.sem = sem32; p.psem = psem32;
  p.Ppad = Ppad; p.per_pixel = nullptr; p.stats = nullptr;
  for (int s = 0; s < NT_MAX_SETS; ++s) p.plus[s] = s < n_sets ? plus[s] : 0;
  const int nslab = dim / NT_BK;
  const size_t fixed = (size_t)2 * nslab * NT_SLAB +
                       (size_t)NT_BM * (2 * NT_MAX_SETS + 1) * 4 + 16 + 34 * 8 + 64;
  const size_t budget = 227 * 1024 - 1024 - 1024 - fixed;
  int nstb = (int)(budget / NT_SLAB);
  if (nstb > 8) nstb = 8;
  HSG_REQUIRE(nstb >= 2, HSG_E_UNSUPPORTED, "nce: shared memory budget");
  p.nstb = nstb;
  u.smem = 1024 + fixed + (size_t)nstb * NT_SLAB;
  u.fixed = fixed;
  int rc;
  if ((rc = encode_2d_f16(&u.ma, ah, (uint64_t)N, (uint64_t)2 * dim, NT_BK, NT_BM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = encode_2d_f16(&u.mb, bh, (uint64_t)Ppad, (uint64_t)2 * dim, NT_BK, NT_BN, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  return HSG_OK;
}

int nce_fwd_tc(const float* e, const float* prototypes, int64_t N, int64_t P, int dim, const int64_t* inst,
               const int64_t* sem, const int64_t* psem, int n_sets, const int32_t* plus, float conc,
               float* per_pixel, float* stats, void* workspace, cudaStream_t st, const int64_t* P_dev) {
  NceTcSetup u;
  int rc = nce_tc_setup(u, P_dev, e, prototypes, N, P, dim, inst, sem, psem, n_sets, plus, conc, workspace, st);
  if (rc) return rc;
  NceTcParams& p = u.p;
  p.per_pixel = per_pixel; p.stats = stats;
  const size_t smem = u.smem;
  const CUtensorMap& ma = u.ma;
  const CUtensorMap& mb = u.mb;
  int64_t grid = num_sms() & ~1;                  // CTA pairs
  const int64_t n_pairs = ceil_div64(N, 2 * NT_BM);
  if (grid > 2 * n_pairs) grid = 2 * n_pairs;
#define HSG_NCE_LAUNCH(NS)                                                                                   \
  do {                                                                                                       \
    HSG_CUDA(cudaFuncSetAttribute(nce_fwd_tc2_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    nce_fwd_tc2_kernel<NS><<<(unsigned)grid, NT_THREADS, smem, st>>>(ma, mb, p);                             \
  } while (0)
  switch (n_sets) {
    case 1: HSG_NCE_LAUNCH(1); break;
    case 2: HSG_NCE_LAUNCH(2); break;
    case 3: HSG_NCE_LAUNCH(3); break;
    default: HSG_NCE_LAUNCH(4); break;
  }
#undef HSG_NCE_LAUNCH
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

// backward: one setup per call (state lives in the caller's workspace), then one launch per pixel chunk
static NceTcSetup* setup_slot(void* host_state) { return reinterpret_cast<NceTcSetup*>(host_state); }
size_t nce_grad_tc_host_state_bytes() { return sizeof(NceTcSetup); }

int nce_grad_tc_prepare(void* host_state, const float* e, const float* prototypes, int64_t N, int64_t P, int dim,
                        const int64_t* inst, const int64_t* sem, const int64_t* psem, int n_sets, const int32_t* plus,
                        float conc, void* workspace, cudaStream_t st, const int64_t* P_dev) {
  return nce_tc_setup(*setup_slot(host_state), P_dev, e, prototypes, N, P, dim, inst, sem, psem, n_sets, plus, conc, workspace, st);
}

const __half* nce_grad_tc_e2(void* host_state, float* scale) {
  *scale = setup_slot(host_state)->t;
  return setup_slot(host_state)->e2;
}

int nce_grad_tc(void* host_state, const float* stats, const float* w, float conc, int64_t i_begin, int64_t i_end,
                float* G, int64_t ldg, __half* G2, const float* gscale, cudaStream_t st) {
  NceTcSetup& u = *setup_slot(host_state);
  HSG_REQUIRE(ldg % 64 == 0 && ldg >= u.p.P, HSG_E_INVALID, "nce_grad_tc: ldg=%lld", (long long)ldg);
  HSG_REQUIRE((G != nullptr) != (G2 != nullptr && gscale != nullptr), HSG_E_INVALID, "nce_grad_tc: exactly one of G, G2");
  NceGradTcParams g;
  g.f = u.p; g.i_begin = i_begin; g.i_end = i_end; g.stats = stats; g.w = w; g.conc = conc; g.G = G; g.ldg = ldg;
  g.G2 = G2; g.gscale = gscale;
  static const int exp_env = getenv("HSG_NCE_EXP") ? atoi(getenv("HSG_NCE_EXP")) : 0;
  g.exp_flags = exp_env;
  const int64_t items = ceil_div64(i_end - i_begin, 2 * NT_BM) * (u.p.Ppad / N2_BN);
  int64_t grid = num_sms() & ~1;
  if (grid > 2 * items) grid = 2 * items;
  size_t smem = u.smem;
  g.stg_off = 0;
  if (G2) {
    // staging area of the coalesced stores (8 epilogue warps), paid for with stages of the prototype ring
    const size_t stg = 8 * (size_t)NG_STG_WARP;
    int nstb = (int)((227 * 1024 - 1024 - 1024 - u.fixed - stg) / NT_SLAB);
    if (nstb > 8) nstb = 8;
    if (nstb >= 2) {
      g.f.nstb = nstb;
      g.stg_off = (uint32_t)(u.fixed + (size_t)nstb * NT_SLAB);
      smem = 1024 + u.fixed + (size_t)nstb * NT_SLAB + stg;
    }
  }
#define HSG_NCE_LAUNCH(NS)                                                                                   \
  do {                                                                                                       \
    HSG_CUDA(cudaFuncSetAttribute(nce_grad_tc2_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    nce_grad_tc2_kernel<NS><<<(unsigned)grid, NT_THREADS, smem, st>>>(u.ma, u.mb, g);                        \
  } while (0)
  switch (u.p.n_sets) {
    case 1: HSG_NCE_LAUNCH(1); break;
    case 2: HSG_NCE_LAUNCH(2); break;
    case 3: HSG_NCE_LAUNCH(3); break;
    default: HSG_NCE_LAUNCH(4); break;
  }
#undef HSG_NCE_LAUNCH
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

}  // namespace hsg
