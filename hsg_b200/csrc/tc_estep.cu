// Tensor-core E-step of the spherical k-means (tcgen05 + TMEM + TMA, sm_100a).
//
// Replaces find_nearest_prototypes (hsg/utils/segsort/common.py:44-64: fp32
// cuBLAS mm -> [n,K] matrix in HBM -> argmax) for the shapes that matter at full
// resolution (D in {64,128,256}, K <= 256).  Per 128-pixel tile:
//
//   TMA        : fp16 pixel tile, one [128 x 64] slab (128B-swizzled) per stage
//   tcgen05.mma: D[128 x K] (TMEM, fp32) += A[128 x 64] * C[K x 64]^T, the fp16
//                centroids of the image resident in shared memory (<= 128 KiB)
//                plus one K=16 MMA on a 16-column tail slab that carries the
//                location features as (hi,lo) fp16 pairs (fp32-grade product) and a
//                marker that pushes padding rows to similarity -4
//   epilogue   : 8 warps read the accumulator back (tcgen05.ld) and keep per pixel
//                the top three values with the centroid index packed into the
//                low mantissa bits (one LOP3 + five FMNMX per value, nothing else)
//
// The similarity matrix never leaves the SM.  A pixel whose gap is below twice
// the rigorous error bound of this pass (fp16 rounding of x and c measured at
// conversion time + accumulation + packing) is appended to the re-decision list
// with its two candidates and an upper bound on everything else; the float64
// kernel (kmeans.cu) settles it.  HBM traffic: 2*D bytes/pixel for the fp16
// copy + 4*L for the location features + 4 (error norm) + 4 (label).
#include "kmeans.cuh"
#include "tc_common.cuh"

#include <float.h>
#include <stdlib.h>
#include <cfloat>

namespace hsg {

constexpr int TC_BM = 128;              // pixels per tile (UMMA M)
constexpr int TC_BK = 64;               // fp16 per slab row (128 bytes, one swizzle atom)
constexpr int TC_STAGE_BYTES = TC_BM * TC_BK * 2;
constexpr int TC_THREADS = 640;         // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4-19 epilogue
constexpr int TC_TAIL_BYTES = TC_BM * HSG_XH_TAIL * 2;   // 16-column tail slab of a pixel tile
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_MAXC = 8;                                // candidates a row may list per column half
constexpr int TC_EX_BYTES = 2 * 3 * TC_BM * 4 + 2 * TC_BM * 4 + 2 * TC_BM * 2 * 4 + 64 + 2 * TC_BM * 2 * TC_MAXC;
constexpr float TC_EPS_CONST = 7.1e-5f; // accumulation (3e-5) + index packing (2^-15 * 1.2) + split tail (1e-6)

struct TcParams {
  int d16;
  int kmax, kpad;            // kpad = centroids per pass (MMA N), a multiple of 16
  int kpad_total;            // centroid rows per segment in the fp16 copy (n_pass * kpad)
  int pass, n_pass;          // K > kpad: the centroids go by in n_pass tiles, one launch each
  float* st_val;             // [N,3] running top-3 (index-packed values) carried between passes
  uint8_t* st_tile;          // [N,4] centroid tile of each of the three
  const float* xerr;
  const float* cerr_max;
  Tiles tiles;
  int sub;                   // 128-pixel sub-tiles per tile
  long long items;           // tiles.bound * sub
  int32_t* keys_out;
  FixList fix;
  int nst;                   // pipeline stages
  float* dbg_sims;           // optional [N,kmax] dump of the screening values
  const __half* xh;          // the fp16 side copy itself (L2 prefetch of whole tiles)
  const float* x32;          // fp32 rows [N,dim32] and centroids [S,kmax,dim32] for the in-kernel re-decision
  const float* c32;          //   (dim32 <= 288; NULL = list everything)
  int dim32;
  long long* dbg_clk;        // [grid,4 roles,3] cycle counters of the timing experiment
  int exp_flags;             // timing experiments (HSG_TC_EXP), 0 in production
  int pf_dist;               // single-pass kernel: pixel tiles prefetched into L2 ahead of the ring
};

__device__ __forceinline__ bool item_rows(const TcParams& p, long long item, int count, int& seg,
                                          int64_t& row0, int& np) {
  const int sh = 31 - __clz(p.sub);                   // sub-tiles per tile: a power of two (checked on the host)
  const int ti = (int)(item >> sh);
  if (ti >= count) return false;
  const int64_t b = p.tiles.begin[ti] + (int64_t)((int)item & (p.sub - 1)) * TC_BM;
  const int64_t e = p.tiles.end[ti];
  if (b >= e) return false;
  seg = p.tiles.seg[ti];
  row0 = b;
  np = (int)min((int64_t)TC_BM, e - b);
  return true;
}

// value with its low 8 mantissa bits replaced by `idx` (an IMAD.HI/IMAD version that moves
// this off the ALU pipe measured slower)
__device__ __forceinline__ float pack_idx(uint32_t bits, int idx) {
  return __uint_as_float((bits & 0xFFFFFF00u) | (uint32_t)idx);         // one LOP3
}

// running top-3 (values carry the centroid index in their low mantissa bits)
__device__ __forceinline__ void upd3(float& m, float& s, float& t, float v) {
  const float a = fminf(m, v);
  m = fmaxf(m, v);
  const float b = fminf(s, a);
  s = fmaxf(s, a);
  t = fmaxf(t, b);
}

template <bool DUMP>
__global__ void __launch_bounds__(TC_THREADS, 1)
estep_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_xt,
                const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_ct,
                const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nslab = p.d16 / TC_BK;
  const uint32_t slab_b_bytes = (uint32_t)p.kpad * 128u;
  const uint32_t tail_b_bytes = (uint32_t)p.kpad * 32u;
  const uint32_t sB = base;                                   // centroids, main slabs
  const uint32_t sBT = sB + nslab * slab_b_bytes;             // centroids, tail slab
  const uint32_t sAT = sBT + 256 * 32;                        // pixel tail slab, 2 buffers
  const uint32_t sA = sAT + 2 * TC_TAIL_BYTES;                // pixel main slabs, nst stages
  const uint32_t sMisc = sA + p.nst * TC_STAGE_BYTES;
  uint8_t* misc = smem_raw + (sMisc - smem_u32(smem_raw));
  float* ex = reinterpret_cast<float*>(misc);                 // [2 accumulators][3 values][128 rows]
  float* ex_thr = ex + 2 * 3 * TC_BM;                         // [2][128]
  int* ex_cnt = reinterpret_cast<int*>(ex_thr + 2 * TC_BM);   // [2][128][2 halves]
  int* ex_flag = ex_cnt + 2 * TC_BM * 2;                      // [2 accumulators][2 parities][4 quadrants]
  uint8_t* ex_list = reinterpret_cast<uint8_t*>(ex_flag + 16); // [2][128][2 halves][TC_MAXC]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ex_list + 2 * TC_BM * 2 * TC_MAXC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
  const uint32_t bar_full = smem_u32(bars);                   // [8]
  const uint32_t bar_empty = bar_full + 8 * 8;                // [8]
  const uint32_t bar_tlfull = bar_empty + 8 * 8;              // [2] tail slab landed
  const uint32_t bar_tlempty = bar_tlfull + 16;               // [2]
  const uint32_t bar_bfull = bar_tlempty + 16;                // [1] centroids landed
  const uint32_t bar_tfull = bar_bfull + 8;                   // [2] accumulator ready
  const uint32_t bar_tempty = bar_tfull + 16;                 // [2] accumulator drained

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nst; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tlfull + 8 * i, 1); mbar_init(bar_tlempty + 8 * i, 1);
      mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 8);
    }
    mbar_init(bar_bfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_xt) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_c) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_ct) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int count = *p.tiles.count;
  const long long i_begin = p.items * blockIdx.x / gridDim.x;
  const long long i_end = p.items * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0, cur_seg = -1, tl = 0, last_tl = 0;
      uint32_t phase = 0, tl_phase = 0, last_tl_phase = 0;
      for (long long item = i_begin; item < i_end; ++item) {
        int seg, np; int64_t row0;
        if (!item_rows(p, item, count, seg, row0, np)) continue;
        if (seg != cur_seg) {
          // the tail MMA is the last one of a tile: once it retired, nothing reads the old centroids
          if (cur_seg >= 0) mbar_wait(bar_tlempty + 8 * last_tl, last_tl_phase);
          mbar_expect_tx(bar_bfull, nslab * slab_b_bytes + tail_b_bytes);
          for (int j = 0; j < nslab; ++j)
            tma_load_2d(sB + j * slab_b_bytes, &tmap_c, j * TC_BK, seg * p.kpad_total + p.pass * p.kpad, bar_bfull);
          tma_load_2d(sBT, &tmap_ct, p.d16, seg * p.kpad_total + p.pass * p.kpad, bar_bfull);
          cur_seg = seg;
        }
        for (int j = 0; j < nslab; ++j) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          mbar_expect_tx(bar_full + 8 * stage, TC_STAGE_BYTES);
          tma_load_2d(sA + stage * TC_STAGE_BYTES, &tmap_x, j * TC_BK, (int)row0, bar_full + 8 * stage);
          if (++stage == p.nst) { stage = 0; phase ^= 1; }
        }
        mbar_wait(bar_tlempty + 8 * tl, tl_phase ^ 1);
        mbar_expect_tx(bar_tlfull + 8 * tl, TC_TAIL_BYTES);
        tma_load_2d(sAT + tl * TC_TAIL_BYTES, &tmap_xt, p.d16, (int)row0, bar_tlfull + 8 * tl);
        last_tl = tl; last_tl_phase = tl_phase;
        if (++tl == 2) { tl = 0; tl_phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // (lane-0 issue kept here: the multi-pass shapes are bound by the float64 re-decision and their re-reads of the
    // pixel copy; the convergent elected-lane form of the other kernels measured the same within 2 %)
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=f16, both K-major, N=kpad, M=128
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.kpad >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0, cur_seg = -1, acc = 0, tl = 0;
      uint32_t phase = 0, bcount = 0, tl_phase = 0, acc_phase0 = 0, acc_phase1 = 0;
      for (long long item = i_begin; item < i_end; ++item) {
        int seg, np; int64_t row0;
        if (!item_rows(p, item, count, seg, row0, np)) continue;
        if (seg != cur_seg) {
          mbar_wait(bar_bfull, bcount & 1);
          ++bcount;
          cur_seg = seg;
        }
        mbar_wait(bar_tempty + 8 * acc, (acc ? acc_phase1 : acc_phase0) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int j = 0; j < nslab; ++j) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t ad = umma_desc(sA + stage * TC_STAGE_BYTES, 1024, 2);
          const uint64_t bd = umma_desc(sB + j * slab_b_bytes, 1024, 2);
#pragma unroll
          for (int k4 = 0; k4 < TC_BK / 16; ++k4)                       // +32 bytes per K=16 step
            tc_mma_f16(d_tmem, ad + 2 * k4, bd + 2 * k4, idesc, (j | k4) ? 1u : 0u);
          tc_commit(bar_empty + 8 * stage);
          if (++stage == p.nst) { stage = 0; phase ^= 1; }
        }
        // tail slab: split location features + padding marker (one K=16 MMA)
        mbar_wait(bar_tlfull + 8 * tl, tl_phase);
        tc_fence_after();
        tc_mma_f16(d_tmem, umma_desc(sAT + tl * TC_TAIL_BYTES, 256, 6), umma_desc(sBT, 256, 6), idesc, 1u);
        tc_commit(bar_tlempty + 8 * tl);
        if (++tl == 2) { tl = 0; tl_phase ^= 1; }
        tc_commit(bar_tfull + 8 * acc);
        if (acc) acc_phase1 ^= 1; else acc_phase0 ^= 1;
        acc ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // 16 warps.  Accumulator g (every second tile of this CTA) belongs to warps with e>>3 == g;
    // inside it a pixel row (TMEM lane) is shared by two threads, one per column half h, in the
    // two warps of the same lane quadrant q.  They merge their top-3 through shared memory and
    // synchronise among themselves only (a 64-thread named barrier per warp pair).  Four warps
    // per scheduler keep the ALU pipe busy while the other accumulator is being refilled.
    const int e = warp - 4;
    const int q = warp & 3;                        // TMEM lane quadrant of this warp
    const int h = (e >> 2) & 1;                    // column half
    const int g = e >> 3;                          // accumulator / tile parity
    const int r = 32 * q + lane;                   // accumulator row = pixel within the tile
    const int nchunk = p.kpad >> 4;
    const int c_mid = (nchunk + 1) >> 1;
    const int c_begin = h ? c_mid : 0, c_end = h ? nchunk : c_mid;
    const int bar_id = 1 + q + 4 * g;
    float* exv = ex + g * 3 * TC_BM;               // (m, s, t3) of the upper column half
    float* thr_row = ex_thr + g * TC_BM;           // per row: collect every value >= this (or +inf)
    int* cnt_row = ex_cnt + (g * TC_BM + r) * 2;
    uint8_t* lst_row = ex_list + (g * TC_BM + r) * 2 * TC_MAXC;
    int cur_seg = -1, seq = 0, par = 0;
    uint32_t acc_phase = 0;
    float cerrmax = 0.f;
    for (long long item = i_begin; item < i_end; ++item) {
      int seg, np; int64_t row0;
      if (!item_rows(p, item, count, seg, row0, np)) continue;
      if (((seq++) & 1) != g) continue;
      if (seg != cur_seg) { cerrmax = p.cerr_max[seg]; cur_seg = seg; }
      const int64_t pix = row0 + r;
      const bool inb = r < np;
      const float xe = (h == 0 && inb) ? p.xerr[pix] : 0.f;    // issued before the wait, used after the sweep
      int* many_flag = ex_flag + (g * 2 + par) * 4 + q;          // one flag per 32-row quadrant, double buffered
      if (h == 0 && lane == 0) *many_flag = 0;                   // ordered before this tile's writers by barrier 1

      // earlier centroid tiles of this pixel (K > kpad): the running top-3 and the tile each came from
      float m = -FLT_MAX, s = -FLT_MAX, t3 = -FLT_MAX;
      float om = -FLT_MAX, os = -FLT_MAX, ot = -FLT_MAX;
      uint32_t otile = 0;
      if (p.pass > 0 && h == 0 && inb) {
        om = p.st_val[pix * 3]; os = p.st_val[pix * 3 + 1]; ot = p.st_val[pix * 3 + 2];
        otile = *reinterpret_cast<const uint32_t*>(p.st_tile + pix * 4);
        m = om; s = os; t3 = ot;
      }
      const int kofs = p.pass * p.kpad;             // global index of this tile's first centroid

      mbar_wait(bar_tfull + 8 * g, acc_phase);
      tc_fence_after();
      const uint32_t trow = tmem_base + g * 256 + ((uint32_t)(32 * q) << 16);
      // software pipelined: the TMEM load of chunk c+1 is in flight while chunk c is reduced
      uint32_t va[16], vb[16];
      if (c_begin < c_end) tc_ld16(trow + c_begin * 16, va);
      for (int c = c_begin; c < c_end; c += 2) {
        tc_ld_wait();
        if (c + 1 < c_end) tc_ld16(trow + (c + 1) * 16, vb);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int k = c * 16 + j;
          if (DUMP) { if (inb && kofs + k < p.kmax) p.dbg_sims[pix * p.kmax + kofs + k] = __uint_as_float(va[j]); }
          upd3(m, s, t3, pack_idx(va[j], 255 - k));
        }
        if (c + 1 < c_end) {
          tc_ld_wait();
          if (c + 2 < c_end) tc_ld16(trow + (c + 2) * 16, va);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = (c + 1) * 16 + j;
            if (DUMP) { if (inb && kofs + k < p.kmax) p.dbg_sims[pix * p.kmax + kofs + k] = __uint_as_float(vb[j]); }
            upd3(m, s, t3, pack_idx(vb[j], 255 - k));
          }
        }
      }
      if (h == 1) { exv[r] = m; exv[TC_BM + r] = s; exv[2 * TC_BM + r] = t3; }
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");                       // barrier 1
      bool amb = false, many = false;
      int kb = 0, ks = 0;
      const bool last_pass = p.pass == p.n_pass - 1;
      if (h == 0) {
        upd3(m, s, t3, exv[r]);
        upd3(m, s, t3, exv[TC_BM + r]);
        upd3(m, s, t3, exv[2 * TC_BM + r]);
        // which centroid tile each survivor came from: a value equal to a carried one IS the carried one
        // (bit-equal packed values from two tiles are an exact tie; both stay candidates below)
        const uint32_t cur = (uint32_t)p.pass;
        auto tile_of = [&](float v) -> uint32_t {
          return v == om ? (otile & 0xFFu) : v == os ? ((otile >> 8) & 0xFFu) : v == ot ? ((otile >> 16) & 0xFFu) : cur;
        };
        const uint32_t tm = p.n_pass > 1 ? tile_of(m) : 0u, ts = p.n_pass > 1 ? tile_of(s) : 0u,
                       tt = p.n_pass > 1 ? tile_of(t3) : 0u;
        if (!last_pass) {
          if (inb) {
            p.st_val[pix * 3] = m; p.st_val[pix * 3 + 1] = s; p.st_val[pix * 3 + 2] = t3;
            *reinterpret_cast<uint32_t*>(p.st_tile + pix * 4) = tm | (ts << 8) | (tt << 16);
          }
          thr_row[r] = FLT_MAX;
        } else {
          kb = (int)tm * p.kpad + 255 - (int)(__float_as_uint(m) & 0xFFu);
          ks = (int)ts * p.kpad + 255 - (int)(__float_as_uint(s) & 0xFFu);
          const float thr = 2.f * (xe * 1.001f + cerrmax * 1.001f + TC_EPS_CONST);
          amb = inb && (m - s <= thr);
          many = amb && (m - t3 <= thr);             // three or more inside the bound
          thr_row[r] = (many && p.n_pass == 1) ? m - thr : FLT_MAX;
          if (many && p.n_pass == 1) *many_flag = 1;
          if (inb) p.keys_out[pix] = seg * p.kmax + kb;
        }
      }
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");                       // barrier 2
      const bool tile_many = *many_flag != 0;
      if (tile_many) {
        // second sweep (rare after the first iterations): rows flagged `many` list every candidate
        const float thr_v = thr_row[r];
        uint8_t* lst = lst_row + h * TC_MAXC;
        int cnt = 0;
        for (int c = c_begin; c < c_end; ++c) {
          uint32_t v[16];
          tc_ld16(trow + c * 16, v);
          tc_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = c * 16 + j;
            const float pv = pack_idx(v[j], 255 - k);
            if (pv >= thr_v) { if (cnt < TC_MAXC) lst[cnt] = (uint8_t)k; ++cnt; }
          }
        }
        cnt_row[h] = cnt;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * g);
      acc_phase ^= 1;
      par ^= 1;
      if (tile_many) asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");        // barrier 3
      // one atomic per warp for the list slots (the counter is a single address shared by every SM)
      const unsigned amb_mask = __ballot_sync(FULL, amb);
      int slot_base = 0;
      if (amb_mask) {
        if (lane == 0) slot_base = atomicAdd(p.fix.count, __popc(amb_mask));
        slot_base = __shfl_sync(FULL, slot_base, 0);
      }
      if (amb) {
        const int slot = slot_base + __popc(amb_mask & ((1u << lane) - 1u));
        if (slot < p.fix.capacity) {
          p.fix.pixels[slot] = (int32_t)pix;
            p.fix.segs[slot] = seg;
          uint16_t* cd = p.fix.cand + (int64_t)slot * FIX_MAX_CAND;
          if (!many) {                               // everything but the top two is provably out of reach
            cd[0] = (uint16_t)kb; cd[1] = (uint16_t)ks; cd[2] = 0xFFFF;
          } else if (p.n_pass > 1) {                 // candidates of earlier tiles are gone: scan every cluster
            cd[0] = 0xFFFF;
            atomicAdd(p.fix.count + 1, 1);
          } else {
            const int c0 = cnt_row[0], c1 = cnt_row[1];
            if (c0 > TC_MAXC || c1 > TC_MAXC || c0 + c1 > FIX_MAX_CAND) {
              cd[0] = 0xFFFF;                        // too many to list: scan every cluster
              atomicAdd(p.fix.count + 1, 1);
            } else {
              int w = 0;
              for (int i = 0; i < c0; ++i) cd[w++] = lst_row[i];
              for (int i = 0; i < c1; ++i) cd[w++] = lst_row[TC_MAXC + i];
              if (w < FIX_MAX_CAND) cd[w] = 0xFFFF;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
  }
}

// ---------------------------------------------------------------- single-pass kernel (K <= 256 per image)
// Same TMA / MMA pipeline as estep_tc_kernel, different epilogue.  The sorted top-3 with packed
// indices costs ~6.5 ALU-pipe instructions per similarity and made the ALU pipe the limiter of the
// whole k-means iteration (profiles/r1_estep_tc.txt: ALU 70 %, tensor 41 %).  Here one thread owns one
// pixel row and sweeps its accumulator lane ONCE, 32 columns at a time:
//
//   chunk max        16 FMNMX3                      (ALU pipe, 0.5 per similarity)
//   running max      run = max(run, chunk max)
//   hit bits         v - (run - thr) on the FMA pipe, its sign funnel-shifted into a 32-bit mask
//                    (one SHF per similarity, ALU pipe)
//
// A chunk's bits are taken against the running maximum at that point, which can only be lower than
// the final one, so they flag a superset of the columns within `thr` of the row maximum; after the
// sweep a chunk whose own maximum is below (final max - thr) is dropped whole.  What is left is exact
// for every chunk from the one holding the maximum onwards and a superset before it -- and an earlier
// chunk only survives when it really holds a second candidate, so no row is sent to the float64
// re-decision that the exact rule would not send (lists can only be longer).  A row with exactly one
// surviving bit is decided; the others are listed with every surviving column as candidate.
// 1.5 ALU-pipe instructions per similarity instead of 6.5, no cross-thread merge, no named barriers.
constexpr int TC1_THREADS = 384;        // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4-11 epilogue
constexpr int TC1_PF_DIST = 0;          // pixel tiles prefetched into L2 ahead of the ring (4 x 70 KB x 148 SMs = 41 MB)
constexpr int TC1_INLINE = 2;           // ambiguous rows a warp settles itself per tile (the rest is listed)
constexpr float TC1_EPS_CONST = 3.3e-5f; // accumulation (3e-5) + split tail (1e-6) + slack for the fp32 threshold arithmetic

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// load chunk c (32 columns; the last chunk may be 16 wide: kpad is a multiple of 16)
__device__ __forceinline__ void tc1_load(uint32_t trow, int c, int kpad, uint32_t (&v)[32]) {
  if (c * 32 + 32 <= kpad) {
    tc_ld32(trow + c * 32, v);
  } else {
    uint32_t lo[16];
    tc_ld16(trow + c * 32, lo);
#pragma unroll
    for (int j = 0; j < 16; ++j) { v[j] = lo[j]; v[16 + j] = 0xFF800000u; }    // -inf: never a hit
  }
}

template <bool DUMP>
__global__ void __launch_bounds__(TC1_THREADS, 1)
estep_tc1_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_xt,
                 const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_ct,
                 const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nslab = p.d16 / TC_BK;
  const uint32_t slab_b_bytes = (uint32_t)p.kpad * 128u;
  const uint32_t tail_b_bytes = (uint32_t)p.kpad * 32u;
  const uint32_t sB = base;                                   // centroids, main slabs
  const uint32_t sBT = sB + nslab * slab_b_bytes;             // centroids, tail slab
  const uint32_t sAT = sBT + 256 * 32;                        // pixel tail slab, 2 buffers
  const uint32_t sA = sAT + 2 * TC_TAIL_BYTES;                // pixel main slabs, nst stages
  const uint32_t sMisc = sA + p.nst * TC_STAGE_BYTES;
  uint8_t* misc = smem_raw + (sMisc - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
  const uint32_t bar_full = smem_u32(bars);                   // [12]
  const uint32_t bar_empty = bar_full + 12 * 8;               // [12]
  const uint32_t bar_tlfull = bar_empty + 12 * 8;             // [2] tail slab landed
  const uint32_t bar_tlempty = bar_tlfull + 16;               // [2]
  const uint32_t bar_bfull = bar_tlempty + 16;                // [1] centroids landed
  const uint32_t bar_tfull = bar_bfull + 8;                   // [2] accumulator ready
  const uint32_t bar_tempty = bar_tfull + 16;                 // [2] accumulator drained

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nst; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tlfull + 8 * i, 1); mbar_init(bar_tlempty + 8 * i, 1);
      mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 4);
    }
    mbar_init(bar_bfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_xt) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_c) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_ct) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int count = *p.tiles.count;
  // CTA -> tiles: one contiguous range per CTA (fewest centroid reloads), or, with HSG_TC_EXP & 4, interleaved
  // (CTA b takes tiles b, b + grid, ...; measured 5 % slower at the benchmark shape)
  const bool contiguous = (p.exp_flags & 4) == 0;
  const long long i_begin = contiguous ? p.items * blockIdx.x / gridDim.x : blockIdx.x;
  const long long i_end = contiguous ? p.items * (blockIdx.x + 1) / gridDim.x : p.items;
  const long long i_step = contiguous ? 1 : gridDim.x;
  // timing experiment (HSG_TC_EXP & 2, tools/estep_timeline.py): cycles each role spends waiting
  const bool timing = (p.exp_flags & 2) && p.dbg_clk;
  long long t_wait0 = 0, t_wait1 = 0;
  const long long t_begin = timing ? clock64() : 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0, cur_seg = -1, tl = 0, last_tl = 0;
      uint32_t phase = 0, tl_phase = 0, last_tl_phase = 0;
      // The shared-memory ring (4 stages next to 136 KiB of centroids) holds too few bytes to cover DRAM
      // latency at full bandwidth (r2 ncu: tensor 58 %, DRAM 51 %, the MMA warp waiting on `full`), so the
      // tiles TC1_PF_DIST ahead are pulled into L2 first: the ring then only has to cover L2 latency.
      auto prefetch_item = [&](long long it2) {
        int seg2, np2; int64_t r2;
        if (it2 < i_end && item_rows(p, it2, count, seg2, r2, np2))      // the tile's rows are one contiguous range
          bulk_prefetch_l2(p.xh + r2 * (p.d16 + HSG_XH_TAIL), (uint32_t)np2 * (uint32_t)(p.d16 + HSG_XH_TAIL) * 2u);
      };
      for (int d = 0; d < p.pf_dist; ++d) prefetch_item(i_begin + d * i_step);
      for (long long item = i_begin; item < i_end; item += i_step) {
        if (p.pf_dist > 0) prefetch_item(item + p.pf_dist * i_step);
        int seg, np; int64_t row0;
        if (!item_rows(p, item, count, seg, row0, np)) continue;
        if (seg != cur_seg) {
          // the tail MMA is the last one of a tile: once it retired, nothing reads the old centroids
          if (cur_seg >= 0) mbar_wait(bar_tlempty + 8 * last_tl, last_tl_phase);
          mbar_expect_tx(bar_bfull, nslab * slab_b_bytes + tail_b_bytes);
          for (int j = 0; j < nslab; ++j)
            tma_load_2d(sB + j * slab_b_bytes, &tmap_c, j * TC_BK, seg * p.kpad_total, bar_bfull);
          tma_load_2d(sBT, &tmap_ct, p.d16, seg * p.kpad_total, bar_bfull);
          cur_seg = seg;
        }
        for (int j = 0; j < nslab; ++j) {
          const long long c0 = timing ? clock64() : 0;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (timing) t_wait0 += clock64() - c0;
          mbar_expect_tx(bar_full + 8 * stage, TC_STAGE_BYTES);
          tma_load_2d(sA + stage * TC_STAGE_BYTES, &tmap_x, j * TC_BK, (int)row0, bar_full + 8 * stage);
          if (++stage == p.nst) { stage = 0; phase ^= 1; }
        }
        mbar_wait(bar_tlempty + 8 * tl, tl_phase ^ 1);
        if (p.exp_flags & 1) {                         // timing experiment only: no tail slab
          mbar_arrive(bar_tlfull + 8 * tl);
        } else {
          mbar_expect_tx(bar_tlfull + 8 * tl, TC_TAIL_BYTES);
          tma_load_2d(sAT + tl * TC_TAIL_BYTES, &tmap_xt, p.d16, (int)row0, bar_tlfull + 8 * tl);
        }
        last_tl = tl; last_tl_phase = tl_phase;
        if (++tl == 2) { tl = 0; tl_phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop and an elected lane issues (tc_mma_f16_elect): r2 timeline -- the issuing
    // thread was busy 90 % of a 3 950-cycle tile while the tensor pipe needs 2 176, because every MMA and commit
    // issued from inside a lane-0 branch went through an ELECT + R2UR.BROADCAST retry loop.
    {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.kpad >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0, cur_seg = -1, acc = 0, tl = 0;
      uint32_t phase = 0, bcount = 0, tl_phase = 0, acc_phase0 = 0, acc_phase1 = 0;
      for (long long item = i_begin; item < i_end; item += i_step) {
        int seg = 0, np; int64_t row0;
        const bool have = item_rows(p, item, count, seg, row0, np);
        if (!uniform_i32(have ? 1 : 0)) continue;             // the table entries come from warp-uniform addresses
        seg = uniform_i32(seg);
        if (seg != cur_seg) {
          mbar_wait(bar_bfull, bcount & 1);
          ++bcount;
          cur_seg = seg;
        }
        long long c0 = timing ? clock64() : 0;
        mbar_wait(bar_tempty + 8 * acc, (acc ? acc_phase1 : acc_phase0) ^ 1);
        if (timing) t_wait1 += clock64() - c0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int j = 0; j < nslab; ++j) {
          c0 = timing ? clock64() : 0;
          mbar_wait(bar_full + 8 * stage, phase);
          if (timing) t_wait0 += clock64() - c0;
          tc_fence_after();
          const uint64_t ad = umma_desc(sA + stage * TC_STAGE_BYTES, 1024, 2);
          const uint64_t bd = umma_desc(sB + j * slab_b_bytes, 1024, 2);
          if (!(p.exp_flags & 32)) {                       // timing experiment (32): the data path without the products
#pragma unroll
            for (int k4 = 0; k4 < TC_BK / 16; ++k4)
              tc_mma_f16_elect(d_tmem, ad + 2 * k4, bd + 2 * k4, idesc, (j | k4) ? 1u : 0u);
          }
          tc_commit_elect(bar_empty + 8 * stage);
          if (++stage == p.nst) { stage = 0; phase ^= 1; }
        }
        mbar_wait(bar_tlfull + 8 * tl, tl_phase);
        tc_fence_after();
        if (!(p.exp_flags & (1 | 32)))
          tc_mma_f16_elect(d_tmem, umma_desc(sAT + tl * TC_TAIL_BYTES, 256, 6), umma_desc(sBT, 256, 6), idesc, 1u);
        tc_commit_elect(bar_tlempty + 8 * tl);
        if (++tl == 2) { tl = 0; tl_phase ^= 1; }
        tc_commit_elect(bar_tfull + 8 * acc);
        if (acc) acc_phase1 ^= 1; else acc_phase0 ^= 1;
        acc ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: one thread per pixel row =====================
    const int e = warp - 4;
    const int q = warp & 3;                        // TMEM lane quadrant of this warp
    const int g = e >> 2;                          // accumulator / tile parity
    const int r = 32 * q + lane;                   // accumulator row = pixel within the tile
    const int kpad = p.kpad;
    // The loop is software-pipelined by one tile: the next tile's rows, its error bound (a global load whose
    // latency the sweep used to wait for) and the centroid error are fetched while the list slots of this tile
    // are still on their way back from the atomic.
    int seq = 0;
    uint32_t acc_phase = 0;
    long long item = i_begin;
    auto next_tile = [&](int& seg_, int64_t& row0_, int& np_) -> bool {     // the next tile of this warp group
      while (item < i_end) {
        const long long it = item;
        item += i_step;
        if (!item_rows(p, it, count, seg_, row0_, np_)) continue;
        if (((seq++) & 1) != g) continue;
        return true;
      }
      return false;
    };
    int seg = 0, np = 0; int64_t row0 = 0;
    bool have = next_tile(seg, row0, np);
    float xe = 0.f, cerrmax = 0.f;
    if (have) { xe = r < np ? p.xerr[row0 + r] : 0.f; cerrmax = p.cerr_max[seg]; }
    while (have) {
      const int64_t pix = row0 + r;
      const bool inb = r < np;
      const float thr = 2.f * (xe * 1.001f + cerrmax * 1.001f + TC1_EPS_CONST);

      long long c0 = timing ? clock64() : 0;
      mbar_wait(bar_tfull + 8 * g, acc_phase);
      if (timing) { const long long c1 = clock64(); t_wait0 += c1 - c0; c0 = c1; }
      tc_fence_after();
      const uint32_t trow = tmem_base + g * 256 + ((uint32_t)(32 * q) << 16);
      const int kpad_eff = (p.exp_flags & 8) ? kpad / 2 : kpad;   // timing experiment: half of the TMEM sweep
      float run = -FLT_MAX;
      float cm[8];
      uint32_t mk[8];
      uint32_t va[32], vb[32];
      tc1_load(trow, 0, kpad, va);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        cm[c] = -FLT_MAX; mk[c] = 0xFFFFFFFFu;
        if (c * 32 < kpad_eff) {                                 // warp-uniform
          uint32_t (&v)[32] = (c & 1) ? vb : va;
          uint32_t (&nx)[32] = (c & 1) ? va : vb;
          tc_ld_wait();
          if ((c + 1) * 32 < kpad_eff) tc1_load(trow, c + 1, kpad, nx);   // in flight while chunk c is reduced
          if ((p.exp_flags & 16) && (c & 1)) continue;            // timing experiment: TMEM loads without the arithmetic
          if (DUMP) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int k = c * 32 + j;
              if (inb && k < p.kmax) p.dbg_sims[pix * p.kmax + k] = __uint_as_float(v[j]);
            }
          }
          float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]);
#pragma unroll
          for (int j = 2; j < 32; j += 4) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
            if (j + 3 < 32) m1 = fmaxf(m1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
          }
          const float cmax = fmaxf(m0, m1);
          cm[c] = cmax;
          run = fmaxf(run, cmax);
          const float tp = run - thr;
          uint32_t ma = 0xFFFFFFFFu, mb = 0xFFFFFFFFu;           // two chains of 16
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            ma = __funnelshift_l(__float_as_uint(__uint_as_float(v[j]) - tp), ma, 1);
            mb = __funnelshift_l(__float_as_uint(__uint_as_float(v[16 + j]) - tp), mb, 1);
          }
          mk[c] = (ma << 16) | (mb & 0xFFFFu);                   // bit 31-j clear = column c*32+j is a hit
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * g);
      acc_phase ^= 1;
      if (timing) t_wait1 += clock64() - c0;          // sweep
      const bool post = !(p.exp_flags & 64);          // timing experiment (64): no labels, no list
      unsigned amb_mask = 0;
      int slot_base = 0, cnt = 0;
      bool amb = false;
      if (post) {
        const float t_final = run - thr;
        int first = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t hits = (c * 32 < kpad && cm[c] >= t_final) ? ~mk[c] : 0u;
          mk[c] = hits;
          if (cnt == 0 && hits) first = c * 32 + __clz(hits);
          cnt += __popc(hits);
        }
        amb = inb && cnt > 1;
        if (inb) p.keys_out[pix] = seg * p.kmax + first;
        // Ambiguous rows, optional in-kernel re-decision (HSG_ESTEP_INLINE; off by default: measured slower): up
        // to TC1_INLINE rows per warp and tile are settled right here (same rule as estep_fixup: float64 dot
        // products of the fp32 row with the fp32 centroids of the listed candidates, fixed order, ties to the
        // lowest index); the rest goes to the list.
        if (p.x32) {
          unsigned todo = __ballot_sync(FULL, amb && cnt <= FIX_MAX_CAND);
          for (int n_in = 0; todo && n_in < TC1_INLINE; ++n_in) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t apix = __shfl_sync(FULL, pix, src);
            const float* xrow = p.x32 + apix * p.dim32;
            const float* cbase = p.c32 + (int64_t)seg * p.kmax * p.dim32;
            float xr[9];
#pragma unroll
            for (int m = 0; m < 9; ++m) { const int d = lane + 32 * m; xr[m] = d < p.dim32 ? ld_stream(xrow + d) : 0.f; }
            double bv = -DBL_MAX;
            int bi = 0x7fffffff;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              uint32_t hits = __shfl_sync(FULL, mk[c], src);
              while (hits) {
                const int jb = __clz(hits);
                hits &= ~(0x80000000u >> jb);
                const int k = c * 32 + jb;
                const float* cr = cbase + (int64_t)min(k, p.kmax - 1) * p.dim32;
                double sacc = 0.0;
#pragma unroll
                for (int m = 0; m < 9; ++m) { const int d = lane + 32 * m; if (d < p.dim32) sacc = fma((double)xr[m], (double)cr[d], sacc); }
                sacc = warp_sum(sacc);
                if (k < p.kmax && (sacc > bv || (sacc == bv && k < bi))) { bv = sacc; bi = k; }
              }
            }
            if (lane == src) { p.keys_out[pix] = seg * p.kmax + bi; amb = false; }
          }
        }
        // one atomic per warp for the list slots (the counter is a single address shared by every SM); its
        // result is not needed before the next tile's loads below are on their way
        amb_mask = __ballot_sync(FULL, amb);
        if (amb_mask && lane == 0) slot_base = atomicAdd(p.fix.count, __popc(amb_mask));
      }
      // the next tile of this warp group: rows and error bounds (in flight during the list writes and the wait)
      int nseg = 0, nnp = 0; int64_t nrow0 = 0;
      const bool nhave = next_tile(nseg, nrow0, nnp);
      float nxe = 0.f, ncm = 0.f;
      if (nhave) { nxe = r < nnp ? p.xerr[nrow0 + r] : 0.f; ncm = p.cerr_max[nseg]; }
      if (amb_mask) {
        slot_base = __shfl_sync(FULL, slot_base, 0);
        if (amb) {
          const int slot = slot_base + __popc(amb_mask & ((1u << lane) - 1u));
          if (slot < p.fix.capacity) {
            p.fix.pixels[slot] = (int32_t)pix;
            p.fix.segs[slot] = seg;
            uint16_t* cd = p.fix.cand + (int64_t)slot * FIX_MAX_CAND;
            if (cnt > FIX_MAX_CAND) {
              cd[0] = 0xFFFF;                        // too many to list: scan every cluster
              atomicAdd(p.fix.count + 1, 1);
            } else {
              int w = 0;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                uint32_t hits = mk[c];
                while (hits) {
                  const int j = __clz(hits);
                  hits &= ~(0x80000000u >> j);
                  cd[w++] = (uint16_t)(c * 32 + j);
                }
              }
              if (w < FIX_MAX_CAND) cd[w] = 0xFFFF;
            }
          }
        }
      }
      seg = nseg; row0 = nrow0; np = nnp; xe = nxe; cerrmax = ncm; have = nhave;
    }
  }

  if (timing && lane == 0 && (warp <= 1 || warp == 4 || warp == 8)) {
    const int role = warp <= 1 ? warp : (warp == 4 ? 2 : 3);     // producer, MMA, epilogue acc 0, epilogue acc 1
    long long* o = p.dbg_clk + ((long long)blockIdx.x * 4 + role) * 3;
    o[0] = clock64() - t_begin; o[1] = t_wait0; o[2] = t_wait1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
  }
}

// ---------------------------------------------------------------- CTA-pair kernel (kpad a multiple of 32)
// r2 timeline of the single-CTA kernel above (tools/estep_timeline.py): the MMA issuer waits < 12 % of the
// time, the epilogue can be cut in half without changing the tile time, and yet a tile takes 3 850 cycles
// instead of the 2 176 seventeen 128x256x16 MMAs need: every MMA streams 12 KB of operands out of shared
// memory (4 KB of pixels + 8 KB of centroids) and that read rate, not the tensor pipe, paces it.  A
// cta_group::2 tile of 256 pixels x 256 centroids reads 8 KB per CTA instead: each CTA keeps its own 128
// pixel rows and HALF of the centroids, the tensor cores fetch the other half from the peer's shared
// memory.  It also frees 68 KB for the pixel ring (9 stages instead of 4).  Plumbing as in the NCE forward
// (nce_tc.cu): the leader (cluster rank 0) issues every MMA, both CTAs run a TMA producer crediting the
// leader's "full" barriers, commits are multicast, each CTA's epilogue drains its own 128 TMEM lanes.
__device__ __forceinline__ bool pair_item_rows(const TcParams& p, long long item, int count, uint32_t rank,
                                               int& seg, int64_t& row0, int& np) {
  const int sub2 = p.sub >> 1;                       // 256-pixel items per tile
  const int ti = (int)(item / sub2);
  if (ti >= count) return false;
  const int64_t b = p.tiles.begin[ti] + (int64_t)(item % sub2) * (2 * TC_BM);
  const int64_t e = p.tiles.end[ti];
  if (b >= e) return false;                          // decided on the pair's first row: both CTAs agree
  seg = p.tiles.seg[ti];
  row0 = b + (int64_t)rank * TC_BM;
  const int64_t left = e - row0;
  np = left <= 0 ? 0 : (int)min((int64_t)TC_BM, left);
  return true;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC1_THREADS, 1)
estep_tc2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_xt,
                 const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_ct,
                 const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nslab = p.d16 / TC_BK;
  const int khalf = p.kpad >> 1;                              // centroid rows held by this CTA
  const uint32_t slab_b_bytes = (uint32_t)khalf * 128u;
  const uint32_t tail_b_bytes = (uint32_t)khalf * 32u;
  const uint32_t sB = base;                                   // centroids (own half), main slabs
  const uint32_t sBT = sB + nslab * slab_b_bytes;             // centroids (own half), tail slab
  const uint32_t sAT = sBT + 128 * 32;                        // pixel tail slab, 2 buffers
  const uint32_t sA = sAT + 2 * TC_TAIL_BYTES;                // pixel main slabs, nst stages
  const uint32_t sMisc = sA + p.nst * TC_STAGE_BYTES;
  uint8_t* misc = smem_raw + (sMisc - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
  const uint32_t bar_full = smem_u32(bars);                   // [12] (waited on in the leader only)
  const uint32_t bar_empty = bar_full + 12 * 8;               // [12]
  const uint32_t bar_tlfull = bar_empty + 12 * 8;             // [2] tail slab landed (leader)
  const uint32_t bar_tlempty = bar_tlfull + 16;               // [2]
  const uint32_t bar_bfull = bar_tlempty + 16;                // [1] centroids landed (leader)
  const uint32_t bar_tfull = bar_bfull + 8;                   // [2] accumulator ready
  const uint32_t bar_tempty = bar_tfull + 16;                 // [2] accumulator drained (leader; 8 arrivals)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nst; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tlfull + 8 * i, 1); mbar_init(bar_tlempty + 8 * i, 1);
      mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 8);
    }
    mbar_init(bar_bfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_xt) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_c) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_ct) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                         // the peer's barriers exist before anything remote touches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int count = *p.tiles.count;
  const long long pitems_all = p.items >> 1;                  // 256-pixel items
  // pair -> items: one contiguous range per pair, as in the single-CTA kernel.  Interleaved (pair b takes items b,
  // b + pairs, ...: HSG_TC_EXP & 4) every pair walks through EVERY image, and each image boundary drains the MMA
  // pipeline and reloads 68 KB of centroids per CTA -- 48 times per pair at the benchmark shape.
  const bool contiguous = (p.exp_flags & 4) == 0;
  const long long n_pairs = gridDim.x >> 1, pair_id = blockIdx.x >> 1;
  const long long i_begin = contiguous ? pitems_all * pair_id / n_pairs : pair_id;
  const long long pitems = contiguous ? pitems_all * (pair_id + 1) / n_pairs : pitems_all;
  const long long i_step = contiguous ? 1 : n_pairs;
  const bool timing = (p.exp_flags & 2) && p.dbg_clk;
  long long t_wait0 = 0, t_wait1 = 0;
  const long long t_begin = timing ? clock64() : 0;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      const uint32_t l_full = mapa_cluster(bar_full, 0);
      const uint32_t l_tlfull = mapa_cluster(bar_tlfull, 0);
      const uint32_t l_bfull = mapa_cluster(bar_bfull, 0);
      int stage = 0, cur_seg = -1, tl = 0, last_tl = 0;
      uint32_t phase = 0, tl_phase = 0, last_tl_phase = 0;
      for (long long item = i_begin; item < pitems; item += i_step) {
        int seg, np; int64_t row0;
        if (!pair_item_rows(p, item, count, rank, seg, row0, np)) continue;
        if (seg != cur_seg) {
          // the tail MMA is the last one of a tile: once it retired, nothing reads the old centroids
          if (cur_seg >= 0) mbar_wait(bar_tlempty + 8 * last_tl, last_tl_phase);
          if (leader) mbar_expect_tx(bar_bfull, 2 * (nslab * slab_b_bytes + tail_b_bytes));
          const int crow = seg * p.kpad_total + (int)rank * khalf;
          for (int j = 0; j < nslab; ++j) tma_load_2d_pair(sB + j * slab_b_bytes, &tmap_c, j * TC_BK, crow, l_bfull);
          tma_load_2d_pair(sBT, &tmap_ct, p.d16, crow, l_bfull);
          cur_seg = seg;
        }
        for (int j = 0; j < nslab; ++j) {
          const long long c0 = timing ? clock64() : 0;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (timing) t_wait0 += clock64() - c0;
          if (leader) mbar_expect_tx(bar_full + 8 * stage, 2 * TC_STAGE_BYTES);
          tma_load_2d_pair(sA + stage * TC_STAGE_BYTES, &tmap_x, j * TC_BK, (int)row0, l_full + 8 * stage);
          if (++stage == p.nst) { stage = 0; phase ^= 1; }
        }
        mbar_wait(bar_tlempty + 8 * tl, tl_phase ^ 1);
        if (leader) mbar_expect_tx(bar_tlfull + 8 * tl, 2 * TC_TAIL_BYTES);
        tma_load_2d_pair(sAT + tl * TC_TAIL_BYTES, &tmap_xt, p.d16, (int)row0, l_tlfull + 8 * tl);
        last_tl = tl; last_tl_phase = tl_phase;
        if (++tl == 2) { tl = 0; tl_phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    // the whole warp walks the loop, an elected lane issues (tc_common.cuh: tc_mma_f16_elect)
    if (uniform_i32(leader ? 1 : 0)) {
      // instruction descriptor: D=f32, A=B=f16, both K-major, N=kpad, M=256 (two CTAs x 128)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.kpad >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
      int stage = 0, cur_seg = -1, acc = 0, tl = 0;
      uint32_t phase = 0, bcount = 0, tl_phase = 0, acc_phase0 = 0, acc_phase1 = 0;
      for (long long item = i_begin; item < pitems; item += i_step) {
        int seg = 0, np; int64_t row0;
        const bool have = pair_item_rows(p, item, count, rank, seg, row0, np);
        if (!uniform_i32(have ? 1 : 0)) continue;
        seg = uniform_i32(seg);
        if (seg != cur_seg) {
          mbar_wait(bar_bfull, bcount & 1);
          ++bcount;
          cur_seg = seg;
        }
        long long c0 = timing ? clock64() : 0;
        mbar_wait(bar_tempty + 8 * acc, (acc ? acc_phase1 : acc_phase0) ^ 1);
        if (timing) t_wait1 += clock64() - c0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int j = 0; j < nslab; ++j) {
          c0 = timing ? clock64() : 0;
          mbar_wait(bar_full + 8 * stage, phase);
          if (timing) t_wait0 += clock64() - c0;
          tc_fence_after();
          const uint64_t ad = umma_desc(sA + stage * TC_STAGE_BYTES, 1024, 2);
          const uint64_t bd = umma_desc(sB + j * slab_b_bytes, 1024, 2);
#pragma unroll
          for (int k4 = 0; k4 < TC_BK / 16; ++k4)
            tc_mma_f16_pair_elect(d_tmem, ad + 2 * k4, bd + 2 * k4, idesc, (j | k4) ? 1u : 0u);
          tc_commit_pair_elect(bar_empty + 8 * stage);
          if (++stage == p.nst) { stage = 0; phase ^= 1; }
        }
        mbar_wait(bar_tlfull + 8 * tl, tl_phase);
        tc_fence_after();
        tc_mma_f16_pair_elect(d_tmem, umma_desc(sAT + tl * TC_TAIL_BYTES, 256, 6), umma_desc(sBT, 256, 6), idesc, 1u);
        tc_commit_pair_elect(bar_tlempty + 8 * tl);
        if (++tl == 2) { tl = 0; tl_phase ^= 1; }
        tc_commit_pair_elect(bar_tfull + 8 * acc);
        if (acc) acc_phase1 ^= 1; else acc_phase0 ^= 1;
        acc ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 TMEM lanes): one thread per pixel row =====================
    const int e = warp - 4;
    const int q = warp & 3;
    const int g = e >> 2;
    const int r = 32 * q + lane;
    const int kpad = p.kpad;
    const uint32_t l_tempty = mapa_cluster(bar_tempty, 0);
    int cur_seg = -1, seq = 0;
    uint32_t acc_phase = 0;
    float cerrmax = 0.f;
    for (long long item = i_begin; item < pitems; item += i_step) {
      int seg, np; int64_t row0;
      if (!pair_item_rows(p, item, count, rank, seg, row0, np)) continue;
      if (((seq++) & 1) != g) continue;
      if (seg != cur_seg) { cerrmax = p.cerr_max[seg]; cur_seg = seg; }
      const int64_t pix = row0 + r;
      const bool inb = r < np;
      const float xe = inb ? p.xerr[pix] : 0.f;
      const float thr = 2.f * (xe * 1.001f + cerrmax * 1.001f + TC1_EPS_CONST);

      long long c0 = timing ? clock64() : 0;
      mbar_wait(bar_tfull + 8 * g, acc_phase);
      if (timing) { const long long c1 = clock64(); t_wait0 += c1 - c0; c0 = c1; }
      tc_fence_after();
      const uint32_t trow = tmem_base + g * 256 + ((uint32_t)(32 * q) << 16);
      float run = -FLT_MAX;
      float cm[8];
      uint32_t mk[8];
      uint32_t va[32], vb[32];
      tc_ld32(trow, va);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        cm[c] = -FLT_MAX; mk[c] = 0xFFFFFFFFu;
        if (c * 32 < kpad) {                                     // warp-uniform; kpad is a multiple of 32 here
          uint32_t (&v)[32] = (c & 1) ? vb : va;
          uint32_t (&nx)[32] = (c & 1) ? va : vb;
          tc_ld_wait();
          if ((c + 1) * 32 < kpad) tc_ld32(trow + (c + 1) * 32, nx);
          float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]);
#pragma unroll
          for (int j = 2; j < 32; j += 4) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
            if (j + 3 < 32) m1 = fmaxf(m1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
          }
          const float cmax = fmaxf(m0, m1);
          cm[c] = cmax;
          run = fmaxf(run, cmax);
          const float tp = run - thr;
          uint32_t ma = 0xFFFFFFFFu, mb = 0xFFFFFFFFu;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            ma = __funnelshift_l(__float_as_uint(__uint_as_float(v[j]) - tp), ma, 1);
            mb = __funnelshift_l(__float_as_uint(__uint_as_float(v[16 + j]) - tp), mb, 1);
          }
          mk[c] = (ma << 16) | (mb & 0xFFFFu);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(l_tempty + 8 * g);
      acc_phase ^= 1;
      if (timing) t_wait1 += clock64() - c0;

      const float t_final = run - thr;
      int cnt = 0, first = 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t hits = (c * 32 < kpad && cm[c] >= t_final) ? ~mk[c] : 0u;
        mk[c] = hits;
        if (cnt == 0 && hits) first = c * 32 + __clz(hits);
        cnt += __popc(hits);
      }
      const bool amb = inb && cnt > 1;
      if (inb) p.keys_out[pix] = seg * p.kmax + first;
      const unsigned amb_mask = __ballot_sync(FULL, amb);
      if (amb_mask) {
        int slot_base = 0;
        if (lane == 0) slot_base = atomicAdd(p.fix.count, __popc(amb_mask));
        slot_base = __shfl_sync(FULL, slot_base, 0);
        if (amb) {
          const int slot = slot_base + __popc(amb_mask & ((1u << lane) - 1u));
          if (slot < p.fix.capacity) {
            p.fix.pixels[slot] = (int32_t)pix;
            p.fix.segs[slot] = seg;
            uint16_t* cd = p.fix.cand + (int64_t)slot * FIX_MAX_CAND;
            if (cnt > FIX_MAX_CAND) {
              cd[0] = 0xFFFF;
              atomicAdd(p.fix.count + 1, 1);
            } else {
              int w = 0;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                uint32_t hits = mk[c];
                while (hits) {
                  const int j = __clz(hits);
                  hits &= ~(0x80000000u >> j);
                  cd[w++] = (uint16_t)(c * 32 + j);
                }
              }
              if (w < FIX_MAX_CAND) cd[w] = 0xFFFF;
            }
          }
        }
      }
    }
  }

  if (timing && lane == 0 && (warp <= 1 || warp == 4 || warp == 8)) {
    const int role = warp <= 1 ? warp : (warp == 4 ? 2 : 3);
    long long* o = p.dbg_clk + ((long long)blockIdx.x * 4 + role) * 3;
    o[0] = clock64() - t_begin; o[1] = t_wait0; o[2] = t_wait1;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // neither CTA leaves (or frees TMEM) while the pair still works
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
  }
}

// ---------------------------------------------------------------- centroid conversion
__global__ void tc_convert_kernel(const float* __restrict__ cent, const int32_t* __restrict__ seg_k, int S,
                                  int kmax, int kpad, int dim, int d16, __half* __restrict__ ch,
                                  float* __restrict__ cerr, float* __restrict__ cerr_max) {
  // row layout mirrors the pixel side (prep.cu): [0,d16) fp16(c_d); d16+3l+{0,1,2} = (hi, lo, hi) of the
  // l-th trailing feature; d16+15 = 0 for a real centroid, -4 for a padding row (similarity -4: never wins)
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (int64_t)S * kpad) return;
  const int s = (int)(row / kpad), k = (int)(row % kpad);
  const int K = seg_k ? seg_k[s] : kmax;
  const int wx = d16 + HSG_XH_TAIL;
  __half* dst = ch + row * wx;
  const bool real = k < K && k < kmax;
  const float* src = cent + ((int64_t)s * kmax + (real ? k : 0)) * dim;
  float e2 = 0.f;
  for (int d = lane; d < d16; d += 32) {
    const float v = real ? src[d] : 0.f;
    const __half hv = __float2half_rn(v);
    const float rr = v - __half2float(hv);
    e2 = fmaf(rr, rr, e2);
    dst[d] = hv;
  }
  if (lane < HSG_XH_TAIL) {
    const int L = dim - d16;
    const int l = lane / 3, part = lane % 3;
    float o = 0.f;
    if (lane == HSG_XH_TAIL - 1) {
      o = real ? 0.f : -4.f;
    } else if (real && l < L) {
      const float v = src[d16 + l];
      const float hi = __half2float(__float2half_rn(v));
      o = part == 1 ? v - hi : hi;
    }
    dst[d16 + lane] = __float2half_rn(o);
  }
  e2 = warp_sum(e2);
  if (lane == 0 && k < kmax) {
    const float e = real ? sqrtf(e2) * 1.0001f + 1e-30f : 0.f;
    cerr[(int64_t)s * kmax + k] = e;
    atomicMax(reinterpret_cast<int*>(cerr_max + s), __float_as_int(e));   // non-negative floats order like ints
  }
}

// ---------------------------------------------------------------- host side
// centroids per pass: the fp16 tile [ktile x (d16+16)] has to fit next to the pixel ring (<= 136 KiB)
// D = 512: a whole 256-centroid tile (270 KB of fp16) does not fit one SM; the CTA-pair kernel holds half of it
// per CTA, so K <= 256 stays a single pass there too (it was two passes with the running top-3 carried through
// HBM: 16.7 ms per iteration at N = 1e7, profiles/r1_kmeans_flat_sweep.txt)
static bool tc_pair_only(int d16, int kpad) { return d16 == 512 && kpad > 128 && kpad <= 256 && kpad % 32 == 0; }
static int tc_ktile(int d16, int kpad) { return (d16 <= 256 || tc_pair_only(d16, kpad)) ? 256 : 128; }

bool tc_shape_supported(int dim, int d16, int kmax) {
  if (!(d16 == 64 || d16 == 128 || d16 == 256 || d16 == 512)) return false;
  if (dim < d16 || dim - d16 > HSG_XH_MAX_TRAILING) return false;
  if (kmax < 1) return false;
  const int kpad = (kmax + 15) / 16 * 16;
  return kpad <= 255 * tc_ktile(d16, kpad);        // tile ids are carried as bytes
}

void tc_carve(Carver& c, TcState& t, int S, int kmax, int d16, int64_t N) {
  t.enabled = false;
  t.d16 = d16;
  const int kpad = (kmax + 15) / 16 * 16, ktile = tc_ktile(d16, kpad);
  t.n_pass = (kpad + ktile - 1) / ktile;
  t.kpad = t.n_pass == 1 ? kpad : ktile;
  t.kpad_total = t.n_pass * t.kpad;
  t.ch = c.take<__half>((size_t)S * t.kpad_total * (d16 + HSG_XH_TAIL));
  t.cerr = c.take<float>((size_t)S * kmax);
  t.cerr_max = c.take<float>(S);
  t.st_val = t.n_pass > 1 ? c.take<float>((size_t)N * 3) : nullptr;
  t.st_tile = t.n_pass > 1 ? c.take<uint8_t>((size_t)N * 4) : nullptr;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp16 tensor map over a row-major [rows, row_elems] array; box = box_cols x box_rows
int encode_2d_f16(void* out, const void* ptr, uint64_t rows, uint64_t row_elems, uint32_t box_cols,
                  uint32_t box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  HSG_REQUIRE(fn, HSG_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {row_elems, rows};
  cuuint64_t strides[1] = {row_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HSG_REQUIRE(r == CUDA_SUCCESS, HSG_E_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box=%ux%u",
              (int)r, (unsigned long long)rows, (unsigned long long)row_elems, box_cols, box_rows);
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  memcpy(out, &m, sizeof(m));
  return HSG_OK;
}

int tc_prepare(TcState& t, int64_t N, int S) {
  HSG_REQUIRE(((uintptr_t)t.xh & 15) == 0, HSG_E_INVALID, "tensor-core E-step: fp16 copy must be 16-byte aligned");
  const uint64_t wx = (uint64_t)t.d16 + HSG_XH_TAIL;
  int rc;
  if ((rc = encode_2d_f16(t.tmap_x, t.xh, (uint64_t)N, wx, TC_BK, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = encode_2d_f16(t.tmap_xt, t.xh, (uint64_t)N, wx, HSG_XH_TAIL, TC_BM, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  if ((rc = encode_2d_f16(t.tmap_c, t.ch, (uint64_t)S * t.kpad_total, wx, TC_BK, (uint32_t)t.kpad, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = encode_2d_f16(t.tmap_ct, t.ch, (uint64_t)S * t.kpad_total, wx, HSG_XH_TAIL, (uint32_t)t.kpad, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  if (t.n_pass == 1 && t.kpad % 32 == 0) {         // CTA-pair kernel: each CTA loads half of the centroid rows
    if ((rc = encode_2d_f16(t.tmap_c2, t.ch, (uint64_t)S * t.kpad_total, wx, TC_BK, (uint32_t)t.kpad / 2, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = encode_2d_f16(t.tmap_ct2, t.ch, (uint64_t)S * t.kpad_total, wx, HSG_XH_TAIL, (uint32_t)t.kpad / 2, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  }
  t.enabled = true;
  return HSG_OK;
}

int tc_convert_centroids(const EStepArgs& a, const TcState& t, cudaStream_t st) {
  HSG_CUDA(cudaMemsetAsync(t.cerr_max, 0, sizeof(float) * a.S, st));
  const int64_t rows = (int64_t)a.S * t.kpad_total;
  tc_convert_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, st>>>(a.centroids, a.seg_k, a.S, a.kmax, t.kpad_total, a.dim,
                                                                  t.d16, t.ch, t.cerr, t.cerr_max);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

float* g_tc_debug_sims = nullptr;
long long* g_tc_debug_clk = nullptr;   // set by hsg_debug_set_tc_clock (tools only)   // set by hsg_debug_set_tc_dump (tests only)

int estep_tc(const EStepArgs& a, const TcState& t, cudaStream_t st) {
  HSG_REQUIRE(t.enabled, HSG_E_INVALID, "tensor-core E-step used before tc_prepare");
  TcParams p;
  p.d16 = t.d16; p.kmax = a.kmax; p.kpad = t.kpad; p.kpad_total = t.kpad_total; p.n_pass = t.n_pass; p.pass = 0;
  p.st_val = t.st_val; p.st_tile = t.st_tile; p.xerr = t.xerr; p.cerr_max = t.cerr_max;
  p.tiles = a.tiles; p.sub = (int)(a.tiles.tile / TC_BM); p.items = (long long)a.tiles.bound * p.sub;
  HSG_REQUIRE(p.sub > 0 && (p.sub & (p.sub - 1)) == 0, HSG_E_INVALID, "tensor-core E-step: tile of %lld rows", (long long)a.tiles.tile);
  p.keys_out = a.keys_out; p.fix = a.fix; p.dbg_sims = g_tc_debug_sims; p.xh = t.xh; p.pf_dist = 0; p.exp_flags = 0; p.dbg_clk = g_tc_debug_clk;
  static const bool no_inline = getenv("HSG_ESTEP_INLINE") == nullptr;      // opt-in: measured 2x SLOWER
  p.x32 = (a.dim <= 288 && !no_inline) ? a.x : nullptr; p.c32 = a.centroids; p.dim32 = a.dim;
  const int nslab = t.d16 / TC_BK;
  const size_t fixed = (size_t)nslab * t.kpad * 128 + 256 * 32 + 2 * TC_TAIL_BYTES   // centroids + tails
                       + TC_EX_BYTES + 32 * 8 + 64;                                  // exchange, barriers, tmem slot
  const size_t budget = 227 * 1024 - 1024 /*static*/ - 1024 /*alignment slack*/ - fixed;
  int nst = (int)(budget / TC_STAGE_BYTES);
  if (nst > 8) nst = 8;
  p.nst = nst;                                   // checked where the multi-pass kernel is launched
  const size_t smem = 1024 + fixed + (size_t)nst * TC_STAGE_BYTES;
  CUtensorMap mx, mxt, mc, mct;
  memcpy(&mx, t.tmap_x, sizeof(mx));
  memcpy(&mxt, t.tmap_xt, sizeof(mxt));
  memcpy(&mc, t.tmap_c, sizeof(mc));
  memcpy(&mct, t.tmap_ct, sizeof(mct));
  long long grid = num_sms();
  if (grid > p.items) grid = p.items;
  if (grid < 1) grid = 1;
  static const bool legacy = getenv("HSG_ESTEP_LEGACY") != nullptr;     // A/B switch for profiling only
  static const bool no_pair = getenv("HSG_ESTEP_PAIR") == nullptr;         // the pair kernel is opt-in (profiling): see DESIGN.md
  static const int exp_all = getenv("HSG_TC_EXP") ? atoi(getenv("HSG_TC_EXP")) : 0;
  const bool need_pair = tc_pair_only(t.d16, t.kpad);
  HSG_REQUIRE(!need_pair || (!p.dbg_sims && p.sub % 2 == 0 && num_sms() >= 2), HSG_E_UNSUPPORTED,
              "tensor-core E-step: D=512 with more than 128 centroids runs on CTA pairs only (no similarity dump)");
  if (p.n_pass == 1 && !legacy && (!no_pair || need_pair) && !p.dbg_sims && t.kpad >= 64 && t.kpad % 32 == 0 &&
      p.sub % 2 == 0 && num_sms() >= 2) {
    // CTA pairs: half of the centroids per CTA (estep_tc2_kernel)
    const int khalf = t.kpad / 2;
    const size_t fixed2 = (size_t)nslab * khalf * 128 + 128 * 32 + 2 * TC_TAIL_BYTES + 40 * 8 + 64;
    int nst2 = (int)((227 * 1024 - 1024 - 1024 - fixed2) / TC_STAGE_BYTES);
    if (nst2 > 12) nst2 = 12;
    HSG_REQUIRE(nst2 >= 2, HSG_E_UNSUPPORTED, "tensor-core E-step: shared memory budget (kpad=%d d16=%d)", t.kpad, t.d16);
    p.nst = nst2;
    p.exp_flags = exp_all;
    const size_t smem2 = 1024 + fixed2 + (size_t)nst2 * TC_STAGE_BYTES;
    CUtensorMap mc2, mct2;
    memcpy(&mc2, t.tmap_c2, sizeof(mc2));
    memcpy(&mct2, t.tmap_ct2, sizeof(mct2));
    long long grid2 = num_sms() & ~1;
    const long long pitems = p.items / 2;
    if (grid2 > 2 * pitems) grid2 = 2 * pitems;
    if (grid2 < 2) grid2 = 2;
    HSG_CUDA(cudaFuncSetAttribute(estep_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    estep_tc2_kernel<<<(unsigned)grid2, TC1_THREADS, smem2, st>>>(mx, mxt, mc2, mct2, p);
    HSG_LAUNCH_CHECK();
    return HSG_OK;
  }
  if (p.n_pass == 1 && !legacy) {
    // single-pass shapes (K <= 256 per image): one thread per pixel row, single sweep (estep_tc1_kernel)
    const size_t fixed1 = (size_t)nslab * t.kpad * 128 + 256 * 32 + 2 * TC_TAIL_BYTES + 40 * 8 + 64;
    int nst1 = (int)((227 * 1024 - 1024 - 1024 - fixed1) / TC_STAGE_BYTES);
    if (nst1 > 12) nst1 = 12;
    HSG_REQUIRE(nst1 >= 2, HSG_E_UNSUPPORTED, "tensor-core E-step: shared memory budget (kpad=%d d16=%d)", t.kpad, t.d16);
    p.nst = nst1;
    static const int pf_env = getenv("HSG_TC_PF") ? atoi(getenv("HSG_TC_PF")) : TC1_PF_DIST;    // tuning knob
    p.pf_dist = pf_env < 0 ? 0 : (pf_env > 16 ? 16 : pf_env);
    static const int exp_env = getenv("HSG_TC_EXP") ? atoi(getenv("HSG_TC_EXP")) : 0;
    p.exp_flags = exp_env;
    const size_t smem1 = 1024 + fixed1 + (size_t)nst1 * TC_STAGE_BYTES;
    if (p.dbg_sims) {
      HSG_CUDA(cudaFuncSetAttribute(estep_tc1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
      estep_tc1_kernel<true><<<(unsigned)grid, TC1_THREADS, smem1, st>>>(mx, mxt, mc, mct, p);
    } else {
      HSG_CUDA(cudaFuncSetAttribute(estep_tc1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
      estep_tc1_kernel<false><<<(unsigned)grid, TC1_THREADS, smem1, st>>>(mx, mxt, mc, mct, p);
    }
    HSG_LAUNCH_CHECK();
    return HSG_OK;
  }
  HSG_REQUIRE(nst >= 2 && fixed < 225 * 1024, HSG_E_UNSUPPORTED, "tensor-core E-step: shared memory budget (kpad=%d d16=%d)", t.kpad, t.d16);
  for (p.pass = 0; p.pass < p.n_pass; ++p.pass) {
    if (p.dbg_sims) {
      HSG_CUDA(cudaFuncSetAttribute(estep_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      estep_tc_kernel<true><<<(unsigned)grid, TC_THREADS, smem, st>>>(mx, mxt, mc, mct, p);
    } else {
      HSG_CUDA(cudaFuncSetAttribute(estep_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      estep_tc_kernel<false><<<(unsigned)grid, TC_THREADS, smem, st>>>(mx, mxt, mc, mct, p);
    }
    HSG_LAUNCH_CHECK();
  }
  return HSG_OK;
}

}  // namespace hsg

// test hook: dump the screening similarities of the next tensor-core E-steps into
// a caller buffer [N,kmax] (NULL switches it off)
extern "C" int hsg_debug_set_tc_clock(long long* clk) {
  hsg::g_tc_debug_clk = clk;
  return HSG_OK;
}

extern "C" int hsg_debug_set_tc_dump(float* sims) {
  hsg::g_tc_debug_sims = sims;
  return HSG_OK;
}
