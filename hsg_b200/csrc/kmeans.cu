// K1: spherical k-means (kmeans_with_initial_labels, hsg/utils/segsort/common.py:67-97)
// batched over independent segments (the per-image loop of segment_by_kmeans,
// :337-372), plus the single-step and segmented-reduction entry points.
//
// M-step: deterministic segmented sum + normalise (segreduce.cu).
// E-step: a cheap pass (fp32 CUDA cores here, fp16 tcgen05 in tc_estep.cu)
//         finds the best centroid and the gap to the runner-up; pixels whose gap
//         is inside that pass's rigorous error bound go to a float64 re-decision
//         (estep_fixup), so the label is always the arg-max of the float64 dot
//         products of the fp32 inputs, ties to the lowest index -- whichever
//         pass produced it.
#include "kmeans.cuh"

#include <float.h>
#include <stdlib.h>

namespace hsg {

// ---------------------------------------------------------------- SIMT E-step
constexpr int ES_TP = 64;     // pixels per CTA
constexpr int ES_TK = 64;     // centroids per k-block
constexpr int ES_DC = 32;     // feature chunk
constexpr int ES_THREADS = 256;

__device__ __forceinline__ void merge_best(float& bv, int& bi, float& sv, float ov, int oi, float osv) {
  if (ov > bv || (ov == bv && oi < bi)) {
    sv = fmaxf(bv, fmaxf(sv, osv));
    bv = ov;
    bi = oi;
  } else {
    sv = fmaxf(ov, fmaxf(sv, osv));
  }
}

__global__ void __launch_bounds__(ES_THREADS) estep_simt_kernel(const EStepArgs a, const float thr) {
  __shared__ float Xs[ES_TP][ES_DC + 1];
  __shared__ float Cs[ES_TK][ES_DC + 1];
  __shared__ float best_v[ES_TP], second_v[ES_TP];
  __shared__ int best_i[ES_TP];

  const int ti = blockIdx.x;
  if (ti >= *a.tiles.count) return;
  const int64_t p0 = a.tiles.begin[ti] + (int64_t)blockIdx.y * ES_TP;
  const int64_t pe = a.tiles.end[ti];
  if (p0 >= pe) return;
  const int np = (int)min((int64_t)ES_TP, pe - p0);
  const int seg = a.tiles.seg[ti];
  const int K = a.seg_k ? a.seg_k[seg] : a.kmax;
  const float* cbase = a.centroids + (int64_t)seg * a.kmax * a.dim;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  if (tid < ES_TP) { best_v[tid] = -FLT_MAX; second_v[tid] = -FLT_MAX; best_i[tid] = 0x7fffffff; }

  for (int kb = 0; kb < K; kb += ES_TK) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int d0 = 0; d0 < a.dim; d0 += ES_DC) {
      __syncthreads();
#pragma unroll
      for (int r = 0; r < (ES_TP * ES_DC) / ES_THREADS; ++r) {
        const int idx = tid + ES_THREADS * r;
        const int row = idx >> 5, dd = idx & 31;
        const int d = d0 + dd;
        Xs[row][dd] = (row < np && d < a.dim) ? a.x[(p0 + row) * a.dim + d] : 0.f;
        const int k = kb + row;
        Cs[row][dd] = (k < K && d < a.dim) ? cbase[(int64_t)k * a.dim + d] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int dd = 0; dd < ES_DC; ++dd) {
        float xa[4], cb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xa[i] = Xs[ty + 16 * i][dd];
#pragma unroll
        for (int j = 0; j < 4; ++j) cb[j] = Cs[tx + 16 * j][dd];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], cb[j], acc[i][j]);
      }
    }
    // per pixel: best / runner-up over this thread's 4 centroids, then over the 16 tx lanes
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float bv = -FLT_MAX, sv = -FLT_MAX;
      int bi = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = kb + tx + 16 * j;
        if (k < K) merge_best(bv, bi, sv, acc[i][j], k, -FLT_MAX);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(FULL, bv, o);
        const int oi = __shfl_xor_sync(FULL, bi, o);
        const float osv = __shfl_xor_sync(FULL, sv, o);
        merge_best(bv, bi, sv, ov, oi, osv);
      }
      if (tx == 0) {
        const int px = ty + 16 * i;
        float rb = best_v[px], rs = second_v[px];
        int ri = best_i[px];
        merge_best(rb, ri, rs, bv, bi, sv);
        best_v[px] = rb; second_v[px] = rs; best_i[px] = ri;
      }
    }
  }
  __syncthreads();
  if (tid < np) {
    const int64_t pix = p0 + tid;
    a.keys_out[pix] = seg * a.kmax + best_i[tid];
    if (best_v[tid] - second_v[tid] <= thr) {
      const int slot = atomicAdd(a.fix.count, 1);
      if (slot < a.fix.capacity) {
        a.fix.pixels[slot] = (int32_t)pix;
        a.fix.segs[slot] = seg;
        if (a.fix.cand) a.fix.cand[(int64_t)slot * FIX_MAX_CAND] = 0xFFFF;
        atomicAdd(a.fix.count + 1, 1);
      }
    }
  }
}

// ---------------------------------------------------------------- float64 re-decision
constexpr int FIX_WARPS = 4;

__device__ __forceinline__ int upper_bound_off(const int64_t* a, int n, int64_t v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// One warp per listed pixel.  The pixel row is read once into registers (all
// loads in flight together); centroid rows are read four at a time so a scan of
// every cluster costs K/4 L2 round trips, not K.  Every dot product is the same
// fixed-order float64 sum (lane-strided partials, xor tree), so the decision
// does not depend on which pass listed the pixel.
template <int FIX_NV>
__device__ __forceinline__ double fix_dot(const float (&xr)[FIX_NV], const float* __restrict__ cr, int dim, int lane) {
  float cv[FIX_NV];
#pragma unroll
  for (int m = 0; m < FIX_NV; ++m) {
    const int d = lane + 32 * m;
    cv[m] = d < dim ? cr[d] : 0.f;
  }
  double s = 0.0;
#pragma unroll
  for (int m = 0; m < FIX_NV; ++m)
    if (lane + 32 * m < dim) s = fma((double)xr[m], (double)cv[m], s);
  return s;
}

template <int FIX_NV>
__global__ void __launch_bounds__(FIX_WARPS * 32) estep_fixup_kernel(const EStepArgs a) {
  const int lane = threadIdx.x & 31;
  const int total = min((int64_t)*a.fix.count, a.fix.capacity);
  const int warps = gridDim.x * FIX_WARPS;
  for (int e = blockIdx.x * FIX_WARPS + (threadIdx.x >> 5); e < total; e += warps) {
    const int64_t pix = a.fix.pixels[e];
    const float* xrow = a.x + pix * a.dim;
    float xr[FIX_NV];                               // issued first: the segment search below overlaps their latency
#pragma unroll
    for (int m = 0; m < FIX_NV; ++m) {
      const int d = lane + 32 * m;
      xr[m] = d < a.dim ? ld_stream(xrow + d) : 0.f;
    }
    const uint16_t* cand = a.fix.cand ? a.fix.cand + (int64_t)e * FIX_MAX_CAND : nullptr;
    const bool all = !cand || cand[0] == 0xFFFF;
    const int seg = upper_bound_off(a.seg_offsets, a.S + 1, pix) - 1;
    const int K = a.seg_k ? a.seg_k[seg] : a.kmax;
    const float* cbase = a.centroids + (int64_t)seg * a.kmax * a.dim;
    double bv = -DBL_MAX;
    int bi = 0x7fffffff;
    if (all) {
      for (int k0 = 0; k0 < K; k0 += 4) {
        double s4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          s4[u] = fix_dot(xr, cbase + (int64_t)min(k0 + u, K - 1) * a.dim, a.dim, lane);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double s = warp_sum(s4[u]);
          const int k = k0 + u;
          if (k < K && (s > bv || (s == bv && k < bi))) { bv = s; bi = k; }
        }
      }
    } else {
      // candidates two at a time (the usual entry lists exactly two): both centroid rows in flight together
      for (int c = 0; c < FIX_MAX_CAND; c += 2) {
        const int k0 = cand[c];
        if (k0 == 0xFFFF) break;
        const int k1 = cand[c + 1];
        const bool two = k1 != 0xFFFF;
        const double p0 = fix_dot(xr, cbase + (int64_t)min(k0, K - 1) * a.dim, a.dim, lane);
        const double p1 = fix_dot(xr, cbase + (int64_t)min(two ? k1 : k0, K - 1) * a.dim, a.dim, lane);
        const double s0 = warp_sum(p0), s1 = warp_sum(p1);
        if (k0 < K && (s0 > bv || (s0 == bv && k0 < bi))) { bv = s0; bi = k0; }
        if (two && k1 < K && (s1 > bv || (s1 == bv && k1 < bi))) { bv = s1; bi = k1; }
        if (!two) break;
      }
    }
    if (lane == 0) a.keys_out[pix] = seg * a.kmax + bi;
  }
}

// Eight lanes per listed pixel, four pixels per warp.  The warp-per-pixel kernel above is latency
// bound (r2 measurement at the benchmark shape: 1.3 pixels/ns = 1.4 TB/s of row reads, each warp
// walking pixel id -> segment search -> row -> candidate rows one dependent round trip after the
// other).  Here the segment comes with the list entry, a warp has four rows and their candidate
// centroid rows in flight at once, and rows are read as float2 when dim is even.  Same rule: fixed
// order float64 dot products (lane-strided partials, xor tree over the 8 lanes), ties to the lowest
// index.  Entries that ask for a scan of every cluster are handed to the whole warp afterwards.
template <int NV, int VEC>
__device__ __forceinline__ void fix8_load(const float* __restrict__ row, int dim, int sub, bool on, float (&v)[NV * VEC]) {
#pragma unroll
  for (int m = 0; m < NV; ++m) {
    const int i = (sub + 8 * m) * VEC;
    if (VEC == 2) {
      float2 t = make_float2(0.f, 0.f);
      if (on && i < dim) t = *reinterpret_cast<const float2*>(row + i);     // dim even: i + 1 < dim too
      v[2 * m] = t.x; v[2 * m + 1] = t.y;
    } else {
      v[m] = (on && i < dim) ? row[i] : 0.f;
    }
  }
}

// fp32 dot product and sum of |x_d c_d| over the 8 lanes of a group (lane-strided FMA chains, xor tree)
template <int NV, int VEC>
__device__ __forceinline__ void fix8_dot32(const float (&x)[NV * VEC], const float (&c)[NV * VEC], float& dot, float& mag) {
  float s = 0.f, a = 0.f;
#pragma unroll
  for (int m = 0; m < NV * VEC; ++m) { s = fmaf(x[m], c[m], s); a = fmaf(fabsf(x[m]), fabsf(c[m]), a); }
  s += __shfl_xor_sync(FULL, s, 4); a += __shfl_xor_sync(FULL, a, 4);
  s += __shfl_xor_sync(FULL, s, 2); a += __shfl_xor_sync(FULL, a, 2);
  s += __shfl_xor_sync(FULL, s, 1); a += __shfl_xor_sync(FULL, a, 1);
  dot = s; mag = a;
}

// Two stages.  (1) fp32: every candidate's dot product comes with a rigorous bound on its rounding error,
// gamma_n * sum|x_d c_d| with n = the longest chain of roundings (NV*VEC fused multiply-adds + 3 tree adds),
// computed from the actual rows (no unit-norm assumption).  When the best candidate's interval lies strictly
// above every other one's, it is the arg-max of the exact -- hence of the float64 -- products and the row is
// decided without a single fp32->fp64 conversion (r2: those conversions, 816 per row on the quarter-rate XU
// pipe, were 45 % of this kernel).  (2) the few rows that stay undecided (gap below ~5e-6, exact duplicates,
// "scan every cluster" entries) go through fixed-order float64 dot products, ties to the lowest index.
template <int NV, int VEC>
__global__ void __launch_bounds__(FIX_WARPS * 32, NV > 17 ? 1 : 4) estep_fixup8_kernel(const EStepArgs a, const int exp_flags) {
  const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
  const int total = (int)min((int64_t)*a.fix.count, a.fix.capacity);
  const int stride = gridDim.x * FIX_WARPS * 4;
  constexpr float GAMMA = (float)(NV * VEC + 3 + 2) * 5.9604645e-8f * 1.01f;   // +2: the bound itself is computed in fp32
  for (int e0 = (blockIdx.x * FIX_WARPS + (threadIdx.x >> 5)) * 4; e0 < total; e0 += stride) {
    const int e = e0 + grp;
    const bool act = e < total;
    const int64_t pix = act ? a.fix.pixels[e] : 0;
    const int seg = act ? a.fix.segs[e] : 0;
    float xr[NV * VEC];
    fix8_load<NV, VEC>(a.x + pix * a.dim, a.dim, sub, act && !(exp_flags & 1), xr);
    const uint16_t* cand = a.fix.cand + (int64_t)e * FIX_MAX_CAND;
    // the eight candidate slots of the entry: lane `sub` of the group reads slot `sub`
    int mine = (act && a.fix.cand) ? (int)cand[sub] : 0xFFFF;
    const int first = __shfl_sync(FULL, mine, grp * 8);
    const bool all = act && first == 0xFFFF;
    // slots after the first 0xFFFF are not part of the list
    const unsigned ends = __ballot_sync(FULL, mine == 0xFFFF) >> (grp * 8) & 0xFFu;
    const int ncand = all ? 0 : (ends ? __ffs(ends) - 1 : FIX_MAX_CAND);
    const int K = a.seg_k ? a.seg_k[seg] : a.kmax;
    const float* cbase = a.centroids + (int64_t)seg * a.kmax * a.dim;
    float bv = -FLT_MAX, be = 0.f, others_hi = -FLT_MAX;     // best value, its error bound, upper end of the rest
    int bi = 0x7fffffff;
    for (int c = 0; c < FIX_MAX_CAND; c += 2) {
      if (!__any_sync(FULL, c < ncand)) break;
      const int k0 = __shfl_sync(FULL, mine, grp * 8 + c);
      const int k1 = __shfl_sync(FULL, mine, grp * 8 + c + 1);
      const bool on0 = c < ncand && k0 < K, on1 = c + 1 < ncand && k1 < K;
      float c0[NV * VEC], c1[NV * VEC];
      fix8_load<NV, VEC>(cbase + (int64_t)min(k0, K - 1) * a.dim, a.dim, sub, on0 && !(exp_flags & 2), c0);
      fix8_load<NV, VEC>(cbase + (int64_t)min(k1, K - 1) * a.dim, a.dim, sub, on1 && !(exp_flags & 2), c1);
      float s0, a0, s1, a1;
      fix8_dot32<NV, VEC>(xr, c0, s0, a0);
      fix8_dot32<NV, VEC>(xr, c1, s1, a1);
      const float e0b = a0 * GAMMA + 1e-37f, e1b = a1 * GAMMA + 1e-37f;
      if (on0) {
        if (s0 > bv) { others_hi = fmaxf(others_hi, bv + be); bv = s0; be = e0b; bi = k0; }
        else others_hi = fmaxf(others_hi, s0 + e0b);
      }
      if (on1) {
        if (s1 > bv) { others_hi = fmaxf(others_hi, bv + be); bv = s1; be = e1b; bi = k1; }
        else others_hi = fmaxf(others_hi, s1 + e1b);
      }
    }
    const bool decided = act && !all && bi != 0x7fffffff && (bv - be > others_hi);
    if (decided && sub == 0) a.keys_out[pix] = seg * a.kmax + bi;

    // the rest (rare): float64, the whole warp on one entry at a time
    unsigned slow = __ballot_sync(FULL, act && !decided && sub == 0);
    while (slow) {
      const int src = __ffs(slow) - 1;
      slow &= slow - 1;
      const int64_t apix = __shfl_sync(FULL, pix, src);
      const int aseg = __shfl_sync(FULL, seg, src);
      const bool aall = __shfl_sync(FULL, (int)all, src) != 0;
      const int anc = __shfl_sync(FULL, ncand, src);
      const int aK = a.seg_k ? a.seg_k[aseg] : a.kmax;
      const float* ab = a.centroids + (int64_t)aseg * a.kmax * a.dim;
      const float* xrow = a.x + apix * a.dim;
      double abv = -DBL_MAX;
      int abi = 0x7fffffff;
      const int n_it = aall ? aK : anc;
      for (int it = 0; it < n_it; ++it) {
        const int k = aall ? it : __shfl_sync(FULL, mine, src + it);      // src is lane 0 of its group
        if (k >= aK) continue;                                             // warp-uniform
        const float* cr = ab + (int64_t)k * a.dim;
        double sacc = 0.0;
        for (int d = lane; d < a.dim; d += 32) sacc = fma((double)xrow[d], (double)cr[d], sacc);
        sacc = warp_sum(sacc);
        if (sacc > abv || (sacc == abv && k < abi)) { abv = sacc; abi = k; }
      }
      if (lane == 0) a.keys_out[apix] = aseg * a.kmax + abi;
    }
  }
}

int estep_simt(const EStepArgs& a, cudaStream_t st) {
  // |fl(dot) - dot| <= gamma_dim * sum|x_d c_d| <= dim*2^-24*(1+tiny) for unit rows;
  // two such errors can reorder a pair, so re-decide below twice that (plus slack).
  const float thr = 2.f * (a.dim + 2) * 5.9604645e-8f * 1.02f + 1e-7f;
  dim3 grid((unsigned)a.tiles.bound, (unsigned)ceil_div64(a.tiles.tile, ES_TP));
  estep_simt_kernel<<<grid, ES_THREADS, 0, st>>>(a, thr);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int estep_fixup(const EStepArgs& a, cudaStream_t st) {
  static const bool legacy = getenv("HSG_FIXUP_LEGACY") != nullptr;     // A/B switch for profiling only
  static const int fexp = getenv("HSG_FIXUP_EXP") ? atoi(getenv("HSG_FIXUP_EXP")) : 0;     // timing experiments only
  const bool aligned8 = (a.dim % 2 == 0) && ((uintptr_t)a.x % 8 == 0) && ((uintptr_t)a.centroids % 8 == 0);
  if (!legacy && a.fix.cand && aligned8 && a.dim <= 272) {
    estep_fixup8_kernel<17, 2><<<num_sms() * 4, FIX_WARPS * 32, 0, st>>>(a, fexp);
    HSG_LAUNCH_CHECK();
    return HSG_OK;
  }
  if (!legacy && a.fix.cand && a.dim <= 136) {
    estep_fixup8_kernel<17, 1><<<num_sms() * 4, FIX_WARPS * 32, 0, st>>>(a, fexp);
    HSG_LAUNCH_CHECK();
    return HSG_OK;
  }
  if (!legacy && a.fix.cand && aligned8 && a.dim <= 528) {
    // D = 512 (+ trailing features): the same two-stage rule with rows of 66 floats per lane; two blocks per SM
    // (three rows in registers).  The all-float64 kernel below took 1.2 of 4.9 ms per iteration at N = 1e7, K = 256
    // and 103 of 137 ms at K = 2048.
    estep_fixup8_kernel<33, 2><<<num_sms() * 2, FIX_WARPS * 32, 0, st>>>(a, fexp);
    HSG_LAUNCH_CHECK();
    return HSG_OK;
  }
  if (a.dim <= 32 * 9)
    estep_fixup_kernel<9><<<num_sms() * 16, FIX_WARPS * 32, 0, st>>>(a);
  else
    estep_fixup_kernel<20><<<num_sms() * 8, FIX_WARPS * 32, 0, st>>>(a);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

__global__ void copy_count_kernel(const int32_t* c, int64_t* out) { out[0] = c[0]; out[1] = c[1]; }

// ---------------------------------------------------------------- orchestration
// Incremental M-step.  After the first iterations only a few per cent of the labels change per
// iteration (measured on the benchmark's iid data: 78 %, 34 %, 14 %, 9 %, 7 %, ... 2.4 %), so the
// cluster sums are kept as running float64 sums and updated from the rows that moved
// (sum[old] -= x, sum[new] += x) instead of re-reading every row.  The moved rows become a list
// of signed entries that goes through the same deterministic sort + run-sum machinery as a full
// pass (float64 partials), so results do not depend on scheduling.  Whether an iteration takes the
// delta or the full pass is decided on the device (entries <= capacity); both are enqueued and
// the other one returns at once (Gate).
struct DeltaPlan {
  SegReducePlan sr;        // entries as the sorted objects; shares perm/hist/bin arrays with the full plan
  int32_t* flag;           // [2]
  int32_t* keys_prev;      // [N]
  int32_t* tile_entries;   // [full tiles bound]
  int64_t* eoff;           // [S+1]
  uint32_t* erow;          // [cap]
  int32_t* ekey;           // [cap]
  int64_t cap;
  double* sums;            // [S*kmax*dim] running sums
  int32_t* members;        // [S*kmax]
  long long* local_sums;   // [S*kmax*dim] exact mode: this shard's running int64 sums
};

struct KmPlan {
  SegReducePlan sr;
  float* centroids;      // [S*kmax*dim]
  FixList fix;
  TcState tc;
  DeltaPlan d;
};

constexpr int64_t KM_DELTA_MIN_ROWS = 1 << 18;
constexpr double KM_DELTA_MAX_FRACTION = 0.4;   // of the rows may have moved (2 entries each) for a delta pass

static void km_carve(Carver& c, KmPlan& p, int64_t N, int dim, int S, int kmax, int64_t max_seg_len,
                     int d16, bool exact = false) {
  // exact: int64 fixed-point sums (the row-sharded loop, hsg_kmeans_dist_*)
  sr_carve(c, p.sr, N, dim, S, kmax, max_seg_len, exact);
  p.centroids = c.take<float>((int64_t)S * kmax * dim);
  p.fix.count = c.take<int32_t>(2);    // [0] listed pixels, [1] of those: scans over every cluster
  p.fix.capacity = N;
  p.fix.pixels = c.take<int32_t>(N);
  p.fix.segs = c.take<int32_t>(N);
  p.fix.cand = c.take<uint16_t>(N * FIX_MAX_CAND);
  tc_carve(c, p.tc, S, kmax, d16 > 0 ? d16 : 64, N);
  // delta plan: own tiles / keys / float64 pieces, everything else aliases the full plan (only
  // one of the two passes runs in an iteration)
  DeltaPlan& d = p.d;
  d.cap = (int64_t)(2.0 * KM_DELTA_MAX_FRACTION * (double)N);
  d.sr = p.sr;
  d.sr.N = d.cap;
  const int64_t dseg = max_seg_len * 2 < d.cap ? max_seg_len * 2 : d.cap;
  d.sr.tiles.tile = sr_tile_size(dseg > 0 ? dseg : 1);
  d.sr.tiles.bound = sr_tiles_bound(d.cap, S, d.sr.tiles.tile);
  if (d.sr.tiles.bound > p.sr.tiles.bound) d.sr.tiles.bound = p.sr.tiles.bound;   // tile_hist is shared
  d.sr.tiles.seg = c.take<int32_t>(d.sr.tiles.bound);
  d.sr.tiles.begin = c.take<int64_t>(d.sr.tiles.bound);
  d.sr.tiles.end = c.take<int64_t>(d.sr.tiles.bound);
  d.sr.tiles.seg_first = c.take<int32_t>(S + 1);
  d.sr.tiles.count = c.take<int32_t>(1);
  d.flag = c.take<int32_t>(2);
  d.keys_prev = c.take<int32_t>(N);
  d.tile_entries = c.take<int32_t>(p.sr.tiles.bound);
  d.eoff = c.take<int64_t>(S + 1);
  d.erow = c.take<uint32_t>(d.cap + 2);
  d.ekey = c.take<int32_t>(d.cap + 2);
  d.sr.keys = d.ekey;
  d.sr.pieces = reinterpret_cast<float*>(c.take<double>((ceil_div64(d.cap, SR_RUN) + p.sr.bins + 2) * dim));
  d.sums = c.take<double>(p.sr.bins * dim);
  d.members = c.take<int32_t>(p.sr.bins);
  d.local_sums = exact ? c.take<long long>(p.sr.bins * dim) : nullptr;
}

// First M-step from the run sums the prep kernel emitted (hsg_prep_sums_f32): one warp per (segment, cluster)
// scans the segment's run table in slot order and adds the matching partial sums in float64 -- fixed order,
// no re-read of the rows.  Runs only when the table is complete (*overflow == 0); otherwise the gated
// ordinary pass has produced the sums.
struct RunSums {
  const float* sums;        // [S*runs_per_segment, dim]
  const int32_t* cluster;   // [S*runs_per_segment]
  const int32_t* count;     // [S*runs_per_segment]
  const int32_t* overflow;  // [1]
  int64_t runs_per_segment;
};

template <int NV>
__global__ void __launch_bounds__(256) runsum_combine_kernel(const RunSums r, int S, int kmax, int dim,
                                                             double* __restrict__ sums, int32_t* __restrict__ members,
                                                             float* __restrict__ out) {
  if (*r.overflow != 0) return;
  const int lane = threadIdx.x & 31;
  const int64_t key = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (key >= (int64_t)S * kmax) return;
  const int seg = (int)(key / kmax), k = (int)(key % kmax);
  const int64_t base = (int64_t)seg * r.runs_per_segment;
  double acc[NV];
#pragma unroll
  for (int m = 0; m < NV; ++m) acc[m] = 0.0;
  int mem = 0;
  // four slots per lane and step (runs_per_segment is a multiple of HSG_PREP_RUNS = 4: 16-byte loads)
  for (int64_t j0 = 0; j0 < r.runs_per_segment; j0 += 128) {
    const int64_t j = j0 + 4 * lane;
    int4 c4 = make_int4(-1, -1, -1, -1);
    if (j < r.runs_per_segment) c4 = *reinterpret_cast<const int4*>(r.cluster + base + j);
    const int cs[4] = {c4.x, c4.y, c4.z, c4.w};
    unsigned any = __ballot_sync(FULL, cs[0] == k || cs[1] == k || cs[2] == k || cs[3] == k);
    while (any) {                                  // lanes in slot order, their four slots in order
      const int b = __ffs(any) - 1;
      any &= any - 1;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (__shfl_sync(FULL, cs[u], b) != k) continue;
        const int64_t slot = base + j0 + 4 * b + u;
        mem += r.count[slot];
        const float* row = r.sums + slot * dim;
#pragma unroll
        for (int m = 0; m < NV; ++m) {
          const int d = lane + 32 * m;
          if (d < dim) acc[m] += (double)row[d];
        }
      }
    }
  }
  if (lane == 0) members[key] = mem;
  float ss = 0.f;
  float f[NV];
#pragma unroll
  for (int m = 0; m < NV; ++m) {
    const int d = lane + 32 * m;
    f[m] = 0.f;
    if (d < dim) {
      const double a = mem == 0 ? 0.0 : acc[m];
      sums[key * dim + d] = a;
      f[m] = (float)a;
      ss = fmaf(f[m], f[m], ss);
    }
  }
  const float n = safe_norm(warp_sum(ss));
#pragma unroll
  for (int m = 0; m < NV; ++m) {
    const int d = lane + 32 * m;
    if (d < dim) out[key * dim + d] = f[m] / n;
  }
}

// M-step for the launch-bound regime (training shapes: 12 x 28x28 pixels, K = 16): ONE launch instead
// of hist / scan / scatter / gather / combine.  It also resets the re-decision list counters the E-step that
// follows appends to.  Reads: every row once, the keys K times (K * 4 bytes per pixel).
// One (segment, cluster) bin per CTA of 8 warps: warp w scans the w-th eighth of the segment's keys 32 at a time and
// adds its members' rows in index order (fp32 over runs of 32 members, folded into float64); the eight partial
// sums are added in warp order -- a fixed order, the accuracy class of the general path.
template <int NV>
__device__ __forceinline__ void mstep_small_bin(const float* __restrict__ x, int dim, int64_t b, int64_t e, int key,
                                                const int32_t* keys, float* out_row,
                                                double (*part)[NV * 32]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t len = e - b;
  const int64_t per = ((len + 7) / 8 + 31) / 32 * 32;            // rows per warp, a multiple of 32
  const int64_t wb = b + warp * per, we = min(e, wb + per);
  double acc[NV];
  float run[NV];
#pragma unroll
  for (int m = 0; m < NV; ++m) { acc[m] = 0.0; run[m] = 0.f; }
  int in_run = 0;
  for (int64_t g0 = wb; g0 < we; g0 += 128) {
   int kq[4];                                               // four groups of 32 keys in flight
#pragma unroll
   for (int q4 = 0; q4 < 4; ++q4) {
     const int64_t i = g0 + 32 * q4 + lane;
     kq[q4] = i < we ? __ldcg(keys + i) : -1;               // L2: another SM may have written it (persistent kernel)
   }
#pragma unroll
   for (int q4 = 0; q4 < 4; ++q4) {
    const int64_t i0 = g0 + 32 * q4;
    unsigned hit = __ballot_sync(FULL, kq[q4] == key);
    while (hit) {
      // two members' rows in flight (the loop is a chain of L2 round trips); added in index order as before
      const int j0 = __ffs(hit) - 1;
      hit &= hit - 1;
      const int j1 = hit ? __ffs(hit) - 1 : -1;
      if (j1 >= 0) hit &= hit - 1;
      const float* row0 = x + (i0 + j0) * dim;
      const float* row1 = x + (i0 + (j1 >= 0 ? j1 : j0)) * dim;
      float v0[NV], v1[NV];
#pragma unroll
      for (int m = 0; m < NV; ++m) {
        const int d = lane + 32 * m;
        v0[m] = d < dim ? row0[d] : 0.f;
        v1[m] = (d < dim && j1 >= 0) ? row1[d] : 0.f;
      }
#pragma unroll
      for (int m = 0; m < NV; ++m) run[m] += v0[m];
      if (++in_run == 32) {
#pragma unroll
        for (int m = 0; m < NV; ++m) { acc[m] += (double)run[m]; run[m] = 0.f; }
        in_run = 0;
      }
      if (j1 >= 0) {
#pragma unroll
        for (int m = 0; m < NV; ++m) run[m] += v1[m];
        if (++in_run == 32) {
#pragma unroll
          for (int m = 0; m < NV; ++m) { acc[m] += (double)run[m]; run[m] = 0.f; }
          in_run = 0;
        }
      }
    }
   }
  }
#pragma unroll
  for (int m = 0; m < NV; ++m) part[warp][lane + 32 * m] = acc[m] + (double)run[m];
  __syncthreads();
  if (warp == 0) {
    float ss = 0.f;
    float f[NV];
#pragma unroll
    for (int m = 0; m < NV; ++m) {
      double a = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) a += part[w][lane + 32 * m];
      f[m] = (float)a;
      if (lane + 32 * m < dim) ss = fmaf(f[m], f[m], ss);
    }
    const float n = safe_norm(warp_sum(ss));
#pragma unroll
    for (int m = 0; m < NV; ++m) {
      const int d = lane + 32 * m;
      if (d < dim) out_row[d] = f[m] / n;
    }
  }
  __syncthreads();
}

template <int NV>
__global__ void __launch_bounds__(256) mstep_small_kernel(const float* __restrict__ x, int dim,
                                                          const int64_t* __restrict__ seg_offsets, int S, int kmax,
                                                          const int32_t* __restrict__ keys, float* __restrict__ out,
                                                          int32_t* __restrict__ fix_count) {
  __shared__ double part[8][NV * 32];
  if (blockIdx.x == 0 && threadIdx.x < 2) fix_count[threadIdx.x] = 0;
  const int64_t bin = blockIdx.x;
  const int seg = (int)(bin / kmax);
  mstep_small_bin<NV>(x, dim, seg_offsets[seg], seg_offsets[seg + 1], (int)bin, keys, out + bin * dim, part);
}

constexpr int64_t KM_SMALL_MAX_SEG = 16384;    // rows per segment up to which the one-launch M-step is used

// ---------------------------------------------------------------- the whole loop in ONE launch (launch-bound regime)
// Training shapes (12 x 28x28 pixels, D = 128, K = 16, T = 15: SURVEY 3.1) spend their time between kernels: 54 launches,
// 0.76 ms per call with the one-launch M-step above.  Here every iteration of every image runs inside one cooperative
// kernel: keys from the initial labels | (M-step | E-step) x T | labels, separated by grid barriers; the data (5 MB)
// stays in L2.  Same arithmetic as the multi-launch path, so the same labels bit for bit: the M-step is
// mstep_small_bin, the E-step the fp32 tile product of estep_simt_kernel, and a pixel whose top-2 gap is inside the
// rigorous fp32 bound is re-decided on the spot by the float64 scan of estep_fixup_kernel (same summation order).
// Anything another SM wrote during the kernel (keys, centroids) is read through L2 (ld.global.cg).
struct SmallArgs {
  EStepArgs a;                 // a.centroids is written here (M-step) and read (E-step)
  float* centroids;
  const int64_t* init_labels;
  int64_t* labels_out;
  float* centroids_out;        // may be NULL
  int iterations;
  float thr;
  unsigned* bar;               // [1] zeroed before the launch
  int exp_flags;               // timing experiments (HSG_SMALL_EXP), 0 in production: 1 no M-step, 2 no E-step
  int64_t max_seg_len;         // longest segment: sub-tiles beyond it are empty in every tile
};

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    while (v < target) {
      __nanosleep(64);                                     // a few hundred pollers on one address slow the arrivals down
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    }
    __threadfence();
  }
  __syncthreads();
}

// One item = 64 pixels of one image against all of its centroids.  The whole [64 x dim] pixel tile and (up to 64 of)
// the centroids are staged in shared memory with ONE round of loads -- an item is a chain of L2 round trips, not a
// throughput problem (r2: the 32-dim chunks of estep_simt_kernel cost 5 dependent rounds per item, 18 us per phase).
// NJ = centroid columns per thread (K <= 16: one).
struct SmallTileSmem {
  float best_v[ES_TP], second_v[ES_TP];
  int best_i[ES_TP];
};

template <int NV, int NJ>
__device__ __forceinline__ void estep_small_tile(const EStepArgs& a, int ti, int sub, float thr, SmallTileSmem& sm,
                                                 float* __restrict__ Xs, float* __restrict__ Cs, int dp) {
  const int64_t p0 = a.tiles.begin[ti] + (int64_t)sub * ES_TP;
  const int64_t pe = a.tiles.end[ti];
  if (p0 >= pe) return;                                    // block-uniform
  const int np = (int)min((int64_t)ES_TP, pe - p0);
  const int seg = a.tiles.seg[ti];
  const int K = a.seg_k ? a.seg_k[seg] : a.kmax;
  const float* cbase = a.centroids + (int64_t)seg * a.kmax * a.dim;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int ld = dp + 1;
  constexpr int CK = 16 * NJ;                              // centroid rows staged per block
  __syncthreads();                                         // the previous item's readers are done with the arrays
  if (tid < ES_TP) { sm.best_v[tid] = -FLT_MAX; sm.second_v[tid] = -FLT_MAX; sm.best_i[tid] = 0x7fffffff; }
  // pixel rows: asynchronous 4-byte copies, all in flight together (read-only data: L1 is fine)
  const int wrp = tid >> 5, ln = tid & 31;
  for (int row = wrp; row < ES_TP; row += ES_THREADS / 32) {
    for (int d = ln; d < dp; d += 32) {
      float* dst = Xs + row * ld + d;
      if (row < np && d < a.dim) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(a.x + (p0 + row) * a.dim + d) : "memory");
      } else {
        *dst = 0.f;
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int kb = 0; kb < K; kb += CK) {
    if (kb > 0) __syncthreads();                           // the previous block's products are done with Cs
    // centroid rows through L2 (another SM wrote them in the M-step phase): every load of a row pair issued first
    for (int row = wrp; row < CK; row += 2 * (ES_THREADS / 32)) {
      float v0[NV], v1[NV];
      const int k0 = kb + row, k1 = k0 + ES_THREADS / 32;
#pragma unroll
      for (int m = 0; m < NV; ++m) {
        const int d = ln + 32 * m;
        v0[m] = (k0 < K && d < a.dim) ? __ldcg(cbase + (int64_t)k0 * a.dim + d) : 0.f;
        v1[m] = (k1 < K && d < a.dim && row + ES_THREADS / 32 < CK) ? __ldcg(cbase + (int64_t)k1 * a.dim + d) : 0.f;
      }
#pragma unroll
      for (int m = 0; m < NV; ++m) {
        const int d = ln + 32 * m;
        if (d < dp) {
          Cs[row * ld + d] = v0[m];
          if (row + ES_THREADS / 32 < CK) Cs[(row + ES_THREADS / 32) * ld + d] = v1[m];
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    float acc[4][NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int dd = 0; dd < dp; ++dd) {
      float xa[4], cb[NJ];
#pragma unroll
      for (int i = 0; i < 4; ++i) xa[i] = Xs[(ty + 16 * i) * ld + dd];
#pragma unroll
      for (int j = 0; j < NJ; ++j) cb[j] = Cs[(tx + 16 * j) * ld + dd];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(xa[i], cb[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float bv = -FLT_MAX, sv = -FLT_MAX;
      int bi = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int k = kb + tx + 16 * j;
        if (k < K) merge_best(bv, bi, sv, acc[i][j], k, -FLT_MAX);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(FULL, bv, o);
        const int oi = __shfl_xor_sync(FULL, bi, o);
        const float osv = __shfl_xor_sync(FULL, sv, o);
        merge_best(bv, bi, sv, ov, oi, osv);
      }
      if (tx == 0) {
        const int px = ty + 16 * i;
        float rb = sm.best_v[px], rs = sm.second_v[px];
        int ri = sm.best_i[px];
        merge_best(rb, ri, rs, bv, bi, sv);
        sm.best_v[px] = rb; sm.second_v[px] = rs; sm.best_i[px] = ri;
      }
    }
  }
  __syncthreads();
  // labels; near-ties by the float64 scan of every cluster (estep_fixup_kernel's rule and summation order)
  const int lane = tid & 31;
  for (int px = tid >> 5; px < np; px += ES_THREADS / 32) {
    const int64_t pix = p0 + px;
    int bi = sm.best_i[px];
    if (sm.best_v[px] - sm.second_v[px] <= thr) {          // warp-uniform
      const float* xrow = a.x + pix * a.dim;
      float xr[NV];
#pragma unroll
      for (int m = 0; m < NV; ++m) { const int d = lane + 32 * m; xr[m] = d < a.dim ? xrow[d] : 0.f; }
      double bv = -DBL_MAX;
      bi = 0x7fffffff;
      for (int k0 = 0; k0 < K; k0 += 4) {                  // four centroid rows in flight: K/4 L2 round trips, not K
        float cv[4][NV];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float* cr = cbase + (int64_t)min(k0 + u, K - 1) * a.dim;
#pragma unroll
          for (int m = 0; m < NV; ++m) { const int d = lane + 32 * m; cv[u][m] = d < a.dim ? __ldcg(cr + d) : 0.f; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < NV; ++m) if (lane + 32 * m < a.dim) s = fma((double)xr[m], (double)cv[u][m], s);
          s = warp_sum(s);
          const int k = k0 + u;
          if (k < K && (s > bv || (s == bv && k < bi))) { bv = s; bi = k; }
        }
      }
    }
    if (lane == 0) a.keys_out[pix] = seg * a.kmax + bi;
  }
}

template <int NV, int NJ>
__global__ void __launch_bounds__(ES_THREADS) kmeans_small_persistent_kernel(const SmallArgs s) {
  // dynamic shared memory: the M-step's float64 partials [8][NV*32], or the E-step's pixel tile [64][dp+1] and
  // centroid block [16 NJ][dp+1]
  extern __shared__ __align__(16) unsigned char raw[];
  __shared__ SmallTileSmem tsm;
  double (*part)[NV * 32] = reinterpret_cast<double (*)[NV * 32]>(raw);
  const int dp = (s.a.dim + 31) / 32 * 32;
  float* Xs = reinterpret_cast<float*>(raw);
  float* Cs = Xs + ES_TP * (dp + 1);
  const EStepArgs& a = s.a;
  unsigned target = 0;
  const int n_tiles = *a.tiles.count;
  const int n_sub = (int)((min(a.tiles.tile, s.max_seg_len) + ES_TP - 1) / ES_TP);      // non-empty 64-pixel items per tile
  const int64_t bins = (int64_t)a.S * a.kmax;

  for (int ti = blockIdx.x; ti < n_tiles; ti += gridDim.x) {                   // keys from the initial labels
    const int seg = a.tiles.seg[ti];
    for (int64_t i = a.tiles.begin[ti] + threadIdx.x; i < a.tiles.end[ti]; i += blockDim.x) {
      int64_t k = s.init_labels[i];
      k = k < 0 ? 0 : (k >= a.kmax ? a.kmax - 1 : k);
      a.keys_out[i] = seg * a.kmax + (int)k;
    }
  }
  grid_barrier(s.bar, target);
  for (int it = 0; it < s.iterations; ++it) {
    for (int64_t bin = blockIdx.x; bin < bins && !(s.exp_flags & 1); bin += gridDim.x) {
      const int seg = (int)(bin / a.kmax);
      mstep_small_bin<NV>(a.x, a.dim, a.seg_offsets[seg], a.seg_offsets[seg + 1], (int)bin, a.keys_out,
                          s.centroids + bin * a.dim, part);
    }
    grid_barrier(s.bar, target);
    for (int64_t w = blockIdx.x; w < (int64_t)n_tiles * n_sub && !(s.exp_flags & 2); w += gridDim.x)
      estep_small_tile<NV, NJ>(a, (int)(w / n_sub), (int)(w % n_sub), s.thr, tsm, Xs, Cs, dp);
    grid_barrier(s.bar, target);
  }
  for (int ti = blockIdx.x; ti < n_tiles; ti += gridDim.x) {                   // labels out
    const int seg = a.tiles.seg[ti];
    for (int64_t i = a.tiles.begin[ti] + threadIdx.x; i < a.tiles.end[ti]; i += blockDim.x)
      s.labels_out[i] = (int64_t)(__ldcg(a.keys_out + i) - seg * a.kmax);
  }
  if (s.centroids_out) {
    const int64_t n = bins * a.dim;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      s.centroids_out[i] = __ldcg(s.centroids + i);
  }
}

// returns HSG_OK and sets *done when the persistent kernel took the call
template <int NV, int NJ>
static int launch_small_persistent(const SmallArgs& s, int64_t work, cudaStream_t st, bool* done) {
  const int dp = (s.a.dim + 31) / 32 * 32;
  size_t smem = (size_t)(ES_TP + 16 * NJ) * (dp + 1) * sizeof(float);
  if (smem < sizeof(double) * 8 * NV * 32) smem = sizeof(double) * 8 * NV * 32;
  if (smem > 200 * 1024) return HSG_OK;
  HSG_CUDA(cudaFuncSetAttribute(kmeans_small_persistent_kernel<NV, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  HSG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kmeans_small_persistent_kernel<NV, NJ>, ES_THREADS, smem));
  if (per_sm < 1) return HSG_OK;
  static const int per_env = getenv("HSG_SMALL_CTAS") ? atoi(getenv("HSG_SMALL_CTAS")) : 2;     // tuning knob
  if (per_sm > per_env) per_sm = per_env;
  int64_t grid = (int64_t)num_sms() * per_sm;
  if (grid > work) grid = work;
  if (grid < 1) grid = 1;
  HSG_CUDA(cudaMemsetAsync(s.bar, 0, sizeof(unsigned), st));
  void* args[] = {(void*)&s};
  HSG_CUDA(cudaLaunchCooperativeKernel((const void*)kmeans_small_persistent_kernel<NV, NJ>, dim3((unsigned)grid),
                                       dim3(ES_THREADS), args, smem, st));
  HSG_LAUNCH_CHECK();
  *done = true;
  return HSG_OK;
}

// ---------------------------------------------------------------- the whole loop inside one thread-block cluster per image
// At the training shapes an image is small enough to LIVE in shared memory: 28x28 rows of 130 floats are 408 KB,
// an eighth of that per CTA of an 8-CTA cluster.  Images are independent, so nothing has to cross the cluster: the
// rows are loaded once, and every iteration is  partial sums (local) | cluster barrier | centroids (distributed
// shared memory) | cluster barrier | assignment (local)  -- hardware barriers of ~1 us instead of grid barriers through
// L2, no traffic between iterations at all, no co-residency requirement (clusters queue like ordinary CTAs).
// The arithmetic is that of mstep_small_bin / estep_small_tile: CTA r of the cluster owns exactly the row range that
// warp r of the one-bin-per-CTA M-step sums (members in index order, fp32 runs of 32 folded into float64, the eight
// partials added in rank order), the products are the same fmaf chains and a near-tie takes the same float64 scan,
// so the labels are bit-identical to the other paths.
constexpr int KC_RANKS = 8;             // CTAs per cluster = row ranges per segment (portable cluster size)

__device__ __forceinline__ uint32_t kc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t kc_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t kc_mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void kc_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double kc_ld_f64(uint32_t caddr) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(caddr) : "memory");
  return v;
}
__device__ __forceinline__ float kc_lds(uint32_t saddr) {      // explicit shared-space load: a generic pointer into shared
  float v;                                                     // memory costs an S2R of the cluster window base per access
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void kc_st_f32(uint32_t caddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(caddr), "f"(v) : "memory");
}

struct ClusterKmArgs {
  const float* x;              // [N,dim]
  int dim, kmax, S;
  const int64_t* seg_offsets;
  const int32_t* seg_k;        // may be NULL
  const int64_t* init_labels;
  int64_t* labels_out;
  float* centroids_out;        // may be NULL: [S,kmax,dim]
  int iterations;
  float thr;
  int per_max;                 // rows per CTA the shared-memory carve-up was sized for
  int exp_flags;               // timing experiments (HSG_CLUSTER_EXP), 0 in production: 1 no partial sums, 2 no centroids, 4 no assignment
};

__host__ __device__ inline int64_t kc_rows_per_rank(int64_t len) { return ((len + 7) / 8 + 31) / 32 * 32; }

// shared-memory carve-up (floats unless noted), the same in every CTA of a launch:
//   XT  [dp][ldr]      the rows of this rank, TRANSPOSED (ldr = rows + 1, odd): the products read consecutive rows of one
//                      feature (a broadcast per row group), the partial sums and the float64 scan read a row across
//                      features -- lane stride ldr, conflict-free because it is odd
//   Cs  [CK][ld]       centroids, row-major (float64 scan; written by the ranks that finish them)
//   CT  [dp][CK]       centroids transposed (products)
//   part[kmax][NV*32]  float64 partial sums of every bin over this rank's rows
//   keys[rows], best_v/second_v/best_i[KC_TILE]
constexpr int KC_TILE = 128;            // rows per pass of the products: 32 row groups x 4 rows, 8 column groups x 2 NJ columns

template <int NV, int NJ>
struct KcLayout {
  int dp, ld, ldr, rows_pad;
  size_t xt, cs, ct, part, keys, best, total;        // byte offsets
  __host__ __device__ KcLayout(int dim, int kmax, int per) {
    constexpr int CK = 16 * NJ;
    dp = (dim + 31) / 32 * 32; ld = dp + 1;
    rows_pad = (per + 3) / 4 * 4;
    ldr = rows_pad + 1;
    xt = 0;
    cs = xt + sizeof(float) * (size_t)dp * ldr;
    ct = cs + sizeof(float) * (size_t)CK * ld;
    part = (ct + sizeof(float) * (size_t)dp * CK + 15) & ~(size_t)15;
    keys = part + sizeof(double) * (size_t)kmax * NV * 32;
    best = keys + sizeof(int) * (size_t)rows_pad;
    total = best + 3 * sizeof(float) * KC_TILE + 16;
  }
};

template <int NV, int NJ>
__global__ void __launch_bounds__(ES_THREADS) kmeans_cluster_kernel(const ClusterKmArgs s) {
  extern __shared__ __align__(16) unsigned char raw[];
  constexpr int CK = 16 * NJ;
  const KcLayout<NV, NJ> L(s.dim, s.kmax, s.per_max);
  const int dp = L.dp, ld = L.ld, ldr = L.ldr;
  float* XT = reinterpret_cast<float*>(raw + L.xt);
  float* Cs = reinterpret_cast<float*>(raw + L.cs);
  float* CT = reinterpret_cast<float*>(raw + L.ct);
  double* part = reinterpret_cast<double*>(raw + L.part);
  int* keys = reinterpret_cast<int*>(raw + L.keys);
  float* best_v = reinterpret_cast<float*>(raw + L.best);
  float* second_v = best_v + KC_TILE;
  int* best_i = reinterpret_cast<int*>(second_v + KC_TILE);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = kc_rank();
  const int seg = blockIdx.x / KC_RANKS;
  const int64_t b = s.seg_offsets[seg], e = s.seg_offsets[seg + 1];
  const int64_t per = kc_rows_per_rank(e - b);
  const int64_t wb = min(e, b + (int64_t)rank * per), we = min(e, wb + per);
  const int nrows = (int)(we - wb);
  const int K = s.seg_k ? s.seg_k[seg] : s.kmax;

  // ---- the rows of this rank, once: coalesced global reads (a warp per row), transposed into XT
  for (int i = tid; i < dp * ldr; i += ES_THREADS) XT[i] = 0.f;
  for (int i = tid; i < CK * ld; i += ES_THREADS) Cs[i] = 0.f;
  __syncthreads();
  for (int row0 = 0; row0 < nrows; row0 += 4 * (ES_THREADS / 32)) {
    float v[4][NV];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = row0 + warp + (ES_THREADS / 32) * u;
#pragma unroll
      for (int m = 0; m < NV; ++m) {
        const int d = lane + 32 * m;
        v[u][m] = (row < nrows && d < s.dim) ? s.x[(wb + row) * s.dim + d] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = row0 + warp + (ES_THREADS / 32) * u;
      if (row < nrows) {
#pragma unroll
        for (int m = 0; m < NV; ++m) {
          const int d = lane + 32 * m;
          if (d < s.dim) XT[(size_t)d * ldr + row] = v[u][m];
        }
      }
    }
  }
  for (int r = tid; r < L.rows_pad; r += ES_THREADS) {
    int64_t k = r < nrows ? s.init_labels[wb + r] : -1;
    if (r < nrows) k = k < 0 ? 0 : (k >= s.kmax ? s.kmax - 1 : k);
    keys[r] = (int)k;
  }
  __syncthreads();

  const uint32_t part_s = kc_smem_u32(part), cs_s = kc_smem_u32(Cs), xt_s = kc_smem_u32(XT);
  const int tx = tid & 7, ty = tid >> 3;
  for (int it = 0; it < s.iterations; ++it) {
    // ---- partial sums of this rank's rows, one warp per bin (mstep_small_bin's warp loop)
    for (int k = warp; k < s.kmax && !(s.exp_flags & 1); k += ES_THREADS / 32) {
      double acc[NV];
      float run[NV];
#pragma unroll
      for (int m = 0; m < NV; ++m) { acc[m] = 0.0; run[m] = 0.f; }
      int in_run = 0;
      for (int g0 = 0; g0 < nrows; g0 += 32) {
        unsigned hit = __ballot_sync(FULL, g0 + lane < nrows && keys[g0 + lane] == k);
        if (s.exp_flags & 8) hit = 0;
        while (hit) {
          // four members' rows in flight (the loop is a chain of shared-memory latencies); added in index order
          int jj[4];
          float v[4][NV];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            jj[u] = hit ? __ffs(hit) - 1 : -1;
            hit &= hit - 1;
            const uint32_t col = xt_s + 4u * (uint32_t)(lane * ldr + g0 + (jj[u] >= 0 ? jj[u] : 0));
#pragma unroll
            for (int m = 0; m < NV; ++m) {
              // unconditional loads (a conditional asm load becomes a branch): an absent member re-reads row g0 and is
              // not added; feature blocks beyond dp (inside the shared-memory window) are zeroed by the select
              const float t = kc_lds(col + 128u * (uint32_t)(min(32 * m, dp - 32) / 32 * ldr));
              v[u][m] = 32 * m < dp ? t : 0.f;                       // features dim..dp-1 hold zeros
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (jj[u] >= 0) {
#pragma unroll
              for (int m = 0; m < NV; ++m) run[m] += v[u][m];
              if (++in_run == 32) {
#pragma unroll
                for (int m = 0; m < NV; ++m) { acc[m] += (double)run[m]; run[m] = 0.f; }
                in_run = 0;
              }
            }
          }
        }
      }
#pragma unroll
      for (int m = 0; m < NV; ++m) part[(size_t)k * NV * 32 + lane + 32 * m] = acc[m] + (double)run[m];
    }
    kc_sync();
    // ---- centroids: bin k is finished by rank k % 8 (partials in rank order) and written into every rank's copy
    for (int k = (int)rank + KC_RANKS * warp; k < s.kmax && !(s.exp_flags & 2); k += KC_RANKS * (ES_THREADS / 32)) {
      float f[NV];
      float ss = 0.f;
#pragma unroll
      for (int m = 0; m < NV; ++m) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < KC_RANKS; ++w)
          a += kc_ld_f64(kc_mapa(part_s + (uint32_t)(((size_t)k * NV * 32 + lane + 32 * m) * sizeof(double)), w));
        f[m] = (float)a;
        if (lane + 32 * m < s.dim) ss = fmaf(f[m], f[m], ss);
      }
      const float n = safe_norm(warp_sum(ss));
#pragma unroll
      for (int m = 0; m < NV; ++m) {
        const int d = lane + 32 * m;
        if (d < s.dim) {
          const float c = f[m] / n;
          if (k < CK) {
#pragma unroll
            for (int w = 0; w < KC_RANKS; ++w) kc_st_f32(kc_mapa(cs_s + (uint32_t)(((size_t)k * ld + d) * sizeof(float)), w), c);
          }
          if (s.centroids_out && it == s.iterations - 1) s.centroids_out[((int64_t)seg * s.kmax + k) * s.dim + d] = c;
        }
      }
    }
    kc_sync();
    for (int i = tid; i < dp * CK; i += ES_THREADS) CT[i] = Cs[(size_t)(i % CK) * ld + i / CK];      // local transpose
    __syncthreads();
    // ---- assignment of this rank's rows, 128 at a time: thread (ty, tx) = rows 4 ty .. 4 ty + 3, columns 16 j + 2 tx, + 1
    // (the same fmaf chain over the features as estep_small_tile, so the same products)
    for (int r0 = 0; r0 < nrows && !(s.exp_flags & 4); r0 += KC_TILE) {
      const int np = min(KC_TILE, nrows - r0);
      float acc[4][2 * NJ];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2 * NJ; ++j) acc[i][j] = 0.f;
      const float* xp = XT + min(r0 + 4 * ty, L.rows_pad - 4);      // row groups beyond the rank's rows recompute the last one
      const float* cp = CT + 2 * tx;
#pragma unroll 4
      for (int dd = 0; dd < dp; ++dd) {
        const float* xq = xp + (size_t)dd * ldr;
        const float xs[4] = {xq[0], xq[1], xq[2], xq[3]};
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const float2 cb = *reinterpret_cast<const float2*>(cp + (size_t)dd * CK + 16 * j);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][2 * j] = fmaf(xs[i], cb.x, acc[i][2 * j]);
            acc[i][2 * j + 1] = fmaf(xs[i], cb.y, acc[i][2 * j + 1]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float bv = -FLT_MAX, sv = -FLT_MAX;
        int bi = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < 2 * NJ; ++j) {
          const int k = 16 * (j >> 1) + 2 * tx + (j & 1);
          if (k < K) merge_best(bv, bi, sv, acc[i][j], k, -FLT_MAX);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(FULL, bv, o);
          const int oi = __shfl_xor_sync(FULL, bi, o);
          const float osv = __shfl_xor_sync(FULL, sv, o);
          merge_best(bv, bi, sv, ov, oi, osv);
        }
        if (tx == 0) {
          const int px = 4 * ty + i;
          best_v[px] = bv; second_v[px] = sv; best_i[px] = bi;
        }
      }
      __syncthreads();
      for (int px = warp; px < np; px += ES_THREADS / 32) {
        int bi = best_i[px];
        if (best_v[px] - second_v[px] <= s.thr) {            // warp-uniform: the float64 scan of every cluster
          float xr[NV];
#pragma unroll
          for (int m = 0; m < NV; ++m) { const int d = lane + 32 * m; xr[m] = d < s.dim ? XT[(size_t)d * ldr + r0 + px] : 0.f; }
          double bv = -DBL_MAX;
          bi = 0x7fffffff;
          for (int k = 0; k < K; ++k) {
            double sum = 0.0;
#pragma unroll
            for (int m = 0; m < NV; ++m)
              if (lane + 32 * m < s.dim) sum = fma((double)xr[m], (double)Cs[(size_t)k * ld + lane + 32 * m], sum);
            sum = warp_sum(sum);
            if (sum > bv || (sum == bv && k < bi)) { bv = sum; bi = k; }
          }
        }
        if (lane == 0) keys[r0 + px] = bi;
      }
      __syncthreads();
    }
  }
  for (int r = tid; r < nrows; r += ES_THREADS) s.labels_out[wb + r] = (int64_t)keys[r];
  kc_sync();                    // no CTA leaves while another may still read its partials
}

template <int NV, int NJ>
static int launch_cluster_kmeans(const ClusterKmArgs& s, int64_t max_seg_len, cudaStream_t st, bool* done) {
  const int per = (int)kc_rows_per_rank(max_seg_len);
  const size_t smem = KcLayout<NV, NJ>(s.dim, s.kmax, per).total;
  if (smem > 220 * 1024) return HSG_OK;
  auto kern = kmeans_cluster_kernel<NV, NJ>;
  HSG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(s.S * KC_RANKS));
  cfg.blockDim = dim3(ES_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = KC_RANKS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg) != cudaSuccess || clusters < 1) {
    cudaGetLastError();
    return HSG_OK;              // no room for such a cluster on this device: the caller takes another path
  }
  ClusterKmArgs a = s;
  a.per_max = per;
  HSG_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  HSG_LAUNCH_CHECK();
  *done = true;
  return HSG_OK;
}

// one M-step of the k-means loop: full pass on the first iteration, afterwards delta or full
// as decided on the device
static int km_mstep(KmPlan& p, const float* x, const int64_t* seg_offsets, int it, bool incremental,
                    cudaStream_t st, const RunSums* runs = nullptr, bool small = false) {
  int rc;
  if (it == 0 && runs) {
    // usual case: the run table is complete and only the combine below does any work
    const Gate fallback{runs->overflow, 1};
    if ((rc = sr_sort_and_sum_gated(p.sr, x, seg_offsets, fallback, nullptr, nullptr, st))) return rc;
    if ((rc = sr_combine64(p.sr, p.sr.pieces, nullptr, nullptr, p.d.sums, p.d.members, p.centroids, st, fallback))) return rc;
    {
      ProfRange prof(PROF_MSTEP_COMBINE, st);
      const unsigned grid = (unsigned)ceil_div64(p.sr.bins, 8);
      if (p.sr.dim <= 32 * 9)
        runsum_combine_kernel<9><<<grid, 256, 0, st>>>(*runs, p.sr.S, p.sr.kmax, p.sr.dim, p.d.sums, p.d.members, p.centroids);
      else
        runsum_combine_kernel<20><<<grid, 256, 0, st>>>(*runs, p.sr.S, p.sr.kmax, p.sr.dim, p.d.sums, p.d.members, p.centroids);
      HSG_LAUNCH_CHECK();
    }
    if (incremental)
      HSG_CUDA(cudaMemcpyAsync(p.d.keys_prev, p.sr.keys, sizeof(int32_t) * p.sr.N, cudaMemcpyDeviceToDevice, st));
    return HSG_OK;
  }
  if (!incremental && small) {
    ProfRange prof(PROF_MSTEP_GATHER, st);
    const unsigned grid = (unsigned)p.sr.bins;
    if (p.sr.dim <= 32 * 5)
      mstep_small_kernel<5><<<grid, 256, 0, st>>>(x, p.sr.dim, seg_offsets, p.sr.S, p.sr.kmax, p.sr.keys, p.centroids, p.fix.count);
    else
      mstep_small_kernel<9><<<grid, 256, 0, st>>>(x, p.sr.dim, seg_offsets, p.sr.S, p.sr.kmax, p.sr.keys, p.centroids, p.fix.count);
    HSG_LAUNCH_CHECK();
    return HSG_OK;
  }
  if (it == 0 || !incremental) {
    if ((rc = sr_sort_and_sum(p.sr, x, seg_offsets, st))) return rc;
    if ((rc = sr_combine64(p.sr, p.sr.pieces, nullptr, nullptr, p.d.sums, p.d.members, p.centroids, st))) return rc;
    if (incremental)
      HSG_CUDA(cudaMemcpyAsync(p.d.keys_prev, p.sr.keys, sizeof(int32_t) * p.sr.N, cudaMemcpyDeviceToDevice, st));
    return HSG_OK;
  }
  DeltaPlan& d = p.d;
  if ((rc = sr_delta_build(p.sr, seg_offsets, d.keys_prev, d.tile_entries, d.eoff, d.flag, d.cap, d.erow,
                           d.ekey, st))) return rc;
  if ((rc = sr_build_tiles(d.sr, d.eoff, st))) return rc;
  if ((rc = sr_sort_and_sum_gated(p.sr, x, seg_offsets, Gate{d.flag, 0}, nullptr, nullptr, st))) return rc;
  if ((rc = sr_sort_and_sum_gated(d.sr, x, seg_offsets, Gate{d.flag, 1}, d.eoff, d.erow, st))) return rc;
  // the two plans share bin_start / bin_count; pieces differ (float vs float64)
  return sr_combine64(p.sr, p.sr.pieces, reinterpret_cast<const double*>(d.sr.pieces), d.flag, d.sums, d.members,
                      p.centroids, st);
}

static int check_common(const float* x, int64_t N, int dim, const int64_t* seg_offsets, int S,
                        int64_t max_seg_len, int kmax) {
  HSG_REQUIRE(N >= 0 && N < (1ll << 31), HSG_E_UNSUPPORTED, "kmeans: N=%lld rows (max 2^31-1)", (long long)N);
  HSG_REQUIRE(dim > 0 && S > 0 && kmax > 0, HSG_E_INVALID, "kmeans: bad shape dim=%d S=%d kmax=%d", dim, S, kmax);
  HSG_REQUIRE(dim <= 32 * 20, HSG_E_UNSUPPORTED, "kmeans: dim %d (max 640)", dim);
  HSG_REQUIRE(max_seg_len >= 0 && max_seg_len <= N, HSG_E_INVALID, "kmeans: max_seg_len %lld outside [0,N]", (long long)max_seg_len);
  HSG_REQUIRE(kmax <= SR_MAX_KEYS, HSG_E_UNSUPPORTED, "kmeans: kmax=%d (max %d)", kmax, SR_MAX_KEYS);
  HSG_REQUIRE((int64_t)S * kmax < (1ll << 31), HSG_E_UNSUPPORTED, "kmeans: S*kmax overflows int32");
  HSG_REQUIRE(N == 0 || (x && seg_offsets), HSG_E_INVALID, "kmeans: null pointer");
  return HSG_OK;
}

static int run_estep(EStepArgs& ea, KmPlan& p, bool use_tc, cudaStream_t st, bool counters_cleared = false) {
  if (!counters_cleared) HSG_CUDA(cudaMemsetAsync(p.fix.count, 0, 2 * sizeof(int32_t), st));
  if (use_tc) {
    int rc;
    {
      ProfRange prof(PROF_CONVERT, st);
      rc = tc_convert_centroids(ea, p.tc, st);
    }
    if (rc) return rc;
    ProfRange prof(PROF_ESTEP, st);
    rc = estep_tc(ea, p.tc, st);
    if (rc) return rc;
  } else {
    ProfRange prof(PROF_ESTEP, st);
    int rc = estep_simt(ea, st);
    if (rc) return rc;
  }
  ProfRange prof(PROF_ESTEP_FIXUP, st);
  return estep_fixup(ea, st);
}

static int decide_tc(int flags, int dim, const void* xh, int d16, const float* xerr, int kmax, bool* use_tc) {
  const bool possible = xh && xerr && tc_shape_supported(dim, d16, kmax);
  if ((flags & 3) == HSG_KMEANS_FORCE_TC) {
    HSG_REQUIRE(possible, HSG_E_UNSUPPORTED,
                "kmeans: tensor-core E-step needs the fp16 copy, d16 in {64,128,256,512}, at most 5 trailing features "
                "(got dim=%d d16=%d kmax=%d, xh=%p)", dim, d16, kmax, xh);
    *use_tc = true;
  } else if ((flags & 3) == HSG_KMEANS_FORCE_SIMT) {
    *use_tc = false;
  } else {
    *use_tc = possible;
  }
  return HSG_OK;
}

}  // namespace hsg

using namespace hsg;

extern "C" {

size_t hsg_kmeans_workspace_bytes(int64_t N, int dim, int S, int kmax, int64_t max_seg_len) {
  size_t need = 0;
  for (int d16 = 64; d16 <= 512; d16 *= 2) {     // whichever fp16 side copy the caller brings
    Carver c(nullptr);
    KmPlan p;
    km_carve(c, p, N, dim, S, kmax, max_seg_len, d16);
    if (c.used() > need) need = c.used();
  }
  return need + 1024;
}

static int kmeans_impl(const float* x, int64_t N, int dim, const void* xh, int d16, const float* xerr,
                       const int64_t* seg_offsets, int S, int64_t max_seg_len, const int32_t* seg_k,
                       int kmax, const int64_t* init_labels, int iterations, int64_t* labels_out,
                       float* centroids_out, int flags, void* workspace, size_t workspace_bytes,
                       void* stream, const RunSums* runs) {
  int rc = check_common(x, N, dim, seg_offsets, S, max_seg_len, kmax);
  if (rc) return rc;
  HSG_REQUIRE(iterations >= 0, HSG_E_INVALID, "kmeans: iterations=%d", iterations);
  if (N == 0) return HSG_OK;
  HSG_REQUIRE(init_labels && labels_out, HSG_E_INVALID, "kmeans: null labels");
  HSG_REQUIRE(workspace && workspace_bytes >= hsg_kmeans_workspace_bytes(N, dim, S, kmax, max_seg_len),
              HSG_E_WORKSPACE, "kmeans: workspace too small (%zu < %zu)", workspace_bytes,
              hsg_kmeans_workspace_bytes(N, dim, S, kmax, max_seg_len));
  bool use_tc = false;
  rc = decide_tc(flags, dim, xh, d16, xerr, kmax, &use_tc);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  ProfRange prof_call(PROF_KMEANS, st);
  Carver c(workspace);
  KmPlan p;
  km_carve(c, p, N, dim, S, kmax, max_seg_len, d16);
  if (use_tc) {
    p.tc.xh = (const __half*)xh; p.tc.xerr = xerr; p.tc.d16 = d16;
    rc = tc_prepare(p.tc, N, S);
    if (rc) return rc;
  }
  if ((rc = sr_build_tiles(p.sr, seg_offsets, st))) return rc;
  if ((rc = sr_labels_to_keys(p.sr, init_labels, nullptr, st))) return rc;

  EStepArgs ea;
  ea.x = x; ea.N = N; ea.dim = dim; ea.centroids = p.centroids; ea.seg_offsets = seg_offsets;
  ea.S = S; ea.seg_k = seg_k; ea.kmax = kmax; ea.tiles = p.sr.tiles; ea.keys_out = p.sr.keys;
  ea.fix = p.fix;

  // below ~2.6e5 rows a full re-sum is cheaper than the dozen extra launches of the delta machinery
  const bool incremental = !(flags & HSG_KMEANS_FULL_MSTEP) && N >= KM_DELTA_MIN_ROWS;
  // launch-bound regime: one-launch M-step (it also clears the re-decision counters)
  const bool small = !incremental && !(flags & HSG_KMEANS_FULL_MSTEP) && max_seg_len <= KM_SMALL_MAX_SEG && dim <= 32 * 9 &&
                     (int64_t)S * kmax <= (1 << 20);
  static const bool no_persistent = getenv("HSG_KMEANS_NO_PERSISTENT") != nullptr;       // A/B switch for profiling only
  static const bool no_cluster = getenv("HSG_KMEANS_NO_CLUSTER") != nullptr;              // A/B switch for profiling only
  if (small && !runs && iterations > 0 && !no_cluster && (!use_tc || kmax <= 64) && kmax <= 64 && S <= (1 << 20)) {
    // an image per thread-block cluster, resident in shared memory for all iterations
    ClusterKmArgs ca;
    ca.x = x; ca.dim = dim; ca.kmax = kmax; ca.S = S; ca.seg_offsets = seg_offsets; ca.seg_k = seg_k;
    ca.init_labels = init_labels; ca.labels_out = labels_out; ca.centroids_out = centroids_out;
    ca.iterations = iterations; ca.per_max = 0;
    static const int cluster_exp = getenv("HSG_CLUSTER_EXP") ? atoi(getenv("HSG_CLUSTER_EXP")) : 0;
    ca.exp_flags = cluster_exp;
    ca.thr = 2.f * (dim + 2) * 5.9604645e-8f * 1.02f + 1e-7f;             // estep_simt's bound
    bool done = false;
    if (dim <= 32 * 5) rc = kmax <= 16 ? launch_cluster_kmeans<5, 1>(ca, max_seg_len, st, &done) : launch_cluster_kmeans<5, 4>(ca, max_seg_len, st, &done);
    else rc = kmax <= 16 ? launch_cluster_kmeans<9, 1>(ca, max_seg_len, st, &done) : launch_cluster_kmeans<9, 4>(ca, max_seg_len, st, &done);
    if (rc) return rc;
    if (done) return HSG_OK;
  }
  if (small && !runs && iterations > 0 && !no_persistent && (!use_tc || kmax <= 64)) {
    // launch-bound regime: the whole loop in one cooperative launch (fp32 CUDA-core E-step: at K <= 64 the products
    // of such a call are a few GFLOP, the launches were the cost)
    SmallArgs sa;
    sa.a = ea; sa.centroids = p.centroids; sa.init_labels = init_labels; sa.labels_out = labels_out;
    sa.centroids_out = centroids_out; sa.iterations = iterations;
    sa.thr = 2.f * (dim + 2) * 5.9604645e-8f * 1.02f + 1e-7f;             // estep_simt's bound
    sa.bar = reinterpret_cast<unsigned*>(p.fix.count);
    static const int small_exp = getenv("HSG_SMALL_EXP") ? atoi(getenv("HSG_SMALL_EXP")) : 0;
    sa.exp_flags = small_exp;
    sa.max_seg_len = max_seg_len;
    const int64_t work = max((int64_t)S * kmax, ceil_div64(N, ES_TP) + S);
    bool done = false;
    if (dim <= 32 * 5) rc = kmax <= 16 ? launch_small_persistent<5, 1>(sa, work, st, &done) : launch_small_persistent<5, 4>(sa, work, st, &done);
    else rc = kmax <= 16 ? launch_small_persistent<9, 1>(sa, work, st, &done) : launch_small_persistent<9, 4>(sa, work, st, &done);
    if (rc) return rc;
    if (done) return HSG_OK;
  }
  for (int it = 0; it < iterations; ++it) {
    if ((rc = km_mstep(p, x, seg_offsets, it, incremental, st, runs, small))) return rc;
    if ((rc = run_estep(ea, p, use_tc, st, small && !(it == 0 && runs)))) return rc;   // the small M-step cleared the counters
  }
  if ((rc = sr_keys_to_labels(p.sr, p.sr.keys, labels_out, st))) return rc;
  if (centroids_out && iterations > 0)
    HSG_CUDA(cudaMemcpyAsync(centroids_out, p.centroids, sizeof(float) * S * kmax * dim,
                             cudaMemcpyDeviceToDevice, st));
  return HSG_OK;
}

int hsg_kmeans_f32(const float* x, int64_t N, int dim, const void* xh, int d16, const float* xerr,
                   const int64_t* seg_offsets, int S, int64_t max_seg_len, const int32_t* seg_k,
                   int kmax, const int64_t* init_labels, int iterations, int64_t* labels_out,
                   float* centroids_out, int flags, void* workspace, size_t workspace_bytes,
                   void* stream) {
  return kmeans_impl(x, N, dim, xh, d16, xerr, seg_offsets, S, max_seg_len, seg_k, kmax, init_labels, iterations,
                     labels_out, centroids_out, flags, workspace, workspace_bytes, stream, nullptr);
}

int hsg_kmeans_presummed_f32(const float* x, int64_t N, int dim, const void* xh, int d16, const float* xerr,
                             const int64_t* seg_offsets, int S, int64_t max_seg_len, const int32_t* seg_k,
                             int kmax, const int64_t* init_labels, int iterations, int64_t* labels_out,
                             float* centroids_out, int flags, void* workspace, size_t workspace_bytes,
                             const float* run_sums, const int32_t* run_cluster, const int32_t* run_count,
                             const int32_t* run_overflow, int64_t runs_per_segment, void* stream) {
  HSG_REQUIRE(run_sums && run_cluster && run_count && run_overflow && runs_per_segment > 0, HSG_E_INVALID,
              "kmeans_presummed: null run buffers");
  RunSums r{run_sums, run_cluster, run_count, run_overflow, runs_per_segment};
  return kmeans_impl(x, N, dim, xh, d16, xerr, seg_offsets, S, max_seg_len, seg_k, kmax, init_labels, iterations,
                     labels_out, centroids_out, flags, workspace, workspace_bytes, stream, &r);
}

int hsg_kmeans_mstep_f32(const float* x, int64_t N, int dim, const int64_t* seg_offsets, int S,
                         int64_t max_seg_len, const int32_t* seg_k, int kmax, const int64_t* labels,
                         float* centroids_out, void* workspace, size_t workspace_bytes, void* stream) {
  (void)seg_k;
  int rc = check_common(x, N, dim, seg_offsets, S, max_seg_len, kmax);
  if (rc) return rc;
  HSG_REQUIRE(centroids_out, HSG_E_INVALID, "mstep: null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    HSG_CUDA(cudaMemsetAsync(centroids_out, 0, sizeof(float) * S * kmax * dim, st));
    return HSG_OK;
  }
  HSG_REQUIRE(labels, HSG_E_INVALID, "mstep: null labels");
  HSG_REQUIRE(workspace && workspace_bytes >= sr_workspace_bytes(N, dim, S, kmax, max_seg_len),
              HSG_E_WORKSPACE, "mstep: workspace too small");
  Carver c(workspace);
  SegReducePlan p;
  sr_carve(c, p, N, dim, S, kmax, max_seg_len);
  if ((rc = sr_build_tiles(p, seg_offsets, st))) return rc;
  if ((rc = sr_labels_to_keys(p, labels, nullptr, st))) return rc;
  if ((rc = sr_sort_and_sum(p, x, seg_offsets, st))) return rc;
  return sr_combine(p, p.bins, nullptr, HSG_REDUCE_NORMALIZE, centroids_out, nullptr, nullptr, st);
}

int hsg_kmeans_estep_f32(const float* x, int64_t N, int dim, const void* xh, int d16,
                         const float* xerr, const int64_t* seg_offsets, int S, int64_t max_seg_len,
                         const int32_t* seg_k, int kmax, const float* centroids, int64_t* labels_out,
                         int64_t* num_rechecked_out, int flags, void* workspace,
                         size_t workspace_bytes, void* stream) {
  int rc = check_common(x, N, dim, seg_offsets, S, max_seg_len, kmax);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    if (num_rechecked_out) HSG_CUDA(cudaMemsetAsync(num_rechecked_out, 0, 2 * sizeof(int64_t), st));
    return HSG_OK;
  }
  HSG_REQUIRE(centroids && labels_out, HSG_E_INVALID, "estep: null pointer");
  HSG_REQUIRE(workspace && workspace_bytes >= hsg_kmeans_workspace_bytes(N, dim, S, kmax, max_seg_len),
              HSG_E_WORKSPACE, "estep: workspace too small");
  bool use_tc = false;
  rc = decide_tc(flags, dim, xh, d16, xerr, kmax, &use_tc);
  if (rc) return rc;
  Carver c(workspace);
  KmPlan p;
  km_carve(c, p, N, dim, S, kmax, max_seg_len, d16);
  if (use_tc) {
    p.tc.xh = (const __half*)xh; p.tc.xerr = xerr; p.tc.d16 = d16;
    rc = tc_prepare(p.tc, N, S);
    if (rc) return rc;
  }
  if ((rc = sr_build_tiles(p.sr, seg_offsets, st))) return rc;
  EStepArgs ea;
  ea.x = x; ea.N = N; ea.dim = dim; ea.centroids = centroids; ea.seg_offsets = seg_offsets;
  ea.S = S; ea.seg_k = seg_k; ea.kmax = kmax; ea.tiles = p.sr.tiles; ea.keys_out = p.sr.keys;
  ea.fix = p.fix;
  if ((rc = run_estep(ea, p, use_tc, st))) return rc;
  if (num_rechecked_out) {
    copy_count_kernel<<<1, 1, 0, st>>>(p.fix.count, num_rechecked_out);
    HSG_LAUNCH_CHECK();
  }
  return sr_keys_to_labels(p.sr, p.sr.keys, labels_out, st);
}

// ---------------------------------------------------------------- row-sharded flat k-means, one iteration at a time
// (kmeans_with_initial_labels over rows sharded across GPUs, SURVEY 8e / BASELINE configs[4]): the caller all-reduces
// the int64 sums between `local` and `assign`; everything else (keys, previous keys, tiles) lives in the workspace,
// which the caller keeps untouched between the calls of one loop.
size_t hsg_kmeans_dist_workspace_bytes(int64_t N, int dim, int kmax) {
  size_t need = 0;
  for (int d16 = 64; d16 <= 512; d16 *= 2) {
    Carver c(nullptr);
    KmPlan p;
    km_carve(c, p, N, dim, 1, kmax, N, d16, true);
    if (c.used() > need) need = c.used();
  }
  return need + 1024;
}

}  // extern "C"

namespace hsg {
// centroids = normalise(sums * 2^-36); an empty cluster has an exactly zero sum and gets the zero centroid
__global__ void __launch_bounds__(256) dist_centroids_kernel(const long long* __restrict__ sums, int kmax, int dim,
                                                             float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (k >= kmax) return;
  float ss = 0.f;
  for (int d = lane; d < dim; d += 32) {
    const float f = (float)((double)sums[(int64_t)k * dim + d] * (1.0 / 68719476736.0));
    out[(int64_t)k * dim + d] = f;
    ss = fmaf(f, f, ss);
  }
  const float n = safe_norm(warp_sum(ss));
  for (int d = lane; d < dim; d += 32) out[(int64_t)k * dim + d] /= n;
}

// This shard's contribution to the running sums of the loop, whichever pass ran (flag[0] == 1: the delta pass over
// the rows that moved, its pieces ARE the contribution; otherwise the full pass: contribution = new sums - the
// shard's previous sums).  Integer arithmetic: both give the same numbers.  `local` keeps the shard's own sums.
__global__ void __launch_bounds__(256) dist_contribution_kernel(
    int64_t bins, int dim, const int64_t* __restrict__ bin_start, const int32_t* __restrict__ bin_count,
    const long long* __restrict__ pieces_full, const long long* __restrict__ pieces_delta,
    const int32_t* __restrict__ flag, long long* __restrict__ local, long long* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t key = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (key >= bins) return;
  const bool delta = flag[0] == 1;
  const long long* pieces = delta ? pieces_delta : pieces_full;
  const int cnt = bin_count[key];
  const int64_t start = bin_start[key];
  const int64_t r0 = start / SR_RUN, r1 = cnt > 0 ? (start + cnt - 1) / SR_RUN : r0 - 1;
  for (int d = lane; d < dim; d += 32) {
    long long a = 0;
    for (int64_t r = r0; r <= r1; ++r) a += pieces[(r + key) * dim + d];
    const long long before = local[key * dim + d];
    out[key * dim + d] = delta ? a : a - before;
    local[key * dim + d] = delta ? before + a : a;
  }
}
}  // namespace hsg

extern "C" {

// phase 0: this shard's contribution to the running sums.  first != 0: keys from init_labels, full exact pass
// (sums_out = the shard's sums).  Otherwise: exact sums over the rows whose label changed in the last `assign`
// (sum[new] += x, sum[old] -= x), to be ADDED to the running sums after the all-reduce.
int hsg_kmeans_dist_local_i64(const float* x, int64_t N, int dim, int d16, const int64_t* seg_offsets /* {0,N} on the device */,
                              int kmax, int first, const int64_t* init_labels, long long* sums_out,
                              void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(x, N, dim, seg_offsets, 1, N, kmax);
  if (rc) return rc;
  HSG_REQUIRE(sums_out, HSG_E_INVALID, "kmeans_dist_local: null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    HSG_CUDA(cudaMemsetAsync(sums_out, 0, sizeof(long long) * kmax * dim, st));
    return HSG_OK;
  }
  HSG_REQUIRE(workspace && workspace_bytes >= hsg_kmeans_dist_workspace_bytes(N, dim, kmax), HSG_E_WORKSPACE,
              "kmeans_dist_local: workspace too small");
  Carver c(workspace);
  KmPlan p;
  km_carve(c, p, N, dim, 1, kmax, N, d16, true);       // the three calls of a loop carve the same layout (same d16)
  if (first) {
    HSG_REQUIRE(init_labels, HSG_E_INVALID, "kmeans_dist_local: null labels");
    if ((rc = sr_build_tiles(p.sr, seg_offsets, st))) return rc;
    if ((rc = sr_labels_to_keys(p.sr, init_labels, nullptr, st))) return rc;
    if ((rc = sr_sort_and_sum(p.sr, x, seg_offsets, st))) return rc;
    if ((rc = sr_combine_exact(p.sr, kmax, nullptr, sums_out, st))) return rc;
    HSG_CUDA(cudaMemcpyAsync(p.d.local_sums, sums_out, sizeof(long long) * kmax * dim, cudaMemcpyDeviceToDevice, st));
    HSG_CUDA(cudaMemcpyAsync(p.d.keys_prev, p.sr.keys, sizeof(int32_t) * N, cudaMemcpyDeviceToDevice, st));
    return HSG_OK;
  }
  // delta or full pass, decided on the device as in the per-image loop (both enqueued, one returns at once)
  DeltaPlan& d = p.d;
  if ((rc = sr_delta_build(p.sr, seg_offsets, d.keys_prev, d.tile_entries, d.eoff, d.flag, d.cap, d.erow, d.ekey, st))) return rc;
  if ((rc = sr_build_tiles(d.sr, d.eoff, st))) return rc;
  if ((rc = sr_sort_and_sum_gated(p.sr, x, seg_offsets, Gate{d.flag, 0}, nullptr, nullptr, st))) return rc;
  if ((rc = sr_sort_and_sum_gated(d.sr, x, seg_offsets, Gate{d.flag, 1}, d.eoff, d.erow, st))) return rc;
  dist_contribution_kernel<<<(unsigned)ceil_div64(p.sr.bins, 8), 256, 0, st>>>(
      p.sr.bins, dim, p.sr.bin_start, p.sr.bin_count, reinterpret_cast<const long long*>(p.sr.pieces),
      reinterpret_cast<const long long*>(d.sr.pieces), d.flag, d.local_sums, sums_out);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

// phase 1: centroids from the all-reduced running sums, then the E-step; the new keys stay in the workspace
int hsg_kmeans_dist_assign_f32(const float* x, int64_t N, int dim, const void* xh, int d16, const float* xerr,
                               const int64_t* seg_offsets, int kmax, const long long* sums, int flags,
                               void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(x, N, dim, seg_offsets, 1, N, kmax);
  if (rc) return rc;
  if (N == 0) return HSG_OK;
  HSG_REQUIRE(sums, HSG_E_INVALID, "kmeans_dist_assign: null sums");
  HSG_REQUIRE(workspace && workspace_bytes >= hsg_kmeans_dist_workspace_bytes(N, dim, kmax), HSG_E_WORKSPACE,
              "kmeans_dist_assign: workspace too small");
  bool use_tc = false;
  if ((rc = decide_tc(flags, dim, xh, d16, xerr, kmax, &use_tc))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(workspace);
  KmPlan p;
  km_carve(c, p, N, dim, 1, kmax, N, d16, true);
  if (use_tc) {
    p.tc.xh = (const __half*)xh; p.tc.xerr = xerr; p.tc.d16 = d16;
    if ((rc = tc_prepare(p.tc, N, 1))) return rc;
  }
  dist_centroids_kernel<<<(unsigned)ceil_div64(kmax, 8), 256, 0, st>>>(sums, kmax, dim, p.centroids);
  HSG_LAUNCH_CHECK();
  EStepArgs ea;
  ea.x = x; ea.N = N; ea.dim = dim; ea.centroids = p.centroids; ea.seg_offsets = seg_offsets;
  ea.S = 1; ea.seg_k = nullptr; ea.kmax = kmax; ea.tiles = p.sr.tiles; ea.keys_out = p.sr.keys;
  ea.fix = p.fix;
  return run_estep(ea, p, use_tc, st);
}

// phase 2: the labels of the last `assign`
int hsg_kmeans_dist_labels_i64(int64_t N, int dim, int d16, int kmax, int64_t* labels_out, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (N == 0) return HSG_OK;
  HSG_REQUIRE(labels_out && workspace && workspace_bytes >= hsg_kmeans_dist_workspace_bytes(N, dim, kmax), HSG_E_WORKSPACE,
              "kmeans_dist_labels: workspace too small");
  Carver c(workspace);
  KmPlan p;
  km_carve(c, p, N, dim, 1, kmax, N, d16, true);
  return sr_keys_to_labels(p.sr, p.sr.keys, labels_out, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- K3 segmented reduction
size_t hsg_segment_reduce_workspace_bytes(int64_t N, int dim, int64_t P, int S, int kmax,
                                          int64_t max_seg_len) {
  (void)P;
  return sr_workspace_bytes(N, dim, S > 0 ? S : 1, kmax, max_seg_len) + 1024;
}

int hsg_segment_reduce_f32(const float* x, int64_t N, int dim, const int64_t* labels, int64_t P,
                           const int64_t* seg_offsets, int S, int64_t max_seg_len,
                           const int64_t* seg_base, int kmax, int mode, float* out, float* sums_out,
                           float* counts_out, void* workspace, size_t workspace_bytes, void* stream) {
  HSG_REQUIRE(P >= 0 && dim > 0 && N >= 0, HSG_E_INVALID, "segment_reduce: bad shape");
  HSG_REQUIRE(mode >= 0 && mode <= 2, HSG_E_INVALID, "segment_reduce: bad mode %d", mode);
  HSG_REQUIRE(seg_offsets && seg_base && S > 0, HSG_E_INVALID,
              "segment_reduce: seg_offsets/seg_base are required (one segment: offsets {0,N}, base {0}, kmax=P)");
  if (P == 0) return HSG_OK;
  HSG_REQUIRE(out, HSG_E_INVALID, "segment_reduce: null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    HSG_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * P * dim, st));
    if (sums_out) HSG_CUDA(cudaMemsetAsync(sums_out, 0, sizeof(float) * P * dim, st));
    if (counts_out) HSG_CUDA(cudaMemsetAsync(counts_out, 0, sizeof(float) * P, st));
    return HSG_OK;
  }
  int rc = check_common(x, N, dim, seg_offsets, S, max_seg_len, kmax);
  if (rc) return rc;
  HSG_REQUIRE(labels, HSG_E_INVALID, "segment_reduce: null labels");
  HSG_REQUIRE(workspace && workspace_bytes >= hsg_segment_reduce_workspace_bytes(N, dim, P, S, kmax, max_seg_len),
              HSG_E_WORKSPACE, "segment_reduce: workspace too small");
  Carver c(workspace);
  SegReducePlan p;
  sr_carve(c, p, N, dim, S, kmax, max_seg_len);
  ProfRange prof(PROF_POOL, st);
  ProfSuppress inner;
  if ((rc = sr_build_tiles(p, seg_offsets, st))) return rc;
  if ((rc = sr_labels_to_keys(p, labels, seg_base, st))) return rc;
  if ((rc = sr_sort_and_sum(p, x, seg_offsets, st))) return rc;
  return sr_combine(p, P, seg_base, mode, out, sums_out, counts_out, st);
}

size_t hsg_segment_sum_exact_workspace_bytes(int64_t N, int dim, int64_t P, int S, int kmax, int64_t max_seg_len) {
  (void)P;
  return sr_workspace_bytes(N, dim, S > 0 ? S : 1, kmax, max_seg_len, true) + 1024;
}

int hsg_segment_sum_exact_i64(const float* x, int64_t N, int dim, const int64_t* labels, int64_t P,
                              const int64_t* seg_offsets, int S, int64_t max_seg_len, const int64_t* seg_base,
                              int kmax, long long* sums_out, void* workspace, size_t workspace_bytes, void* stream) {
  HSG_REQUIRE(P >= 0 && dim > 0 && N >= 0, HSG_E_INVALID, "segment_sum_exact: bad shape");
  HSG_REQUIRE(seg_offsets && seg_base && S > 0, HSG_E_INVALID, "segment_sum_exact: seg_offsets/seg_base are required");
  if (P == 0) return HSG_OK;
  HSG_REQUIRE(sums_out, HSG_E_INVALID, "segment_sum_exact: null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    HSG_CUDA(cudaMemsetAsync(sums_out, 0, sizeof(long long) * P * dim, st));
    return HSG_OK;
  }
  int rc = check_common(x, N, dim, seg_offsets, S, max_seg_len, kmax);
  if (rc) return rc;
  HSG_REQUIRE(labels, HSG_E_INVALID, "segment_sum_exact: null labels");
  HSG_REQUIRE(workspace && workspace_bytes >= hsg_segment_sum_exact_workspace_bytes(N, dim, P, S, kmax, max_seg_len),
              HSG_E_WORKSPACE, "segment_sum_exact: workspace too small");
  Carver c(workspace);
  SegReducePlan p;
  sr_carve(c, p, N, dim, S, kmax, max_seg_len, true);
  ProfRange prof(PROF_POOL, st);
  ProfSuppress inner;
  if ((rc = sr_build_tiles(p, seg_offsets, st))) return rc;
  if ((rc = sr_labels_to_keys(p, labels, seg_base, st))) return rc;
  if ((rc = sr_sort_and_sum(p, x, seg_offsets, st))) return rc;
  return sr_combine_exact(p, P, seg_base, sums_out, st);
}

}  // extern "C"

// backward: per-bin gradient of the finishing step, then a row gather
namespace hsg {

__global__ void segreduce_bin_grad_kernel(const float* __restrict__ g, const float* __restrict__ out,
                                          const float* __restrict__ sums, const float* __restrict__ counts,
                                          int64_t P, int dim, int mode, float* __restrict__ gs) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= P) return;
  const float* gr = g + p * dim;
  float* o = gs + p * dim;
  if (mode == HSG_REDUCE_SUM) {
    for (int d = lane; d < dim; d += 32) o[d] = gr[d];
  } else if (mode == HSG_REDUCE_MEAN) {
    const float c = counts[p] > 0.f ? counts[p] : 1.f;
    for (int d = lane; d < dim; d += 32) o[d] = gr[d] / c;
  } else {
    const float* sr = sums + p * dim;
    const float* pr = out + p * dim;
    float ss = 0.f, dot = 0.f;
    for (int d = lane; d < dim; d += 32) {
      ss = fmaf(sr[d], sr[d], ss);
      dot = fmaf(pr[d], gr[d], dot);
    }
    ss = warp_sum(ss);
    dot = warp_sum(dot);
    const float n = sqrtf(ss);
    if (n >= 1e-12f) {
      for (int d = lane; d < dim; d += 32) o[d] = (gr[d] - pr[d] * dot) / n;
    } else {
      for (int d = lane; d < dim; d += 32) o[d] = gr[d] / 1e-12f;
    }
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx,
                                   int64_t N, int dim, int64_t P, float* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N) return;
  int64_t k = idx[i];
  k = k < 0 ? 0 : (k >= P ? P - 1 : k);
  const float* s = src + k * dim;
  float* d = dst + i * dim;
  for (int j = lane; j < dim; j += 32) d[j] = s[j];
}

}  // namespace hsg

extern "C" int hsg_segment_reduce_bwd_f32(const float* grad_out, const float* out, const float* sums,
                                          const float* counts, const int64_t* labels, int64_t N,
                                          int dim, int64_t P, int mode, float* grad_x,
                                          void* workspace, size_t workspace_bytes, void* stream) {
  HSG_REQUIRE(P >= 0 && dim > 0 && N >= 0 && mode >= 0 && mode <= 2, HSG_E_INVALID, "segment_reduce_bwd: bad argument");
  if (N == 0 || P == 0) return HSG_OK;
  HSG_REQUIRE(grad_out && labels && grad_x, HSG_E_INVALID, "segment_reduce_bwd: null pointer");
  HSG_REQUIRE(mode != HSG_REDUCE_NORMALIZE || (out && sums), HSG_E_INVALID, "segment_reduce_bwd: NORMALIZE needs out and sums");
  HSG_REQUIRE(mode != HSG_REDUCE_MEAN || counts, HSG_E_INVALID, "segment_reduce_bwd: MEAN needs counts");
  HSG_REQUIRE(workspace && workspace_bytes >= sizeof(float) * P * dim, HSG_E_WORKSPACE,
              "segment_reduce_bwd: workspace needs P*dim floats");
  cudaStream_t st = (cudaStream_t)stream;
  float* gs = (float*)workspace;
  segreduce_bin_grad_kernel<<<(unsigned)ceil_div64(P, 8), 256, 0, st>>>(grad_out, out, sums, counts, P, dim, mode, gs);
  HSG_LAUNCH_CHECK();
  gather_rows_kernel<<<(unsigned)ceil_div64(N, 8), 256, 0, st>>>(gs, labels, N, dim, P, grad_x);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}
