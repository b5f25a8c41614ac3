// Deterministic segmented row sums (see segreduce.cuh for the scheme).
#include "segreduce.cuh"

#include <type_traits>

namespace hsg {

// ---------------------------------------------------------------- helpers
__device__ __forceinline__ int upper_bound_i64(const int64_t* a, int n, int64_t v) {
  // first index with a[idx] > v
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// exclusive scan over one value per thread of a 1024-thread CTA (int64)
__device__ __forceinline__ int64_t block_scan_excl_1024(int64_t v, int64_t* warp_tot, int64_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int64_t w = lane < (blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(FULL, w, o);
      if (lane >= o) w += t;
    }
    warp_tot[lane] = w;
  }
  __syncthreads();
  const int64_t res = (warp ? warp_tot[warp - 1] : 0) + incl - v;
  *total = warp_tot[(blockDim.x >> 5) - 1];
  __syncthreads();
  return res;
}

// ---------------------------------------------------------------- tiles
__global__ void __launch_bounds__(1024) build_tiles_kernel(const int64_t* __restrict__ off, int S,
                                                           Tiles t) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < S; base += 1024) {
    const int s = base + threadIdx.x;
    int64_t nt = 0;
    if (s < S) nt = (off[s + 1] - off[s] + t.tile - 1) / t.tile;
    int64_t total;
    const int64_t first = carry_s + block_scan_excl_1024(nt, warp_tot, &total);
    if (s < S) {
      t.seg_first[s] = (int32_t)min(first, t.bound);
      for (int64_t j = 0; j < nt; ++j) {
        const int64_t idx = first + j;
        if (idx < t.bound) {
          t.seg[idx] = s;
          const int64_t b = off[s] + j * t.tile;
          t.begin[idx] = b;
          t.end[idx] = min(off[s + 1], b + t.tile);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_s += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    t.seg_first[S] = (int32_t)min(carry_s, t.bound);
    *t.count = (int32_t)min(carry_s, t.bound);
  }
}

// rows [r0,r1) of segment `seg` handled by this CTA (grid = (tiles.bound, tile/chunk_rows))
__device__ __forceinline__ bool cta_rows(const Tiles& t, int chunk_rows, int64_t& r0, int64_t& r1, int& seg) {
  const int ti = blockIdx.x;
  if (ti >= *t.count) return false;
  r0 = t.begin[ti] + (int64_t)blockIdx.y * chunk_rows;
  const int64_t e = t.end[ti];
  if (r0 >= e) return false;
  r1 = min(e, r0 + chunk_rows);
  seg = t.seg[ti];
  return true;
}

__global__ void labels_to_keys_kernel(Tiles t, const int64_t* __restrict__ labels,
                                      const int64_t* __restrict__ seg_base, int kmax,
                                      int32_t* __restrict__ keys) {
  int64_t r0, r1; int seg;
  if (!cta_rows(t, SR_TILE_MIN, r0, r1, seg)) return;
  const int64_t base = seg_base ? seg_base[seg] : 0;
  for (int64_t i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
    int64_t k = labels[i] - base;
    k = k < 0 ? 0 : (k >= kmax ? kmax - 1 : k);      // out-of-range ids are clamped, never scattered out of bounds
    keys[i] = seg * kmax + (int)k;
  }
}

__global__ void keys_to_labels_kernel(Tiles t, const int32_t* __restrict__ keys, int kmax,
                                      int64_t* __restrict__ labels) {
  int64_t r0, r1; int seg;
  if (!cta_rows(t, SR_TILE_MIN, r0, r1, seg)) return;
  for (int64_t i = r0 + threadIdx.x; i < r1; i += blockDim.x)
    labels[i] = (int64_t)(keys[i] - seg * kmax);
}

// ---------------------------------------------------------------- hist / scan / scatter
__global__ void __launch_bounds__(512) hist_kernel(Tiles t, const int32_t* __restrict__ keys, int kmax,
                                                   uint32_t* __restrict__ tile_hist, Gate gate) {
  extern __shared__ uint32_t hist[];
  const int ti = blockIdx.x;
  if (gate.closed() || ti >= *t.count) return;
  for (int k = threadIdx.x; k < kmax; k += blockDim.x) hist[k] = 0;
  __syncthreads();
  const int seg = t.seg[ti];
  const int64_t b = t.begin[ti], e = t.end[ti];
  const int kbase = seg * kmax;
  for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) atomicAdd(&hist[keys[i] - kbase], 1u);
  __syncthreads();
  for (int k = threadIdx.x; k < kmax; k += blockDim.x) tile_hist[(int64_t)ti * kmax + k] = hist[k];
}

__global__ void __launch_bounds__(1024) scan_kernel(Tiles t, const int64_t* __restrict__ off, int kmax,
                                                    uint32_t* __restrict__ tile_hist,
                                                    int64_t* __restrict__ bin_start,
                                                    int32_t* __restrict__ bin_count, Gate gate) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry_s;
  if (gate.closed()) return;
  const int s = blockIdx.x;
  const int t0 = t.seg_first[s], t1 = t.seg_first[s + 1];
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < kmax; base += 1024) {
    const int k = base + threadIdx.x;
    int64_t run = 0;
    if (k < kmax) {
      for (int ti = t0; ti < t1; ++ti) {
        const int64_t idx = (int64_t)ti * kmax + k;
        const uint32_t v = tile_hist[idx];
        tile_hist[idx] = (uint32_t)run;
        run += v;
      }
    }
    int64_t total;
    const int64_t before = carry_s + block_scan_excl_1024(run, warp_tot, &total);
    if (k < kmax) {
      bin_start[(int64_t)s * kmax + k] = off[s] + before;
      bin_count[(int64_t)s * kmax + k] = (int32_t)run;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_s += total;
    __syncthreads();
  }
}

template <int SCATTER_WARPS>
__global__ void __launch_bounds__(SCATTER_WARPS * 32) scatter_kernel(
    Tiles t, const int64_t* __restrict__ off, const int32_t* __restrict__ keys, int kmax,
    const uint32_t* __restrict__ tile_hist, const int64_t* __restrict__ bin_start,
    uint32_t* __restrict__ perm, Gate gate) {
  extern __shared__ uint32_t cursors[];          // [SCATTER_WARPS][kmax]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ti = blockIdx.x * SCATTER_WARPS + warp;
  if (gate.closed() || ti >= *t.count) return;   // only __syncwarp below
  uint32_t* cur = cursors + (size_t)warp * kmax;
  const int seg = t.seg[ti];
  const int64_t so = off[seg];
  const int kbase = seg * kmax;
  for (int k = lane; k < kmax; k += 32)
    cur[k] = (uint32_t)(bin_start[(int64_t)kbase + k] - so) + tile_hist[(int64_t)ti * kmax + k];
  __syncwarp();
  const int64_t b = t.begin[ti], e = t.end[ti];
  const unsigned lt = (1u << lane) - 1u;
  // the placement is serial per tile (one warp walks it in order), so the key loads are issued eight steps
  // ahead: the walk was a chain of 64 dependent load latencies per tile
  constexpr int AHEAD = 8;
  for (int64_t base0 = b; base0 < e; base0 += 32 * AHEAD) {
    int kk[AHEAD];
#pragma unroll
    for (int u = 0; u < AHEAD; ++u) {
      const int64_t i = base0 + 32 * u + lane;
      kk[u] = i < e ? keys[i] - kbase : -1 - lane;
    }
#pragma unroll
    for (int u = 0; u < AHEAD; ++u) {
      const int64_t i = base0 + 32 * u + lane;
      if (base0 + 32 * u >= e) break;            // warp-uniform
      const bool valid = i < e;
      const int kl = kk[u];
      const unsigned m = __match_any_sync(FULL, kl);
      uint32_t pos = 0;
      if (valid) pos = cur[kl] + __popc(m & lt);
      __syncwarp();
      if (valid) {
        perm[so + pos] = (uint32_t)(i - so);
        if ((m & lt) == 0) cur[kl] += __popc(m);   // lowest lane of the group advances the cursor
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------- gather-sum
constexpr int GATHER_WARPS = 8;

// DELTA = false: entry j of the key-sorted permutation is row off[s] + perm[j] (key keys[row]).
// DELTA = true : the sorted objects are the signed entries of a DeltaList (kmeans.cuh): entry
//                e = eoff[s] + perm[j] adds (+) or removes (-) row off[s] + (erow[e] & 0x7fffffff)
//                to / from key keys[e]; sums are accumulated and emitted in float64.
//                emitted in float64.
// MODE 2 (exact): rows as in mode 0, but every element is converted to 2^-36 fixed point and summed in int64:
//                integer sums do not depend on the order or on how the rows are split over launches / GPUs.
constexpr float SR_FIXED_SCALE = 68719476736.f;     // 2^36: |x| <= 1 rows, up to 2^26 rows per bin

template <int NV, int MODE>
__global__ void __launch_bounds__(GATHER_WARPS * 32) gather_sum_kernel(
    const float* __restrict__ x, int dim, int64_t N, const int64_t* __restrict__ off, int S,
    const int32_t* __restrict__ keys, const uint32_t* __restrict__ perm,
    float* __restrict__ pieces, int32_t* __restrict__ piece_cnt, Gate gate,
    const int64_t* __restrict__ eoff, const uint32_t* __restrict__ erow) {
  constexpr bool DELTA = MODE == 1 || MODE == 3;       // MODE 3: signed entries AND int64 fixed point (exact delta)
  constexpr bool EXACT = MODE == 2 || MODE == 3;
  typedef typename std::conditional<EXACT, long long, typename std::conditional<DELTA, double, float>::type>::type acc_t;
  if (gate.closed()) return;
  const int lane = threadIdx.x & 31;
  const int64_t run = (int64_t)blockIdx.x * GATHER_WARPS + (threadIdx.x >> 5);
  const int64_t j0 = run * SR_RUN;
  if (DELTA) N = eoff[S];
  if (j0 >= N) return;
  const int n = (int)min((int64_t)SR_RUN, N - j0);

  int64_t pix[2];          // DELTA: bit 62 = "remove"
  int key[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t j = j0 + h * 32 + lane;
    pix[h] = 0; key[h] = -1;
    if (j < j0 + n) {
      if (DELTA) {
        const int s = upper_bound_i64(eoff, S + 1, j) - 1;
        const int64_t e = eoff[s] + perm[j];
        const uint32_t r = erow[e];
        pix[h] = (off[s] + (int64_t)(r & 0x7fffffffu)) | ((int64_t)(r >> 31) << 62);
        key[h] = keys[e];
      } else {
        const int s = upper_bound_i64(off, S + 1, j) - 1;
        pix[h] = off[s] + perm[j];
        key[h] = keys[pix[h]];
      }
    }
  }

  acc_t acc[NV];
#pragma unroll
  for (int m = 0; m < NV; ++m) acc[m] = 0;
  int cur = -1, cnt = 0;

  auto flush = [&]() {
    const int64_t id = run + cur;
    acc_t* dst = reinterpret_cast<acc_t*>(pieces) + id * dim;
#pragma unroll
    for (int m = 0; m < NV; ++m) {
      const int d = lane + 32 * m;
      if (d < dim) dst[d] = acc[m];
    }
    if (lane == 0) piece_cnt[id] = cnt;
  };

  constexpr int U = 4;
  for (int e0 = 0; e0 < n; e0 += U) {
    float v[U][NV];
    int kk[U];
    bool neg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u;
      const int src = e & 31;
      int64_t p = __shfl_sync(FULL, e < 32 ? pix[0] : pix[1], src);
      kk[u] = __shfl_sync(FULL, e < 32 ? key[0] : key[1], src);
      neg[u] = DELTA && ((p >> 62) & 1);
      p &= ~(1ll << 62);
      if (e < n) {
        const float* row = x + p * dim;
#pragma unroll
        for (int m = 0; m < NV; ++m) {
          const int d = lane + 32 * m;
          v[u][m] = d < dim ? ld_stream(row + d) : 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (e0 + u < n) {
        if (kk[u] != cur) {
          if (cur >= 0) flush();
          cur = kk[u];
          cnt = 0;
#pragma unroll
          for (int m = 0; m < NV; ++m) acc[m] = 0;
        }
        if (DELTA && EXACT) {
#pragma unroll
          for (int m = 0; m < NV; ++m) {
            const long long q = __float2ll_rn(v[u][m] * SR_FIXED_SCALE);
            acc[m] += (acc_t)(neg[u] ? -q : q);
          }
          cnt += neg[u] ? -1 : 1;
        } else if (DELTA) {
#pragma unroll
          for (int m = 0; m < NV; ++m) acc[m] += (acc_t)(neg[u] ? -v[u][m] : v[u][m]);
          cnt += neg[u] ? -1 : 1;
        } else if (EXACT) {
#pragma unroll
          for (int m = 0; m < NV; ++m) acc[m] += (acc_t)__float2ll_rn(v[u][m] * SR_FIXED_SCALE);
          ++cnt;
        } else {
#pragma unroll
          for (int m = 0; m < NV; ++m) acc[m] += v[u][m];
          ++cnt;
        }
      }
    }
  }
  if (cur >= 0) flush();
}

// ---------------------------------------------------------------- delta gather, lean inner loop
// gather_sum_kernel<NV, 1> spends ~116 instructions per entry (r2 ncu: 127 M for 1.09 M entries, 2 IPC: the pass is
// issue-bound, not DRAM-bound): a select and an address computation per element, a 64-bit shuffle pair per entry,
// register copies around the bin-change branch.  Same work, same order of additions, for rows of 32*NVF + tail floats
// (the D+2 rows of k-means): the (row, sign) of an entry travels as one 32-bit word, the full vectors load
// unconditionally at immediate offsets, and the sign is a +-1.0 factor of a float64 FMA (acc + (-1)(double)v is the
// same rounding as acc + (double)(-v)), so an element costs load + convert + FMA.
template <int NVF>
__global__ void __launch_bounds__(GATHER_WARPS * 32) gather_delta_lean_kernel(
    const float* __restrict__ x, int dim, const int64_t* __restrict__ off, int S,
    const int32_t* __restrict__ keys, const uint32_t* __restrict__ perm,
    double* __restrict__ pieces, int32_t* __restrict__ piece_cnt, Gate gate,
    const int64_t* __restrict__ eoff, const uint32_t* __restrict__ erow) {
  constexpr int NV = NVF + 1;
  if (gate.closed()) return;
  const int lane = threadIdx.x & 31;
  const int64_t run = (int64_t)blockIdx.x * GATHER_WARPS + (threadIdx.x >> 5);
  const int64_t j0 = run * SR_RUN;
  const int64_t M = eoff[S];
  if (j0 >= M) return;
  const int n = (int)min((int64_t)SR_RUN, M - j0);
  const bool tail_on = 32 * NVF + lane < dim;
  const float* xl = x + lane;

  uint32_t rs[2];          // global row (< 2^31) | remove bit
  int key[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t j = j0 + h * 32 + lane;
    rs[h] = 0; key[h] = -1;
    if (j < j0 + n) {
      const int s = upper_bound_i64(eoff, S + 1, j) - 1;
      const int64_t e = eoff[s] + perm[j];
      const uint32_t r = erow[e];
      rs[h] = (uint32_t)(off[s] + (int64_t)(r & 0x7fffffffu)) | (r & 0x80000000u);
      key[h] = keys[e];
    }
  }

  double acc[NV];
#pragma unroll
  for (int m = 0; m < NV; ++m) acc[m] = 0.0;
  int cur = -1, cnt = 0;
  auto flush = [&]() {
    const int64_t id = run + cur;
    double* dst = pieces + id * dim + lane;
#pragma unroll
    for (int m = 0; m < NVF; ++m) dst[32 * m] = acc[m];
    if (tail_on) dst[32 * NVF] = acc[NVF];
    if (lane == 0) piece_cnt[id] = cnt;
  };

  constexpr int U = 4;
  for (int e0 = 0; e0 < n; e0 += U) {
    float v[U][NV];
    uint32_t rr[U];
    int kk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u;
      rr[u] = __shfl_sync(FULL, e < 32 ? rs[0] : rs[1], e & 31);
      kk[u] = __shfl_sync(FULL, e < 32 ? key[0] : key[1], e & 31);
      if (e < n) {
        const float* row = xl + (size_t)(rr[u] & 0x7fffffffu) * dim;
#pragma unroll
        for (int m = 0; m < NVF; ++m) v[u][m] = ld_stream(row + 32 * m);
        v[u][NVF] = tail_on ? ld_stream(row + 32 * NVF) : 0.f;
      }
    }
    // the usual batch: four entries of the bin that is already open -- straight-line code, the accumulators stay where
    // they are (with the bin-change branch inside, the compiler copies all 18 accumulator registers per entry)
    if (e0 + U <= n && kk[0] == cur && kk[1] == cur && kk[2] == cur && kk[3] == cur) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool neg = rr[u] >> 31;
        const double sg = neg ? -1.0 : 1.0;
        cnt += neg ? -1 : 1;
#pragma unroll
        for (int m = 0; m < NV; ++m) acc[m] = fma((double)v[u][m], sg, acc[m]);
      }
      continue;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (e0 + u < n) {
        if (kk[u] != cur) {
          if (cur >= 0) flush();
          cur = kk[u];
          cnt = 0;
#pragma unroll
          for (int m = 0; m < NV; ++m) acc[m] = 0.0;
        }
        const bool neg = rr[u] >> 31;
        const double sg = neg ? -1.0 : 1.0;
        cnt += neg ? -1 : 1;
#pragma unroll
        for (int m = 0; m < NV; ++m) acc[m] = fma((double)v[u][m], sg, acc[m]);
      }
    }
  }
  if (cur >= 0) flush();
}

// ---------------------------------------------------------------- combine
constexpr int COMBINE_WARPS = 8;

__global__ void __launch_bounds__(COMBINE_WARPS * 32) combine_kernel(
    int64_t P, int dim, int S, int kmax, const int64_t* __restrict__ seg_base,
    const int64_t* __restrict__ bin_start, const int32_t* __restrict__ bin_count,
    const float* __restrict__ pieces, int mode, float* __restrict__ out,
    float* __restrict__ sums_out, float* __restrict__ counts_out) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * COMBINE_WARPS + (threadIdx.x >> 5);
  if (p >= P) return;
  int64_t key = p;
  bool covered = true;
  if (seg_base) {
    const int s = upper_bound_i64(seg_base, S, p) - 1;
    const int64_t k = s >= 0 ? p - seg_base[s] : kmax;
    covered = s >= 0 && k < kmax;
    key = covered ? (int64_t)s * kmax + k : 0;
  }
  const int cnt = covered ? bin_count[key] : 0;
  const int64_t start = covered ? bin_start[key] : 0;
  const int64_t r0 = start / SR_RUN, r1 = cnt > 0 ? (start + cnt - 1) / SR_RUN : r0 - 1;
  float* o = out + p * dim;
  float ss = 0.f;
  for (int d = lane; d < dim; d += 32) {
    float a = 0.f;
    for (int64_t r = r0; r <= r1; ++r) a += pieces[(r + key) * dim + d];
    if (sums_out) sums_out[p * dim + d] = a;
    o[d] = a;
    ss = fmaf(a, a, ss);
  }
  if (counts_out && lane == 0) counts_out[p] = (float)cnt;
  if (mode == HSG_REDUCE_NORMALIZE) {
    const float n = safe_norm(warp_sum(ss));
    for (int d = lane; d < dim; d += 32) o[d] = o[d] / n;
  } else if (mode == HSG_REDUCE_MEAN) {
    const float c = cnt > 0 ? (float)cnt : 1.f;
    for (int d = lane; d < dim; d += 32) o[d] = o[d] / c;
  }
}

// k-means centroids from running float64 sums (kmeans.cu keeps them across iterations):
// a full pass SETS sums = sum of the float pieces, a delta pass ADDS the float64 pieces of the
// signed entries.  Member counts are carried exactly, so a cluster that lost every member is
// reset to an exact zero sum (the reference's empty cluster: zero centroid).
template <int NV>
__global__ void __launch_bounds__(COMBINE_WARPS * 32) combine64_kernel(
    int64_t bins, int dim, const int64_t* __restrict__ bin_start, const int32_t* __restrict__ bin_count,
    const float* __restrict__ pieces, const double* __restrict__ p64, const int32_t* __restrict__ piece_cnt,
    const int32_t* __restrict__ delta_flag, double* __restrict__ sums, int32_t* __restrict__ members,
    float* __restrict__ out, Gate gate) {
  const int lane = threadIdx.x & 31;
  const int64_t key = (int64_t)blockIdx.x * COMBINE_WARPS + (threadIdx.x >> 5);
  if (gate.closed() || key >= bins) return;
  const bool delta = delta_flag && *delta_flag == 1;
  const int cnt = bin_count[key];
  const int64_t start = bin_start[key];
  const int64_t r0 = start / SR_RUN, r1 = cnt > 0 ? (start + cnt - 1) / SR_RUN : r0 - 1;
  int mem = cnt;
  if (delta) {
    // member count: lanes read the pieces' counts in parallel, integer sum (order-free)
    int c = 0;
    for (int64_t r = r0 + lane; r <= r1; r += 32) c += piece_cnt[r + key];
    mem = members[key] + warp_sum(c);
  }
  if (lane == 0) members[key] = mem;
  double* sk = sums + key * dim;
  // the whole row lives in registers: every run contributes NV independent loads (the kernel used to be a chain
  // of dependent load latencies, 30-50 us per launch whatever the batch size); pieces are added in run order
  double a[NV];
#pragma unroll
  for (int m = 0; m < NV; ++m) {
    const int d = lane + 32 * m;
    a[m] = (delta && d < dim) ? sk[d] : 0.0;
  }
  if (delta) {
#pragma unroll 2
    for (int64_t r = r0; r <= r1; ++r) {
      const double* src = p64 + (r + key) * dim + lane;
#pragma unroll
      for (int m = 0; m < NV; ++m)
        if (lane + 32 * m < dim) a[m] += src[32 * m];
    }
  } else {
#pragma unroll 2
    for (int64_t r = r0; r <= r1; ++r) {
      const float* src = pieces + (r + key) * dim + lane;
#pragma unroll
      for (int m = 0; m < NV; ++m)
        if (lane + 32 * m < dim) a[m] += (double)src[32 * m];
    }
  }
  float f[NV];
  float ss = 0.f;
#pragma unroll
  for (int m = 0; m < NV; ++m) {
    const int d = lane + 32 * m;
    if (mem == 0) a[m] = 0.0;
    f[m] = (float)a[m];
    if (d < dim) {
      sk[d] = a[m];
      ss = fmaf(f[m], f[m], ss);
    }
  }
  const float n = safe_norm(warp_sum(ss));
  float* o = out + key * dim;
#pragma unroll
  for (int m = 0; m < NV; ++m) {
    const int d = lane + 32 * m;
    if (d < dim) o[d] = f[m] / n;
  }
}

// ---------------------------------------------------------------- delta list (incremental k-means M-step)
// Rows whose key changed since the previous M-step, as signed entries in row order:
// (row, old key, remove) then (row, new key, add).  Built without host synchronisation;
// flag[0] says whether the list fits the capacity (delta pass) or the full pass has to run.
__global__ void __launch_bounds__(256) delta_count_kernel(Tiles t, const int32_t* __restrict__ keys,
                                                          const int32_t* __restrict__ prev,
                                                          int32_t* __restrict__ tile_entries) {
  __shared__ int wsum[8];
  const int ti = blockIdx.x;
  if (ti >= *t.count) return;
  const int64_t b = t.begin[ti], e = t.end[ti];
  int c = 0;
  for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) c += keys[i] != prev[i];
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < 8; ++w) tot += wsum[w];
    tile_entries[ti] = 2 * tot;
  }
}

__global__ void __launch_bounds__(1024) delta_scan_kernel(Tiles t, int S, int32_t* __restrict__ tile_entries,
                                                          int64_t* __restrict__ eoff, int32_t* __restrict__ flag,
                                                          int64_t cap) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry_s;
  const int nt = *t.count;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nt; base += 1024) {
    const int ti = base + threadIdx.x;
    const int64_t v = ti < nt ? tile_entries[ti] : 0;
    int64_t total;
    const int64_t before = carry_s + block_scan_excl_1024(v, warp_tot, &total);
    if (ti < nt) tile_entries[ti] = (int32_t)min(before, (int64_t)0x7fffffff);
    __syncthreads();
    if (threadIdx.x == 0) carry_s += total;
    __syncthreads();
  }
  const int64_t M = carry_s;
  for (int s = threadIdx.x; s <= S; s += blockDim.x) {
    const int first = t.seg_first[s];
    eoff[s] = (s == S || first >= nt) ? M : (int64_t)tile_entries[first];
  }
  if (threadIdx.x == 0) {
    flag[0] = M <= cap ? 1 : 0;
    flag[1] = (int32_t)min(M / 2, (int64_t)0x7fffffff);
  }
}

__global__ void __launch_bounds__(256) delta_compact_kernel(Tiles t, const int64_t* __restrict__ off,
                                                            const int32_t* __restrict__ keys,
                                                            int32_t* __restrict__ prev,
                                                            const int32_t* __restrict__ tile_entries,
                                                            const int32_t* __restrict__ flag,
                                                            uint32_t* __restrict__ erow, int32_t* __restrict__ ekey) {
  __shared__ int wsum[8];
  __shared__ int carry_s;
  const int ti = blockIdx.x;
  if (ti >= *t.count) return;
  const bool emit = flag[0] == 1;
  const int64_t b = t.begin[ti], e = t.end[ti];
  const int64_t so = off[t.seg[ti]];
  const int64_t base = tile_entries[ti];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t c0 = b; c0 < e; c0 += 256 * 8) {
    const int64_t i0 = c0 + (int64_t)threadIdx.x * 8;
    int kn[8], ko[8];
    int c = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t i = i0 + k;
      kn[k] = ko[k] = 0;
      if (i < e) { kn[k] = keys[i]; ko[k] = prev[i]; }
      c += kn[k] != ko[k];
    }
    const int incl = warp_scan_incl(c, lane);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int before = carry_s + incl - c;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    int64_t pos = base + 2 * (int64_t)before;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (kn[k] != ko[k]) {
        const int64_t i = i0 + k;
        if (emit) {
          const uint32_t r = (uint32_t)(i - so);
          erow[pos] = r | 0x80000000u; ekey[pos] = ko[k];
          erow[pos + 1] = r;           ekey[pos + 1] = kn[k];
          pos += 2;
        }
        prev[i] = kn[k];
      }
    }
    __syncthreads();
    if (threadIdx.x == 255) carry_s = before + c;
    __syncthreads();
  }
}

int sr_delta_build(const SegReducePlan& p, const int64_t* seg_offsets, int32_t* keys_prev,
                   int32_t* tile_entries, int64_t* eoff, int32_t* flag, int64_t cap, uint32_t* erow,
                   int32_t* ekey, cudaStream_t st) {
  ProfRange prof(PROF_MSTEP_SORT, st);
  delta_count_kernel<<<(unsigned)p.tiles.bound, 256, 0, st>>>(p.tiles, p.keys, keys_prev, tile_entries);
  HSG_LAUNCH_CHECK();
  delta_scan_kernel<<<1, 1024, 0, st>>>(p.tiles, p.S, tile_entries, eoff, flag, cap);
  HSG_LAUNCH_CHECK();
  delta_compact_kernel<<<(unsigned)p.tiles.bound, 256, 0, st>>>(p.tiles, seg_offsets, p.keys, keys_prev,
                                                                tile_entries, flag, erow, ekey);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

// ---------------------------------------------------------------- host side
int64_t sr_tile_size(int64_t max_seg_len) {
  int64_t tile = SR_TILE_MIN;
  while (ceil_div64(max_seg_len, tile) > 1024) tile *= 2;
  return tile;
}

int64_t sr_tiles_bound(int64_t N, int S, int64_t tile) { return N / tile + S + 1; }

static int64_t sr_num_pieces(int64_t N, int64_t bins) { return ceil_div64(N, SR_RUN) + bins + 2; }

void sr_carve(Carver& c, SegReducePlan& p, int64_t N, int dim, int S, int kmax, int64_t max_seg_len, bool exact) {
  p.N = N; p.dim = dim; p.S = S; p.kmax = kmax; p.bins = (int64_t)S * kmax; p.exact = exact;
  p.tiles.tile = sr_tile_size(max_seg_len);
  p.tiles.bound = sr_tiles_bound(N, S, p.tiles.tile);
  p.tiles.seg = c.take<int32_t>(p.tiles.bound);
  p.tiles.begin = c.take<int64_t>(p.tiles.bound);
  p.tiles.end = c.take<int64_t>(p.tiles.bound);
  p.tiles.seg_first = c.take<int32_t>(S + 1);
  p.tiles.count = c.take<int32_t>(1);
  p.keys = c.take<int32_t>(N);
  p.perm = c.take<uint32_t>(N);
  p.tile_hist = c.take<uint32_t>(p.tiles.bound * kmax);
  p.bin_start = c.take<int64_t>(p.bins);
  p.bin_count = c.take<int32_t>(p.bins);
  const int64_t np = sr_num_pieces(N, p.bins);
  p.pieces = c.take<float>(np * dim * (exact ? 2 : 1));      // int64 pieces in exact mode
  p.piece_cnt = c.take<int32_t>(np);
}

size_t sr_workspace_bytes(int64_t N, int dim, int S, int kmax, int64_t max_seg_len, bool exact) {
  Carver c(nullptr);
  SegReducePlan p;
  sr_carve(c, p, N, dim, S, kmax, max_seg_len, exact);
  return c.used() + 256;
}

int sr_build_tiles(const SegReducePlan& p, const int64_t* seg_offsets, cudaStream_t st) {
  build_tiles_kernel<<<1, 1024, 0, st>>>(seg_offsets, p.S, p.tiles);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

static dim3 row_grid(const Tiles& t, int chunk) {
  return dim3((unsigned)t.bound, (unsigned)ceil_div64(t.tile, chunk));
}

int sr_labels_to_keys(const SegReducePlan& p, const int64_t* labels, const int64_t* seg_base,
                      cudaStream_t st) {
  labels_to_keys_kernel<<<row_grid(p.tiles, SR_TILE_MIN), 256, 0, st>>>(p.tiles, labels, seg_base,
                                                                         p.kmax, p.keys);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int sr_keys_to_labels(const SegReducePlan& p, const int32_t* keys, int64_t* labels, cudaStream_t st) {
  keys_to_labels_kernel<<<row_grid(p.tiles, SR_TILE_MIN), 256, 0, st>>>(p.tiles, keys, p.kmax, labels);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

template <int NV, int DELTA>
static int launch_gather(const SegReducePlan& p, const float* x, const int64_t* off, Gate gate,
                         const int64_t* eoff, const uint32_t* erow, cudaStream_t st) {
  const int64_t runs = ceil_div64(p.N, SR_RUN);
  gather_sum_kernel<NV, DELTA><<<(unsigned)ceil_div64(runs, GATHER_WARPS), GATHER_WARPS * 32, 0, st>>>(
      x, p.dim, p.N, off, p.S, p.keys, p.perm, p.pieces, p.piece_cnt, gate, eoff, erow);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

template <int DELTA>
static int gather_dispatch(const SegReducePlan& p, const float* x, const int64_t* off, Gate gate,
                           const int64_t* eoff, const uint32_t* erow, cudaStream_t st) {
  const int nv = (p.dim + 31) / 32;
  if (nv <= 1) return launch_gather<1, DELTA>(p, x, off, gate, eoff, erow, st);
  if (nv <= 2) return launch_gather<2, DELTA>(p, x, off, gate, eoff, erow, st);
  if (nv <= 3) return launch_gather<3, DELTA>(p, x, off, gate, eoff, erow, st);
  if (nv <= 5) return launch_gather<5, DELTA>(p, x, off, gate, eoff, erow, st);
  if (nv <= 9) return launch_gather<9, DELTA>(p, x, off, gate, eoff, erow, st);
  if (nv <= 12) return launch_gather<12, DELTA>(p, x, off, gate, eoff, erow, st);
  if (nv <= 17) return launch_gather<17, DELTA>(p, x, off, gate, eoff, erow, st);
  return launch_gather<20, DELTA>(p, x, off, gate, eoff, erow, st);
}

int sr_sort_and_sum_gated(const SegReducePlan& p, const float* x, const int64_t* off, Gate gate,
                          const int64_t* eoff, const uint32_t* erow, cudaStream_t st) {
  HSG_REQUIRE(p.kmax <= SR_MAX_KEYS, HSG_E_UNSUPPORTED, "segment reduce: %d keys per segment (max %d)", p.kmax, SR_MAX_KEYS);
  HSG_REQUIRE(p.dim <= 32 * 20, HSG_E_UNSUPPORTED, "segment reduce: dim %d (max 640)", p.dim);
  const int64_t* sort_off = eoff ? eoff : off;       // the objects being sorted: delta entries or rows
  {
  ProfRange prof(PROF_MSTEP_SORT, st);
  const size_t hist_smem = (size_t)p.kmax * sizeof(uint32_t);
  if (hist_smem > 48 * 1024)
    HSG_CUDA(cudaFuncSetAttribute(hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem));
  hist_kernel<<<(unsigned)p.tiles.bound, 512, hist_smem, st>>>(p.tiles, p.keys, p.kmax, p.tile_hist, gate);
  HSG_LAUNCH_CHECK();
  scan_kernel<<<p.S, 1024, 0, st>>>(p.tiles, sort_off, p.kmax, p.tile_hist, p.bin_start, p.bin_count, gate);
  HSG_LAUNCH_CHECK();
  if (p.kmax <= 8192) {
    const size_t sc_smem = (size_t)4 * p.kmax * sizeof(uint32_t);
    if (sc_smem > 48 * 1024)
      HSG_CUDA(cudaFuncSetAttribute(scatter_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));
    scatter_kernel<4><<<(unsigned)ceil_div64(p.tiles.bound, 4), 128, sc_smem, st>>>(
        p.tiles, sort_off, p.keys, p.kmax, p.tile_hist, p.bin_start, p.perm, gate);
  } else {
    const size_t sc_smem = (size_t)p.kmax * sizeof(uint32_t);
    HSG_CUDA(cudaFuncSetAttribute(scatter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));
    scatter_kernel<1><<<(unsigned)p.tiles.bound, 32, sc_smem, st>>>(
        p.tiles, sort_off, p.keys, p.kmax, p.tile_hist, p.bin_start, p.perm, gate);
  }
  HSG_LAUNCH_CHECK();
  }
  ProfRange prof(PROF_MSTEP_GATHER, st);
  static const bool no_lean = getenv("HSG_GATHER_NO_LEAN") != nullptr;       // A/B switch for profiling only
  if (eoff && !p.exact && !no_lean && p.dim % 32 != 0 && (p.dim / 32 == 2 || p.dim / 32 == 4 || p.dim / 32 == 8)) {
    const unsigned grid = (unsigned)ceil_div64(ceil_div64(p.N, SR_RUN), GATHER_WARPS);
    double* p64 = reinterpret_cast<double*>(p.pieces);
    switch (p.dim / 32) {
      case 2: gather_delta_lean_kernel<2><<<grid, GATHER_WARPS * 32, 0, st>>>(x, p.dim, off, p.S, p.keys, p.perm, p64, p.piece_cnt, gate, eoff, erow); break;
      case 4: gather_delta_lean_kernel<4><<<grid, GATHER_WARPS * 32, 0, st>>>(x, p.dim, off, p.S, p.keys, p.perm, p64, p.piece_cnt, gate, eoff, erow); break;
      default: gather_delta_lean_kernel<8><<<grid, GATHER_WARPS * 32, 0, st>>>(x, p.dim, off, p.S, p.keys, p.perm, p64, p.piece_cnt, gate, eoff, erow); break;
    }
    HSG_LAUNCH_CHECK();
    return HSG_OK;
  }
  if (p.exact) return eoff ? gather_dispatch<3>(p, x, off, gate, eoff, erow, st)
                           : gather_dispatch<2>(p, x, off, gate, nullptr, nullptr, st);
  return eoff ? gather_dispatch<1>(p, x, off, gate, eoff, erow, st)
              : gather_dispatch<0>(p, x, off, gate, nullptr, nullptr, st);
}

int sr_sort_and_sum(const SegReducePlan& p, const float* x, const int64_t* off, cudaStream_t st) {
  return sr_sort_and_sum_gated(p, x, off, Gate{nullptr, 0}, nullptr, nullptr, st);
}

// exact sums: out[p, d] = sum of the int64 fixed-point pieces of bin p (see gather_sum_kernel, mode 2)
__global__ void __launch_bounds__(COMBINE_WARPS * 32) combine_exact_kernel(
    int64_t P, int dim, int S, int kmax, const int64_t* __restrict__ seg_base,
    const int64_t* __restrict__ bin_start, const int32_t* __restrict__ bin_count,
    const long long* __restrict__ pieces, long long* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * COMBINE_WARPS + (threadIdx.x >> 5);
  if (p >= P) return;
  int64_t key = p;
  bool covered = true;
  if (seg_base) {
    const int s = upper_bound_i64(seg_base, S, p) - 1;
    const int64_t k = s >= 0 ? p - seg_base[s] : kmax;
    covered = s >= 0 && k < kmax;
    key = covered ? (int64_t)s * kmax + k : 0;
  }
  const int cnt = covered ? bin_count[key] : 0;
  const int64_t start = covered ? bin_start[key] : 0;
  const int64_t r0 = start / SR_RUN, r1 = cnt > 0 ? (start + cnt - 1) / SR_RUN : r0 - 1;
  for (int d = lane; d < dim; d += 32) {
    long long a = 0;
    for (int64_t r = r0; r <= r1; ++r) a += pieces[(r + key) * dim + d];
    out[p * dim + d] = a;
  }
}

int sr_combine_exact(const SegReducePlan& p, int64_t P, const int64_t* seg_base, long long* out, cudaStream_t st) {
  if (P == 0) return HSG_OK;
  combine_exact_kernel<<<(unsigned)ceil_div64(P, COMBINE_WARPS), COMBINE_WARPS * 32, 0, st>>>(
      P, p.dim, p.S, p.kmax, seg_base, p.bin_start, p.bin_count, reinterpret_cast<const long long*>(p.pieces), out);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int sr_combine64(const SegReducePlan& p, const float* pieces_full, const double* pieces_delta,
                 const int32_t* delta_flag, double* sums, int32_t* members, float* out, cudaStream_t st, Gate gate) {
  ProfRange prof(PROF_MSTEP_COMBINE, st);
  const unsigned grid = (unsigned)ceil_div64(p.bins, COMBINE_WARPS);
#define HSG_COMBINE64(NV) combine64_kernel<NV><<<grid, COMBINE_WARPS * 32, 0, st>>>( \
      p.bins, p.dim, p.bin_start, p.bin_count, pieces_full, pieces_delta, p.piece_cnt, delta_flag, sums, members, out, gate)
  if (p.dim <= 32 * 3) HSG_COMBINE64(3);
  else if (p.dim <= 32 * 5) HSG_COMBINE64(5);
  else if (p.dim <= 32 * 9) HSG_COMBINE64(9);
  else HSG_COMBINE64(20);
#undef HSG_COMBINE64
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int sr_combine(const SegReducePlan& p, int64_t P, const int64_t* seg_base, int mode, float* out,
               float* sums_out, float* counts_out, cudaStream_t st) {
  if (P == 0) return HSG_OK;
  ProfRange prof(PROF_MSTEP_COMBINE, st);
  combine_kernel<<<(unsigned)ceil_div64(P, COMBINE_WARPS), COMBINE_WARPS * 32, 0, st>>>(
      P, p.dim, p.S, p.kmax, seg_base, p.bin_start, p.bin_count, p.pieces, mode, out, sums_out,
      counts_out);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

}  // namespace hsg
