// K0: front half of segment_by_kmeans as one pass over the embedding tensor.
//
// Reference (hsg/utils/segsort/common.py:305-365 + general/common.py:101-120):
// permute NCHW->NHWC + contiguous, normalise, per image cat(loc) + normalise
// again, then nonzero/index_select to drop ignore pixels: >= 5 full-tensor
// temporaries.  Here: one read of the NCHW tensor (transposed through shared
// memory), one write each of x, x_with_loc and the optional fp16 side copy,
// with the compaction offsets coming from a small count + scan pre-pass.
#include "common.cuh"

namespace hsg {

constexpr int PREP_TPX = 64;       // pixels per CTA tile (256 contiguous bytes of every channel row)
constexpr int PREP_THREADS = 512;  // 16 warps
constexpr int PREP_LD = PREP_TPX + 1;   // shared-memory row stride (conflict-free column reads)

// ------------------------------------------------------------ normalize (a1)
__global__ void normalize_kernel(const float* __restrict__ x, float* __restrict__ y,
                                 int64_t rows, int dim) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * dim;
  float ss = 0.f;
  for (int d = lane; d < dim; d += 32) {
    float v = xr[d];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float n = safe_norm(ss);
  float* yr = y + row * dim;
  for (int d = lane; d < dim; d += 32) yr[d] = xr[d] / n;
}

__global__ void normalize_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                     float* __restrict__ gx, int64_t rows, int dim) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * dim;
  const float* gr = gy + row * dim;
  float ss = 0.f, dot = 0.f;
  for (int d = lane; d < dim; d += 32) {
    float v = xr[d];
    ss = fmaf(v, v, ss);
    dot = fmaf(v, gr[d], dot);
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float n = sqrtf(ss);
  float* out = gx + row * dim;
  if (n >= 1e-12f) {
    const float inv = 1.f / n;
    const float proj = dot * inv * inv;   // <xhat, g> / ||x||
    for (int d = lane; d < dim; d += 32) out[d] = (gr[d] - xr[d] * proj) * inv;
  } else {
    for (int d = lane; d < dim; d += 32) out[d] = gr[d] / 1e-12f;
  }
}

// ------------------------------------------------------------ half copy
// fp16 side copy for the tensor-core E-step, HSG_XH_TAIL extra columns per row:
//   [0,d16)            fp16(x_d)
//   d16+3l+{0,1,2}     (hi, hi, lo) of the l-th trailing feature (location), hi = fp16(v), lo = fp16(v - hi);
//                      against the centroid side's (hi, lo, hi) this gives v*c to ~2^-21 relative on the tensor core
//   d16+15             1.0 (multiplies the centroid side's "row is padding" marker)
// lanes 0..15 of a warp write the tail of one row.
__device__ __forceinline__ void write_xh_tail(__half* tail, int lane, int L, float v_l /*feature lane/3*/) {
  if (lane < HSG_XH_TAIL) {
    float o = 0.f;
    const int l = lane / 3, part = lane % 3;
    if (lane == HSG_XH_TAIL - 1) {
      o = 1.f;
    } else if (l < L) {
      const float hi = __half2float(__float2half_rn(v_l));
      o = part < 2 ? hi : v_l - hi;
    }
    tail[lane] = __float2half_rn(o);
  }
}

__global__ void half_copy_kernel(const float* __restrict__ x, int64_t rows, int dim, int d16,
                                 __half* __restrict__ xh, float* __restrict__ xerr) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * dim;
  __half* hr = xh + row * (d16 + HSG_XH_TAIL);
  float e2 = 0.f;
  for (int d = lane; d < d16; d += 32) {
    const float v = xr[d];
    const __half h = __float2half_rn(v);
    const float r = v - __half2float(h);
    e2 = fmaf(r, r, e2);
    hr[d] = h;
  }
  const int L = dim - d16;
  const int l = lane / 3;
  write_xh_tail(hr + d16, lane, L, (lane < HSG_XH_TAIL - 1 && l < L) ? xr[d16 + l] : 0.f);
  e2 = warp_sum(e2);
  if (lane == 0 && xerr) xerr[row] = sqrtf(e2) * 1.0001f + 1e-30f;
}

// ------------------------------------------------------------ compaction pre-pass
__global__ void prep_count_kernel(const int64_t* __restrict__ labels, int64_t ignore_index,
                                  int HW, int tiles_per_image, int64_t n_tiles,
                                  int32_t* __restrict__ tile_count) {
  // one warp per tile of PREP_TPX pixels
  const int lane = threadIdx.x & 31;
  const int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile >= n_tiles) return;
  const int64_t b = tile / tiles_per_image;
  const int t = (int)(tile % tiles_per_image);
  int cnt = 0;
#pragma unroll
  for (int u = 0; u < PREP_TPX / 32; ++u) {
    const int p = t * PREP_TPX + 32 * u + lane;
    bool valid = false;
    if (p < HW) valid = labels[b * HW + p] != ignore_index;
    cnt += __popc(__ballot_sync(FULL, valid));
  }
  if (lane == 0) tile_count[tile] = cnt;
}

// exclusive scan of tile counts (single CTA, 1024 threads, chunked)
__global__ void prep_scan_kernel(const int32_t* __restrict__ tile_count, int64_t n_tiles,
                                 int tiles_per_image, int B, int64_t* __restrict__ tile_base,
                                 int64_t* __restrict__ seg_offsets) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_tiles; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    const int v = i < n_tiles ? tile_count[i] : 0;
    const int incl = warp_scan_incl(v, lane);
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int64_t w = warp_tot[lane];
      // inclusive scan of 32 warp totals
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(FULL, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const int64_t before = carry + (warp ? warp_tot[warp - 1] : 0) + (incl - v);
    if (i < n_tiles) {
      tile_base[i] = before;
      if (i % tiles_per_image == 0) seg_offsets[i / tiles_per_image] = before;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) seg_offsets[B] = carry;
}

__global__ void fill_dense_offsets_kernel(int64_t* seg_offsets, int B, int64_t HW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= B) seg_offsets[i] = (int64_t)i * HW;
}

// a / b rounded to nearest, given rb = RN(1/b): one FMA correction step of the product a*rb
// is correctly rounded (Markstein) -- three FMA-pipe operations instead of the ~10-instruction
// IEEE division sequence, which made this pass instruction-bound
__device__ __forceinline__ float div_rn_by(float a, float b, float rb) {
  const float q = a * rb;
  const float e = fmaf(-q, b, a);
  return fmaf(e, rb, q);
}

// ------------------------------------------------------------ main pass
struct PrepArgs {
  const float* emb;
  int B, D, HW;
  const float* loc;
  int L;
  int64_t loc_image_stride;
  const int64_t* labels;
  int use_ignore;
  int64_t ignore_index;
  const int64_t* init;
  int64_t init_image_stride;
  int64_t batch_base;
  float* x;
  float* xloc;
  __half* xh;
  float* xerr;
  int64_t* labels_out;
  int64_t* clusters_out;
  int64_t* batch_out;
  int64_t* pixel_out;         // optional: flat source pixel b*HW+p of each row
  const int64_t* tile_base;   // NULL when nothing is dropped
  int tiles_per_image;
  // optional: partial sums of the xloc rows over runs of equal initial cluster id inside each tile
  // (the first k-means M-step without re-reading xloc; see hsg_prep_sums_f32)
  float* run_sums;            // [B*tiles_per_image*HSG_PREP_RUNS, D+L]
  int32_t* run_cluster;       // [B*tiles_per_image*HSG_PREP_RUNS] local cluster id, -1 = unused
  int32_t* run_count;         // [B*tiles_per_image*HSG_PREP_RUNS] pixels in the run
  int32_t* run_overflow;      // [1] set to 1 when a tile has more runs than slots
};

__global__ void __launch_bounds__(PREP_THREADS) prep_main_kernel(const PrepArgs a) {
  extern __shared__ float tile[];                 // [D][PREP_LD] ([D+L] when run sums are wanted)
  __shared__ int64_t rows[PREP_TPX];
  __shared__ int ck[PREP_TPX];                    // initial cluster of each kept pixel, -1 = dropped
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int t = blockIdx.x;
  const int p0 = t * PREP_TPX;

  // NCHW read: one channel row of 64 pixels per warp-iteration (two coalesced 128-byte requests),
  // as asynchronous global->shared copies so every request of the CTA is in flight at once
  const float* src = a.emb + (int64_t)b * a.D * a.HW;
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
  for (int d = warp; d < a.D; d += PREP_THREADS / 32) {
    const float* cr = src + (int64_t)d * a.HW + p0;
#pragma unroll
    for (int u = 0; u < PREP_TPX / 32; ++u) {
      const int px = 32 * u + lane;
      const bool in = p0 + px < a.HW;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(tile_s + 4u * (uint32_t)(d * PREP_LD + px)),
                   "l"(in ? cr + px : src), "r"(in ? 4 : 0) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  if (warp < PREP_TPX / 32) {
    // rows of the output: pixels kept in order, ignore pixels dropped (ranks inside the tile)
    const int px = 32 * warp + lane;
    const int p = p0 + px;
    const bool inb = p < a.HW;
    int64_t lab = 0;
    if (inb && a.labels) lab = a.labels[(int64_t)b * a.HW + p];
    const bool valid = inb && !(a.use_ignore && lab == a.ignore_index);
    const unsigned m = __ballot_sync(FULL, valid);
    int before = 0;                                // valid pixels in the lower 32-pixel groups of this tile
    for (int w = 0; w < warp; ++w) {
      const int pw = p0 + 32 * w + lane;
      bool v = pw < a.HW;
      if (v && a.use_ignore) v = a.labels[(int64_t)b * a.HW + pw] != a.ignore_index;
      before += __popc(__ballot_sync(FULL, v));
    }
    const int64_t base = a.tile_base ? a.tile_base[(int64_t)b * a.tiles_per_image + t]
                                     : (int64_t)b * a.HW + p0;
    const int64_t row = valid ? base + before + __popc(m & ((1u << lane) - 1u)) : -1;
    rows[px] = row;
    const int64_t c0 = valid ? a.init[(int64_t)b * a.init_image_stride + p] : -1;
    ck[px] = (int)c0;
    if (valid) {
      a.labels_out[row] = lab;
      a.clusters_out[row] = c0;
      a.batch_out[row] = a.batch_base + b;
      if (a.pixel_out) a.pixel_out[row] = (int64_t)b * a.HW + p;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int Dp = a.D + a.L;
  for (int i = 0; i < PREP_TPX / (PREP_THREADS / 32); ++i) {
    const int px = warp * (PREP_TPX / (PREP_THREADS / 32)) + i;
    const int64_t row = rows[px];
    if (row < 0) {                                  // warp-uniform
      if (a.run_sums) for (int d = lane; d < a.D + a.L; d += 32) tile[d * PREP_LD + px] = 0.f;
      continue;
    }
    float ss = 0.f;
    for (int d = lane; d < a.D; d += 32) {
      const float v = tile[d * PREP_LD + px];
      ss = fmaf(v, v, ss);
    }
    const float n1 = safe_norm(warp_sum(ss));
    const float r1 = __frcp_rn(n1);
    // second normalisation over cat(normalised embedding, local features); the normalised value
    // replaces the raw one in the tile so the last pass divides only once
    float ss2 = 0.f;
    for (int d = lane; d < a.D; d += 32) {
      const float y = div_rn_by(tile[d * PREP_LD + px], n1, r1);
      tile[d * PREP_LD + px] = y;
      ss2 = fmaf(y, y, ss2);
    }
    float lv = 0.f;
    if (lane < a.L) {
      lv = a.loc[(int64_t)b * a.loc_image_stride + (int64_t)(p0 + px) * a.L + lane];
      ss2 = fmaf(lv, lv, ss2);
    }
    const float n2 = safe_norm(warp_sum(ss2));
    const float r2 = __frcp_rn(n2);
    float* xr = a.x + row * a.D;
    float* xl = a.xloc + row * Dp;
    float e2 = 0.f;
    for (int d = lane; d < a.D; d += 32) {
      const float y = tile[d * PREP_LD + px];
      const float z = div_rn_by(y, n2, r2);
      xr[d] = y;
      xl[d] = z;
      if (a.run_sums) tile[d * PREP_LD + px] = z;
      if (a.xh) {
        const __half h = __float2half_rn(z);
        const float r = z - __half2float(h);
        e2 = fmaf(r, r, e2);
        a.xh[row * (a.D + HSG_XH_TAIL) + d] = h;
      }
    }
    if (lane < a.L) {
      xl[a.D + lane] = lv / n2;
      if (a.run_sums) tile[(a.D + lane) * PREP_LD + px] = lv / n2;
    }
    if (a.xh) {
      const int l = lane / 3;
      const float vl = __shfl_sync(FULL, lv / n2, l < a.L ? l : 0);
      write_xh_tail(a.xh + row * (a.D + HSG_XH_TAIL) + a.D, lane, a.L, vl);
    }
    if (a.xh && a.xerr) {
      e2 = warp_sum(e2);
      if (lane == 0) a.xerr[row] = sqrtf(e2) * 1.0001f + 1e-30f;
    }
  }

  if (a.run_sums) {
    // The tile now holds the xloc rows (zero columns for dropped pixels).  Warp 0 finds the runs of equal
    // initial cluster among the kept pixels, then one thread per feature sums each run's pixel range -- fixed
    // order, so the sums do not depend on scheduling.
    __shared__ int run_lo[HSG_PREP_RUNS + 1];
    __shared__ int n_runs_s;
    const int64_t slot0 = ((int64_t)b * a.tiles_per_image + t) * HSG_PREP_RUNS;
    if (warp == 0) {
      // run starts among the kept pixels, found with two 32-pixel ballots (no serial walk)
      const int ca = ck[lane], cb = ck[lane + 32];
      const uint64_t kept = (uint64_t)__ballot_sync(FULL, ca >= 0) | ((uint64_t)__ballot_sync(FULL, cb >= 0) << 32);
      auto prev_cluster = [&](int px) {
        const uint64_t below = kept & ((1ull << px) - 1ull);
        return below ? ck[63 - __clzll((long long)below)] : -2;
      };
      const bool sa = ca >= 0 && ca != prev_cluster(lane);
      const bool sb = cb >= 0 && cb != prev_cluster(lane + 32);
      const uint64_t starts = (uint64_t)__ballot_sync(FULL, sa) | ((uint64_t)__ballot_sync(FULL, sb) << 32);
      const int n_all = __popcll(starts);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int px = lane + 32 * h;
        if (h ? sb : sa) {
          const int run = __popcll(starts & ((1ull << px) - 1ull));
          if (run < HSG_PREP_RUNS) {
            const uint64_t later = px == 63 ? 0ull : (starts >> (px + 1)) << (px + 1);
            const int end = later ? __ffsll((long long)later) - 1 : PREP_TPX;
            const uint64_t range = (end == 64 ? ~0ull : ((1ull << end) - 1ull)) & ~((1ull << px) - 1ull);
            run_lo[run] = px;
            a.run_cluster[slot0 + run] = h ? cb : ca;
            a.run_count[slot0 + run] = __popcll(kept & range);
          }
        }
      }
      const int nr0 = min(n_all, HSG_PREP_RUNS);
      if (lane >= nr0 && lane < HSG_PREP_RUNS) a.run_cluster[slot0 + lane] = -1;
      if (lane == 0) {
        if (n_all > HSG_PREP_RUNS) *a.run_overflow = 1;
        run_lo[nr0] = PREP_TPX;                    // dropped pixels inside a range hold zeros
        n_runs_s = nr0;
      }
    }
    __syncthreads();
    const int nr = n_runs_s;
    for (int d = threadIdx.x; d < Dp; d += PREP_THREADS) {
      const float* col = tile + d * PREP_LD;
      for (int r = 0; r < nr; ++r) {
        float acc = 0.f;
        for (int px = run_lo[r]; px < run_lo[r + 1]; ++px) acc += col[px];
        a.run_sums[(slot0 + r) * Dp + d] = acc;
      }
    }
  }
}

// ------------------------------------------------------------ backward of the prep chain
// Reference autograd chain (hsg/utils/segsort/common.py:305-365): permute -> normalize -> cat(loc) ->
// normalize -> index_select.  With r the raw pixel row, y = r/n1, z = [y, loc]/n2:
//   g_cat = (g_z - z <z,g_z>) / n2          (g_z / n2 when ||cat|| < eps)
//   g_y   = g_x + g_cat[:D]
//   g_r   = (g_y - y <y,g_y>) / n1          (g_y / eps when ||r|| < eps)
// scattered back NHWC -> NCHW, zero for dropped pixels.  One CTA per 64 source pixels, like the forward:
// the raw tile comes in by cp.async, the gradient tile leaves through the same transposed shared tile, so
// both NCHW accesses are coalesced; y, g_x, g_z are read twice (dot products, then the update), the second
// time from L1/L2.
struct PrepBwdArgs {
  const float* emb;           // [B,D,HW] raw input of the forward
  int B, D, HW, L;
  const float* x;             // [N,D]   forward output (y)
  const float* loc;           // local features as in the forward
  int64_t loc_image_stride;
  const int64_t* row_of_pixel;// [B*HW] output row of each source pixel, -1 = dropped; NULL = identity
  const float* gx;            // [N,D] or NULL
  const float* gz;            // [N,D+L] or NULL
  float* gemb;                // [B,D,HW]
};

__global__ void __launch_bounds__(PREP_THREADS) prep_bwd_kernel(const PrepBwdArgs a) {
  extern __shared__ float tile[];                 // [D][PREP_LD]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y, t = blockIdx.x, p0 = t * PREP_TPX;
  const float* src = a.emb + (int64_t)b * a.D * a.HW;
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
  for (int d = warp; d < a.D; d += PREP_THREADS / 32) {
    const float* cr = src + (int64_t)d * a.HW + p0;
#pragma unroll
    for (int u = 0; u < PREP_TPX / 32; ++u) {
      const int px = 32 * u + lane;
      const bool in = p0 + px < a.HW;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(tile_s + 4u * (uint32_t)(d * PREP_LD + px)),
                   "l"(in ? cr + px : src), "r"(in ? 4 : 0) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int Dp = a.D + a.L;
  for (int i = 0; i < PREP_TPX / (PREP_THREADS / 32); ++i) {
    const int px = warp * (PREP_TPX / (PREP_THREADS / 32)) + i;
    const int p = p0 + px;
    int64_t row = -1;
    if (p < a.HW) row = a.row_of_pixel ? a.row_of_pixel[(int64_t)b * a.HW + p] : (int64_t)b * a.HW + p;
    if (row < 0) {                                   // warp-uniform: dropped (or out of range) pixel
      for (int d = lane; d < a.D; d += 32) tile[d * PREP_LD + px] = 0.f;
      continue;
    }
    const float* yr = a.x + row * a.D;
    const float* gxr = a.gx ? a.gx + row * a.D : nullptr;
    const float* gzr = a.gz ? a.gz + row * Dp : nullptr;
    float s_rr = 0.f, s_yy = 0.f, s_ygx = 0.f, s_ygz = 0.f;
    for (int d = lane; d < a.D; d += 32) {
      const float r = tile[d * PREP_LD + px], y = yr[d];
      s_rr = fmaf(r, r, s_rr);
      s_yy = fmaf(y, y, s_yy);
      if (gxr) s_ygx = fmaf(y, gxr[d], s_ygx);
      if (gzr) s_ygz = fmaf(y, gzr[d], s_ygz);
    }
    float lv = 0.f, s_ll = 0.f, s_lg = 0.f;
    if (lane < a.L) {
      lv = a.loc[(int64_t)b * a.loc_image_stride + (int64_t)p * a.L + lane];
      s_ll = lv * lv;
      if (gzr) s_lg = lv * gzr[a.D + lane];
    }
    s_rr = warp_sum(s_rr); s_yy = warp_sum(s_yy); s_ygx = warp_sum(s_ygx); s_ygz = warp_sum(s_ygz);
    s_ll = warp_sum(s_ll); s_lg = warp_sum(s_lg);
    const float raw1 = sqrtf(s_rr), raw2 = sqrtf(s_yy + s_ll);
    const bool big1 = raw1 >= 1e-12f, big2 = raw2 >= 1e-12f;
    const float n1 = big1 ? raw1 : 1e-12f, n2 = big2 ? raw2 : 1e-12f;
    const float dotz = (s_ygz + s_lg) / n2;                          // <z, g_z>
    const float y_gcat = big2 ? (s_ygz - (s_yy / n2) * dotz) / n2 : s_ygz / n2;
    const float doty = s_ygx + y_gcat;                               // <y, g_y>
    for (int d = lane; d < a.D; d += 32) {
      const float y = yr[d];
      const float gzv = gzr ? gzr[d] : 0.f;
      const float gcat = big2 ? (gzv - (y / n2) * dotz) / n2 : gzv / n2;
      const float gy = (gxr ? gxr[d] : 0.f) + gcat;
      tile[d * PREP_LD + px] = big1 ? (gy - y * doty) / n1 : gy / n1;
    }
  }
  __syncthreads();
  float* dst = a.gemb + (int64_t)b * a.D * a.HW;
  for (int d = warp; d < a.D; d += PREP_THREADS / 32) {
#pragma unroll
    for (int u = 0; u < PREP_TPX / 32; ++u) {
      const int px = 32 * u + lane;
      if (p0 + px < a.HW) dst[(int64_t)d * a.HW + p0 + px] = tile[d * PREP_LD + px];
    }
  }
}

}  // namespace hsg

using namespace hsg;

extern "C" {

int hsg_normalize_f32(const float* x, float* y, int64_t rows, int dim, void* stream) {
  HSG_REQUIRE(rows >= 0 && dim > 0, HSG_E_INVALID, "normalize: bad shape rows=%lld dim=%d", (long long)rows, dim);
  if (rows == 0) return HSG_OK;
  HSG_REQUIRE(x && y, HSG_E_INVALID, "normalize: null pointer");
  const int rpb = 8;
  normalize_kernel<<<(unsigned)ceil_div64(rows, rpb), rpb * 32, 0, (cudaStream_t)stream>>>(x, y, rows, dim);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int hsg_normalize_bwd_f32(const float* x, const float* gy, float* gx, int64_t rows, int dim, void* stream) {
  HSG_REQUIRE(rows >= 0 && dim > 0, HSG_E_INVALID, "normalize_bwd: bad shape");
  if (rows == 0) return HSG_OK;
  HSG_REQUIRE(x && gy && gx, HSG_E_INVALID, "normalize_bwd: null pointer");
  const int rpb = 8;
  normalize_bwd_kernel<<<(unsigned)ceil_div64(rows, rpb), rpb * 32, 0, (cudaStream_t)stream>>>(x, gy, gx, rows, dim);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int hsg_make_half_copy_f32(const float* x, int64_t rows, int dim, int d16, void* xh_out,
                           float* xerr_out, void* stream) {
  HSG_REQUIRE(rows >= 0 && dim > 0 && d16 > 0 && d16 <= dim && dim - d16 <= HSG_XH_MAX_TRAILING,
              HSG_E_INVALID, "half_copy: bad shape (at most %d trailing features)", HSG_XH_MAX_TRAILING);
  if (rows == 0) return HSG_OK;
  HSG_REQUIRE(x && xh_out, HSG_E_INVALID, "half_copy: null pointer");
  const int rpb = 8;
  half_copy_kernel<<<(unsigned)ceil_div64(rows, rpb), rpb * 32, 0, (cudaStream_t)stream>>>(
      x, rows, dim, d16, (__half*)xh_out, xerr_out);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

size_t hsg_prep_workspace_bytes(int B, int H, int W) {
  const int64_t tpi = ceil_div64((int64_t)H * W, PREP_TPX);
  return align_up((size_t)B * tpi * sizeof(int32_t), 256) + align_up((size_t)B * tpi * sizeof(int64_t), 256) + 512;
}

static int prep_impl(const float* emb_nchw, int B, int D, int H, int W,
                     const float* loc, int L, int64_t loc_image_stride,
                     const int64_t* labels, int use_ignore, int64_t ignore_index,
                     const int64_t* init_clusters, int64_t init_image_stride,
                     int64_t batch_index_base,
                     float* x_out, float* xloc_out, void* xh_out, float* xerr_out,
                     int64_t* labels_out, int64_t* clusters_out, int64_t* batch_out,
                     int64_t* pixel_out, int64_t* seg_offsets, void* workspace, size_t workspace_bytes,
                     float* run_sums, int32_t* run_cluster, int32_t* run_count, int32_t* run_overflow,
                     void* stream) {
  HSG_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, HSG_E_INVALID, "prep: bad shape B=%d D=%d H=%d W=%d", B, D, H, W);
  HSG_REQUIRE(L >= 0 && L <= 32, HSG_E_UNSUPPORTED, "prep: %d local-feature channels (max 32)", L);
  HSG_REQUIRE(emb_nchw && (loc || L == 0) && init_clusters && x_out && xloc_out && labels_out &&
              clusters_out && batch_out && seg_offsets, HSG_E_INVALID, "prep: null pointer");
  HSG_REQUIRE(!use_ignore || labels, HSG_E_INVALID, "prep: ignore_index without labels");
  HSG_REQUIRE(!xh_out || L <= HSG_XH_MAX_TRAILING, HSG_E_UNSUPPORTED,
              "prep: the fp16 side copy holds at most %d local-feature channels", HSG_XH_MAX_TRAILING);
  HSG_REQUIRE((int64_t)H * W < (1ll << 31) && B <= 65535, HSG_E_UNSUPPORTED, "prep: image too large");
  const size_t smem = (size_t)(D + (run_sums ? L : 0)) * PREP_LD * sizeof(float);
  HSG_REQUIRE(smem <= 200 * 1024, HSG_E_UNSUPPORTED, "prep: embedding_dim %d too large", D);
  HSG_REQUIRE(!run_sums || (run_cluster && run_count && run_overflow), HSG_E_INVALID, "prep: run sums need all four buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  const int tpi = (HW + PREP_TPX - 1) / PREP_TPX;

  ProfRange prof(PROF_PREP, st);
  const int64_t* tile_base = nullptr;
  if (use_ignore) {
    HSG_REQUIRE(workspace && workspace_bytes >= hsg_prep_workspace_bytes(B, H, W), HSG_E_WORKSPACE,
                "prep: workspace too small");
    Carver c(workspace);
    int32_t* cnt = c.take<int32_t>((size_t)B * tpi);
    int64_t* base = c.take<int64_t>((size_t)B * tpi);
    const int64_t n_tiles = (int64_t)B * tpi;
    prep_count_kernel<<<(unsigned)ceil_div64(n_tiles, 8), 256, 0, st>>>(
        labels, ignore_index, HW, tpi, n_tiles, cnt);
    HSG_LAUNCH_CHECK();
    prep_scan_kernel<<<1, 1024, 0, st>>>(cnt, n_tiles, tpi, B, base, seg_offsets);
    HSG_LAUNCH_CHECK();
    tile_base = base;
  } else {
    fill_dense_offsets_kernel<<<(B + 256) / 256, 256, 0, st>>>(seg_offsets, B, HW);
    HSG_LAUNCH_CHECK();
  }

  PrepArgs a;
  a.emb = emb_nchw; a.B = B; a.D = D; a.HW = HW; a.loc = loc; a.L = L;
  a.loc_image_stride = loc_image_stride; a.labels = labels; a.use_ignore = use_ignore;
  a.ignore_index = ignore_index; a.init = init_clusters; a.init_image_stride = init_image_stride;
  a.batch_base = batch_index_base; a.x = x_out; a.xloc = xloc_out; a.xh = (__half*)xh_out;
  a.xerr = xerr_out; a.labels_out = labels_out; a.clusters_out = clusters_out;
  a.batch_out = batch_out; a.pixel_out = pixel_out; a.tile_base = tile_base; a.tiles_per_image = tpi;
  a.run_sums = run_sums; a.run_cluster = run_cluster; a.run_count = run_count; a.run_overflow = run_overflow;
  if (run_sums) HSG_CUDA(cudaMemsetAsync(run_overflow, 0, sizeof(int32_t), (cudaStream_t)stream));
  if (smem > 48 * 1024)
    HSG_CUDA(cudaFuncSetAttribute(prep_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prep_main_kernel<<<dim3(tpi, B), PREP_THREADS, smem, st>>>(a);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int hsg_prep_f32(const float* emb_nchw, int B, int D, int H, int W,
                 const float* loc, int L, int64_t loc_image_stride,
                 const int64_t* labels, int use_ignore, int64_t ignore_index,
                 const int64_t* init_clusters, int64_t init_image_stride,
                 int64_t batch_index_base,
                 float* x_out, float* xloc_out, void* xh_out, float* xerr_out,
                 int64_t* labels_out, int64_t* clusters_out, int64_t* batch_out,
                 int64_t* pixel_out, int64_t* seg_offsets, void* workspace, size_t workspace_bytes,
                 void* stream) {
  return prep_impl(emb_nchw, B, D, H, W, loc, L, loc_image_stride, labels, use_ignore, ignore_index, init_clusters,
                   init_image_stride, batch_index_base, x_out, xloc_out, xh_out, xerr_out, labels_out, clusters_out,
                   batch_out, pixel_out, seg_offsets, workspace, workspace_bytes, nullptr, nullptr, nullptr, nullptr, stream);
}

int64_t hsg_prep_runs_per_image(int H, int W) {
  return ceil_div64((int64_t)H * W, PREP_TPX) * HSG_PREP_RUNS;
}

int hsg_prep_sums_f32(const float* emb_nchw, int B, int D, int H, int W,
                      const float* loc, int L, int64_t loc_image_stride,
                      const int64_t* labels, int use_ignore, int64_t ignore_index,
                      const int64_t* init_clusters, int64_t init_image_stride,
                      int64_t batch_index_base,
                      float* x_out, float* xloc_out, void* xh_out, float* xerr_out,
                      int64_t* labels_out, int64_t* clusters_out, int64_t* batch_out,
                      int64_t* pixel_out, int64_t* seg_offsets, void* workspace, size_t workspace_bytes,
                      float* run_sums, int32_t* run_cluster, int32_t* run_count, int32_t* run_overflow,
                      void* stream) {
  HSG_REQUIRE(run_sums && run_cluster && run_count && run_overflow, HSG_E_INVALID, "prep_sums: null run buffers");
  return prep_impl(emb_nchw, B, D, H, W, loc, L, loc_image_stride, labels, use_ignore, ignore_index, init_clusters,
                   init_image_stride, batch_index_base, x_out, xloc_out, xh_out, xerr_out, labels_out, clusters_out,
                   batch_out, pixel_out, seg_offsets, workspace, workspace_bytes, run_sums, run_cluster, run_count,
                   run_overflow, stream);
}

int hsg_prep_bwd_f32(const float* emb_nchw, int B, int D, int H, int W, const float* x, const float* loc, int L,
                     int64_t loc_image_stride, const int64_t* row_of_pixel, const float* grad_x,
                     const float* grad_xloc, float* grad_emb_nchw, void* stream) {
  HSG_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && L >= 0 && L <= 32, HSG_E_INVALID, "prep_bwd: bad shape");
  HSG_REQUIRE(emb_nchw && x && grad_emb_nchw && (loc || L == 0), HSG_E_INVALID, "prep_bwd: null pointer");
  HSG_REQUIRE((int64_t)H * W < (1ll << 31) && B <= 65535, HSG_E_UNSUPPORTED, "prep_bwd: image too large");
  const size_t smem = (size_t)D * PREP_LD * sizeof(float);
  HSG_REQUIRE(smem <= 200 * 1024, HSG_E_UNSUPPORTED, "prep_bwd: embedding_dim %d too large", D);
  PrepBwdArgs a;
  a.emb = emb_nchw; a.B = B; a.D = D; a.HW = H * W; a.L = L; a.x = x; a.loc = loc; a.loc_image_stride = loc_image_stride;
  a.row_of_pixel = row_of_pixel; a.gx = grad_x; a.gz = grad_xloc; a.gemb = grad_emb_nchw;
  if (smem > 48 * 1024)
    HSG_CUDA(cudaFuncSetAttribute(prep_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tpi = (a.HW + PREP_TPX - 1) / PREP_TPX;
  prep_bwd_kernel<<<dim3(tpi, B), PREP_THREADS, smem, (cudaStream_t)stream>>>(a);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

}  // extern "C"
