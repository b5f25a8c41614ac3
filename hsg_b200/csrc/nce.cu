// K4: pixel-to-prototype NCE ("SegSort+") loss, forward and backward.
//
// Reference: _calculate_log_likelihood (hsg/utils/segsort/loss.py:15-82)
// materialises S = exp(c E P^T) as an [N,P] fp32 matrix plus five more [N,P]
// temporaries, once per label set (Hsg.losses evaluates 3 label sets on the same
// (E,P), hsg/models/predictions/hsg.py:105,130,149).  Here S is produced tile
// by tile, never stored, and every label set is reduced in the same pass.
//
// This file is the exact-fp32 CUDA-core version (round 1).  Per pixel and set:
//   own = S[i,inst_i]
//   A   = sum_{psem_j == sem_i, j != inst_i} S_ij
//   pos = fl(fl(A + own) - own) if psem[inst_i] == sem_i else A - own
//         (restates the reference's "sum over the class, then subtract own",
//          loss.py:64-66, including its fp32 flush of positives far below own)
//   num = pos > 0 ? pos : own  ('segsort+'), num = own otherwise
//   den = sum_{psem_j != sem_i} S_ij + num ;  l = -log(num/den)
// Backward: G_ij = c * sum_s w_si * dl_si/dS_ij * S_ij,
//   dE = G P, dP = G^T E, evaluated over pixel chunks so G stays small.
#include "common.cuh"

#include <float.h>

namespace hsg {

constexpr int NC_T = 64;        // tile edge (pixels and prototypes)
constexpr int NC_DC = 32;       // feature chunk
constexpr int NC_THREADS = 256;
constexpr int NC_MAX_SETS = 4;
constexpr int NC_STATS = 4;     // num, den, own, flags per (set,pixel)

// 64x64 tile of dot products, thread (ty,tx) owns rows ty+16i, cols tx+16j
__device__ __forceinline__ void tile_dots(const float* __restrict__ a, int64_t a_rows, int64_t a0,
                                          const float* __restrict__ b, int64_t b_rows, int64_t b0,
                                          int dim, float (*As)[NC_DC + 1], float (*Bs)[NC_DC + 1],
                                          float acc[4][4]) {
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int d0 = 0; d0 < dim; d0 += NC_DC) {
    __syncthreads();
#pragma unroll
    for (int r = 0; r < (NC_T * NC_DC) / NC_THREADS; ++r) {
      const int idx = tid + NC_THREADS * r;
      const int row = idx >> 5, dd = idx & 31;
      const int d = d0 + dd;
      As[row][dd] = (a0 + row < a_rows && d < dim) ? a[(a0 + row) * dim + d] : 0.f;
      Bs[row][dd] = (b0 + row < b_rows && d < dim) ? b[(b0 + row) * dim + d] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int dd = 0; dd < NC_DC; ++dd) {
      float xa[4], xb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xa[i] = As[ty + 16 * i][dd];
#pragma unroll
      for (int j = 0; j < 4; ++j) xb[j] = Bs[tx + 16 * j][dd];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], xb[j], acc[i][j]);
    }
  }
}

struct NceArgs {
  const float* e;
  const float* p;
  int64_t N, P;
  const int64_t* P_dev;  // optional device-side count of valid prototypes (<= P, the capacity of the arrays)
  int dim;
  const int64_t* inst;
  const int64_t* sem;    // [n_sets,N]
  const int64_t* psem;   // [n_sets,P]
  int n_sets;
  int plus[NC_MAX_SETS];
  float conc;
};

template <int NS>
__global__ void __launch_bounds__(NC_THREADS) nce_fwd_kernel(const NceArgs a, float* __restrict__ per_pixel,
                                                             float* __restrict__ stats) {
  __shared__ float As[NC_T][NC_DC + 1];
  __shared__ float Bs[NC_T][NC_DC + 1];
  __shared__ int64_t psem_s[NS][NC_T];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = (int64_t)blockIdx.x * NC_T;

  int64_t my_inst[4], my_sem[NS][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t pix = i0 + ty + 16 * i;
    my_inst[i] = pix < a.N ? a.inst[pix] : -1;
#pragma unroll
    for (int s = 0; s < NS; ++s) my_sem[s][i] = pix < a.N ? a.sem[(int64_t)s * a.N + pix] : 0;
  }
  float own[4], pos[NS][4], neg[NS][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    own[i] = 0.f;
#pragma unroll
    for (int s = 0; s < NS; ++s) { pos[s][i] = 0.f; neg[s][i] = 0.f; }
  }

  const int64_t P_eff = a.P_dev ? max((int64_t)0, min(a.P, *a.P_dev)) : a.P;
  for (int64_t j0 = 0; j0 < P_eff; j0 += NC_T) {
    float acc[4][4];
    tile_dots(a.e, a.N, i0, a.p, a.P, j0, a.dim, As, Bs, acc);   // begins with __syncthreads()
    if (tid < NC_T) {
#pragma unroll
      for (int s = 0; s < NS; ++s)
        psem_s[s][tid] = j0 + tid < a.P ? a.psem[(int64_t)s * a.P + j0 + tid] : INT64_MIN;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t pj = j0 + tx + 16 * j;
      if (pj >= P_eff) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float sv = expf(acc[i][j] * a.conc);
        const bool is_own = pj == my_inst[i];
        if (is_own) own[i] = sv;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const bool same = psem_s[s][tx + 16 * j] == my_sem[s][i];
          if (!same) neg[s][i] += sv;
          else if (!is_own) pos[s][i] += sv;
        }
      }
    }
  }
  // reduce over the 16 tx lanes (fixed xor tree -> deterministic)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      own[i] += __shfl_xor_sync(FULL, own[i], o);
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        pos[s][i] += __shfl_xor_sync(FULL, pos[s][i], o);
        neg[s][i] += __shfl_xor_sync(FULL, neg[s][i], o);
      }
    }
    const int64_t pix = i0 + ty + 16 * i;
    if (tx == 0 && pix < a.N) {
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int64_t inst = my_inst[i];
        const bool own_same = inst >= 0 && inst < P_eff && a.psem[(int64_t)s * a.P + inst] == my_sem[s][i];
        float num = own[i];
        float flags = own_same ? 2.f : 0.f;
        if (a.plus[s]) {
          const float ps = own_same ? __fsub_rn(__fadd_rn(pos[s][i], own[i]), own[i])
                                    : __fsub_rn(pos[s][i], own[i]);
          if (ps > 0.f) { num = ps; flags += 1.f; }
        }
        const float den = neg[s][i] + num;   // neg already holds own when psem[inst] != sem
        per_pixel[(int64_t)s * a.N + pix] = -logf(num / den);
        if (stats) {
          float* st = stats + ((int64_t)s * a.N + pix) * NC_STATS;
          st[0] = num; st[1] = den; st[2] = own[i]; st[3] = flags;
        }
      }
    }
  }
}

// G[i - i_begin, j] for a chunk of pixels
template <int NS>
__global__ void __launch_bounds__(NC_THREADS) nce_grad_kernel(const NceArgs a, const float* __restrict__ stats,
                                                              const float* __restrict__ w, int64_t i_begin,
                                                              int64_t i_end, float* __restrict__ G, int64_t ldg) {
  __shared__ float As[NC_T][NC_DC + 1];
  __shared__ float Bs[NC_T][NC_DC + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = i_begin + (int64_t)blockIdx.x * NC_T;
  const int64_t j0 = (int64_t)blockIdx.y * NC_T;
  float acc[4][4];
  tile_dots(a.e, i_end, i0, a.p, a.P, j0, a.dim, As, Bs, acc);
  const int64_t P_eff = a.P_dev ? max((int64_t)0, min(a.P, *a.P_dev)) : a.P;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t pix = i0 + ty + 16 * i;
    if (pix >= i_end) continue;
    const int64_t inst = a.inst[pix];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t pj = j0 + tx + 16 * j;
      if (pj >= a.P) continue;
      if (pj >= P_eff) { G[(pix - i_begin) * ldg + pj] = 0.f; continue; }
      const float sv = expf(acc[i][j] * a.conc);
      float coef = 0.f;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float* st = stats + ((int64_t)s * a.N + pix) * NC_STATS;
        const float num = st[0], den = st[1];
        const int fl = (int)st[3];
        const bool use = fl & 1, own_same = fl & 2;
        const bool same = a.psem[(int64_t)s * a.P + pj] == a.sem[(int64_t)s * a.N + pix];
        const bool is_own = pj == inst;
        float wij;
        if (use) wij = (same && !is_own ? 1.f : 0.f) - (is_own && !own_same ? 1.f : 0.f);
        else wij = is_own ? 1.f : 0.f;
        const float dl = (same ? 0.f : 1.f / den) + wij * (1.f / den - 1.f / num);
        coef = fmaf(w[(int64_t)s * a.N + pix], dl, coef);
      }
      G[(pix - i_begin) * ldg + pj] = a.conc * coef * sv;
    }
  }
}

// C[M,N] (+)= op(A) B ; A is [M,K] (TA=false) or [K,M] (TA=true), B [K,N], all row-major.
template <bool TA>
__global__ void __launch_bounds__(NC_THREADS) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                           float* __restrict__ C, int64_t M, int64_t N,
                                                           int64_t K, int accumulate) {
  __shared__ float As[16][NC_T + 1];   // [k][m]
  __shared__ float Bs[16][NC_T + 1];   // [k][n]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * NC_T, n0 = (int64_t)blockIdx.x * NC_T;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t k0 = 0; k0 < K; k0 += 16) {
    __syncthreads();
#pragma unroll
    for (int r = 0; r < (16 * NC_T) / NC_THREADS; ++r) {
      const int idx = tid + NC_THREADS * r;
      if (TA) {
        const int kk = idx >> 6, mm = idx & 63;          // A[k][m]: m contiguous
        As[kk][mm] = (k0 + kk < K && m0 + mm < M) ? A[(k0 + kk) * M + m0 + mm] : 0.f;
      } else {
        const int mm = idx >> 4, kk = idx & 15;          // A[m][k]: k contiguous
        As[kk][mm] = (k0 + kk < K && m0 + mm < M) ? A[(m0 + mm) * K + k0 + kk] : 0.f;
      }
      const int kk = idx >> 6, nn = idx & 63;
      Bs[kk][nn] = (k0 + kk < K && n0 + nn < N) ? B[(k0 + kk) * N + n0 + nn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float xa[4], xb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xa[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) xb[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], xb[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx + 16 * j;
      if (n >= N) continue;
      float* c = C + m * N + n;
      *c = accumulate ? *c + acc[i][j] : acc[i][j];
    }
  }
}

// tensor-core forward (nce_tc.cu)
bool nce_tc_supported(int64_t N, int64_t P, int dim, int n_sets);
size_t nce_tc_workspace_bytes(int64_t N, int64_t P, int dim, int n_sets);
int nce_fwd_tc(const float* e, const float* prototypes, int64_t N, int64_t P, int dim, const int64_t* inst,
               const int64_t* sem, const int64_t* psem, int n_sets, const int32_t* plus, float conc,
               float* per_pixel, float* stats, void* workspace, cudaStream_t st, const int64_t* P_dev);
int g_debug_flags = 0;   // tests: bit 0 keeps the NCE forward on the fp32 CUDA-core kernel, bit 2 the backward GEMMs,
                         // bit 3 the backward's G chunk; bit 6: backward chunks of 4096 pixels

// fp32-grade tensor-core GEMM on pre-split fp16 (hi|lo) operands (gemm_tc.cu)
bool gemm_tc_supported(int N, int K);
int gemm_tc_split(const __half* a2, const __half* b2, int64_t M, int N, int K, float* C, int64_t ldc,
                  const float* inv_scale, float alpha, bool accumulate, cudaStream_t st);
int split_rows(const float* src, int64_t ld, int64_t R, int C, int Kp, float mul, const float* dev_mul, __half* dst,
               cudaStream_t st);
int split_transpose(const float* src, int64_t ld, int64_t R, int C, int64_t Rp, float mul, const float* dev_mul,
                    __half* dst, cudaStream_t st);
// tensor-core G chunk (nce_tc.cu): one prepare per backward call, one launch per pixel chunk
size_t nce_grad_tc_host_state_bytes();
int nce_grad_tc_prepare(void* host_state, const float* e, const float* prototypes, int64_t N, int64_t P, int dim,
                        const int64_t* inst, const int64_t* sem, const int64_t* psem, int n_sets, const int32_t* plus,
                        float conc, void* workspace, cudaStream_t st, const int64_t* P_dev);
int nce_grad_tc(void* host_state, const float* stats, const float* w, float conc, int64_t i_begin, int64_t i_end,
                float* G, int64_t ldg, __half* G2, const float* gscale, cudaStream_t st);
const __half* nce_grad_tc_e2(void* host_state, float* scale);
// dP = G^T E straight from the row-major fp16 copies (MN-major tcgen05 operands, split K) -- gemm_tc.cu
bool gemm_tn_tc_supported(int N);
int gemm_tn_ksplit(int64_t M, int64_t n_k);
int gemm_tn_tc_split(const __half* a2, int64_t a_rows, int a_half, const __half* b2, int64_t b_rows, int b_half,
                     int64_t b_row0, int64_t M, int N, int64_t n_k, int ksplit, float* part, const float* inv_scale,
                     cudaStream_t st);
int sum_splits(const float* part, int64_t n, int ksplit, float* out, cudaStream_t st);
int absmax(const float* x, int64_t n, float* out, cudaStream_t st);

// operand scales of the backward GEMMs: G is multiplied by sc[0] = 2^10 / (conc * n_sets * max|w|) before the
// fp16 split (|G_ij| <= conc * sum_s |w_si|), E and P by 16; sc[1] = 1 / (16 sc[0]) undoes both
__global__ void nce_bwd_scales_kernel(const float* __restrict__ wmax, float conc, int n_sets, float e2_scale,
                                      float* __restrict__ sc) {
  const float bound = fabsf(conc) * n_sets * wmax[0];
  const float sa = bound > 0.f ? 1024.f / bound : 1.f;
  sc[0] = sa;
  sc[1] = 1.f / (16.f * sa);
  sc[2] = 1.f / (e2_scale * sa);     // dP from the forward's fp16 copy of E (scaled by e2_scale)
}

struct BwdTcPlan {
  bool on;                 // GEMMs on tensor cores
  bool g_on;               // G chunk on tensor cores as well
  int64_t ldg;             // row stride of the G chunk
  void* fwd_ws;            // fp16 operand copies / int32 labels of the pair kernels
  int64_t Pp, chunkp;
  float* G;
  __half* G2;
  __half* Gt2;
  __half* Pt2;
  __half* Et2;
  float* scal;     // [0] max|w|, [1] G scale, [2] output scale (E, P scaled by 16), [3] output scale of the direct dP product
  bool direct;     // G written as fp16 (hi | lo) by the G kernel; dP through the MN-major product (no fp32 G, no transposes)
  float* part;     // [8][P, dim] K-split partial sums of dP
};

static void bwd_carve(Carver& c, BwdTcPlan& b, int64_t N, int64_t P, int dim, int n_sets, int64_t chunk, bool tc,
                      bool g_tc = true) {
  b.on = tc;
  b.Pp = (P + 63) / 64 * 64;
  b.chunkp = (chunk + 63) / 64 * 64;
  b.ldg = tc ? b.Pp : P;
  b.g_on = tc && g_tc && nce_tc_supported(N, P, dim, n_sets);
  b.direct = b.g_on && gemm_tn_tc_supported(dim);
  b.G = b.direct ? nullptr : c.take<float>((size_t)chunk * b.ldg);
  b.fwd_ws = b.g_on ? c.take<char>(nce_tc_workspace_bytes(N, P, dim, n_sets)) : nullptr;
  if (!tc) return;
  b.G2 = c.take<__half>((size_t)chunk * 2 * b.Pp);
  b.Pt2 = c.take<__half>((size_t)dim * 2 * b.Pp);
  b.scal = c.take<float>(4);
  if (b.direct) {
    b.Gt2 = b.Et2 = nullptr;
    b.part = c.take<float>((size_t)8 * P * dim);
    return;
  }
  b.part = nullptr;
  b.Gt2 = c.take<__half>((size_t)P * 2 * b.chunkp);
  b.Et2 = c.take<__half>((size_t)dim * 2 * b.chunkp);
}

static bool bwd_tc_shape(int64_t P, int dim) {
  return dim >= 16 && dim <= 256 && dim % 16 == 0 && P >= 1;
}

static int64_t nce_chunk_pixels(int64_t N, int64_t P) {
  // G chunk of up to 2 GB (fp32; its two fp16 splits are as large again): the chunk's rows are the M of the dE GEMM
  // (128-row tiles, one CTA each) and the K of the dP GEMM, so a chunk has to hold a few hundred tiles to fill
  // 148 SMs -- at 256 MB and P = 12288 it held 43 (r1: 197 ms per 1M pixels)
  int64_t c = (int64_t)(512ll << 20) / (P > 0 ? P : 1);
  if (g_debug_flags & 64) c = 4096;                       // tests: several chunks (and a ragged last one) at small sizes
  const int64_t wave = (int64_t)num_sms() * 128;          // rows of one wave of 128-row dE tiles
  if (c > wave) c = c / wave * wave;
  c = c / NC_T * NC_T;
  if (c < NC_T) c = NC_T;
  if (c > N) c = ceil_div64(N, NC_T) * NC_T;
  return c;
}

static int fill_args(NceArgs& a, const float* e, const float* p, int64_t N, int64_t P, int dim,
                     const int64_t* inst, const int64_t* sem, const int64_t* psem, int n_sets,
                     const int32_t* plus, float conc, const int64_t* P_dev) {
  HSG_REQUIRE(N >= 0 && P > 0 && dim > 0, HSG_E_INVALID, "nce: bad shape N=%lld P=%lld dim=%d", (long long)N, (long long)P, dim);
  HSG_REQUIRE(n_sets >= 1 && n_sets <= NC_MAX_SETS, HSG_E_UNSUPPORTED, "nce: %d label sets (1..%d)", n_sets, NC_MAX_SETS);
  HSG_REQUIRE(N == 0 || (e && p && inst && sem && psem && plus), HSG_E_INVALID, "nce: null pointer");
  a.e = e; a.p = p; a.N = N; a.P = P; a.P_dev = P_dev; a.dim = dim; a.inst = inst; a.sem = sem; a.psem = psem;
  a.n_sets = n_sets; a.conc = conc;
  for (int s = 0; s < NC_MAX_SETS; ++s) a.plus[s] = s < n_sets ? plus[s] : 0;
  return HSG_OK;
}

}  // namespace hsg

using namespace hsg;

extern "C" {

size_t hsg_nce_workspace_bytes(int64_t N, int64_t P, int dim, int n_sets) {
  size_t need = 0;
  for (int g_tc = 0; g_tc < 2; ++g_tc) {                                         // backward: one G chunk (+ its fp16 splits),
    Carver cb(nullptr);                                                          // with or without the tensor-core G kernel (tests)
    BwdTcPlan bp;
    bwd_carve(cb, bp, N, P, dim, n_sets, nce_chunk_pixels(N, P), bwd_tc_shape(P, dim), g_tc != 0);
    if (cb.used() + 1024 > need) need = cb.used() + 1024;
  }
  if (nce_tc_supported(N, P, dim, n_sets)) {
    const size_t tc = nce_tc_workspace_bytes(N, P, dim, n_sets);                 // forward: fp16 (hi,lo) copies
    if (tc > need) need = tc;
  }
  return need;
}

int hsg_debug_set_flags(int flags) {
  g_debug_flags = flags;
  return HSG_OK;
}

int hsg_nce_fwd_f32(const float* e, const float* prototypes, int64_t N, int64_t P, int dim,
                    const int64_t* inst, const int64_t* sem, const int64_t* psem, int n_sets,
                    const int32_t* group_plus_host, float concentration, float* per_pixel_out,
                    float* stats_out, void* workspace, size_t workspace_bytes, void* stream) {
  return hsg_nce_fwd_counted_f32(e, prototypes, N, P, nullptr, dim, inst, sem, psem, n_sets, group_plus_host, concentration,
                                 per_pixel_out, stats_out, workspace, workspace_bytes, stream);
}

int hsg_nce_fwd_counted_f32(const float* e, const float* prototypes, int64_t N, int64_t P,
                            const int64_t* num_prototypes_dev, int dim,
                            const int64_t* inst, const int64_t* sem, const int64_t* psem, int n_sets,
                            const int32_t* group_plus_host, float concentration, float* per_pixel_out,
                            float* stats_out, void* workspace, size_t workspace_bytes, void* stream) {
  NceArgs a;
  int rc = fill_args(a, e, prototypes, N, P, dim, inst, sem, psem, n_sets, group_plus_host, concentration, num_prototypes_dev);
  if (rc) return rc;
  if (N == 0) return HSG_OK;
  HSG_REQUIRE(per_pixel_out, HSG_E_INVALID, "nce_fwd: null output");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)ceil_div64(N, NC_T);
  ProfRange prof(PROF_NCE_FWD, st);
  if (!(g_debug_flags & 1) && nce_tc_supported(N, P, dim, n_sets) && workspace &&
      workspace_bytes >= nce_tc_workspace_bytes(N, P, dim, n_sets))
    return nce_fwd_tc(e, prototypes, N, P, dim, inst, sem, psem, n_sets, group_plus_host, concentration,
                      per_pixel_out, stats_out, workspace, st, num_prototypes_dev);
  switch (n_sets) {
    case 1: nce_fwd_kernel<1><<<grid, NC_THREADS, 0, st>>>(a, per_pixel_out, stats_out); break;
    case 2: nce_fwd_kernel<2><<<grid, NC_THREADS, 0, st>>>(a, per_pixel_out, stats_out); break;
    case 3: nce_fwd_kernel<3><<<grid, NC_THREADS, 0, st>>>(a, per_pixel_out, stats_out); break;
    default: nce_fwd_kernel<4><<<grid, NC_THREADS, 0, st>>>(a, per_pixel_out, stats_out); break;
  }
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int hsg_nce_bwd_f32(const float* e, const float* prototypes, int64_t N, int64_t P, int dim,
                    const int64_t* inst, const int64_t* sem, const int64_t* psem, int n_sets,
                    const int32_t* group_plus_host, float concentration, const float* stats,
                    const float* w, float* grad_e, float* grad_p, void* workspace,
                    size_t workspace_bytes, void* stream) {
  return hsg_nce_bwd_counted_f32(e, prototypes, N, P, nullptr, dim, inst, sem, psem, n_sets, group_plus_host, concentration,
                                 stats, w, grad_e, grad_p, workspace, workspace_bytes, stream);
}

int hsg_nce_bwd_counted_f32(const float* e, const float* prototypes, int64_t N, int64_t P,
                            const int64_t* num_prototypes_dev, int dim,
                            const int64_t* inst, const int64_t* sem, const int64_t* psem, int n_sets,
                            const int32_t* group_plus_host, float concentration, const float* stats,
                            const float* w, float* grad_e, float* grad_p, void* workspace,
                            size_t workspace_bytes, void* stream) {
  NceArgs a;
  int rc = fill_args(a, e, prototypes, N, P, dim, inst, sem, psem, n_sets, group_plus_host, concentration, num_prototypes_dev);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  HSG_REQUIRE(grad_p && (N == 0 || (grad_e && stats && w)), HSG_E_INVALID, "nce_bwd: null pointer");
  HSG_CUDA(cudaMemsetAsync(grad_p, 0, sizeof(float) * P * dim, st));
  if (N == 0) return HSG_OK;
  HSG_REQUIRE(workspace && workspace_bytes >= hsg_nce_workspace_bytes(N, P, dim, n_sets), HSG_E_WORKSPACE,
              "nce_bwd: workspace too small");
  ProfRange prof(PROF_NCE_BWD, st);
  const int64_t chunk = nce_chunk_pixels(N, P);
  Carver cw(workspace);
  BwdTcPlan b;
  bwd_carve(cw, b, N, P, dim, n_sets, chunk, bwd_tc_shape(P, dim) && !(g_debug_flags & 4), !(g_debug_flags & 8));
  alignas(64) unsigned char tc_state[1024];
  HSG_REQUIRE(nce_grad_tc_host_state_bytes() <= sizeof(tc_state), HSG_E_UNSUPPORTED, "nce_bwd: host state");
  float* G = b.G;
  const __half* e2 = nullptr;
  int ksplit = 1;
  if (b.on) {
    // dE = G P and dP = G^T E on the tensor cores (three fp16 passes each, gemm_tc.cu)
    if (b.g_on && (rc = nce_grad_tc_prepare(tc_state, e, prototypes, N, P, dim, inst, sem, psem, n_sets, group_plus_host,
                                            concentration, b.fwd_ws, st, num_prototypes_dev))) return rc;
    float e2_scale = 1.f;
    if (b.g_on) e2 = nce_grad_tc_e2(tc_state, &e2_scale);
    if ((rc = absmax(w, (int64_t)n_sets * N, b.scal, st))) return rc;
    nce_bwd_scales_kernel<<<1, 1, 0, st>>>(b.scal, concentration, n_sets, e2_scale, b.scal + 1);
    HSG_LAUNCH_CHECK();
    if ((rc = split_transpose(prototypes, dim, P, dim, b.Pp, 16.f, nullptr, b.Pt2, st))) return rc;
    if (b.direct) HSG_CUDA(cudaMemsetAsync(b.part, 0, sizeof(float) * 8 * P * dim, st));
  }
  for (int64_t i0 = 0; i0 < N; i0 += chunk) {
    const int64_t i1 = i0 + chunk < N ? i0 + chunk : N;
    const int64_t m = i1 - i0;
    if (b.direct) {
      // G chunk as fp16 (hi | lo) rows straight from the accumulator, then both products read that one copy:
      // dE[i0:i1] = G P (K-major A) and dP += G^T E[i0:i1] (MN-major A and B, K split over the SMs)
      if ((rc = nce_grad_tc(tc_state, stats, w, concentration, i0, i1, nullptr, b.ldg, b.G2, b.scal + 1, st))) return rc;
      if ((rc = gemm_tc_split(b.G2, b.Pt2, m, dim, (int)b.Pp, grad_e + i0 * dim, dim, b.scal + 2, 1.f, false, st))) return rc;
      const int64_t n_k = ceil_div64(m, 64);
      ksplit = gemm_tn_ksplit(P, ceil_div64(chunk < N ? chunk : N, 64));      // the same split for every chunk of the call
      if (ksplit > n_k) ksplit = (int)n_k;
      if ((rc = gemm_tn_tc_split(b.G2, m, (int)b.Pp, e2, N, dim, i0, P, dim, n_k, ksplit, b.part, b.scal + 3, st))) return rc;
      continue;
    }
    if (b.g_on) {
      if ((rc = nce_grad_tc(tc_state, stats, w, concentration, i0, i1, G, b.ldg, nullptr, nullptr, st))) return rc;
    } else {
      if (b.on && b.ldg != P) HSG_CUDA(cudaMemsetAsync(G, 0, sizeof(float) * m * b.ldg, st));   // zero pad columns
      dim3 gg((unsigned)ceil_div64(m, NC_T), (unsigned)ceil_div64(P, NC_T));
      switch (n_sets) {
        case 1: nce_grad_kernel<1><<<gg, NC_THREADS, 0, st>>>(a, stats, w, i0, i1, G, b.ldg); break;
        case 2: nce_grad_kernel<2><<<gg, NC_THREADS, 0, st>>>(a, stats, w, i0, i1, G, b.ldg); break;
        case 3: nce_grad_kernel<3><<<gg, NC_THREADS, 0, st>>>(a, stats, w, i0, i1, G, b.ldg); break;
        default: nce_grad_kernel<4><<<gg, NC_THREADS, 0, st>>>(a, stats, w, i0, i1, G, b.ldg); break;
      }
      HSG_LAUNCH_CHECK();
    }
    if (b.on) {
      const int64_t mp = (m + 63) / 64 * 64;
      // dE[i0:i1] = G P          ([m,P] x [P,dim])
      if ((rc = split_rows(G, b.ldg, m, (int)P, (int)b.Pp, 1.f, b.scal + 1, b.G2, st))) return rc;
      if ((rc = gemm_tc_split(b.G2, b.Pt2, m, dim, (int)b.Pp, grad_e + i0 * dim, dim, b.scal + 2, 1.f, false, st))) return rc;
      // dP += G^T E[i0:i1]       ([P,m] x [m,dim])
      if ((rc = split_transpose(G, b.ldg, m, (int)P, mp, 1.f, b.scal + 1, b.Gt2, st))) return rc;
      if ((rc = split_transpose(e + i0 * dim, dim, m, dim, mp, 16.f, nullptr, b.Et2, st))) return rc;
      if ((rc = gemm_tc_split(b.Gt2, b.Et2, P, dim, (int)mp, grad_p, dim, b.scal + 2, 1.f, true, st))) return rc;
      continue;
    }
    // dE[i0:i1] = G P          ([m,P] x [P,dim])
    dim3 g1((unsigned)ceil_div64(dim, NC_T), (unsigned)ceil_div64(m, NC_T));
    sgemm_kernel<false><<<g1, NC_THREADS, 0, st>>>(G, prototypes, grad_e + i0 * dim, m, dim, P, 0);
    HSG_LAUNCH_CHECK();
    // dP += G^T E[i0:i1]       ([P,m] x [m,dim])
    dim3 g2((unsigned)ceil_div64(dim, NC_T), (unsigned)ceil_div64(P, NC_T));
    sgemm_kernel<true><<<g2, NC_THREADS, 0, st>>>(G, e + i0 * dim, grad_p, P, dim, m, 1);
    HSG_LAUNCH_CHECK();
  }
  if (b.direct) return sum_splits(b.part, P * dim, 8, grad_p, st);      // unused splits stay zero
  return HSG_OK;
}

}  // extern "C"
