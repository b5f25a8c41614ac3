// top_k_ranking (hsg/utils/segsort/eval.py:9-52): the k best prototypes of every query row by inner product.
//
// The reference materialises the [N,P] affinity matrix (torch.mm) and fully sorts every row (argsort) to read
// k columns -- every training step, on the prototypes against themselves (hsg/models/predictions/hsg.py:113-118).
// Here: one kernel, fp32 register-tiled 64x64x32 products (exact fp32 like the reference's mm, no [N,P] matrix
// in HBM) and a running sorted top-k (k <= 8) per row that only looks at a block's 64 values once.
// Ties keep the lower prototype index first.
#include "common.cuh"

#include <float.h>

namespace hsg {

constexpr int TK_TP = 64;      // query rows per CTA
constexpr int TK_TK = 64;      // prototypes per block
constexpr int TK_DC = 32;      // feature chunk
constexpr int TK_THREADS = 256;
constexpr int TK_MAXK = 8;

__global__ void __launch_bounds__(TK_THREADS) topk_affinity_kernel(const float* __restrict__ e, int64_t N,
                                                                   const float* __restrict__ p, int64_t M, int dim, int k,
                                                                   int64_t* __restrict__ out_idx, float* __restrict__ out_val) {
  __shared__ float Xs[TK_TP][TK_DC + 1];
  __shared__ float Cs[TK_TK][TK_DC + 1];
  __shared__ float Ss[TK_TP][TK_TK + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t r0 = (int64_t)blockIdx.x * TK_TP;
  const int nr = (int)min((int64_t)TK_TP, N - r0);

  float tv[TK_MAXK];
  int ti[TK_MAXK];
#pragma unroll
  for (int j = 0; j < TK_MAXK; ++j) { tv[j] = -FLT_MAX; ti[j] = 0x7fffffff; }

  for (int64_t kb = 0; kb < M; kb += TK_TK) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int d0 = 0; d0 < dim; d0 += TK_DC) {
      __syncthreads();
#pragma unroll
      for (int r = 0; r < (TK_TP * TK_DC) / TK_THREADS; ++r) {
        const int idx = tid + TK_THREADS * r;
        const int row = idx >> 5, dd = idx & 31;
        const int d = d0 + dd;
        Xs[row][dd] = (row < nr && d < dim) ? e[(r0 + row) * dim + d] : 0.f;
        const int64_t c = kb + row;
        Cs[row][dd] = (c < M && d < dim) ? p[c * dim + d] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int dd = 0; dd < TK_DC; ++dd) {
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
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Ss[ty + 16 * i][tx + 16 * j] = (kb + tx + 16 * j < M) ? acc[i][j] : -FLT_MAX;
    __syncthreads();
    if (tid < nr) {
      const int nc = (int)min((int64_t)TK_TK, M - kb);
      for (int c = 0; c < nc; ++c) {
        const float v = Ss[tid][c];
        if (v > tv[TK_MAXK - 1]) {                       // strict: an equal later value stays behind
          float cv = v;
          int ci = (int)(kb + c);
#pragma unroll
          for (int j = 0; j < TK_MAXK; ++j) {
            if (cv > tv[j]) {
              const float sv = tv[j]; const int si = ti[j];
              tv[j] = cv; ti[j] = ci; cv = sv; ci = si;
            }
          }
        }
      }
    }
  }
  if (tid < nr) {
#pragma unroll
    for (int j = 0; j < TK_MAXK; ++j) {
      if (j < k) {
        out_idx[(r0 + tid) * k + j] = ti[j] == 0x7fffffff ? 0 : ti[j];
        if (out_val) out_val[(r0 + tid) * k + j] = tv[j];
      }
    }
  }
}

}  // namespace hsg

using namespace hsg;

extern "C" int hsg_topk_affinity_f32(const float* embeddings, int64_t N, const float* prototypes, int64_t M, int dim,
                                     int k, int64_t* indices_out, float* values_out, void* stream) {
  HSG_REQUIRE(N >= 0 && M > 0 && dim > 0, HSG_E_INVALID, "topk: bad shape N=%lld M=%lld dim=%d", (long long)N, (long long)M, dim);
  HSG_REQUIRE(k >= 1 && k <= TK_MAXK && k <= M, HSG_E_UNSUPPORTED, "topk: k=%d (1..%d, <= prototypes)", k, TK_MAXK);
  HSG_REQUIRE(M < (1ll << 31), HSG_E_UNSUPPORTED, "topk: too many prototypes");
  if (N == 0) return HSG_OK;
  HSG_REQUIRE(embeddings && prototypes && indices_out, HSG_E_INVALID, "topk: null pointer");
  topk_affinity_kernel<<<(unsigned)ceil_div64(N, TK_TP), TK_THREADS, 0, (cudaStream_t)stream>>>(
      embeddings, N, prototypes, M, dim, k, indices_out, values_out);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}
