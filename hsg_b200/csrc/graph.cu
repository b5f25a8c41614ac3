// k-NN affinity graph of the DMoN regulariser.
//
// Reference: affinity_matrix_as_attention (hsg/utils/graph/common.py:39-125): on the
// [B,n,n] kernel matrix A it masks padded nodes, removes self loops, then -- in a Python double
// loop over batch entries and segment labels with `unique`, `nonzero`, `masked_select`, `topk`
// (two host synchronisations per segment) -- keeps, for every row, only the k largest entries
// among the columns of each segment, and binarises.  Entry (i,j) survives iff fewer than
// k_seg = min(#valid nodes of j's segment, knn) valid columns of j's segment are strictly larger
// in row i (the reference zeroes `A < kth value`).  Here: one launch, one warp per row, no sort.
#include "common.cuh"

namespace hsg {

constexpr int KG_WARPS = 8;

__global__ void __launch_bounds__(KG_WARPS * 32) knn_adjacency_kernel(
    const float* __restrict__ A, const unsigned char* __restrict__ pad, const int64_t* __restrict__ seg,
    int n, int knn, int remove_self, int binarize, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* rowv = sm;                                         // [KG_WARPS][n] masked row values
  int* lab = reinterpret_cast<int*>(rowv + KG_WARPS * n);  // [n] segment label (dense int) or -1 when padded
  int* segcnt = lab + n;                                    // [n] valid nodes in the node's segment
  __shared__ int nvalid_s;
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned char* padb = pad ? pad + (int64_t)b * n : nullptr;
  const int64_t* segb = seg ? seg + (int64_t)b * n : nullptr;
  if (threadIdx.x == 0) nvalid_s = 0;
  __syncthreads();
  int local_valid = 0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {       // lab[j]: first node carrying j's label, -1 when j is padded
    const bool v = !(padb && padb[j]);
    int rep = 0;
    if (v && segb) {
      const int64_t lj = segb[j];
      rep = j;
      for (int t = 0; t < j; ++t)
        if (segb[t] == lj) { rep = t; break; }
    }
    lab[j] = v ? rep : -1;
    local_valid += v;
  }
  atomicAdd(&nvalid_s, local_valid);
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {       // size of j's segment (valid nodes with the same label)
    int c = 0;
    const int lj = lab[j];
    if (lj >= 0)
      for (int t = 0; t < n; ++t) c += lab[t] == lj;
    segcnt[j] = c;
  }
  __syncthreads();
  const bool drop_self = remove_self && nvalid_s > 1;
  const int i = blockIdx.x * KG_WARPS + warp;
  if (i >= n) return;
  const bool pad_i = lab[i] < 0;
  float* rv = rowv + warp * n;
  const float* arow = A + ((int64_t)b * n + i) * n;
  for (int j = lane; j < n; j += 32) {
    float a = arow[j];
    if (pad_i || lab[j] < 0) a = 0.f;                       // common.py:83-85
    if (drop_self && j == i) a = 0.f;                       // :88-97
    rv[j] = a;
  }
  __syncwarp();
  float* orow = out + ((int64_t)b * n + i) * n;
  for (int j = lane; j < n; j += 32) {
    float a = rv[j];
    if (knn > 0 && lab[j] >= 0) {                           // :100-121, columns of valid segments only
      const int lj = lab[j];
      int greater = 0;
      for (int t = 0; t < n; ++t) greater += lab[t] == lj && rv[t] > a;
      const int k = min(segcnt[j], knn);
      if (greater >= k) a = 0.f;                            // a < k-th largest of the segment
    }
    orow[j] = binarize ? (a > 0.f ? 1.f : 0.f) : a;         // :123-125
  }
}

}  // namespace hsg

using namespace hsg;

extern "C" int hsg_knn_adjacency_f32(const float* A, const unsigned char* padding_mask,
                                     const int64_t* segment_labels, int B, int n, int knn,
                                     int remove_self_loop, int binarize, float* out, void* stream) {
  HSG_REQUIRE(B > 0 && n > 0 && B <= 65535, HSG_E_INVALID, "knn_adjacency: bad shape B=%d n=%d", B, n);
  HSG_REQUIRE(A && out, HSG_E_INVALID, "knn_adjacency: null pointer");
  const size_t smem = (size_t)KG_WARPS * n * sizeof(float) + 2 * (size_t)n * sizeof(int);
  HSG_REQUIRE(smem <= 200 * 1024, HSG_E_UNSUPPORTED, "knn_adjacency: %d nodes (shared memory)", n);
  if (smem > 48 * 1024)
    HSG_CUDA(cudaFuncSetAttribute(knn_adjacency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((n + KG_WARPS - 1) / KG_WARPS, B);
  knn_adjacency_kernel<<<grid, KG_WARPS * 32, smem, (cudaStream_t)stream>>>(A, padding_mask, segment_labels, n, knn,
                                                                           remove_self_loop, binarize, out);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}
