// K2: dense relabel of (batch, cluster, label) triples -- the tail of
// segment_by_kmeans (hsg/utils/segsort/common.py:397-405) and
// prepare_prototype_labels (:192-218).
//
// The reference runs two sort-based torch.unique(return_inverse) calls (each a
// host sync).  The ids it produces are the ranks of the distinct triples in
// lexicographic order, so a presence table over the (small) triple space plus a
// prefix scan gives the same ids with no sort and no sync.
#include "common.cuh"

namespace hsg {

__device__ __forceinline__ int lower_bound_i64(const int64_t* a, int n, int64_t v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

struct RelabelArgs {
  const int64_t* batch;
  const int64_t* cluster;
  const int64_t* label;
  int64_t N;
  int64_t batch_base;
  int B, kmax;
  const int64_t* label_values;
  int nl;
  int32_t* flags;   // [T]
  int32_t* rank;    // [T]
  int64_t T;
};

__device__ __forceinline__ int64_t triple_index(const RelabelArgs& a, int64_t i) {
  int64_t b = a.batch[i] - a.batch_base;
  b = b < 0 ? 0 : (b >= a.B ? a.B - 1 : b);
  int64_t k = a.cluster[i];
  k = k < 0 ? 0 : (k >= a.kmax ? a.kmax - 1 : k);
  int lr = lower_bound_i64(a.label_values, a.nl, a.label[i]);
  if (lr >= a.nl) lr = a.nl - 1;
  return (b * a.kmax + k) * a.nl + lr;
}

__global__ void relabel_mark_kernel(const RelabelArgs a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.N; i += stride)
    a.flags[triple_index(a, i)] = 1;
}

__global__ void __launch_bounds__(1024) relabel_scan_kernel(const RelabelArgs a, int64_t* n_protos) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < a.T; base += 1024) {
    const int64_t t = base + threadIdx.x;
    const int v = t < a.T ? a.flags[t] : 0;
    const int incl = warp_scan_incl(v, lane);
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) warp_tot[lane] = warp_scan_incl(warp_tot[lane], lane);
    __syncthreads();
    if (t < a.T) a.rank[t] = carry + (warp ? warp_tot[warp - 1] : 0) + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_protos = carry;
}

__global__ void relabel_assign_kernel(const RelabelArgs a, int64_t* __restrict__ ids) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.N; i += stride)
    ids[i] = a.rank[triple_index(a, i)];
}

__global__ void relabel_describe_kernel(const RelabelArgs a, int64_t* __restrict__ proto_label,
                                        int64_t* __restrict__ proto_batch, int64_t* __restrict__ proto_cluster) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.T; t += stride) {
    if (!a.flags[t]) continue;
    const int r = a.rank[t];
    const int lr = (int)(t % a.nl);
    const int64_t bk = t / a.nl;
    if (proto_label) proto_label[r] = a.label_values[lr];
    if (proto_batch) proto_batch[r] = a.batch_base + bk / a.kmax;
    if (proto_cluster) proto_cluster[r] = bk % a.kmax;
  }
}

}  // namespace hsg

using namespace hsg;

extern "C" {

size_t hsg_relabel_workspace_bytes(int B, int kmax, int64_t n_label_values) {
  const int64_t T = (int64_t)B * kmax * (n_label_values > 0 ? n_label_values : 1);
  return 2 * align_up((size_t)T * sizeof(int32_t), 256) + 512;
}

int hsg_relabel_i64(const int64_t* batch, const int64_t* cluster, const int64_t* label, int64_t N,
                    int64_t batch_base, int B, int kmax, const int64_t* label_values,
                    int64_t n_label_values, int64_t* ids_out, int64_t* proto_label_out,
                    int64_t* proto_batch_out, int64_t* proto_cluster_out, int64_t* n_protos_out,
                    void* workspace, size_t workspace_bytes, void* stream) {
  HSG_REQUIRE(N >= 0 && B > 0 && kmax > 0 && n_label_values > 0, HSG_E_INVALID, "relabel: bad shape");
  const int64_t T = (int64_t)B * kmax * n_label_values;
  HSG_REQUIRE(T < (1ll << 28) && n_label_values < (1ll << 24), HSG_E_UNSUPPORTED,
              "relabel: triple space %lld too large (B=%d kmax=%d labels=%lld)", (long long)T, B, kmax,
              (long long)n_label_values);
  HSG_REQUIRE(n_protos_out && label_values && (N == 0 || (batch && cluster && label && ids_out)),
              HSG_E_INVALID, "relabel: null pointer");
  HSG_REQUIRE(workspace && workspace_bytes >= hsg_relabel_workspace_bytes(B, kmax, n_label_values),
              HSG_E_WORKSPACE, "relabel: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(workspace);
  RelabelArgs a;
  a.batch = batch; a.cluster = cluster; a.label = label; a.N = N; a.batch_base = batch_base;
  a.B = B; a.kmax = kmax; a.label_values = label_values; a.nl = (int)n_label_values; a.T = T;
  a.flags = c.take<int32_t>(T);
  a.rank = c.take<int32_t>(T);
  ProfRange prof(PROF_RELABEL, st);
  HSG_CUDA(cudaMemsetAsync(a.flags, 0, sizeof(int32_t) * T, st));
  const int blocks = num_sms() * 8;
  if (N > 0) {
    relabel_mark_kernel<<<blocks, 256, 0, st>>>(a);
    HSG_LAUNCH_CHECK();
  }
  relabel_scan_kernel<<<1, 1024, 0, st>>>(a, n_protos_out);
  HSG_LAUNCH_CHECK();
  if (N > 0) {
    relabel_assign_kernel<<<blocks, 256, 0, st>>>(a, ids_out);
    HSG_LAUNCH_CHECK();
  }
  if (proto_label_out || proto_batch_out || proto_cluster_out) {
    relabel_describe_kernel<<<blocks, 256, 0, st>>>(a, proto_label_out, proto_batch_out, proto_cluster_out);
    HSG_LAUNCH_CHECK();
  }
  return HSG_OK;
}

}  // extern "C"
