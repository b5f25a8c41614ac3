// Prototype exchange without host reads (the all-gather of prototypes before the loss,
// hsg/models/utils.py:127-217).  Every rank packs a fixed-capacity record
//   [count:int64 | pad:int64 | prototypes cap x d | prototypes_with_loc cap x d2 | sem, inst, batch: 3 x cap int64]
// with its device-side prototype count in the header, the host all-gathers the records (one collective), and
// every rank unpacks them into rank-major compacted arrays of world x cap rows plus the device-side total -- which
// the NCE kernels take as their prototype count (hsg_nce_fwd_counted_f32).  Two launches around the collective,
// no size exchange, no synchronisation.
#include "common.cuh"

namespace hsg {

__host__ __device__ inline size_t xr_float_off() { return 16; }
__host__ __device__ inline size_t xr_int_off(int64_t cap, int d, int d2) {
  return (16 + sizeof(float) * (size_t)cap * (d + d2) + 7) & ~(size_t)7;
}
__host__ __device__ inline size_t xr_bytes(int64_t cap, int d, int d2) { return xr_int_off(cap, d, d2) + 3 * 8 * (size_t)cap; }

__global__ void __launch_bounds__(256) exchange_pack_kernel(
    const float* __restrict__ protos, const float* __restrict__ protos_loc, const int64_t* __restrict__ sem,
    const int64_t* __restrict__ inst, const int64_t* __restrict__ batch, const int64_t* __restrict__ count,
    int64_t cap, int d, int d2, unsigned char* __restrict__ record) {
  const int64_t n = min(cap, max((int64_t)0, *count));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    reinterpret_cast<int64_t*>(record)[0] = n;
    reinterpret_cast<int64_t*>(record)[1] = 0;
  }
  float* f = reinterpret_cast<float*>(record + xr_float_off());
  int64_t* li = reinterpret_cast<int64_t*>(record + xr_int_off(cap, d, d2));
  const int64_t nf1 = cap * d, nf2 = cap * d2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nf1 + nf2 + 3 * cap; i += stride) {
    if (i < nf1) f[i] = i / d < n ? protos[i] : 0.f;
    else if (i < nf1 + nf2) { const int64_t k = i - nf1; f[i] = k / d2 < n ? protos_loc[k] : 0.f; }
    else {
      const int64_t k = i - nf1 - nf2, which = k / cap, r = k % cap;
      const int64_t* src = which == 0 ? sem : which == 1 ? inst : batch;
      li[k] = r < n ? src[r] : -1;
    }
  }
}

// one warp per output row j of the compacted rank-major arrays
__global__ void __launch_bounds__(256) exchange_unpack_kernel(
    const unsigned char* __restrict__ gathered, int world, int rank, int64_t cap, int d, int d2,
    float* __restrict__ protos, float* __restrict__ protos_loc, int64_t* __restrict__ sem, int64_t* __restrict__ inst,
    int64_t* __restrict__ batch, int64_t* __restrict__ total_out, int64_t* __restrict__ offset_out) {
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const size_t rec = xr_bytes(cap, d, d2);
  int64_t start = 0, total = 0, before_rank = 0;
  int src_rank = -1;
  int64_t src_row = 0;
  for (int r = 0; r < world; ++r) {
    const int64_t n = min(cap, max((int64_t)0, *reinterpret_cast<const int64_t*>(gathered + (size_t)r * rec)));
    if (r == rank) before_rank = total;
    if (src_rank < 0 && j < total + n) { src_rank = r; src_row = j - total; start = total; }
    total += n;
  }
  (void)start;
  if (j == 0 && lane == 0) { *total_out = total; *offset_out = before_rank; }
  if (j >= (int64_t)world * cap) return;
  if (src_rank >= 0) {
    const unsigned char* base = gathered + (size_t)src_rank * rec;
    const float* f = reinterpret_cast<const float*>(base + xr_float_off());
    const int64_t* li = reinterpret_cast<const int64_t*>(base + xr_int_off(cap, d, d2));
    for (int k = lane; k < d; k += 32) protos[j * d + k] = f[src_row * d + k];
    for (int k = lane; k < d2; k += 32) protos_loc[j * d2 + k] = f[cap * d + src_row * d2 + k];
    if (lane == 0) { sem[j] = li[src_row]; inst[j] = li[cap + src_row]; batch[j] = li[2 * cap + src_row]; }
  } else {
    for (int k = lane; k < d; k += 32) protos[j * d + k] = 0.f;
    for (int k = lane; k < d2; k += 32) protos_loc[j * d2 + k] = 0.f;
    if (lane == 0) { sem[j] = -1; inst[j] = -1; batch[j] = -1; }
  }
}

}  // namespace hsg

using namespace hsg;

extern "C" {

size_t hsg_exchange_record_bytes(int64_t capacity, int dim, int dim_loc) { return xr_bytes(capacity, dim, dim_loc); }

int hsg_exchange_pack(const float* prototypes, const float* prototypes_with_loc, const int64_t* sem, const int64_t* inst,
                      const int64_t* batch, const int64_t* num_prototypes_dev, int64_t capacity, int dim, int dim_loc,
                      void* record, void* stream) {
  HSG_REQUIRE(capacity > 0 && dim > 0 && dim_loc > 0, HSG_E_INVALID, "exchange_pack: bad shape");
  HSG_REQUIRE(prototypes && prototypes_with_loc && sem && inst && batch && num_prototypes_dev && record, HSG_E_INVALID,
              "exchange_pack: null pointer");
  const int64_t elems = capacity * (dim + dim_loc + 3);
  const int64_t blocks = ceil_div64(elems, 256);
  const unsigned grid = (unsigned)(blocks < 4096 ? blocks : 4096);
  exchange_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(prototypes, prototypes_with_loc, sem, inst, batch,
                                                               num_prototypes_dev, capacity, dim, dim_loc,
                                                               static_cast<unsigned char*>(record));
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

int hsg_exchange_unpack(const void* gathered, int world, int rank, int64_t capacity, int dim, int dim_loc,
                        float* prototypes_out, float* prototypes_with_loc_out, int64_t* sem_out, int64_t* inst_out,
                        int64_t* batch_out, int64_t* total_out, int64_t* offset_out, void* stream) {
  HSG_REQUIRE(world > 0 && rank >= 0 && rank < world && capacity > 0 && dim > 0 && dim_loc > 0, HSG_E_INVALID,
              "exchange_unpack: bad shape");
  HSG_REQUIRE(gathered && prototypes_out && prototypes_with_loc_out && sem_out && inst_out && batch_out && total_out &&
              offset_out, HSG_E_INVALID, "exchange_unpack: null pointer");
  const int64_t rows = (int64_t)world * capacity;
  exchange_unpack_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, (cudaStream_t)stream>>>(
      static_cast<const unsigned char*>(gathered), world, rank, capacity, dim, dim_loc, prototypes_out,
      prototypes_with_loc_out, sem_out, inst_out, batch_out, total_out, offset_out);
  HSG_LAUNCH_CHECK();
  return HSG_OK;
}

}  // extern "C"
