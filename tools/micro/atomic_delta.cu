// Microbenchmark: throughput of the "exact delta M-step by atomics" idea.
// Each warp takes changed rows; per row it reads the fp32 row [258], converts to 2^-36 fixed point and issues
// red.global.add.u64 to sums[new][d] and sums[old][d].  sums = [bins, 258] int64 (bins = 48*256: 25 MB, L2 resident).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o atomic_delta atomic_delta.cu ; ./atomic_delta
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void delta_kernel(const float* __restrict__ x, int dim, const int* __restrict__ rows, const int* __restrict__ knew,
                             const int* __restrict__ kold, int n, unsigned long long* __restrict__ sums, int mode) {
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * (blockDim.x >> 5);
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
    const float* row = x + (size_t)rows[i] * dim;
    unsigned long long* a = sums + (size_t)knew[i] * dim;
    unsigned long long* b = sums + (size_t)kold[i] * dim;
    float v[9];
#pragma unroll
    for (int m = 0; m < 9; ++m) { const int d = lane + 32 * m; v[m] = d < dim ? row[d] : 0.f; }
    if (mode == 1) {          // reads only
      float s = 0; for (int m = 0; m < 9; ++m) s += v[m];
      if (s == 1234.5f) sums[0] = 1;
      continue;
    }
#pragma unroll
    for (int m = 0; m < 9; ++m) {
      const int d = lane + 32 * m;
      if (d < dim) {
        const long long q = __float2ll_rn(v[m] * 68719476736.f);
        atomicAdd(a + d, (unsigned long long)q);
        atomicAdd(b + d, (unsigned long long)(-q));
      }
    }
  }
}

int main() {
  const int dim = 258, bins = 48 * 256;
  const size_t N = 9633792;
  float* x; cudaMalloc(&x, N * dim * sizeof(float)); cudaMemset(x, 0, N * dim * sizeof(float));
  unsigned long long* sums; cudaMalloc(&sums, (size_t)bins * dim * 8); cudaMemset(sums, 0, (size_t)bins * dim * 8);
  for (double frac : {0.33, 0.14, 0.05, 0.025}) {
    const int n = (int)(N * frac);
    int *rows, *kn, *ko;
    cudaMallocManaged(&rows, n * 4); cudaMallocManaged(&kn, n * 4); cudaMallocManaged(&ko, n * 4);
    unsigned s = 12345;
    const int per = 200704;
    for (int i = 0; i < n; ++i) {
      const size_t r = (size_t)((double)i / n * N);          // increasing rows, spread evenly
      rows[i] = (int)r;
      const int img = (int)(r / per);
      s = s * 1664525u + 1013904223u; kn[i] = img * 256 + (s >> 24);
      s = s * 1664525u + 1013904223u; ko[i] = img * 256 + (s >> 24);
    }
    cudaMemPrefetchAsync(rows, n * 4, 0); cudaMemPrefetchAsync(kn, n * 4, 0); cudaMemPrefetchAsync(ko, n * 4, 0);
    cudaDeviceSynchronize();
    for (int mode = 0; mode < 2; ++mode) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      delta_kernel<<<148 * 8, 256>>>(x, dim, rows, kn, ko, n, sums, mode);
      cudaEventRecord(e0);
      for (int it = 0; it < 5; ++it) delta_kernel<<<148 * 8, 256>>>(x, dim, rows, kn, ko, n, sums, mode);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
      printf("changed fraction %.3f (%d rows) %s: %.3f ms  (%.1f GB/s of rows, %.1f G atomics/s)\n", frac, n,
             mode ? "reads only " : "reads+atomics", ms, n * dim * 4.0 / ms / 1e6, mode ? 0.0 : 2.0 * n * dim / ms / 1e6);
    }
    cudaFree(rows); cudaFree(kn); cudaFree(ko);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
