// micro-benchmark: how fast can 534 MB be zero-filled on B200, by access pattern?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fill_patterns fill_patterns.cu && ./fill_patterns
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// (a) flat grid-stride, every warp-instruction writes 512 contiguous bytes, the grid sweeps memory front to back
__global__ void flat(float4* d, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  const float4 z = make_float4(0, 0, 0, 0);
  for (; i + 3 * stride < n4; i += 4 * stride) { __stcs(d + i, z); __stcs(d + i + stride, z); __stcs(d + i + 2 * stride, z); __stcs(d + i + 3 * stride, z); }
  for (; i < n4; i += stride) __stcs(d + i, z);
}
// (b) one warp owns a contiguous tile of `tile4` float4 (10368 B = 648 float4 for 32 rows x 81 classes), tiles dealt round-robin
__global__ void tiles(float4* d, size_t n4, int tile4) {
  const int lane = threadIdx.x & 31;
  const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t nt = (n4 + tile4 - 1) / tile4;
  const float4 z = make_float4(0, 0, 0, 0);
  for (size_t t = w; t < nt; t += nw) {
    float4* p = d + t * tile4;
    size_t left = n4 - t * tile4;
    int m = left < (size_t)tile4 ? (int)left : tile4;
    for (int k = lane; k < m; k += 32) __stcs(p + k, z);
  }
}
// (c) like (b) but a warp takes `run` consecutive tiles (contiguous run per warp)
__global__ void runs(float4* d, size_t n4, int tile4, int run) {
  const int lane = threadIdx.x & 31;
  const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t nt = (n4 + tile4 - 1) / tile4;
  const float4 z = make_float4(0, 0, 0, 0);
  for (size_t t0 = w * run; t0 < nt; t0 += nw * run)
    for (size_t t = t0; t < t0 + run && t < nt; ++t) {
      float4* p = d + t * tile4;
      size_t left = n4 - t * tile4;
      int m = left < (size_t)tile4 ? (int)left : tile4;
      for (int k = lane; k < m; k += 32) __stcs(p + k, z);
    }
}
template <class F> float timeit(F f, int reps = 10) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e9;
  for (int i = 0; i < reps; ++i) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms; }
  return best * 1e3f;
}
int main() {
  const size_t bytes = (size_t)64 * 24564 * (81 * 4);   // grad_conf of SSD512-COCO B=64
  const size_t n4 = bytes / 16;
  float4* d; CK(cudaMalloc(&d, bytes));
  float* other; CK(cudaMalloc(&other, 600u << 20));
  auto report = [&](const char* name, float us) { printf("%-60s %8.1f us  %6.2f TB/s\n", name, us, bytes / us * 1e-6); };
  report("cudaMemsetAsync", timeit([&] { CK(cudaMemsetAsync(d, 0, bytes)); }));
  for (int cps : {4, 8, 16}) for (int th : {256, 512}) {
    char nm[128]; snprintf(nm, 128, "flat  grid=148*%d block=%d", cps, th);
    report(nm, timeit([&] { flat<<<148 * cps, th>>>(d, n4); }));
  }
  for (int warps_per_sm : {16, 20, 32, 64}) {
    char nm[128]; snprintf(nm, 128, "tiles 648 float4 / warp, %d warps/SM", warps_per_sm);
    report(nm, timeit([&] { tiles<<<148 * warps_per_sm / 4, 128>>>(d, n4, 648); }));
  }
  for (int t4 : {162, 324, 1296, 2592}) {
    char nm[128]; snprintf(nm, 128, "tiles %d float4 / warp, 20 warps/SM", t4);
    report(nm, timeit([&] { tiles<<<148 * 5, 128>>>(d, n4, t4); }));
  }
  for (int run : {4, 17}) {
    char nm[128]; snprintf(nm, 128, "runs of %d tiles x 648 float4 per warp, 20 warps/SM", run);
    report(nm, timeit([&] { runs<<<148 * 5, 128>>>(d, n4, 648, run); }));
  }
  return 0;
}
