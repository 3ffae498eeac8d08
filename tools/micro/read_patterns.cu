// micro-benchmark: how fast can 509 MB be READ once on B200, by access pattern?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_patterns read_patterns.cu && ./read_patterns
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// (a) flat grid-stride 16-byte loads (ld.global.nc.L1::no_allocate), 4 in flight per thread
__global__ void flat(const float4* __restrict__ d, size_t n4, float* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 a = __ldcs(d + i), b = __ldcs(d + i + stride), c = __ldcs(d + i + 2 * stride), e = __ldcs(d + i + 3 * stride);
    acc += a.x + b.y + c.z + e.w;
  }
  for (; i < n4; i += stride) acc += __ldcs(d + i).x;
  if (acc == 123.456f) *out = acc;
}
// (b) every CTA owns a contiguous run and reads it front to back (the partition of the TMA ring), plain loads
__global__ void runs(const float4* __restrict__ d, size_t n4, float* out) {
  const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const size_t lo = blockIdx.x * per, hi = lo + per < n4 ? lo + per : n4;
  float acc = 0.f;
  size_t i = lo + threadIdx.x;
  for (; i + 3 * blockDim.x < hi; i += 4 * blockDim.x) {
    float4 a = __ldcs(d + i), b = __ldcs(d + i + blockDim.x), c = __ldcs(d + i + 2 * blockDim.x), e = __ldcs(d + i + 3 * blockDim.x);
    acc += a.x + b.y + c.z + e.w;
  }
  for (; i < hi; i += blockDim.x) acc += __ldcs(d + i).x;
  if (acc == 123.456f) *out = acc;
}
// (c) TMA bulk-copy ring: one producer lane streams the CTA's run through NS stages of `stage` bytes; consumers only release
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(64, 1) ring(const char* __restrict__ d, size_t bytes, int NS, int stage, int interleave, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 16;
  unsigned char* buf = smem + 256;
  const size_t tiles = (bytes + stage - 1) / stage;
  const size_t per = (tiles + gridDim.x - 1) / gridDim.x;
  size_t t0, tstep, n;
  if (interleave) { t0 = blockIdx.x; tstep = gridDim.x; n = t0 < tiles ? (tiles - t0 + gridDim.x - 1) / gridDim.x : 0; }
  else { t0 = blockIdx.x * per; tstep = 1; size_t t1 = t0 + per < tiles ? t0 + per : tiles; n = t1 > t0 ? t1 - t0 : 0; }
  if (threadIdx.x == 0) {
    for (int j = 0; j < NS; ++j) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[j])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[j])), "r"(1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto wait = [&](uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  };
  if (threadIdx.x == 0) {            // producer
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    for (size_t it = 0; it < n; ++it) {
      const int s = it % NS;
      if (it >= (size_t)NS) wait(&empty[s], ((it / NS) - 1) & 1);
      const size_t off = (t0 + it * tstep) * (size_t)stage;
      uint32_t nb = (uint32_t)(bytes - off < (size_t)stage ? bytes - off : stage);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(nb) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                   ::"r"(smem_u32(buf + (size_t)s * stage)), "l"(d + off), "r"(nb), "r"(smem_u32(&full[s])), "l"(policy) : "memory");
    }
  } else if (threadIdx.x == 32) {    // consumer: release at once
    float acc = 0.f;
    for (size_t it = 0; it < n; ++it) {
      const int s = it % NS;
      wait(&full[s], (it / NS) & 1);
      acc += reinterpret_cast<float*>(buf + (size_t)s * stage)[0];
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
    if (acc == 123.456f) *out = acc;
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
  const size_t bytes = (size_t)64 * 24564 * 81 * 4;   // conf of SSD512-COCO B=64: 509.4 MB
  const size_t n4 = bytes / 16;
  char* d; CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 0, bytes));
  float* out; CK(cudaMalloc(&out, 4));
  auto report = [&](const char* name, float us) { printf("%-70s %8.1f us  %6.2f TB/s\n", name, us, bytes / us * 1e-6); };
  for (int cps : {4, 8, 16}) for (int th : {256, 512}) {
    char nm[128]; snprintf(nm, 128, "flat loads  grid=148*%d block=%d", cps, th);
    report(nm, timeit([&] { flat<<<148 * cps, th>>>((const float4*)d, n4, out); }));
  }
  for (int g : {148, 296, 592}) {
    char nm[128]; snprintf(nm, 128, "contiguous run per CTA, plain loads, grid=%d block=512", g);
    report(nm, timeit([&] { runs<<<g, 512>>>((const float4*)d, n4, out); }));
  }
  CK(cudaFuncSetAttribute(ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  for (int inter : {0, 1}) for (int grid : {148, 132}) for (int stage : {41472, 20736, 82944}) for (int NS : {2, 3, 4, 5, 8}) {
    if ((size_t)NS * stage + 256 > 220 * 1024) continue;
    char nm[128]; snprintf(nm, 128, "TMA ring %s grid=%d stage=%d B x %d stages", inter ? "interleaved" : "runs       ", grid, stage, NS);
    report(nm, timeit([&] { ring<<<grid, 64, (size_t)NS * stage + 256>>>(d, bytes, NS, stage, inter, out); }));
  }
  return 0;
}
