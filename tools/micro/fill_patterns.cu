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
// (d) TMA bulk stores (cp.async.bulk.global.shared::cta) of a shared all-zero tile: lane 0 of every warp issues the
// stores of its tiles; wait_each = 1 waits until the source has been read before the next one (what a kernel that
// patches its tile between stores must do), 0 issues back to back.  per_cta = 1: every CTA owns one contiguous range.
__global__ void bulk_fill(char* d, size_t bytes, int tile_bytes, int wait_each, int per_cta) {
  extern __shared__ __align__(128) char z[];
  for (int i = threadIdx.x; i < tile_bytes / 4; i += blockDim.x) reinterpret_cast<float*>(z)[i] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x & 31) return;
  const size_t nt = (bytes + tile_bytes - 1) / tile_bytes;
  const int wpc = blockDim.x >> 5, warp = threadIdx.x >> 5;
  size_t t0, t1, step;
  if (per_cta) {
    const size_t per = (nt + gridDim.x - 1) / gridDim.x;
    t0 = blockIdx.x * per + warp; t1 = (blockIdx.x + 1) * per < nt ? (blockIdx.x + 1) * per : nt; step = wpc;
  } else {
    t0 = (size_t)blockIdx.x * wpc + warp; t1 = nt; step = (size_t)gridDim.x * wpc;
  }
  const unsigned src = (unsigned)__cvta_generic_to_shared(z);
  for (size_t t = t0; t < t1; t += step) {
    const size_t left = bytes - t * tile_bytes;
    const unsigned n = left < (size_t)tile_bytes ? (unsigned)left : (unsigned)tile_bytes;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + t * tile_bytes), "r"(src), "r"(n) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (wait_each) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// (e) the backward's shape: 20 warps per SM, 10 KB bulk stores with wait.read, plus small READS mixed into the write
// stream: reads & 1: 64 B of a 2-byte-per-row array per tile (the selection flags); reads & 2: the same bytes but as
// one 512-byte request every 8 tiles; reads & 4: one 324-byte row of a second 509 MB array per tile (a logits row).
// The loaded values only feed a store that never happens.
__global__ void bulk_fill_reads(char* d, size_t bytes, int tile_bytes, const short* sel, const float* rowsrc, int reads, int* sink) {
  extern __shared__ __align__(128) char z[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int wpc = blockDim.x >> 5;
  char* mine = z + (size_t)(warp % 20) * tile_bytes;
  const size_t nt = (bytes + tile_bytes - 1) / tile_bytes;
  const size_t per = (nt + gridDim.x - 1) / gridDim.x;
  const size_t t1 = (blockIdx.x + 1) * per < nt ? (blockIdx.x + 1) * per : nt;
  if ((reads & 64) && warp >= 20) {        // reader warps: the same reads, from warps that store nothing
    int acc2 = 0;
    for (size_t t = blockIdx.x * per + (warp - 20); t < t1; t += wpc - 20) {
      acc2 += sel[t * 32 + lane];
      const float* r = rowsrc + (t * 32 + (t * 7 & 31)) * 81;
      acc2 += (int)(r[lane] + r[lane + 32] + (lane < 17 ? r[lane + 64] : 0.f));
    }
    if (acc2 == 0x7fffffff) *sink = acc2;
    return;
  }
  for (int i = lane; i < tile_bytes / 4; i += 32) reinterpret_cast<float*>(mine)[i] = (reads & 32) ? __int_as_float(0x3f000000 + i * 2654435 + warp * 977) : 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (reads & 64) wpc = 20;
  const unsigned src = (unsigned)__cvta_generic_to_shared(mine);
  int acc = 0, pre0 = 0, pre1 = 0;
  float prer0 = 0.f, prer1 = 0.f;
  int it = 0;
  for (size_t t = blockIdx.x * per + warp; t < t1; t += wpc, ++it) {
    if (reads & 8) {              // prefetched: the flags / the row requested two tiles ago are consumed now
      acc += pre1 + (int)prer1;
      pre1 = pre0; prer1 = prer0;
      const size_t tn = t + 2 * wpc < t1 ? t + 2 * wpc : t;
      pre0 = sel[tn * 32 + lane];
      if (reads & 16) { const float* r = rowsrc + (tn * 32 + (tn * 7 & 31)) * 81; prer0 = r[lane] + r[lane + 32] + (lane < 17 ? r[lane + 64] : 0.f); }
    }
    if (reads & 128) {            // the same two reads per tile, but from a 64 KB / 2.6 MB window: L2 hits after the first touch
      const size_t tw = t & 1023;
      acc += sel[tw * 32 + lane];
      const float* r = rowsrc + (tw * 32 + (tw * 7 & 31)) * 81; acc += (int)(r[lane] + r[lane + 32] + (lane < 17 ? r[lane + 64] : 0.f));
    }
    if (reads & 1) acc += sel[t * 32 + lane];
    if ((reads & 2) && (it & 7) == 0) { const int4 v = reinterpret_cast<const int4*>(sel)[(t * 32) / 8 + lane]; acc += v.x + v.w; }
    if (reads & 4) { const float* r = rowsrc + (t * 32 + (t * 7 & 31)) * 81; acc += (int)(r[lane] + r[lane + 32] + (lane < 17 ? r[lane + 64] : 0.f)); }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    if (acc == 0x7fffffff) reinterpret_cast<int*>(mine)[lane] = acc;
    if (lane == 0) {
      const size_t left = bytes - t * tile_bytes;
      const unsigned n = left < (size_t)tile_bytes ? (unsigned)left : (unsigned)tile_bytes;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + t * tile_bytes), "r"(src), "r"(n) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (acc == 0x7fffffff) *sink = acc;
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
  {
    short* sel; CK(cudaMalloc(&sel, (size_t)64 * 24564 * 2 + 4096)); CK(cudaMemset(sel, 0, (size_t)64 * 24564 * 2 + 4096));
    float* rowsrc; CK(cudaMalloc(&rowsrc, bytes + 4096)); CK(cudaMemset(rowsrc, 0, bytes + 4096));
    int* sink; CK(cudaMalloc(&sink, 4));
    CK(cudaFuncSetAttribute(bulk_fill_reads, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 10368));
    for (int reads : {0, 5, 128}) {
      char nm[128]; snprintf(nm, 128, "20 warps x 10368 B bulk stores + reads mode %d", reads);
      report(nm, timeit([&] { bulk_fill_reads<<<148, (reads & 64) ? 768 : 640, 20 * 10368>>>((char*)d, bytes, 10368, sel, rowsrc, reads, sink); }));
    }
  }
  CK(cudaFuncSetAttribute(bulk_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
  for (int per_cta : {1}) for (int wait_each : {1}) for (int tile : {10368}) for (int warps : {4, 20}) {
    char nm[128]; snprintf(nm, 128, "bulk stores of %d B, %d warps/SM, %s, %s", tile, warps, wait_each ? "wait.read each" : "back to back", per_cta ? "range per CTA" : "round robin");
    report(nm, timeit([&] { bulk_fill<<<148, warps * 32, tile>>>((char*)d, bytes, tile, wait_each, per_cta); }));
  }
  return 0;
}
