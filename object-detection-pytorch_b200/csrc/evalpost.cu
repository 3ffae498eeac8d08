// Eval post-processing right after DetectOut:
// lib/utils/evaluate_utils.py:63-70 (rescale by the image size), :127-139 / :175-190
// (convert_ssd_result: append image / class / coco id, keep rows with score > 0, reorder columns)
// and :193-203 (EvalCOCO.post_proc: x2,y2 -> w,h and the COCO result column order).
//
// The reference builds three broadcast [B,C,K,1] index tensors, concatenates them to the detections,
// masked_select's the [B,C,K,8] tensor and permutes columns.  Here: 2 launches over the [B,C,K,5]
// tensor that DetectOut just wrote (20.7 MB at SSD512-COCO, L2 resident):
//   compact_count_kernel  one warp per (image, class) segment counts its rows with score > 0; the
//                         last CTA turns the B*C counts into exclusive offsets (+ total)
//   compact_write_kernel  one warp per segment writes its rows, in k order, at its offset
// Row order = (image, class, k) row-major = the order masked_select produces.
#include "ops.h"
#include "ssdbox_dev.cuh"

namespace ssdbox {

constexpr int kCompactThreads = 1024;

struct CompactArgs {
  const float* det;       // [B,C,K,5] (score, x1, y1, x2, y2)
  const float* extra;     // [B,2] (h, w) nullable: no rescale
  const float* image_ids; // [B] nullable (COCO ids as fp32, like torch.Tensor(self.dataset.ids[...]))
  int B, C, K, mode;
  int32_t* cnt;           // [B*C]
  int32_t* offsets;       // [B*C+1]
  uint32_t* ticket;       // [1] zero before the launch, zero again after it
  float* out;
  int32_t* total;         // [1]
  long long capacity;     // rows `out` can hold
};

__global__ void __launch_bounds__(kCompactThreads) compact_count_kernel(CompactArgs a) {
  __shared__ int s_scan[33];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nseg = a.B * a.C;
  const int seg = blockIdx.x * (kCompactThreads / 32) + warp;
  if (seg < nseg) {
    const float* d = a.det + (size_t)seg * a.K * 5;
    int n = 0;
    for (int k = lane; k < a.K; k += 32) n += d[(size_t)k * 5] > 0.0f ? 1 : 0;     // evaluate_utils.py:131 det[...,0].gt(0.)
    n = warp_sum(n);
    if (lane == 0) a.cnt[seg] = n;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA: exclusive scan of the segment counts in segment order
  int running = 0;
  for (int base = 0; base < nseg; base += kCompactThreads) {
    int i = base + threadIdx.x;
    int v = i < nseg ? __ldcg(&a.cnt[i]) : 0;
    int total;
    int ex = block_exclusive_scan(v, s_scan, &total);
    if (i < nseg) a.offsets[i] = running + ex;
    running += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    a.offsets[nseg] = running;
    *a.total = running;
    *a.ticket = 0u;
  }
}

__global__ void __launch_bounds__(kCompactThreads) compact_write_kernel(CompactArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nseg = a.B * a.C;
  const int seg = blockIdx.x * (kCompactThreads / 32) + warp;
  if (seg >= nseg) return;
  const int n = a.cnt[seg];
  if (n == 0) return;
  const int b = seg / a.C, c = seg - b * a.C;
  float h = 1.0f, w = 1.0f;
  if (a.extra) {
    h = a.extra[(size_t)b * 2];            // evaluate_utils.py:63-64
    w = a.extra[(size_t)b * 2 + 1];
  }
  const float img = a.image_ids ? a.image_ids[b] : 0.0f;
  const int ncol = a.mode == 1 ? 8 : 7;
  const float* d = a.det + (size_t)seg * a.K * 5;
  long long row = a.offsets[seg];
  for (int k0 = 0; k0 < a.K; k0 += 32) {
    const int k = k0 + lane;
    float s = 0.f, x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
    if (k < a.K) {
      const float* r = d + (size_t)k * 5;
      s = r[0];
      x1 = r[1]; y1 = r[2]; x2 = r[3]; y2 = r[4];
    }
    const bool on = k < a.K && s > 0.0f;
    const uint32_t m = __ballot_sync(SSDBOX_FULL_MASK, on);
    if (on) {
      const long long dst = row + __popc(m & ((1u << lane) - 1u));
      if (dst < a.capacity) {
        if (a.extra) {                     // :65-68  det[...,1] *= w ; [3] *= w ; [2] *= h ; [4] *= h
          x1 = __fmul_rn(x1, w); x2 = __fmul_rn(x2, w);
          y1 = __fmul_rn(y1, h); y2 = __fmul_rn(y2, h);
        }
        float* o = a.out + dst * ncol;
        if (a.mode == 2) {                 // :193-199  cocoid, x1, y1, w, h, score, cls
          o[0] = img; o[1] = x1; o[2] = y1; o[3] = __fsub_rn(x2, x1); o[4] = __fsub_rn(y2, y1); o[5] = s; o[6] = (float)c;
        } else {                           // :135 / :188  xmin, ymin, xmax, ymax, score, image, cls(, cocoid)
          o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2; o[4] = s; o[5] = (float)b; o[6] = (float)c;
          if (a.mode == 1) o[7] = img;
        }
      }
    }
    row += __popc(m);
  }
}

}  // namespace ssdbox

using namespace ssdbox;

extern "C" int ssdbox_detections_compact(const float* det, int32_t B, int32_t C, int32_t K, const float* extra,
                                         const float* image_ids, int32_t mode, float* out, int64_t capacity_rows,
                                         int32_t* total, int32_t* seg_offsets, void* ws, size_t ws_bytes,
                                         ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(B >= 0 && C >= 1 && K >= 1, SSDBOX_EINVAL, "compact: bad shape");
  SSDBOX_REQUIRE(mode >= 0 && mode <= 2, SSDBOX_EINVAL, "compact: mode must be 0 (VOC), 1 (COCO convert) or 2 (COCO result rows)");
  SSDBOX_REQUIRE((long long)B * C < (1ll << 31) && (long long)B * C * K < (1ll << 31), SSDBOX_ESHAPE, "compact: B*C*K must be < 2^31");
  SSDBOX_REQUIRE(total, SSDBOX_EINVAL, "compact: null total");
  SSDBOX_REQUIRE(mode == 0 || image_ids, SSDBOX_EINVAL, "compact: the COCO modes need image_ids");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B == 0) {
    SSDBOX_CUDA(cudaMemsetAsync(total, 0, sizeof(int32_t), st));
    return SSDBOX_OK;
  }
  SSDBOX_REQUIRE(det && out && ws && capacity_rows >= 0, SSDBOX_EINVAL, "compact: null pointer");
  SSDBOX_REQUIRE(ws_bytes >= compact_ws_bytes(B, C), SSDBOX_EWORKSPACE, "compact: workspace too small");
  Carver cv(ws);
  CompactArgs a{};
  a.det = det; a.extra = extra; a.image_ids = image_ids;
  a.B = B; a.C = C; a.K = K; a.mode = mode;
  a.cnt = cv.take<int32_t>((size_t)B * C);
  int32_t* own_offsets = cv.take<int32_t>((size_t)B * C + 1);
  a.ticket = cv.take<uint32_t>(1);
  a.offsets = seg_offsets ? seg_offsets : own_offsets;
  a.out = out; a.total = total; a.capacity = capacity_rows;
  SSDBOX_CUDA(cudaMemsetAsync(a.ticket, 0, sizeof(uint32_t), st));
  const int per = kCompactThreads / 32;
  const int grid = (B * C + per - 1) / per;
  compact_count_kernel<<<grid, kCompactThreads, 0, st>>>(a);
  SSDBOX_LAUNCH_OK("compact_count_kernel");
  compact_write_kernel<<<grid, kCompactThreads, 0, st>>>(a);
  SSDBOX_LAUNCH_OK("compact_write_kernel");
  return SSDBOX_OK;
}
