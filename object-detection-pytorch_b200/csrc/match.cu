// Ground-truth <-> prior matching for a whole batch: lib/layers/box_utils.py:92-133 (match),
// :51-70 (jaccard), :6-15 (point_form) and the python loop of multibox_loss.py:69-74.
//
// One CTA owns 1024 priors of one image.  The image's truths (+areas, +labels) sit in shared
// memory; every thread keeps 4 priors in registers and walks the truths once, so the [G,P] IoU
// matrix never exists in memory.  Per prior: running (max, first index) over truths.  Per truth:
// (max IoU, lowest prior index) as a packed 64-bit key  iou_bits<<32 | ~prior , reduced with two
// REDUX per warp, one shared-memory atomicMax per warp and one global atomicMax per CTA.
// The CTA that finishes an image last replays the reference's sequential forced assignment
// ("for j: best_truth_idx[best_prior_idx[j]] = j", last truth wins) for that image.
#include "ops.h"
#include "ssdbox_dev.cuh"

namespace ssdbox {

constexpr int kMatchThreads = 256;
constexpr int kMatchPPT = 4;
constexpr int kMatchTile = kMatchThreads * kMatchPPT;

__global__ void init_kernel(unsigned long long* best, size_t nbest, uint32_t* z0, size_t n0, uint32_t* z1,
                            size_t n1, uint32_t* z2, size_t n2) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = i; k < nbest; k += stride) best[k] = kBestInit;
  for (size_t k = i; k < n0; k += stride) z0[k] = 0u;
  for (size_t k = i; k < n1; k += stride) z1[k] = 0u;
  for (size_t k = i; k < n2; k += stride) z2[k] = 0u;
}

int launch_init(unsigned long long* best, size_t nbest, uint32_t* z0, size_t n0, uint32_t* z1, size_t n1,
                uint32_t* z2, size_t n2, cudaStream_t st) {
  size_t m = nbest;
  if (n0 > m) m = n0;
  if (n1 > m) m = n1;
  if (n2 > m) m = n2;
  if (m == 0) return SSDBOX_OK;
  int blocks = (int)((m + 255) / 256);
  if (blocks > 592) blocks = 592;
  SSDBOX_CARVE(init_kernel);
  {
    TimerScope ts__(KID_INIT, st);
    init_kernel<<<blocks, 256, 0, st>>>(best, nbest, z0, n0, z1, n1, z2, n2);
  }
  SSDBOX_LAUNCH_OK("init_kernel");
  return SSDBOX_OK;
}

#ifdef SSDBOX_PHASE_TIMING
__device__ long long g_mphase[32];
#define MPHASE(k) do { __syncthreads(); if (blockIdx.y == 1 && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) g_mphase[(blockIdx.x == 0 ? 0 : 16) + (k)] = clock64(); } while (0)
#else
#define MPHASE(k) do { } while (0)
#endif

__global__ void __launch_bounds__(kMatchThreads)
match_kernel(MatchArgs a, unsigned long long* __restrict__ gt_best, uint32_t* __restrict__ done,
             int16_t* __restrict__ lab, int16_t* __restrict__ tidx, float* __restrict__ overlap, int gpad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* s_best = reinterpret_cast<unsigned long long*>(smem_raw);
  float4* s_box = reinterpret_cast<float4*>(s_best + gpad);
  float* s_area = reinterpret_cast<float*>(s_box + gpad);
  int* s_lab = reinterpret_cast<int*>(s_area + gpad);
  __shared__ int s_last;

  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31;
  MPHASE(0);
  const int g0 = a.gt_offsets[b];
  int G = a.gt_offsets[b + 1] - g0;
  G = G < 0 ? 0 : (G > a.gmax ? a.gmax : G);

  for (int g = tid; g < G; g += kMatchThreads) {
    const float* r = a.gt + (size_t)(g0 + g) * 5;
    Box t;
    t.x1 = r[0]; t.y1 = r[1]; t.x2 = r[2]; t.y2 = r[3];
    s_box[g] = make_float4(t.x1, t.y1, t.x2, t.y2);
    s_area[g] = box_area(t);
    s_lab[g] = a.binarize ? 1 : (int)(r[4] + 1.0f);   // box_utils.py:129 labels + 1
    s_best[g] = kBestInit;
  }
  __syncthreads();
  MPHASE(1);

  const size_t img_off = (size_t)b * (size_t)a.P;
  const float* pri = a.priors + (size_t)b * (size_t)a.prior_stride;
  const float* anc = a.anchors_xyxy ? a.anchors_xyxy + (size_t)b * (size_t)a.prior_stride : nullptr;

  // each thread owns 4 CONSECUTIVE priors (the anchors of one feature-map cell for SSD heads), so a
  // warp covers 128 consecutive priors = a short run of neighbouring cells with a tight bounding box
  Box box[kMatchPPT];
  float area[kMatchPPT];
  float bt_ov[kMatchPPT];
  int bt_idx[kMatchPPT];
  int pidx[kMatchPPT];
  bool valid[kMatchPPT];
  float wx1 = INFINITY, wy1 = INFINITY, wx2 = -INFINITY, wy2 = -INFINITY;
#pragma unroll
  for (int k = 0; k < kMatchPPT; ++k) {
    int p = (gridDim.x - 1 - blockIdx.x) * kMatchTile + tid * kMatchPPT + k;   // heavy coarse-layer tiles first
    pidx[k] = p;
    valid[k] = p < a.P;
    if (valid[k]) {
      if (a.rf.arm_loc) {
        box[k] = decode_box(*reinterpret_cast<const float4*>(a.rf.arm_loc + (img_off + (size_t)p) * 4),
                            *reinterpret_cast<const float4*>(pri + (size_t)p * 4), a.rf.var0, a.rf.var1);
      } else if (anc) {
        float4 v = *reinterpret_cast<const float4*>(anc + (size_t)p * 4);
        box[k].x1 = v.x; box[k].y1 = v.y; box[k].x2 = v.z; box[k].y2 = v.w;
      } else {
        box[k] = point_form(*reinterpret_cast<const float4*>(pri + (size_t)p * 4));
      }
      wx1 = fminf(wx1, box[k].x1); wy1 = fminf(wy1, box[k].y1);
      wx2 = fmaxf(wx2, box[k].x2); wy2 = fmaxf(wy2, box[k].y2);
    } else {
      box[k].x1 = box[k].y1 = box[k].x2 = box[k].y2 = 0.0f;
    }
    area[k] = box_area(box[k]);
    bt_ov[k] = 0.0f;     // max over truths of IoU >= 0; all-zero column -> truth 0 (first index)
    bt_idx[k] = 0;
  }
  // warp-wide bounding box of the 128 priors: a truth that does not overlap it has IoU == 0 with
  // every prior of the warp and can change neither running maximum (both updates are strict >)
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    wx1 = fminf(wx1, __shfl_xor_sync(SSDBOX_FULL_MASK, wx1, d));
    wy1 = fminf(wy1, __shfl_xor_sync(SSDBOX_FULL_MASK, wy1, d));
    wx2 = fmaxf(wx2, __shfl_xor_sync(SSDBOX_FULL_MASK, wx2, d));
    wy2 = fmaxf(wy2, __shfl_xor_sync(SSDBOX_FULL_MASK, wy2, d));
  }

  MPHASE(2);
  for (int g = 0; g < G; ++g) {
    float4 tv = s_box[g];
    if (!(tv.x < wx2 && tv.z > wx1 && tv.y < wy2 && tv.w > wy1)) continue;   // warp-uniform
    Box t;
    t.x1 = tv.x; t.y1 = tv.y; t.x2 = tv.z; t.y2 = tv.w;
    float ta = s_area[g];
    float lm = 0.0f;
    uint32_t lp = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < kMatchPPT; ++k) {
      float iou = valid[k] ? iou_jaccard(t, ta, box[k], area[k]) : 0.0f;
      if (iou > bt_ov[k]) {           // strict: first truth wins ties (box_utils.py:118)
        bt_ov[k] = iou;
        bt_idx[k] = g;
      }
      if (iou > lm) {                 // strict + ascending p: lowest prior wins ties (:116)
        lm = iou;
        lp = (uint32_t)pidx[k];
      }
    }
    uint32_t mb = __reduce_max_sync(SSDBOX_FULL_MASK, __float_as_uint(lm));
    if (mb != 0u) {
      uint32_t cand = (__float_as_uint(lm) == mb) ? lp : 0xffffffffu;
      uint32_t pm = __reduce_min_sync(SSDBOX_FULL_MASK, cand);
      if (lane == 0) {
        unsigned long long key = ((unsigned long long)mb << 32) | (unsigned long long)(uint32_t)(~pm);
        if (key > *reinterpret_cast<volatile unsigned long long*>(&s_best[g])) atomicMax(&s_best[g], key);
      }
    }
  }

  MPHASE(3);
#pragma unroll
  for (int k = 0; k < kMatchPPT; ++k) {
    if (valid[k]) {
      float ov = G > 0 ? bt_ov[k] : 0.0f;
      int lb = (G > 0 && !(ov < a.threshold)) ? s_lab[bt_idx[k]] : 0;   // :130
      lab[img_off + pidx[k]] = (int16_t)lb;
      tidx[img_off + pidx[k]] = (int16_t)bt_idx[k];
      if (overlap) overlap[img_off + pidx[k]] = ov;
    }
  }

  // publish this CTA's per-truth candidates, then elect the last CTA of the image
  MPHASE(4);
  __syncthreads();
  for (int g = tid; g < G; g += kMatchThreads) {
    unsigned long long v = s_best[g];
    if (v > kBestInit) atomicMax(&gt_best[(size_t)b * gpad + g], v);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned t = atomicAdd(&done[b], 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  MPHASE(5);
  if (!s_last) return;
  __threadfence();

  // forced assignment, box_utils.py:123-127: overlap := 2, truth := j, sequentially (last j wins)
  for (int g = tid; g < G; g += kMatchThreads) {
    s_best[g] = __ldcg(&gt_best[(size_t)b * gpad + g]);
    gt_best[(size_t)b * gpad + g] = kBestInit;          // the image's state is handed back initialised
  }
  if (tid == 0) done[b] = 0u;
  __syncthreads();
  for (int j = tid; j < G; j += kMatchThreads) {
    uint32_t pj = ~(uint32_t)(s_best[j] & 0xffffffffull);
    bool winner = true;
    for (int j2 = j + 1; j2 < G; ++j2) {
      if (~(uint32_t)(s_best[j2] & 0xffffffffull) == pj) {
        winner = false;
        break;
      }
    }
    if (winner && pj < (uint32_t)a.P) {
      lab[img_off + pj] = (int16_t)s_lab[j];
      tidx[img_off + pj] = (int16_t)j;
      if (overlap) overlap[img_off + pj] = 2.0f;
    }
  }
}

int launch_match(const MatchArgs& a, const MatchWs& w, int16_t* lab, int16_t* tidx, float* overlap,
                 cudaStream_t st) {
  if (a.B == 0 || a.P == 0) return SSDBOX_OK;
  int gpad = gt_pad(a.gmax);
  size_t smem = (size_t)gpad * 32;
  if (smem > 48 * 1024) {
    SSDBOX_CUDA(cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid((a.P + kMatchTile - 1) / kMatchTile, a.B);
{
    TimerScope ts__(KID_MATCH, st);
    match_kernel<<<grid, kMatchThreads, smem, st>>>(a, w.gt_best, w.done, lab, tidx, overlap, gpad);
  }
  SSDBOX_LAUNCH_OK("match_kernel");
  return SSDBOX_OK;
}

__global__ void materialize_kernel(MatchArgs a, float var0, float var1, const int16_t* __restrict__ lab,
                                   const int16_t* __restrict__ tidx, float* __restrict__ loc_t,
                                   int64_t* __restrict__ conf_t, int32_t* __restrict__ match_idx) {
  size_t n = (size_t)a.B * a.P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int b = (int)(i / a.P);
    int p = (int)(i - (size_t)b * a.P);
    int g0 = a.gt_offsets[b];
    int G = a.gt_offsets[b + 1] - g0;
    int t = tidx[i];
    if (conf_t) conf_t[i] = (int64_t)lab[i];
    if (match_idx) match_idx[i] = t;
    if (loc_t) {
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (G > 0) {
        const float* row = a.gt + (size_t)(g0 + t) * 5;
        Box m;
        m.x1 = row[0]; m.y1 = row[1]; m.x2 = row[2]; m.y2 = row[3];
        float4 pr = refine_center(a.rf, *reinterpret_cast<const float4*>(a.priors + (size_t)b * (size_t)a.prior_stride + (size_t)p * 4), i);
        r = encode_box(m, pr, var0, var1);
      }
      *reinterpret_cast<float4*>(loc_t + i * 4) = r;
    }
  }
}

int launch_materialize(const MatchArgs& a, float var0, float var1, const int16_t* lab, const int16_t* tidx,
                       float* loc_t, int64_t* conf_t, int32_t* match_idx, cudaStream_t st) {
  size_t n = (size_t)a.B * a.P;
  if (n == 0) return SSDBOX_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
{
    TimerScope ts__(KID_MATERIALIZE, st);
    materialize_kernel<<<blocks, 256, 0, st>>>(a, var0, var1, lab, tidx, loc_t, conf_t, match_idx);
  }
  SSDBOX_LAUNCH_OK("materialize_kernel");
  return SSDBOX_OK;
}

}  // namespace ssdbox

using namespace ssdbox;

#ifdef SSDBOX_PHASE_TIMING
extern "C" __attribute__((visibility("default"))) int ssdbox_debug_match_phases(long long* out32) {
  return cudaMemcpyFromSymbol(out32, ssdbox::g_mphase, sizeof(long long) * 32) == cudaSuccess ? 0 : -5;
}
#endif

extern "C" int ssdbox_match_encode(const float* gt, const int32_t* gt_offsets, int32_t gmax, const float* priors,
                                   int64_t prior_batch_stride, const float* anchors_xyxy, int32_t B, int32_t P,
                                   float threshold, float var0, float var1, int32_t binarize_labels, float* loc_t,
                                   int64_t* conf_t, int32_t* match_idx, float* overlap, void* ws, size_t ws_bytes,
                                   ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(B >= 0 && P >= 0 && gmax >= 0, SSDBOX_EINVAL, "match: negative size");
  SSDBOX_REQUIRE(gmax <= kGmaxLimit, SSDBOX_ESHAPE, "match: gmax %d > %d", gmax, kGmaxLimit);
  SSDBOX_REQUIRE(B <= 65535, SSDBOX_ESHAPE, "match: batch %d > 65535", B);
  if (B == 0 || P == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(gt_offsets && priors && ws && (gt || gmax == 0), SSDBOX_EINVAL, "match: null pointer");
  SSDBOX_REQUIRE(prior_batch_stride == 0 || prior_batch_stride == (int64_t)P * 4, SSDBOX_EINVAL,
                 "match: prior_batch_stride must be 0 or 4*P");
  SSDBOX_REQUIRE(aligned16(priors) && (!anchors_xyxy || aligned16(anchors_xyxy)) && (!loc_t || aligned16(loc_t)),
                 SSDBOX_EALIGN, "match: box pointers must be 16-byte aligned");
  SSDBOX_REQUIRE(ws_bytes >= match_ws_bytes(B, P, gmax), SSDBOX_EWORKSPACE, "match: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(ws);
  MatchWs w;
  carve_match_core(c, B, gmax, &w);
  w.lab = c.take<int16_t>((size_t)B * P);
  w.tidx = c.take<int16_t>((size_t)B * P);
  MatchArgs a{gt, gt_offsets, gmax, priors, (long long)prior_batch_stride, anchors_xyxy, B, P, threshold,
              binarize_labels};
  int rc = launch_init(w.gt_best, (size_t)(B + 1) * gt_pad(gmax), w.done, (size_t)B + 1, nullptr, 0, nullptr, 0, st);
  if (rc) return rc;
  rc = launch_match(a, w, w.lab, w.tidx, overlap, st);
  if (rc) return rc;
  if (loc_t || conf_t || match_idx) rc = launch_materialize(a, var0, var1, w.lab, w.tidx, loc_t, conf_t, match_idx, st);
  return rc;
}
