// PASCAL VOC evaluation and the crop-sampling IoU.
//
// ssdbox_voc_eval replaces the chain evaluate_detections -> write_voc_results_file -> do_python_eval
// -> voc_eval -> voc_ap of lib/datasets/voc_eval.py (:58-75, :78-106, :109-242, :244-262) for ALL
// classes at once, on the flat rows that ssdbox_detections_compact emits:
//
//   voc_match_kernel    one warp per (image, class) segment, one lane per detection.  The text round
//                       trip of :70-74 / :170-175 is arithmetic here: '{:.3f}' of an fp32 score is
//                       rint(score * 1000) / 1000 and '{:.1f}' of an fp32 coordinate is
//                       rint(x * 10) / 10 -- exact, because an fp32 value times 1000 (10) is exact in
//                       fp64, rint rounds half to even like the correctly rounded formatter, and the
//                       fp64 quotient of two exact integers is the double float() parses.  IoU in fp64
//                       with the reference's operation order (:190-203), first-index arg-max.  The
//                       sequential "first detection to reach a truth claims it" rule (:207-215) becomes
//                       an atomicMin on the truth of the detection's sort key (score bin, row): the
//                       claimant is the detection that comes first in the class's sorted order.
//   radix_*_kernel      stable LSD radix sort (8-bit digits) of the rows by (class, descending
//                       quantised score); equal scores keep the file order = (image, row) order.
//                       This is the canonical form of the reference's np.argsort(-confidence) (:178),
//                       whose order among equal scores is whatever numpy's introsort leaves.
//   voc_curve_kernel    one CTA per class walks its sorted list: tp / fp flags, running sums, rec and
//                       prec (:218-223) in fp64, the 11-point AP (:85-93, p/11 accumulated in threshold
//                       order) or the area under the precision envelope (:95-105).
//
// ssdbox_crop_overlaps is the data-parallel part of one RandomSampleCrop trial
// (lib/utils/augmentations.py:13-37, 250-268) batched over images and candidate rects.
#include <cfloat>

#include <math_constants.h>

#include "ops.h"
#include "ssdbox_dev.cuh"

namespace ssdbox {

constexpr int kVocWarps = 8;
constexpr int kRadixThreads = 256;
constexpr int kRadixWarps = kRadixThreads / 32;
constexpr int kRadixRounds = 16;                                       // items per thread
constexpr int kRadixTile = kRadixThreads * kRadixRounds;               // 4096 items per CTA
constexpr int kScoreBits = 10;                                         // 1000 - k, k = 0..1000
constexpr int kCurveThreads = 1024;
constexpr int kCurveItems = 4;                                         // 4096 positions per block scan: counts fit 16 bits

struct VocArgs {
  const float* rows;
  int row_stride;
  const int32_t* seg;
  const float* gt_boxes;
  const int32_t* gt_labels;
  const uint8_t* gt_difficult;
  const int32_t* gt_offsets;
  int I, C, N, M, use07;
  double ovthresh;
  unsigned long long* claim;   // [M]  min sort key of the detections that reach the truth
  uint32_t* skey;              // [N]  class << 10 | (1000 - k)
  int32_t* code;               // [N]  -2 false positive, -1 neither (difficult truth), >= 0 truth index
  int32_t* cnt;                // [C]  detections per class
  int32_t* npos;               // [C]  non-difficult truths per class (:163)
  int32_t* status;             // [1]  rows whose quantised score is outside [0, 1]
  const int32_t* order;        // [N]  sorted position -> row
  int32_t* cls_offsets;        // [C+1]
  uint8_t* tpfp;               // [N]  sorted order: 1 tp, 2 fp, 0 neither
  double* rec;
  double* prec;
  double* ap;                  // [C]
};

// float('{:.{d}f}'.format(v)) for fp32 v and scale = 10^d <= 1000 (see the header comment)
__device__ __forceinline__ double text_round_trip(float v, double scale) {
  return __ddiv_rn(rint(__dmul_rn((double)v, scale)), scale);
}

__global__ void __launch_bounds__(kVocWarps * 32) voc_match_kernel(VocArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long s = (long long)blockIdx.x * kVocWarps + warp;
  if (s >= (long long)a.I * a.C) return;
  const int i = (int)(s / a.C), c = (int)(s - (long long)i * a.C);
  int r0 = a.seg[s], r1 = a.seg[s + 1];
  // malformed segments (not monotone / beyond the rows) never index out of bounds: they are clamped here and
  // reported through the status word (same 2^30 flag voc_curve_kernel raises for seg[-1] != num_rows)
  if (r0 < 0 || r1 < r0 || r1 > a.N) {
    if (lane == 0) atomicOr(reinterpret_cast<unsigned int*>(a.status), 1u << 30);
    r0 = r0 < 0 ? 0 : (r0 > a.N ? a.N : r0);
    r1 = r1 < r0 ? r0 : (r1 > a.N ? a.N : r1);
  }
  if (c == 0) {                                                        // background rows (DetectOut emits none) are never
    if (lane == 0 && r1 > r0) atomicAdd(&a.cnt[0], r1 - r0);           // evaluated: they sort in front, flagged "neither"
    for (int r = r0 + lane; r < r1; r += 32) {
      a.skey[r] = 0u;
      a.code[r] = -1;
    }
    return;
  }
  const int g0 = a.gt_offsets[i], g1 = a.gt_offsets[i + 1];
  int np = 0;
  for (int g = g0 + lane; g < g1; g += 32) np += (a.gt_labels[g] == c && !a.gt_difficult[g]) ? 1 : 0;
  np = warp_sum(np);
  if (lane == 0) {
    if (np) atomicAdd(&a.npos[c], np);                                 // voc_eval.py:163
    if (r1 > r0) atomicAdd(&a.cnt[c], r1 - r0);
  }
  const float4* gtb = reinterpret_cast<const float4*>(a.gt_boxes);
  for (int r = r0 + lane; r < r1; r += 32) {
    const float* row = a.rows + (size_t)r * a.row_stride;
    // :70-74  '{:.3f}' score, '{:.1f}' of (float32 coordinate + 1)
    const double bx1 = text_round_trip(__fadd_rn(row[0], 1.0f), 10.0);
    const double by1 = text_round_trip(__fadd_rn(row[1], 1.0f), 10.0);
    const double bx2 = text_round_trip(__fadd_rn(row[2], 1.0f), 10.0);
    const double by2 = text_round_trip(__fadd_rn(row[3], 1.0f), 10.0);
    const double ks = rint(__dmul_rn((double)row[4], 1000.0));
    int k;
    if (ks >= 0.0 && ks <= 1000.0) {
      k = (int)ks;
    } else {
      atomicAdd(a.status, 1);
      k = ks > 1000.0 ? 1000 : 0;
    }
    const uint32_t bin = (uint32_t)(1000 - k);
    a.skey[r] = ((uint32_t)c << kScoreBits) | bin;
    const double barea = __dmul_rn(__dsub_rn(bx2, bx1), __dsub_rn(by2, by1));
    double best = -CUDART_INF;
    int j = -1;
    bool nan = false;
    for (int g = g0; g < g1; ++g) {                                    // :190-205
      if (a.gt_labels[g] != c) continue;
      const float4 t = gtb[g];
      const double tx1 = t.x, ty1 = t.y, tx2 = t.z, ty2 = t.w;
      const double iw = fmax(__dsub_rn(fmin(tx2, bx2), fmax(tx1, bx1)), 0.0);
      const double ih = fmax(__dsub_rn(fmin(ty2, by2), fmax(ty1, by1)), 0.0);
      const double inters = __dmul_rn(iw, ih);
      const double uni = __dsub_rn(__dadd_rn(barea, __dmul_rn(__dsub_rn(tx2, tx1), __dsub_rn(ty2, ty1))), inters);
      const double ov = __ddiv_rn(inters, uni);
      if (ov != ov) nan = true;                                        // np.max propagates NaN -> not > ovthresh
      else if (ov > best) { best = ov; j = g; }                        // np.argmax: first maximum
    }
    int code;
    if (nan || !(best > a.ovthresh)) {
      code = -2;                                                       // :215-216
    } else if (a.gt_difficult[j]) {
      code = -1;                                                       // :208
    } else {
      code = j;                                                        // :209-214, resolved in voc_curve_kernel
      atomicMin(&a.claim[j], ((unsigned long long)bin << 32) | (uint32_t)r);
    }
    a.code[r] = code;
  }
}

// ---- stable LSD radix sort, 8-bit digits ------------------------------------------------------------
__global__ void __launch_bounds__(kRadixThreads) radix_hist_kernel(const uint32_t* keys, int n, int shift,
                                                                    uint32_t* blockhist, int nblk) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * kRadixTile;
#pragma unroll 4
  for (int j = 0; j < kRadixRounds; ++j) {
    const int i = base + j * kRadixThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  blockhist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];    // digit-major
}

// exclusive scan of blockhist in (digit, block) order, in place (one CTA): every thread sums a run of
// consecutive entries, one block scan of the run totals, then the run is rewritten with its prefixes
// (two barriers in all; the table is a few hundred KB and stays in L2)
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t* blockhist, int total) {
  __shared__ int s_scan[33];
  const int per = (total + 1023) / 1024;
  const int lo = min(threadIdx.x * per, total), hi = min(lo + per, total);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += (int)blockhist[i];
  int all;
  int run = block_exclusive_scan(sum, s_scan, &all);
  for (int i = lo; i < hi; ++i) {
    const int v = (int)blockhist[i];
    blockhist[i] = (uint32_t)run;
    run += v;
  }
}

// every warp owns a contiguous run of kRadixRounds * 32 items; destination = digit base of the CTA +
// items of the digit in the lower warps + items of the digit earlier in this warp + rank among the
// lanes of this round -> input order is preserved inside every digit (stable).
__global__ void __launch_bounds__(kRadixThreads) radix_scatter_kernel(const uint32_t* keys_in, const uint32_t* vals_in,
                                                                       int n, int shift, const uint32_t* bases, int nblk,
                                                                       uint32_t* keys_out, uint32_t* vals_out) {
  __shared__ uint32_t wc[kRadixWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = threadIdx.x; t < kRadixWarps * 256; t += kRadixThreads) (&wc[0][0])[t] = 0;
  __syncthreads();
  const int first = blockIdx.x * kRadixTile + warp * (kRadixRounds * 32);
  uint32_t key[kRadixRounds];
#pragma unroll
  for (int j = 0; j < kRadixRounds; ++j) {
    const int i = first + j * 32 + lane;
    const bool valid = i < n;
    key[j] = valid ? keys_in[i] : 0u;
    const uint32_t d = valid ? (key[j] >> shift) & 255u : 256u;
    const uint32_t m = __match_any_sync(SSDBOX_FULL_MASK, d);
    if (valid && lane == __ffs(m) - 1) wc[warp][d] += __popc(m);
    __syncwarp();
  }
  __syncthreads();
  {
    const int d = threadIdx.x;
    uint32_t run = bases[(size_t)d * nblk + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kRadixWarps; ++w) {
      const uint32_t t = wc[w][d];
      wc[w][d] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kRadixRounds; ++j) {
    const int i = first + j * 32 + lane;
    const bool valid = i < n;
    const uint32_t d = valid ? (key[j] >> shift) & 255u : 256u;
    const uint32_t m = __match_any_sync(SSDBOX_FULL_MASK, d);
    if (valid) {
      const uint32_t pos = wc[warp][d] + __popc(m & ((1u << lane) - 1u));
      keys_out[pos] = key[j];
      vals_out[pos] = vals_in ? vals_in[i] : (uint32_t)i;
    }
    __syncwarp();
    if (valid && lane == __ffs(m) - 1) wc[warp][d] += __popc(m);
    __syncwarp();
  }
}

// ---- per-class curves ----------------------------------------------------------------------------
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(SSDBOX_FULL_MASK, v, d));
  return v;
}

// inclusive running max over the block's threads in thread-id order; *total = max over the block.
__device__ __forceinline__ double block_inclusive_max(double v, double* scratch, double* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double t = __shfl_up_sync(SSDBOX_FULL_MASK, v, d);
    if (lane >= d) v = fmax(v, t);
  }
  __syncthreads();
  if (lane == 31) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double w = lane < nwarp ? scratch[lane] : 0.0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double t = __shfl_up_sync(SSDBOX_FULL_MASK, w, d);
      if (lane >= d) w = fmax(w, t);
    }
    // exclusive: max over the lower warps (values are >= 0, so 0 is the identity)
    const double ex = __shfl_up_sync(SSDBOX_FULL_MASK, w, 1);
    scratch[lane] = lane == 0 ? 0.0 : ex;
    if (lane == 31) scratch[32] = w;
  }
  __syncthreads();
  const double res = fmax(v, scratch[warp]);
  *total = scratch[32];
  return res;
}

__global__ void __launch_bounds__(kCurveThreads) voc_curve_kernel(VocArgs a) {
  __shared__ int s_scan[33];
  __shared__ double s_d[33];
  __shared__ double s_m[11][33];
  __shared__ int s_begin, s_n;
  const int c = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    int off = 0;
    for (int cc = 0; cc < c; ++cc) off += a.cnt[cc];
    s_begin = off;
    s_n = a.cnt[c];
    if (c == 0) {                                                      // the first CTA publishes the class offsets
      int run = 0;
      for (int cc = 0; cc < a.C; ++cc) {
        a.cls_offsets[cc] = run;
        run += a.cnt[cc];
      }
      a.cls_offsets[a.C] = run;
      if (run != a.N) atomicOr(reinterpret_cast<unsigned int*>(a.status), 1u << 30);                   // the segments do not cover the rows exactly
    }
  }
  __syncthreads();
  const int begin = s_begin, n = s_n;
  if (c == 0) {                                                        // background rows: no curve
    for (int p = tid; p < n; p += kCurveThreads) {
      a.rec[begin + p] = 0.0;
      a.prec[begin + p] = 0.0;
      a.tpfp[begin + p] = 0;
    }
    if (tid == 0) a.ap[0] = -1.0;
    return;
  }
  if (n == 0) {
    if (tid == 0) a.ap[c] = -1.0;                                      // :238-241
    return;
  }
  const double npos = (double)a.npos[c];
  double m[11];
#pragma unroll
  for (int q = 0; q < 11; ++q) m[q] = 0.0;
  int carry_tp = 0, carry_fp = 0;
  // kCurveItems consecutive sorted positions per thread: the three dependent gathers (order -> code /
  // key -> claim) of the items overlap and one block scan serves kCurveItems * 1024 positions
  for (int base = 0; base < n; base += kCurveThreads * kCurveItems) {
    const int p0 = base + tid * kCurveItems;
    uint32_t row[kCurveItems];
    int code[kCurveItems];
    uint32_t bin[kCurveItems];
    int flag[kCurveItems];                                             // (tp << 16) | fp
#pragma unroll
    for (int q = 0; q < kCurveItems; ++q) row[q] = p0 + q < n ? (uint32_t)a.order[begin + p0 + q] : 0u;
#pragma unroll
    for (int q = 0; q < kCurveItems; ++q) {
      const bool valid = p0 + q < n;
      code[q] = valid ? a.code[row[q]] : -1;
      bin[q] = valid ? a.skey[row[q]] & ((1u << kScoreBits) - 1u) : 0u;
    }
    int mine_sum = 0;
#pragma unroll
    for (int q = 0; q < kCurveItems; ++q) {
      int f = 0;
      if (code[q] == -2) {
        f = 1;
      } else if (code[q] >= 0) {                                       // :209-214: the first detection in sorted order claims the truth
        const unsigned long long key = ((unsigned long long)bin[q] << 32) | row[q];
        f = a.claim[code[q]] == key ? (1 << 16) : 1;
      }
      flag[q] = f;
      mine_sum += f;
    }
    int sum;
    int run = block_exclusive_scan(mine_sum, s_scan, &sum);
#pragma unroll
    for (int q = 0; q < kCurveItems; ++q) {
      run += flag[q];
      if (p0 + q < n) {
        const double t = (double)(carry_tp + (run >> 16));             // np.cumsum of 0/1 doubles: exact integers
        const double f = (double)(carry_fp + (run & 0xffff));
        const double r = __ddiv_rn(t, npos);                           // :220
        const double pr = __ddiv_rn(t, fmax(__dadd_rn(t, f), DBL_EPSILON));   // :223
        a.rec[begin + p0 + q] = r;
        a.prec[begin + p0 + q] = pr;
        a.tpfp[begin + p0 + q] = (flag[q] >> 16) ? 1 : (flag[q] ? 2 : 0);
#pragma unroll
        for (int k = 0; k < 11; ++k)
          if (r >= __dmul_rn((double)k, 0.1)) m[k] = fmax(m[k], pr);   // np.arange(0., 1.1, 0.1)[k] == k * 0.1
      }
    }
    carry_tp += sum >> 16;
    carry_fp += sum & 0xffff;
    __syncthreads();
  }
  if (a.use07) {                                                       // :85-93
#pragma unroll
    for (int q = 0; q < 11; ++q) {
      const double v = warp_max(m[q]);
      if (lane == 0) s_m[q][warp] = v;
    }
    __syncthreads();
    if (tid == 0) {
      double ap = 0.0;
      for (int q = 0; q < 11; ++q) {
        double best = 0.0;
        for (int w = 0; w < kCurveThreads / 32; ++w) best = fmax(best, s_m[q][w]);
        ap = __dadd_rn(ap, __ddiv_rn(best, 11.0));
      }
      a.ap[c] = ap;
    }
    return;
  }
  // :95-105  mrec = [0, rec, 1], mpre = [0, prec, 0] made non-increasing from the right; sum over the
  // positions where mrec changes of (mrec[i+1] - mrec[i]) * mpre[i+1].  Thread t of a chunk takes the
  // position top-1-t, so the running max in thread order is the envelope.
  double carry_max = 0.0, acc = 0.0;
  for (int top = n; top > 0; top -= kCurveThreads) {
    const int p = top - 1 - tid;
    const double pr = p >= 0 ? a.prec[begin + p] : 0.0;
    double chunk_max;
    const double env = fmax(block_inclusive_max(pr, s_d, &chunk_max), carry_max);
    if (p >= 0) {
      const double r = a.rec[begin + p];
      const double rprev = p > 0 ? a.rec[begin + p - 1] : 0.0;
      if (r != rprev) acc = __dadd_rn(acc, __dmul_rn(__dsub_rn(r, rprev), env));
    }
    carry_max = fmax(carry_max, chunk_max);
    __syncthreads();
  }
  const double total = block_sum(acc, s_d);
  if (tid == 0) {
    const double rl = a.rec[begin + n - 1];
    double ap = total;
    if (1.0 != rl) ap = __dadd_rn(ap, __dmul_rn(__dsub_rn(1.0, rl), 0.0));    // the closing sentinel: adds 0 (NaN if rec is not finite)
    a.ap[c] = ap;
  }
}

// ---- crop-sampling IoU ---------------------------------------------------------------------------
struct CropArgs {
  const double* boxes;       // [sum G, 4] absolute xyxy (float64 like the numpy pipeline)
  const int32_t* offsets;    // [B+1]
  const long long* rects;    // [B, T, 4] int64 (augmentations.py:248)
  int B, T;
  double* overlap;           // per image [T, G_b] at offsets[b] * T, nullable
  double* minmax;            // [B, T, 2]
  uint8_t* center_mask;      // like overlap, nullable
};

__global__ void __launch_bounds__(kVocWarps * 32) crop_overlaps_kernel(CropArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long u = (long long)blockIdx.x * kVocWarps + warp;
  if (u >= (long long)a.B * a.T) return;
  const int b = (int)(u / a.T), t = (int)(u - (long long)b * a.T);
  const int g0 = a.offsets[b], G = a.offsets[b + 1] - g0;
  const long long* rc = a.rects + u * 4;
  const long long rx1 = rc[0], ry1 = rc[1], rx2 = rc[2], ry2 = rc[3];
  const double dx1 = (double)rx1, dy1 = (double)ry1, dx2 = (double)rx2, dy2 = (double)ry2;
  const double area_b = (double)((rx2 - rx1) * (ry2 - ry1));            // int64 product, then promoted (:33-34)
  double lo = CUDART_INF, hi = -CUDART_INF;
  bool nan = false;
  for (int g = lane; g < G; g += 32) {
    const double* bx = a.boxes + (size_t)(g0 + g) * 4;
    const double x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
    const double w = fmax(__dsub_rn(fmin(x2, dx2), fmax(x1, dx1)), 0.0);   // :14-17 np.clip(max_xy - min_xy, 0, inf)
    const double h = fmax(__dsub_rn(fmin(y2, dy2), fmax(y1, dy1)), 0.0);
    const double inter = __dmul_rn(w, h);
    const double area_a = __dmul_rn(__dsub_rn(x2, x1), __dsub_rn(y2, y1));
    const double ov = __ddiv_rn(inter, __dsub_rn(__dadd_rn(area_a, area_b), inter));   // :35-36
    if (ov != ov) nan = true;
    lo = fmin(lo, ov);
    hi = fmax(hi, ov);
    const size_t o = (size_t)g0 * a.T + (size_t)t * G + g;
    if (a.overlap) a.overlap[o] = ov;
    if (a.center_mask) {                                               // :257-268
      const double cx = __ddiv_rn(__dadd_rn(x1, x2), 2.0), cy = __ddiv_rn(__dadd_rn(y1, y2), 2.0);
      a.center_mask[o] = (dx1 < cx && dy1 < cy && dx2 > cx && dy2 > cy) ? 1 : 0;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(SSDBOX_FULL_MASK, lo, d));
    hi = fmax(hi, __shfl_xor_sync(SSDBOX_FULL_MASK, hi, d));
  }
  nan = __any_sync(SSDBOX_FULL_MASK, nan);
  if (lane == 0) {
    const double q = nan ? __longlong_as_double(0x7ff8000000000000ll) : 0.0;
    a.minmax[u * 2] = nan ? q : lo;                                    // overlap.min() / .max() (:254): NaN propagates
    a.minmax[u * 2 + 1] = nan ? q : hi;
  }
}

}  // namespace ssdbox

using namespace ssdbox;

extern "C" int ssdbox_voc_eval(const ssdbox_voc_eval_cfg* cfg, const float* rows, const int32_t* seg_offsets,
                               const float* gt_boxes, const int32_t* gt_labels, const uint8_t* gt_difficult,
                               const int32_t* gt_offsets, int32_t* order, int32_t* cls_offsets, uint8_t* tpfp,
                               double* rec, double* prec, double* ap, int32_t* npos, int32_t* status, void* ws,
                               size_t ws_bytes, ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(cfg, SSDBOX_EINVAL, "voc_eval: null cfg");
  const int I = cfg->num_images, C = cfg->num_classes, M = cfg->num_gt;
  const long long N = cfg->num_rows;
  SSDBOX_REQUIRE(I >= 0 && C >= 1 && M >= 0 && N >= 0 && cfg->row_stride >= 5, SSDBOX_EINVAL, "voc_eval: bad sizes");
  SSDBOX_REQUIRE(C <= 1024, SSDBOX_ESHAPE, "voc_eval: at most 1024 classes");
  SSDBOX_REQUIRE(N < (1ll << 31) - kRadixTile && (long long)I * C < (1ll << 31) - 1, SSDBOX_ESHAPE, "voc_eval: too many rows / segments");
  SSDBOX_REQUIRE(cls_offsets && ap && npos && status, SSDBOX_EINVAL, "voc_eval: null output");
  SSDBOX_REQUIRE(N == 0 || (rows && order && tpfp && rec && prec), SSDBOX_EINVAL, "voc_eval: null pointer");
  SSDBOX_REQUIRE(I == 0 || (seg_offsets && gt_offsets), SSDBOX_EINVAL, "voc_eval: null offsets");
  SSDBOX_REQUIRE(M == 0 || (gt_boxes && gt_labels && gt_difficult), SSDBOX_EINVAL, "voc_eval: null truths");
  SSDBOX_REQUIRE(aligned16(gt_boxes), SSDBOX_EALIGN, "voc_eval: gt_boxes must be 16-byte aligned");
  SSDBOX_REQUIRE(ws && ws_bytes >= voc_eval_ws_bytes((int)N, M, C), SSDBOX_EWORKSPACE, "voc_eval: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver cv(ws);
  VocArgs a{};
  a.rows = rows; a.row_stride = cfg->row_stride; a.seg = seg_offsets;
  a.gt_boxes = gt_boxes; a.gt_labels = gt_labels; a.gt_difficult = gt_difficult; a.gt_offsets = gt_offsets;
  a.I = I; a.C = C; a.N = (int)N; a.M = M; a.use07 = cfg->use_07_metric; a.ovthresh = cfg->ovthresh;
  a.claim = cv.take<unsigned long long>((size_t)M + 1);
  a.skey = cv.take<uint32_t>((size_t)N + 1);
  a.code = cv.take<int32_t>((size_t)N + 1);
  a.cnt = cv.take<int32_t>((size_t)C);
  const int nblk = (int)((N + kRadixTile - 1) / kRadixTile);
  uint32_t* kb[2] = {cv.take<uint32_t>((size_t)N + 1), cv.take<uint32_t>((size_t)N + 1)};
  uint32_t* vb[2] = {cv.take<uint32_t>((size_t)N + 1), cv.take<uint32_t>((size_t)N + 1)};
  uint32_t* blockhist = cv.take<uint32_t>((size_t)256 * (nblk > 0 ? nblk : 1));
  a.npos = npos; a.status = status; a.order = order; a.cls_offsets = cls_offsets; a.tpfp = tpfp;
  a.rec = rec; a.prec = prec; a.ap = ap;
  SSDBOX_CUDA(cudaMemsetAsync(a.claim, 0xff, ((size_t)M + 1) * 8, st));
  SSDBOX_CUDA(cudaMemsetAsync(a.skey, 0, ((size_t)N + 1) * 4, st));      // rows outside every segment (caller error, reported
  SSDBOX_CUDA(cudaMemsetAsync(a.code, 0xff, ((size_t)N + 1) * 4, st));   // through *status): class 0, "neither", never dereferenced
  SSDBOX_CUDA(cudaMemsetAsync(a.cnt, 0, (size_t)C * 4, st));
  SSDBOX_CUDA(cudaMemsetAsync(npos, 0, (size_t)C * 4, st));
  SSDBOX_CUDA(cudaMemsetAsync(status, 0, 4, st));
  SSDBOX_CUDA(cudaMemsetAsync(ap, 0, (size_t)C * 8, st));
  if ((long long)I * C > 0) {
    const int grid = (int)(((long long)I * C + kVocWarps - 1) / kVocWarps);
    voc_match_kernel<<<grid, kVocWarps * 32, 0, st>>>(a);
    SSDBOX_LAUNCH_OK("voc_match_kernel");
  }
  if (N > 0) {
    int cbits = 0;
    while ((1 << cbits) < C) ++cbits;
    const int passes = (kScoreBits + cbits + 7) / 8;
    const uint32_t* kin = a.skey;
    const uint32_t* vin = nullptr;
    for (int p = 0; p < passes; ++p) {
      uint32_t* kout = kb[p & 1];
      uint32_t* vout = p == passes - 1 ? reinterpret_cast<uint32_t*>(order) : vb[p & 1];
      radix_hist_kernel<<<nblk, kRadixThreads, 0, st>>>(kin, (int)N, 8 * p, blockhist, nblk);
      SSDBOX_LAUNCH_OK("radix_hist_kernel");
      radix_scan_kernel<<<1, 1024, 0, st>>>(blockhist, 256 * nblk);
      SSDBOX_LAUNCH_OK("radix_scan_kernel");
      radix_scatter_kernel<<<nblk, kRadixThreads, 0, st>>>(kin, vin, (int)N, 8 * p, blockhist, nblk, kout, vout);
      SSDBOX_LAUNCH_OK("radix_scatter_kernel");
      kin = kout;
      vin = vout;
    }
  }
  voc_curve_kernel<<<C, kCurveThreads, 0, st>>>(a);
  SSDBOX_LAUNCH_OK("voc_curve_kernel");
  return SSDBOX_OK;
}

extern "C" int ssdbox_crop_overlaps(const double* boxes, const int32_t* box_offsets, const int64_t* rects, int32_t B,
                                    int32_t T, double* overlap, double* minmax, uint8_t* center_mask,
                                    ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(B >= 0 && T >= 0, SSDBOX_EINVAL, "crop_overlaps: bad sizes");
  if ((long long)B * T == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(box_offsets && rects && minmax, SSDBOX_EINVAL, "crop_overlaps: null pointer");
  SSDBOX_REQUIRE((long long)B * T < (1ll << 31) - kVocWarps, SSDBOX_ESHAPE, "crop_overlaps: B*T must be < 2^31");
  CropArgs a{};
  a.boxes = boxes; a.offsets = box_offsets; a.rects = reinterpret_cast<const long long*>(rects);
  a.B = B; a.T = T; a.overlap = overlap; a.minmax = minmax; a.center_mask = center_mask;
  const int grid = (int)(((long long)B * T + kVocWarps - 1) / kVocWarps);
  crop_overlaps_kernel<<<grid, kVocWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
  SSDBOX_LAUNCH_OK("crop_overlaps_kernel");
  return SSDBOX_OK;
}
