// DetectOut.forward (lib/layers/functions/detection.py:25-64) and nms (box_utils.py:279-343).
//
// Detect = 4 launches on the caller's stream (5 the first time a workspace is used: + init_kernel):
//   detect_stream_kernel   THE HBM-bound kernel: scores [B*P, C] streamed once through the TMA
//                          bulk-copy ring; one thread per prior row takes the row maximum over the
//                          foreground classes; the few rows with a hit are re-scanned by the warp
//                          class-major (lane = class, one atomicAdd per class per 32 rows) and append
//                          (ordered score << 32 | prior) to the (image, class) candidate list.  With
//                          raw logits (SSDBOX_DETECT_LOGITS) the softmax is fused in.  The transposed
//                          [C,P] view of detection.py:38-39 is never materialised; decode runs only
//                          on candidates.
//   detect_segments_kernel every list that fits the capacity: one WARP per (image, class) finishes lists of
//                          <= 32 candidates (the normal case) entirely in registers -- shuffle bitonic sort,
//                          decode, NMS, output rows; longer lists go to a CTA-local queue and are finished by
//                          the whole CTA (bitonic sort by (score desc, prior desc) = the reference's visiting
//                          order, top_k cut, decode of the survivors, upper-triangular suppression bit-matrix
//                          built with warp ballots in shared memory, one-warp greedy sweep)
//   detect_overflow_chunk_kernel  only for lists that exceeded the candidate capacity (dense scores):
//                          per (image, 8 classes) a range-adaptive radix select straight from
//                          `scores` rewrites each list with the <= 1024 best keys and queues it
//   detect_segment_kernel  one CTA per rewritten list; both exit on one load when nothing overflowed.
// (top_k > 256: detect_overflow_kernel, one CTA per overflowed list, exact column select + NMS.)
// Every kernel resets the counters of the lists it finishes: the workspace state is clean after the call.
#include "ops.h"
#include "peer.cuh"
#include "ring.cuh"
#include "select.cuh"
#include "ssdbox_dev.cuh"

namespace ssdbox {

// ------------------------------------------------------------------------------------------------
// candidate pass
// ------------------------------------------------------------------------------------------------
struct DetStreamArgs {
  RingPlan ring;            // scores [B*P, C]
  const uint8_t* keep;      // [B*P] nullable
  RefineArgs rf;            // RefineDet fused: the score mask is the ARM objectness
  uint32_t* cnt;            // [B*C]
  unsigned long long* cand; // [B*C, cap]
  int P;
  int cap;
  float thr;
  float* row_m;             // [B*P] logits mode: row maximum ...
  float* row_s;             // [B*P] ... and sum_c expf(x_c - max) of rows that can hold a candidate, else 0
};

// LOGITS: the rows are raw class logits.  Every row pays one cheap softmax denominator (ex2.approx,
// like the loss kernel) to decide conservatively whether any foreground class can exceed the
// threshold; rows that can are re-scanned by the whole warp with expf and a correctly rounded division.
constexpr float kDetLog2e = 1.4426950408889634f;

template <int CT, bool LOGITS>
__global__ void __launch_bounds__(kRingThreads, 1) detect_stream_kernel(DetStreamArgs a) {
  extern __shared__ __align__(128) unsigned char smem_ring[];
  RingCtx rc = ring_setup(a.ring, smem_ring);
  const int C = CT > 0 ? CT : a.ring.C;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == kRingConsumerWarps) {
    ring_produce(a.ring, rc);
    return;
  }
  const int wg = warp / kRingGroupWarps;
  const int rbase = tid % (32 * kRingGroupWarps);
  const int KR = a.ring.KR;
  const int R = a.ring.R, NS = a.ring.NS;
  const float thr = a.thr;
  uint32_t full = 0u, full_b = 0xffffffffu;     // saturated class slots of this lane, and the image they belong to
  for (int it = wg; it < rc.n_local; it += kRingGroups) {
    const int s = it % NS, j = it % (2 * NS), ph = it / (2 * NS);
    const float* st = rc.stages + (size_t)s * rc.stage_floats;
    mbar_wait(&rc.full[j], (uint32_t)(ph & 1));
    for (int k = 0; k < KR; ++k) {
      const int r = k * (32 * kRingGroupWarps) + rbase;
      const long long row = (rc.t0 + it * rc.tstep) * R + r;
      const bool valid = (r < R) && (row < a.ring.rows);
      bool kept = true;
      if (valid && (a.keep || a.rf.arm_conf)) kept = refine_member(a.rf, a.keep, (size_t)row);   // RefineDet: filtered anchors score 0
      const float* rp = st + (size_t)r * C;
      float m = -INFINITY;        // scores: max over the foreground classes; logits: max over all classes
      bool hit = false;
      if (valid) {
        if (LOGITS) {
          float mfg = -INFINITY, sum = 0.f;
          if (CT > 1) {
            float v[CT > 1 ? CT : 1];
  #pragma unroll
            for (int c = 0; c < CT; ++c) v[c] = rp[c];
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  #pragma unroll
            for (int c = 1; c < CT; ++c) m4[c & 3] = fmaxf(m4[c & 3], v[c]);
            mfg = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            m = fmaxf(mfg, v[0]);
            const float nml = -m * kDetLog2e;
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
  #pragma unroll
            for (int c = 0; c < CT; ++c) s4[c & 3] += ex2_approx(fmaf(v[c], kDetLog2e, nml));
            sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
          } else {
            for (int c = 1; c < C; ++c) mfg = fmaxf(mfg, rp[c]);
            m = fmaxf(mfg, rp[0]);
            const float nml = -m * kDetLog2e;
            for (int c = 0; c < C; ++c) sum += ex2_approx(fmaf(rp[c], kDetLog2e, nml));
          }
          // conservative: the approximate ratio is within ~1e-5 of the exact one
          hit = kept ? ex2_approx((mfg - m) * kDetLog2e) > thr * sum * 0.9999f : 0.0f > thr;
          if (!hit) a.row_s[row] = 0.0f;
        } else {
          if (CT > 1) {
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  #pragma unroll
            for (int c = 1; c < CT; ++c) m4[c & 3] = fmaxf(m4[c & 3], rp[c]);
            m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          } else {
            for (int c = 1; c < C; ++c) m = fmaxf(m, rp[c]);
          }
          if (!kept) m = 0.0f;
          hit = m > thr;                                     // detection.py:48 strict >
        }
      }
      const uint32_t hits = __ballot_sync(SSDBOX_FULL_MASK, valid && hit);
      if (hits) {
        const uint32_t keptmask = __ballot_sync(SSDBOX_FULL_MASK, kept);
        const int wrow0 = r - lane;                          // the warp's 32 rows are consecutive
        const long long grow0 = row - lane;
        float my_s = 0.f;                                    // LOGITS: softmax denominator of my row (m holds its max)
        if (LOGITS) {
          uint32_t h = hits;
          while (h) {
            const int src = __ffs(h) - 1;
            h &= h - 1;
            const float* hp = st + (size_t)(wrow0 + src) * C;
            const float hm = __shfl_sync(SSDBOX_FULL_MASK, m, src);
            float part = 0.f;
            for (int c = lane; c < C; c += 32) part += expf(hp[c] - hm);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(SSDBOX_FULL_MASK, part, d);
            if (lane == src) my_s = part;
            if (lane == 0) {
              a.row_m[grow0 + src] = hm;
              a.row_s[grow0 + src] = part;
            }
          }
        }
        // Class-major over the hit rows: lane = class, ONE atomicAdd per (class, 32 rows) reserves the
        // slots of all its candidates (dense scores would otherwise serialise an atomic round trip per
        // candidate).  All hit rows of a warp belong to one image except at an image boundary.
        const uint32_t b_lo = (uint32_t)(grow0 + (__ffs(hits) - 1)) / (uint32_t)a.P;
        const uint32_t b_hi = (uint32_t)(grow0 + (31 - __clz(hits))) / (uint32_t)a.P;
        for (uint32_t b = b_lo; b <= b_hi; ++b) {
          // hit rows of image b
          uint32_t hb = hits;
          if (b_lo != b_hi) {
            hb = 0u;
            uint32_t h = hits;
            while (h) {
              const int rb = __ffs(h) - 1;
              h &= h - 1;
              if ((uint32_t)(grow0 + rb) / (uint32_t)a.P == b) hb |= 1u << rb;
            }
            if (hb == 0u) continue;
          }
          if (b != full_b) {
            full_b = b;
            full = 0u;
          }
          const int p0 = (int)(grow0 - (long long)b * a.P);   // prior index of the warp's first row (may be < 0)
          for (int sb = 0; 1 + 32 * sb < C; ++sb) {
            const int c = 1 + lane + 32 * sb;
            const bool act = c < C && !(sb < 32 && ((full >> sb) & 1u));
            if (!__any_sync(SSDBOX_FULL_MASK, act)) continue;      // every class of this slot is saturated (dense scores)
            const int cc = act ? c : 0;
            int cnt = 0;
            uint32_t h = hb;
            while (h) {
              const int rb = __ffs(h) - 1;
              h &= h - 1;
              float v = st[(size_t)(wrow0 + rb) * C + cc];
              if (LOGITS) v = __fdiv_rn(expf(v - __shfl_sync(SSDBOX_FULL_MASK, m, rb)), __shfl_sync(SSDBOX_FULL_MASK, my_s, rb));
              if (!((keptmask >> rb) & 1u)) v = 0.0f;
              cnt += (act && v > thr) ? 1 : 0;                // detection.py:48 strict >
            }
            uint32_t base = 0u;
            if (cnt) {
              base = atomicAdd(&a.cnt[(size_t)b * C + c], (uint32_t)cnt);
              if (base >= (uint32_t)a.cap) {                 // list already full: its extra candidates are never read
                if (sb < 32) full |= 1u << sb;
                cnt = 0;
              }
            }
            if (__any_sync(SSDBOX_FULL_MASK, cnt != 0)) {
              h = hb;
              while (h) {
                const int rb = __ffs(h) - 1;
                h &= h - 1;
                float v = st[(size_t)(wrow0 + rb) * C + cc];
                if (LOGITS) v = __fdiv_rn(expf(v - __shfl_sync(SSDBOX_FULL_MASK, m, rb)), __shfl_sync(SSDBOX_FULL_MASK, my_s, rb));
                if (!((keptmask >> rb) & 1u)) v = 0.0f;
                if (cnt && v > thr) {
                  if (base < (uint32_t)a.cap)
                    a.cand[((size_t)b * C + c) * a.cap + base] = ((unsigned long long)f2ord(v) << 32) | (uint32_t)(p0 + rb);
                  ++base;
                }
              }
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&rc.empty[j]);
  }
}

// ------------------------------------------------------------------------------------------------
// NMS on candidates already sorted in visiting order (keys descending)
// ------------------------------------------------------------------------------------------------
struct NmsSmem {
  float4* box;      // [top_k]
  float* area;      // [top_k]
  uint32_t* mask;   // [top_k * W]
  uint16_t* keep;   // [top_k]
  int* cnt;         // [1]
};

static inline size_t nms_smem_bytes(int top_k) {
  int W = (top_k + 31) / 32;
  return (size_t)top_k * 16 + (size_t)top_k * 4 + (size_t)top_k * W * 4 + align_up((size_t)top_k * 2, 16) + 16;
}

__device__ __forceinline__ NmsSmem carve_nms(unsigned char* base, int top_k) {
  NmsSmem s;
  int W = (top_k + 31) / 32;
  s.box = reinterpret_cast<float4*>(base);
  s.area = reinterpret_cast<float*>(base + (size_t)top_k * 16);
  s.mask = reinterpret_cast<uint32_t*>(base + (size_t)top_k * 20);
  s.keep = reinterpret_cast<uint16_t*>(base + (size_t)top_k * 20 + (size_t)top_k * W * 4);
  s.cnt = reinterpret_cast<int*>(base + (size_t)top_k * 20 + (size_t)top_k * W * 4 + (((size_t)top_k * 2 + 15) / 16) * 16);
  return s;
}

// s.box / s.area hold the m sorted boxes.  Greedy sweep of box_utils.py:311-342:
// box j (later in the order) is dropped by a kept box i iff !(IoU(j | i) <= thr).
__device__ __forceinline__ int nms_sweep(const NmsSmem& s, int m, float thr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int W = (m + 31) / 32;
  // suppression bits of row i, words at or right of the diagonal only (box j > i); rows are dealt to
  // the warps round-robin
  for (int i = warp; i < m; i += nwarp) {
    const float4 bi = s.box[i];
    Box I;
    I.x1 = bi.x; I.y1 = bi.y; I.x2 = bi.z; I.y2 = bi.w;
    const float ai = s.area[i];
    for (int wd = i >> 5; wd < W; ++wd) {
      const int j = wd * 32 + lane;
      bool bit = false;
      if (j > i && j < m) {
        const float4 bj = s.box[j];
        Box J;
        J.x1 = bj.x; J.y1 = bj.y; J.x2 = bj.z; J.y2 = bj.w;
        bit = nms_suppresses(I, ai, J, s.area[j], thr);   // :342 keeps IoU <= overlap
      }
      const uint32_t word = __ballot_sync(SSDBOX_FULL_MASK, bit);
      if (lane == 0) s.mask[i * W + wd] = word;
    }
  }
  __syncthreads();
  if (warp == 0) {
    // Greedy sweep, 32 boxes at a time.  Inside a chunk the decision "box i survives" depends on the boxes before
    // it; it is resolved by every lane redundantly on the chunk's 32 x 32 diagonal block, whose rows are fetched
    // with 32 shuffles that do not depend on the running state (they pipeline) -- the dependent chain per box is a
    // test and an OR instead of a shuffle plus a shared-memory load.  The kept boxes of the chunk then OR their
    // rows into the later words (independent loads).
    uint32_t removed = 0u;   // lane l owns bits 32l .. 32l+31
    int cnt = 0;
    for (int c = 0; c < W; ++c) {
      const int base = c * 32;
      const int nb = m - base < 32 ? m - base : 32;
      const uint32_t diag = lane < nb ? s.mask[(base + lane) * W + c] : 0u;
      uint32_t rem = __shfl_sync(SSDBOX_FULL_MASK, removed, c);
      if (nb < 32) rem |= ~0u << nb;                     // beyond the list
      uint32_t kept = 0u;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const uint32_t di = __shfl_sync(SSDBOX_FULL_MASK, diag, i);
        if (!((rem >> i) & 1u)) {
          kept |= 1u << i;
          rem |= di;
        }
      }
      if (lane < nb && ((kept >> lane) & 1u)) s.keep[cnt + __popc(kept & ((1u << lane) - 1u))] = (uint16_t)(base + lane);
      cnt += __popc(kept);
      if (lane > c && lane < W) {
        uint32_t k = kept, acc = 0u;
        while (k) {
          const int i = __ffs(k) - 1;
          k &= k - 1;
          acc |= s.mask[(base + i) * W + lane];
        }
        removed |= acc;
      }
    }
    if (lane == 0) *s.cnt = cnt;
  }
  __syncthreads();
  return *s.cnt;
}

struct DetSegArgs {
  int B, P, C, top_k, cap;
  float nms_thr, conf_thr, var0, var1;
  long long prior_stride;
  const float* loc;
  const float* scores;
  const float* priors;
  const uint8_t* keep;
  RefineArgs rf;         // RefineDet fused: boxes are decoded against anchors refined on the fly, scores masked by the ARM objectness
  uint32_t* cnt;
  unsigned long long* cand;
  uint32_t* ovf_count; // [1] number of (image, class) lists that overflowed
  int32_t* ovf_list;   // [B*C] their segment ids
  uint32_t* big_count; // [1] number of lists rewritten by the overflow select (<= cap candidates each)
  int32_t* big_list;   // [B*C]
  uint32_t* tail_ticket; // [1] CTAs of detect_segment_kernel that are done (the last one resets the bookkeeping)
  int has_fin;           // ssdbox_detect_peers: block 0 of detect_segments_kernel also completes a deferred loss reduction
  PeerFinishArgs fin;
  uint32_t* scratch;   // [kOverflowSlots, P]
  const float* row_m;  // logits mode (nullptr otherwise): softmax row max / denominator from the stream pass
  const float* row_s;
  float* out;
  int32_t* counts;
};

// keys[0..n) hold the candidates (unsorted; n <= npad, npad power of two, padding = 0):
// sort -> top_k -> decode -> NMS -> write the [top_k,5] rows of this (image, class).
__device__ void segment_finish(const DetSegArgs& a, int b, int seg, unsigned long long* keys, int n, int npad,
                               const NmsSmem& ns) {
  const int tid = threadIdx.x;
  bitonic_sort_desc(keys, npad);
  const int m = n < a.top_k ? n : a.top_k;
  const float* pri = a.priors + (size_t)b * (size_t)a.prior_stride;
  for (int i = tid; i < m; i += blockDim.x) {
    uint32_t p = (uint32_t)(keys[i] & 0xffffffffull);
    Box bx = decode_box(*reinterpret_cast<const float4*>(a.loc + ((size_t)b * a.P + p) * 4),
                        refine_center(a.rf, *reinterpret_cast<const float4*>(pri + (size_t)p * 4), (size_t)b * a.P + p),
                        a.var0, a.var1);   // detection.py:43
    ns.box[i] = make_float4(bx.x1, bx.y1, bx.x2, bx.y2);
    ns.area[i] = box_area(bx);                                                                    // box_utils.py:298
  }
  __syncthreads();
  const int count = nms_sweep(ns, m, a.nms_thr);
  float* o = a.out + (size_t)seg * a.top_k * 5;
  for (int e = tid; e < a.top_k * 5; e += blockDim.x) {
    int k = e / 5, f = e - k * 5;
    float v = 0.0f;
    if (k < count) {
      int i = ns.keep[k];
      if (f == 0) v = ord2f((uint32_t)(keys[i] >> 32));
      else {
        float4 bx = ns.box[i];
        v = f == 1 ? bx.x : (f == 2 ? bx.y : (f == 3 ? bx.z : bx.w));
      }
    }
    o[e] = v;                                                    // detection.py:57-59
  }
  if (tid == 0 && a.counts) a.counts[seg] = count;
}

// ------------------------------------------------------------------------------------------------
// (image, class) segments.  Small segments (<= 32 candidates: the normal case with a trained
// detector, ~12 on average at SSD512-COCO) are finished by ONE WARP entirely in registers:
// shuffle bitonic sort, decode, suppression bits via broadcast + ballot sweep.  Larger ones are
// queued for the CTA-wide kernel (<= capacity) or the overflow kernel (> capacity).
// ------------------------------------------------------------------------------------------------
constexpr int kSmallThreads = 256;       // 8 warps = 8 (image, class) segments per CTA
constexpr int kSmallSegs = kSmallThreads / 32;

__device__ __forceinline__ void warp_zero_rows(float* o, int nfloat, int lane) {
  if ((reinterpret_cast<uintptr_t>(o) & 15u) == 0 && (nfloat & 3) == 0) {
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = lane; e < (nfloat >> 2); e += 32) reinterpret_cast<float4*>(o)[e] = z;
  } else {
    for (int e = lane; e < nfloat; e += 32) o[e] = 0.0f;
  }
}

// One launch finishes every list that fits the candidate capacity:
//   phase 1  one warp per (image, class): lists of <= 32 candidates entirely in registers (the normal case);
//            longer lists are noted in a CTA-local queue, overflowed ones (> capacity, dense scores) in the
//            global overflow list for detect_overflow_chunk_kernel
//   phase 2  the CTA works off its own queue (sort in shared memory, decode, bit-matrix NMS)
// The kernel also leaves the workspace state clean for the next call: the counter of every list it finishes
// goes back to zero (SSDBOX_DETECT_WS_CLEAN, include/ssdbox.h).
__global__ void __launch_bounds__(kSmallThreads) detect_segments_kernel(DetSegArgs a) {
  extern __shared__ __align__(16) unsigned char smem_seg[];
  __shared__ int s_q[kSmallSegs], s_qn[kSmallSegs], s_nq;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_nq = 0;
  // multi-GPU step: the wait for the other ranks' loss sums (posted by the mining kernel a whole detect_stream
  // ago) rides on this launch instead of a kernel of its own (ssdbox_detect_peers)
  if (a.has_fin && blockIdx.x == 0 && threadIdx.x < 32) peer_finish_warp(a.fin, threadIdx.x);
  __syncthreads();
  const int seg = blockIdx.x * kSmallSegs + (threadIdx.x >> 5);
  if (seg < a.B * a.C) {
    const int b = seg / a.C, c = seg - b * a.C;
    const uint32_t total = c == 0 ? 0u : a.cnt[seg];
    // the first 32 candidates are requested together with the counter (one round trip instead of two)
    const unsigned long long key_first = a.cand[(size_t)seg * a.cap + lane];
    if (total > 32u) {
      if (lane == 0) {
        if (total > (uint32_t)a.cap) {
          a.ovf_list[atomicAdd(a.ovf_count, 1u)] = seg;      // its counter stays: the overflow kernels read it
        } else {
          const int q = atomicAdd(&s_nq, 1);
          s_q[q] = seg;
          s_qn[q] = (int)total;
          a.cnt[seg] = 0u;
        }
      }
    } else {
      float* o = a.out + (size_t)seg * a.top_k * 5;
      warp_zero_rows(o, a.top_k * 5, lane);    // background plane / no candidate: zeros (detection.py:37,50-51)
      if (total == 0u) {
        if (lane == 0 && a.counts) a.counts[seg] = 0;
      } else {
        if (lane == 0) a.cnt[seg] = 0u;
        const int n = (int)total;
        unsigned long long key = lane < n ? key_first : 0ull;
        // bitonic sort, descending (= visiting order: score desc, higher prior first); padding sinks
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
          for (int j = k >> 1; j > 0; j >>= 1) {
            unsigned long long other = __shfl_xor_sync(SSDBOX_FULL_MASK, key, j);
            bool desc = (lane & k) == 0;
            bool lower = (lane & j) == 0;
            bool keep_max = lower == desc;
            key = keep_max ? (key > other ? key : other) : (key < other ? key : other);
          }
        }
        const int m = n < a.top_k ? n : a.top_k;           // box_utils.py:301 idx[-top_k:]
        Box me;
        me.x1 = me.y1 = me.x2 = me.y2 = 0.f;
        if (lane < m) {
          uint32_t p = (uint32_t)(key & 0xffffffffull);
          me = decode_box(*reinterpret_cast<const float4*>(a.loc + ((size_t)b * a.P + p) * 4),
                          refine_center(a.rf, *reinterpret_cast<const float4*>(a.priors + (size_t)b * (size_t)a.prior_stride + (size_t)p * 4),
                                        (size_t)b * a.P + p),
                          a.var0, a.var1);
        }
        const float my_area = box_area(me);
        uint32_t supp_by = 0u;                              // bit i: box i (earlier in the order) suppresses me
        for (int i = 0; i + 1 < m; ++i) {
          Box bi;
          bi.x1 = __shfl_sync(SSDBOX_FULL_MASK, me.x1, i);
          bi.y1 = __shfl_sync(SSDBOX_FULL_MASK, me.y1, i);
          bi.x2 = __shfl_sync(SSDBOX_FULL_MASK, me.x2, i);
          bi.y2 = __shfl_sync(SSDBOX_FULL_MASK, me.y2, i);
          float ai = __shfl_sync(SSDBOX_FULL_MASK, my_area, i);
          if (lane > i && lane < m && !(iou_nms(bi, ai, me, my_area) <= a.nms_thr)) supp_by |= 1u << i;
        }
        uint32_t removed = 0u, kept = 0u;
        for (int i = 0; i < m; ++i) {
          uint32_t col = __ballot_sync(SSDBOX_FULL_MASK, (supp_by >> i) & 1u);
          if (!((removed >> i) & 1u)) {
            kept |= 1u << i;
            removed |= col;
          }
        }
        __syncwarp();                                        // orders the zero fill before the row writes
        if (lane < m && ((kept >> lane) & 1u)) {
          float* r = o + __popc(kept & ((1u << lane) - 1u)) * 5;
          r[0] = ord2f((uint32_t)(key >> 32));
          r[1] = me.x1; r[2] = me.y1; r[3] = me.x2; r[4] = me.y2;
        }
        if (lane == 0 && a.counts) a.counts[seg] = __popc(kept);
      }
    }
  }
  __syncthreads();
  const int nq = s_nq;
  if (nq == 0) return;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_seg);
  NmsSmem ns = carve_nms(smem_seg + (size_t)a.cap * 8, a.top_k);
  for (int q = 0; q < nq; ++q) {
    const int sg = s_q[q], n = s_qn[q];                    // 32 < n <= cap
    int npad = 64;
    while (npad < n) npad <<= 1;
    const unsigned long long* src = a.cand + (size_t)sg * a.cap;
    for (int i = threadIdx.x; i < npad; i += kSmallThreads) keys[i] = i < n ? src[i] : 0ull;
    __syncthreads();
    segment_finish(a, sg / a.C, sg, keys, n, npad, ns);
    __syncthreads();
  }
}

constexpr int kSegThreads = 256;

// Lists rewritten by the overflow select (dense scores).  Exits at once when nothing overflowed -- the normal case;
// otherwise the CTA that finishes last resets the overflow bookkeeping (the workspace state is clean again).
__global__ void __launch_bounds__(kSegThreads) detect_segment_kernel(DetSegArgs a) {
  extern __shared__ __align__(16) unsigned char smem_seg[];
  __shared__ int s_last;
  if (*a.ovf_count == 0u) return;                     // written by detect_segments_kernel; nothing to do, nothing to clean
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_seg);
  NmsSmem ns = carve_nms(smem_seg + (size_t)a.cap * 8, a.top_k);
  const int tid = threadIdx.x;
  const int nbig = (int)*a.big_count;                 // written by detect_overflow_chunk_kernel
  for (int w = blockIdx.x; w < nbig; w += gridDim.x) {
    const int seg = a.big_list[w];
    const int b = seg / a.C;
    const int n = (int)a.cnt[seg];                    // <= cap
    int npad = 64;
    while (npad < n) npad <<= 1;
    const unsigned long long* src = a.cand + (size_t)seg * a.cap;
    for (int i = tid; i < npad; i += kSegThreads) keys[i] = i < n ? src[i] : 0ull;
    __syncthreads();
    if (tid == 0) a.cnt[seg] = 0u;
    segment_finish(a, b, seg, keys, n, npad, ns);
    __syncthreads();
  }
  if (tid == 0) {
    __threadfence();
    const unsigned t = atomicAdd(a.tail_ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && tid == 0) {
    *a.ovf_count = 0u;
    *a.big_count = 0u;
    *a.tail_ticket = 0u;
  }
}

constexpr int kOvfThreads = 1024;

__global__ void __launch_bounds__(kOvfThreads, 1) detect_overflow_kernel(DetSegArgs a) {
  extern __shared__ __align__(16) unsigned char smem_ovf[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_ovf);          // [1024]
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_ovf + 8192);                      // 2048
  int* s_iscr = reinterpret_cast<int*>(smem_ovf + 16384);                               // 64
  int* s_res = s_iscr + 64;                                                             // 8: res[0..1], [4] = counter
  NmsSmem ns = carve_nms(smem_ovf + 16384 + 288, a.top_k);
  uint32_t* uk = a.scratch + (size_t)blockIdx.x * a.P;
  const int tid = threadIdx.x;
  const int novf = (int)*a.ovf_count;      // written by detect_segment_kernel (previous launch)
  for (int w = blockIdx.x; w < novf; w += gridDim.x) {
    const int seg = a.ovf_list[w];
    const int b = seg / a.C, c = seg - b * a.C;
    // ordered scores of the class column; 0 = not a candidate
    for (int p = tid; p < a.P; p += kOvfThreads) {
      size_t row = (size_t)b * a.P + p;
      float v = a.scores[row * a.C + c];
      if (a.row_s) {           // logits: the same expression as the candidate pass
        const float sden = a.row_s[row];
        v = sden > 0.0f ? __fdiv_rn(expf(v - a.row_m[row]), sden) : 0.0f;
      }
      if (!refine_member(a.rf, a.keep, row)) v = 0.0f;
      uk[p] = v > a.conf_thr ? f2ord(v) : 0u;
    }
    if (tid == 0) s_res[4] = 0;
    __syncthreads();
    uint32_t Tu = cta_select_threshold<true>(uk, a.P, a.top_k, nullptr, s_hist, s_iscr, s_res);
    __syncthreads();
    for (int p = tid; p < a.P; p += kOvfThreads) {
      uint32_t u = uk[p];
      if (u != 0u && u >= Tu) {
        int slot = atomicAdd(&s_res[4], 1);
        if (slot < 1024) keys[slot] = ((unsigned long long)u << 32) | (uint32_t)p;
      }
    }
    __syncthreads();
    int n = s_res[4];
    if (n > 1024) n = 1024;
    int npad = 32;
    while (npad < n) npad <<= 1;
    for (int i = n + tid; i < npad; i += kOvfThreads) keys[i] = 0ull;
    __syncthreads();
    segment_finish(a, b, seg, keys, n, npad, ns);
    if (tid == 0) a.cnt[seg] = 0u;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Dense scores, chunked: one work item = (image, 8 consecutive classes).  The per-(image, class)
// kernel above reads a strided score column per segment (24564 sectors for 24564 floats); here the 8
// classes of a chunk share the sectors of every row, their radix-select histograms are built in the
// same passes (8 x 2048 bins in shared memory, one warp per class finds the digit), and all
// keys >= the class threshold are collected and sorted: the sort order (score desc, prior desc) IS
// the reference's visiting order, so ties at the cut need no special handling as long as they fit the
// 1024-entry list; a class whose ties do not fit (e.g. all scores equal) takes the column path.
// ------------------------------------------------------------------------------------------------
constexpr int kOvfClasses = 8;

__device__ __forceinline__ float ovf_score(const DetSegArgs& a, size_t row, int c) {
  float v = a.scores[row * a.C + c];
  if (a.row_s) {           // logits: the same expression as the candidate pass
    const float sden = a.row_s[row];
    v = sden > 0.0f ? __fdiv_rn(expf(v - a.row_m[row]), sden) : 0.0f;
  }
  if (!refine_member(a.rf, a.keep, row)) v = 0.0f;
  return v;
}

// the scores of classes c0 .. c0+7 of one row: all loads issued together, row-wise state read once
__device__ __forceinline__ void ovf_row_scores(const DetSegArgs& a, size_t row, int c0, float (&v)[kOvfClasses]) {
  const float* x = a.scores + row * a.C + c0;
#pragma unroll
  for (int k = 0; k < kOvfClasses; ++k) v[k] = c0 + k < a.C ? x[k] : 0.0f;
  const bool kept = refine_member(a.rf, a.keep, row);
  if (a.row_s) {           // logits: the same expression as the candidate pass
    const float sden = a.row_s[row], m = a.row_m[row];
#pragma unroll
    for (int k = 0; k < kOvfClasses; ++k) v[k] = sden > 0.0f ? __fdiv_rn(expf(v[k] - m), sden) : 0.0f;
  }
  if (!kept) {
#pragma unroll
    for (int k = 0; k < kOvfClasses; ++k) v[k] = 0.0f;
  }
}

// one warp: digit d with  above = sum_{bin > d} hist < K <= above + hist[d]  (d = -1: fewer than K entries)
__device__ __forceinline__ void warp_find_digit(const uint32_t* hist, int nbins, int K, int lane, int* d_out, int* above_out) {
  const int span = nbins / 32;
  const int hi = nbins - 1 - lane * span;       // lane owns bins hi, hi-1, ..., hi-span+1 (descending)
  int local = 0;
  for (int e = 0; e < span; ++e) local += (int)hist[hi - e];
  const int incl = warp_inclusive_scan(local, lane);
  int above = incl - local;
  const bool mine = above < K && incl >= K;
  int d = -1;
  if (mine) {
    for (int e = 0; e < span; ++e) {
      const int h = (int)hist[hi - e];
      if (above + h >= K) {
        d = hi - e;
        break;
      }
      above += h;
    }
  }
  const uint32_t who = __ballot_sync(SSDBOX_FULL_MASK, mine);
  if (who == 0u) {
    *d_out = -1;
    *above_out = 0;
    return;
  }
  const int src = __ffs(who) - 1;
  *d_out = __shfl_sync(SSDBOX_FULL_MASK, d, src);
  *above_out = __shfl_sync(SSDBOX_FULL_MASK, above, src);
}

// the per-segment column path (any tie pattern): ordered scores -> scratch, exact select, collect
__device__ void overflow_column_segment(const DetSegArgs& a, int b, int c, int seg, unsigned long long* keys,
                                        uint32_t* s_hist, int* s_iscr, int* s_res, const NmsSmem& ns, uint32_t* uk) {
  const int tid = threadIdx.x;
  for (int p = tid; p < a.P; p += kOvfThreads) {
    float v = ovf_score(a, (size_t)b * a.P + p, c);
    uk[p] = v > a.conf_thr ? f2ord(v) : 0u;
  }
  if (tid == 0) s_res[4] = 0;
  __syncthreads();
  uint32_t Tu = cta_select_threshold<true>(uk, a.P, a.top_k, nullptr, s_hist, s_iscr, s_res);
  __syncthreads();
  for (int p = tid; p < a.P; p += kOvfThreads) {
    uint32_t u = uk[p];
    if (u != 0u && u >= Tu) {
      int slot = atomicAdd(&s_res[4], 1);
      if (slot < 1024) keys[slot] = ((unsigned long long)u << 32) | (uint32_t)p;
    }
  }
  __syncthreads();
  int n = s_res[4];
  if (n > 1024) n = 1024;
  int npad = 32;
  while (npad < n) npad <<= 1;
  for (int i = n + tid; i < npad; i += kOvfThreads) keys[i] = 0ull;
  __syncthreads();
  segment_finish(a, b, seg, keys, n, npad, ns);
  if (tid == 0) a.cnt[seg] = 0u;
  __syncthreads();
}

__global__ void __launch_bounds__(kOvfThreads, 1) detect_overflow_chunk_kernel(DetSegArgs a) {
  extern __shared__ __align__(16) unsigned char smem_ovc[];
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_ovc);                                   // [8][2048]
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem_ovc + 65536);         // [8][1024]
  int* s_iscr = reinterpret_cast<int*>(smem_ovc + 131072);                                      // 64
  int* s_res = s_iscr + 64;                                                                     // 8
  int* s_cls = s_res + 8;                                                                       // 8 x {active, K left, range lo, n collected, range size - 1, shift}
  NmsSmem ns = carve_nms(smem_ovc + 131072 + 288 + 192, a.top_k);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nchunk = (a.C - 1 + kOvfClasses - 1) / kOvfClasses;
  int* s_act = s_cls;
  int* s_kleft = s_cls + 8;
  uint32_t* s_pre = reinterpret_cast<uint32_t*>(s_cls + 16);     // key prefix fixed so far, finally the threshold key
  int* s_n = s_cls + 24;
  uint32_t* s_spanm1 = reinterpret_cast<uint32_t*>(s_cls + 32);   // current range size - 1 (0: the threshold key is s_pre)
  int* s_shift = s_cls + 40;
  uint32_t* uk = a.scratch + (size_t)blockIdx.x * a.P;
  if (*a.ovf_count == 0u) return;       // no list overflowed (the normal case): written by detect_segment_small_kernel
  for (int item = blockIdx.x; item < a.B * nchunk; item += gridDim.x) {
    const int b = item / nchunk, c0 = 1 + (item - b * nchunk) * kOvfClasses;
    __syncthreads();
    if (tid < kOvfClasses) {
      const int c = c0 + tid;
      s_act[tid] = (c < a.C && a.cnt[(size_t)b * a.C + c] > (uint32_t)a.cap) ? 1 : 0;
      s_kleft[tid] = a.top_k;
      const uint32_t lo = f2ord(a.conf_thr), one = f2ord(1.0f);
      const uint32_t nominal = one > lo ? one - lo : 0u;       // keys of the scores in (conf_thr, 1]
      const int bits = nominal ? 32 - __clz(nominal) : 0;
      s_pre[tid] = lo;
      s_spanm1[tid] = 0xffffffffu - lo;                        // first level: everything above the threshold
      s_shift[tid] = bits > 11 ? bits - 11 : 0;                // (nominal >> shift) <= 2047: the last bin also catches keys above 1.0
      s_n[tid] = 0;
    }
    __syncthreads();
    int actmask = 0;
#pragma unroll
    for (int k = 0; k < kOvfClasses; ++k) actmask |= s_act[k] << k;
    if (actmask == 0) continue;

    // Range-adaptive radix select, all active classes per pass: a level buckets the keys of the current
    // range [lo, lo + span) into 2048 bins of 2^shift keys.  The first range is (key(conf_thr), key(1.0)]
    // -- softmax scores -- with the last bin catching anything above, so the bins are spread over the
    // scores' log-range (bucketing the top key bits would put every score in ~30 bins and serialise the
    // shared-memory atomics).  Three levels in the normal case, at most five.
    for (int level = 0; level < 5; ++level) {
      if (level > 0) {
        int pending = 0;
#pragma unroll
        for (int k = 0; k < kOvfClasses; ++k) pending |= ((actmask >> k) & 1) && s_spanm1[k] != 0u;
        if (!pending) break;                      // uniform: shared state read after a barrier
      }
      for (int i = tid; i < kOvfClasses * 2048; i += kOvfThreads) s_hist[i] = 0u;
      __syncthreads();
      uint32_t lo_r[kOvfClasses], spanm1_r[kOvfClasses];
      int shift_r[kOvfClasses];
#pragma unroll
      for (int k = 0; k < kOvfClasses; ++k) {
        lo_r[k] = s_pre[k];
        spanm1_r[k] = ((actmask >> k) & 1) ? s_spanm1[k] : 0u;     // 0: class inactive or already resolved
        shift_r[k] = s_shift[k];
      }
      for (int p = tid; p < a.P; p += kOvfThreads) {
        float v[kOvfClasses];
        ovf_row_scores(a, (size_t)b * a.P + p, c0, v);
#pragma unroll
        for (int k = 0; k < kOvfClasses; ++k) {
          const uint32_t u = f2ord(v[k]);
          const uint32_t off = u - lo_r[k];
          if (spanm1_r[k] != 0u && v[k] > a.conf_thr && u >= lo_r[k] && off <= spanm1_r[k]) {
            const uint32_t bin = off >> shift_r[k];
            atomicAdd(&s_hist[k * 2048 + (bin < 2047u ? bin : 2047u)], 1u);
          }
        }
      }
      __syncthreads();
      if (warp < kOvfClasses && ((actmask >> warp) & 1) && s_spanm1[warp] != 0u) {
        int d, above;
        warp_find_digit(s_hist + warp * 2048, 2048, s_kleft[warp], lane, &d, &above);
        if (lane == 0) {
          if (d < 0) {               // cannot happen for an overflowed list (more than cap > top_k candidates)
            d = 0;
            above = 0;
          }
          s_kleft[warp] -= above;
          const int sh = s_shift[warp];
          const uint32_t lo = s_pre[warp] + ((uint32_t)d << sh);
          // bin 2047 of the first level is open-ended; every other bin holds exactly 2^shift keys
          uint32_t spanm1 = (level == 0 && d == 2047) ? (0xffffffffu - lo) : ((1u << sh) - 1u);
          // everything above the bin plus the whole bin fits the list: stop refining, the sort of
          // segment_finish orders the bin and keeps the first top_k
          if ((a.top_k - s_kleft[warp]) + (int)s_hist[warp * 2048 + d] <= 1024) spanm1 = 0u;
          int bits = spanm1 ? 32 - __clz(spanm1) : 0;
          s_pre[warp] = lo;
          s_spanm1[warp] = spanm1;
          s_shift[warp] = bits > 11 ? bits - 11 : 0;
        }
      }
      __syncthreads();
    }
    // s_pre[k] is now a key <= the top_k-th largest score with at most 1024 keys >= it (or the exact
    // key of the top_k-th largest score): collect every key >= it
    uint32_t thr_r[kOvfClasses];
#pragma unroll
    for (int k = 0; k < kOvfClasses; ++k) thr_r[k] = ((actmask >> k) & 1) ? s_pre[k] : 0xffffffffu;
    for (int p = tid; p < a.P; p += kOvfThreads) {
      float v[kOvfClasses];
      ovf_row_scores(a, (size_t)b * a.P + p, c0, v);
#pragma unroll
      for (int k = 0; k < kOvfClasses; ++k) {
        const uint32_t u = f2ord(v[k]);
        if (v[k] > a.conf_thr && u >= thr_r[k] && thr_r[k] != 0xffffffffu) {
          const int slot = atomicAdd(&s_n[k], 1);
          if (slot < 1024) s_keys[k * 1024 + slot] = ((unsigned long long)u << 32) | (uint32_t)p;
        }
      }
    }
    __syncthreads();
    for (int k = 0; k < kOvfClasses; ++k) {
      if (!((actmask >> k) & 1)) continue;
      const int c = c0 + k, seg = b * a.C + c;
      const int n = s_n[k];
      if (n > 1024) {              // ties at the cut do not fit: exact index-ordered selection on the column
        overflow_column_segment(a, b, c, seg, s_keys + k * 1024, s_hist, s_iscr, s_res, ns, uk);
        continue;
      }
      // hand the selected keys to detect_segment_kernel (next launch): several of its small CTAs share an
      // SM, so the serial NMS sweep of one segment overlaps the IoU matrix of others
      const unsigned long long* keys = s_keys + k * 1024;
      unsigned long long* dst = a.cand + (size_t)seg * a.cap;
      for (int i = tid; i < n; i += kOvfThreads) dst[i] = keys[i];
      if (tid == 0) {
        a.cnt[seg] = (uint32_t)n;
        a.big_list[atomicAdd(a.big_count, 1u)] = seg;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// stand-alone nms(boxes, scores, overlap, top_k), box_utils.py:279-343 -- one CTA
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kOvfThreads, 1)
nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, int n, float thr, int top_k,
           int64_t* __restrict__ keep, int32_t* __restrict__ count, uint32_t* uk) {
  extern __shared__ __align__(16) unsigned char smem_nms[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_nms);
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_nms + 8192);
  int* s_iscr = reinterpret_cast<int*>(smem_nms + 16384);
  int* s_res = s_iscr + 64;
  NmsSmem ns = carve_nms(smem_nms + 16384 + 288, top_k);
  const int tid = threadIdx.x;
  for (int i = tid; i < n; i += kOvfThreads) {
    uk[i] = f2ord(scores[i]);
    keep[i] = 0;                                           // :291 zero-initialised keep
  }
  if (tid == 0) s_res[4] = 0;
  __syncthreads();
  const int K = n < top_k ? n : top_k;                     // :301 idx[-top_k:]
  uint32_t Tu = cta_select_threshold<true>(uk, n, K, nullptr, s_hist, s_iscr, s_res);
  __syncthreads();
  for (int i = tid; i < n; i += kOvfThreads) {
    uint32_t u = uk[i];
    if (u != 0u && u >= Tu) {
      int slot = atomicAdd(&s_res[4], 1);
      if (slot < 1024) keys[slot] = ((unsigned long long)u << 32) | (uint32_t)i;
    }
  }
  __syncthreads();
  int m = s_res[4];
  if (m > 1024) m = 1024;
  int npad = 32;
  while (npad < m) npad <<= 1;
  for (int i = m + tid; i < npad; i += kOvfThreads) keys[i] = 0ull;
  __syncthreads();
  bitonic_sort_desc(keys, npad);
  for (int i = tid; i < m; i += kOvfThreads) {
    uint32_t idx = (uint32_t)(keys[i] & 0xffffffffull);
    const float* bp = boxes + (size_t)idx * 4;
    Box bx;
    bx.x1 = bp[0]; bx.y1 = bp[1]; bx.x2 = bp[2]; bx.y2 = bp[3];
    ns.box[i] = make_float4(bx.x1, bx.y1, bx.x2, bx.y2);
    ns.area[i] = box_area(bx);
  }
  __syncthreads();
  const int cnt = nms_sweep(ns, m, thr);
  for (int k = tid; k < cnt; k += kOvfThreads) keep[k] = (int64_t)(keys[ns.keep[k]] & 0xffffffffull);
  if (tid == 0) *count = cnt;
}

// ------------------------------------------------------------------------------------------------
// Any top_k (> kTopKLimit): the reference takes whatever top_k it is given (box_utils.py:299-301).  The fast paths
// above keep a list's sort keys and its top_k x top_k suppression bits in shared memory; beyond 1024 that does not fit,
// and nobody runs SSD with such a top_k for speed.  These kernels trade speed for generality: one CTA per list,
// the keys sorted in GLOBAL memory (same bitonic network), and a greedy sweep that needs no matrix -- 32 boxes at a
// time, the kept ones of a chunk are tested against every later box that is still alive.  Same visiting order, same
// suppression test, same results as the fast paths wherever both apply (tests compare them at top_k <= 1024 too).
// ------------------------------------------------------------------------------------------------
constexpr int kLargeThreads = 1024;

struct LargeSmem {
  uint32_t hist[2048];
  uint32_t removed[kLargeTopKLimit / 32];
  float4 kbox[32];
  float karea[32];
  int iscr[64 + 8];
  uint32_t kept;
  int cnt, m;
};

// box / area: the m boxes in visiting order (global memory).  keep[0..cnt) <- their positions, returns cnt.
__device__ int nms_sweep_stream(const float4* __restrict__ box, const float* __restrict__ area, int m, float thr,
                                int32_t* __restrict__ keep, LargeSmem& s) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x;
  const int W = (m + 31) >> 5;
  for (int w = tid; w < W; w += T) s.removed[w] = 0u;
  if (tid == 0) s.cnt = 0;
  __syncthreads();
  for (int c = 0; c < W; ++c) {
    const int base = c << 5;
    const int nb = m - base < 32 ? m - base : 32;
    if (warp == 0) {
      float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
      float aj = 0.f;
      if (lane < nb) {
        bj = box[base + lane];
        aj = area[base + lane];
      }
      Box J;
      J.x1 = bj.x; J.y1 = bj.y; J.x2 = bj.z; J.y2 = bj.w;
      uint32_t rem = s.removed[c];                       // warp-uniform from here on
      if (nb < 32) rem |= ~0u << nb;
      uint32_t kept = 0u;
      for (int i = 0; i < nb; ++i) {
        if ((rem >> i) & 1u) continue;
        kept |= 1u << i;
        Box I;
        I.x1 = __shfl_sync(SSDBOX_FULL_MASK, bj.x, i); I.y1 = __shfl_sync(SSDBOX_FULL_MASK, bj.y, i);
        I.x2 = __shfl_sync(SSDBOX_FULL_MASK, bj.z, i); I.y2 = __shfl_sync(SSDBOX_FULL_MASK, bj.w, i);
        const float ai = __shfl_sync(SSDBOX_FULL_MASK, aj, i);
        const bool bit = lane > i && lane < nb && nms_suppresses(I, ai, J, aj, thr);     // box_utils.py:342 keeps IoU <= overlap
        rem |= __ballot_sync(SSDBOX_FULL_MASK, bit);
      }
      const int cnt = s.cnt;
      if ((kept >> lane) & 1u) {
        const int k = __popc(kept & ((1u << lane) - 1u));
        keep[cnt + k] = base + lane;
        s.kbox[k] = bj;
        s.karea[k] = aj;
      }
      __syncwarp();
      if (lane == 0) {
        s.kept = kept;
        s.cnt = cnt + __popc(kept);
      }
    }
    __syncthreads();
    const int nk = __popc(s.kept);
    for (int j = base + 32 + tid; j < m && nk > 0; j += T) {
      if ((s.removed[j >> 5] >> (j & 31)) & 1u) continue;
      const float4 b4 = box[j];
      Box J;
      J.x1 = b4.x; J.y1 = b4.y; J.x2 = b4.z; J.y2 = b4.w;
      const float aj = area[j];
      bool sup = false;
      for (int k = 0; k < nk && !sup; ++k) {
        const float4 bi = s.kbox[k];
        Box I;
        I.x1 = bi.x; I.y1 = bi.y; I.x2 = bi.z; I.y2 = bi.w;
        sup = nms_suppresses(I, s.karea[k], J, aj, thr);
      }
      if (sup) atomicOr(&s.removed[j >> 5], 1u << (j & 31));
    }
    __syncthreads();
  }
  return s.cnt;
}

// uk[0..n): ordered keys (0 = not a candidate).  Leaves the min(K, candidates) best in keys[0..) sorted in visiting
// order (key descending, higher index first) and returns their number.
__device__ int large_select_sort(uint32_t* uk, int n, int K, unsigned long long* keys, LargeSmem& s) {
  const int tid = threadIdx.x, T = blockDim.x;
  if (tid == 0) s.m = 0;
  __syncthreads();
  const uint32_t Tu = cta_select_threshold<true>(uk, n, K, nullptr, s.hist, s.iscr, s.iscr + 64);
  __syncthreads();
  for (int i = tid; i < n; i += T) {
    const uint32_t u = uk[i];
    if (u != 0u && u >= Tu) keys[atomicAdd(&s.m, 1)] = ((unsigned long long)u << 32) | (uint32_t)i;
  }
  __syncthreads();
  const int m = s.m;
  int npad = 32;
  while (npad < m) npad <<= 1;
  for (int i = m + tid; i < npad; i += T) keys[i] = 0ull;
  bitonic_sort_desc(keys, npad);
  return m < K ? m : K;
}

// stand-alone nms, any top_k: ws = uk [n] | keys [pow2 >= n] | box [K] | area [K] | pos [K]
__global__ void __launch_bounds__(kLargeThreads, 1)
nms_large_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, int n, float thr, int top_k,
                 int64_t* __restrict__ keep, int32_t* __restrict__ count, uint32_t* uk, unsigned long long* keys,
                 float4* gbox, float* garea, int32_t* gpos) {
  __shared__ LargeSmem s;
  const int tid = threadIdx.x;
  for (int i = tid; i < n; i += kLargeThreads) {
    uk[i] = f2ord(scores[i]);
    keep[i] = 0;                                           // box_utils.py:291 zero-initialised keep
  }
  __syncthreads();
  const int K = n < top_k ? n : top_k;                     // :301 idx[-top_k:]
  const int m = large_select_sort(uk, n, K, keys, s);
  for (int i = tid; i < m; i += kLargeThreads) {
    const float* bp = boxes + (size_t)(keys[i] & 0xffffffffull) * 4;
    Box bx;
    bx.x1 = bp[0]; bx.y1 = bp[1]; bx.x2 = bp[2]; bx.y2 = bp[3];
    gbox[i] = make_float4(bx.x1, bx.y1, bx.x2, bx.y2);
    garea[i] = box_area(bx);
  }
  __syncthreads();
  const int cnt = nms_sweep_stream(gbox, garea, m, thr, gpos, s);
  for (int k = tid; k < cnt; k += kLargeThreads) keep[k] = (int64_t)(keys[gpos[k]] & 0xffffffffull);
  if (tid == 0) *count = cnt;
}

struct DetLargeArgs {
  int B, P, C, top_k;
  float nms_thr, conf_thr, var0, var1;
  long long prior_stride;
  const float* loc;
  const float* scores;
  const float* priors;
  const uint8_t* keep;
  RefineArgs rf;
  float* out;
  int32_t* counts;
  unsigned char* ws;
  size_t slice_bytes, off_keys, off_box, off_area, off_pos;
};

// DetectOut, any top_k: persistent CTAs, one (image, class) list at a time, each CTA in its own workspace slice
__global__ void __launch_bounds__(kLargeThreads, 1) detect_large_kernel(DetLargeArgs a) {
  __shared__ LargeSmem s;
  const int tid = threadIdx.x;
  unsigned char* mine = a.ws + (size_t)blockIdx.x * a.slice_bytes;
  uint32_t* uk = reinterpret_cast<uint32_t*>(mine);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(mine + a.off_keys);
  float4* gbox = reinterpret_cast<float4*>(mine + a.off_box);
  float* garea = reinterpret_cast<float*>(mine + a.off_area);
  int32_t* gpos = reinterpret_cast<int32_t*>(mine + a.off_pos);
  for (int seg = blockIdx.x; seg < a.B * a.C; seg += gridDim.x) {
    const int b = seg / a.C, c = seg - b * a.C;
    float* o = a.out + (size_t)seg * a.top_k * 5;
    int cnt = 0;
    if (c > 0 && a.P > 0) {                                // detection.py:45 for cl in range(1, num_classes)
      for (int p = tid; p < a.P; p += kLargeThreads) {
        const size_t row = (size_t)b * a.P + p;
        const float v = a.scores[row * a.C + c];
        const bool cand = v > a.conf_thr && refine_member(a.rf, a.keep, row);       // detection.py:47 c_mask = scores.gt(conf_thresh)
        uk[p] = cand ? f2ord(v) : 0u;
      }
      __syncthreads();
      const int m = large_select_sort(uk, a.P, a.top_k, keys, s);
      const float* pri = a.priors + (size_t)b * (size_t)a.prior_stride;
      for (int i = tid; i < m; i += kLargeThreads) {
        const uint32_t p = (uint32_t)(keys[i] & 0xffffffffull);
        const Box bx = decode_box(*reinterpret_cast<const float4*>(a.loc + ((size_t)b * a.P + p) * 4),
                                  refine_center(a.rf, *reinterpret_cast<const float4*>(pri + (size_t)p * 4), (size_t)b * a.P + p),
                                  a.var0, a.var1);         // detection.py:43
        gbox[i] = make_float4(bx.x1, bx.y1, bx.x2, bx.y2);
        garea[i] = box_area(bx);
      }
      __syncthreads();
      cnt = nms_sweep_stream(gbox, garea, m, a.nms_thr, gpos, s);
    }
    for (int e = tid; e < a.top_k * 5; e += kLargeThreads) {
      const int k = e / 5, f = e - k * 5;
      float v = 0.0f;
      if (k < cnt) {
        const int i = gpos[k];
        if (f == 0) v = ord2f((uint32_t)(keys[i] >> 32));
        else {
          const float4 bx = gbox[i];
          v = f == 1 ? bx.x : (f == 2 ? bx.y : (f == 3 ? bx.z : bx.w));
        }
      }
      o[e] = v;                                            // detection.py:57-59
    }
    if (tid == 0 && a.counts) a.counts[seg] = cnt;
    __syncthreads();
  }
}

}  // namespace ssdbox

using namespace ssdbox;

extern "C" int ssdbox_detect(const ssdbox_detect_cfg* cfg, const float* loc, const float* scores, const float* priors,
                             const uint8_t* score_keep, float* out, int32_t* counts, void* ws, size_t ws_bytes,
                             ssdbox_stream_t stream) {
  return ssdbox_detect_peers(cfg, loc, scores, priors, score_keep, out, counts, nullptr, nullptr, nullptr, ws, ws_bytes, stream);
}

static int detect_impl(const ssdbox_detect_cfg* cfg, const float* loc, const float* scores, const float* priors,
                       const uint8_t* score_keep, const ssdbox_refine* refine, float* out, int32_t* counts,
                       const ssdbox_peer_group* peers, double* loss_sums, float* losses, void* ws, size_t ws_bytes,
                       ssdbox_stream_t stream);

extern "C" int ssdbox_detect_peers(const ssdbox_detect_cfg* cfg, const float* loc, const float* scores, const float* priors,
                                   const uint8_t* score_keep, float* out, int32_t* counts, const ssdbox_peer_group* peers,
                                   double* loss_sums, float* losses, void* ws, size_t ws_bytes, ssdbox_stream_t stream) {
  return detect_impl(cfg, loc, scores, priors, score_keep, nullptr, out, counts, peers, loss_sums, losses, ws, ws_bytes, stream);
}

extern "C" int ssdbox_detect_refine(const ssdbox_detect_cfg* cfg, const float* loc, const float* scores, const float* priors,
                                    const ssdbox_refine* refine, float* out, int32_t* counts, void* ws, size_t ws_bytes,
                                    ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(refine, SSDBOX_EINVAL, "detect: null refine descriptor");
  return detect_impl(cfg, loc, scores, priors, nullptr, refine, out, counts, nullptr, nullptr, nullptr, ws, ws_bytes, stream);
}

static int detect_impl(const ssdbox_detect_cfg* cfg, const float* loc, const float* scores, const float* priors,
                       const uint8_t* score_keep, const ssdbox_refine* refine, float* out, int32_t* counts,
                       const ssdbox_peer_group* peers, double* loss_sums, float* losses, void* ws, size_t ws_bytes,
                       ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(cfg, SSDBOX_EINVAL, "detect: null cfg");
  RefineArgs rf{};
  if (refine) {
    SSDBOX_REQUIRE(cfg->prior_batch_stride == 0, SSDBOX_EINVAL, "refine: priors must be the shared [P,4] tensor (prior_batch_stride 0)");
    SSDBOX_REQUIRE((cfg->B == 0 || cfg->P == 0 || refine->arm_loc) && aligned16(refine->arm_loc) && aligned16(refine->arm_conf),
                   SSDBOX_EINVAL, "refine: arm_loc null / arm tensors not 16-byte aligned");
    rf.arm_loc = refine->arm_loc;
    rf.arm_conf = refine->arm_conf;
    rf.theta = refine->theta;
    rf.var0 = cfg->var0;
    rf.var1 = cfg->var1;
  }
  if (peers) {
    SSDBOX_REQUIRE(peers->world >= 1 && peers->world <= SSDBOX_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world && loss_sums,
                   SSDBOX_EINVAL, "detect: bad peer group / null loss_sums");
    if (cfg->B == 0)     // nothing to launch here: complete the reduction with the stand-alone kernel
      return ssdbox_multibox_loss_peer_finish(peers, loss_sums, losses, stream);
  }
  SSDBOX_REQUIRE(cfg->nms_thresh > 0.0f, SSDBOX_EINVAL, "nms_threshold must be non negative.");  // detection.py:19-20
  const int B = cfg->B, P = cfg->P, C = cfg->C, top_k = cfg->top_k;
  SSDBOX_REQUIRE(B >= 0 && P >= 0 && C >= 1, SSDBOX_EINVAL, "detect: negative size");
  SSDBOX_REQUIRE(top_k >= 1 && top_k <= kLargeTopKLimit, SSDBOX_ESHAPE, "detect: top_k %d outside 1..%d", top_k, kLargeTopKLimit);
  SSDBOX_REQUIRE((long long)B * P < (1ll << 31) && (long long)B * C < (1ll << 31), SSDBOX_ESHAPE, "detect: B*P and B*C must be < 2^31");
  SSDBOX_REQUIRE(cfg->prior_batch_stride == 0 || cfg->prior_batch_stride == (int64_t)P * 4, SSDBOX_EINVAL,
                 "detect: prior_batch_stride must be 0 or 4*P");
  if (B == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(out && ws, SSDBOX_EINVAL, "detect: null pointer");
  SSDBOX_REQUIRE(P == 0 || (loc && scores && priors), SSDBOX_EINVAL, "detect: null pointer");
  SSDBOX_REQUIRE(aligned16(loc) && aligned16(priors), SSDBOX_EALIGN, "detect: box pointers must be 16-byte aligned");
  SSDBOX_REQUIRE(ws_bytes >= detect_ws_bytes(B, P, C, top_k), SSDBOX_EWORKSPACE, "detect: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DevInfo dev;
  int rc = get_dev_info(&dev);
  if (rc) return rc;
  if (top_k > kTopKLimit) {              // the generality path (one CTA per list, no shared-memory suppression matrix)
    SSDBOX_REQUIRE(!(cfg->flags & SSDBOX_DETECT_LOGITS), SSDBOX_ESHAPE, "detect: top_k > %d needs softmaxed scores (no fused softmax)", kTopKLimit);
    DetLargeArgs la{};
    la.B = B; la.P = P; la.C = C; la.top_k = top_k;
    la.nms_thr = cfg->nms_thresh; la.conf_thr = cfg->conf_thresh; la.var0 = cfg->var0; la.var1 = cfg->var1;
    la.prior_stride = (long long)cfg->prior_batch_stride;
    la.loc = loc; la.scores = scores; la.priors = priors; la.keep = score_keep; la.rf = rf;
    la.out = out; la.counts = counts;
    la.ws = static_cast<unsigned char*>(ws);
    const size_t k = (size_t)(top_k < P ? top_k : P);
    la.slice_bytes = large_list_bytes(P, top_k);
    la.off_keys = align_up((size_t)P * 4);
    la.off_box = la.off_keys + align_up(next_pow2((size_t)P) * 8);
    la.off_area = la.off_box + align_up(k * 16);
    la.off_pos = la.off_area + align_up(k * 4);
    int grid = dev.sm_count < kLargeCtas ? dev.sm_count : kLargeCtas;
    if (grid > B * C) grid = B * C;
    {
      TimerScope ts__(KID_DET_SEGMENT_BIG, st);
      detect_large_kernel<<<grid, kLargeThreads, 0, st>>>(la);
    }
    SSDBOX_LAUNCH_OK("detect_large_kernel");
    if (peers) return ssdbox_multibox_loss_peer_finish(peers, loss_sums, losses, stream);
    return SSDBOX_OK;
  }
  const int cap = detect_cand_cap(top_k);
  Carver cv(ws);
  uint32_t* cnt = cv.take<uint32_t>((size_t)B * C + 4);   // [B*C] counters + overflow count / rewritten-list count / tail ticket
  int32_t* ovf_list = cv.take<int32_t>((size_t)B * C);
  int32_t* big_list = cv.take<int32_t>((size_t)B * C);
  unsigned long long* cand = cv.take<unsigned long long>((size_t)B * C * cap);
  uint32_t* scratch = cv.take<uint32_t>((size_t)kOverflowSlots * P);
  float* row_m = cv.take<float>((size_t)B * P);
  float* row_s = cv.take<float>((size_t)B * P);
  const bool logits = (cfg->flags & SSDBOX_DETECT_LOGITS) != 0;

  // State = the counters.  Every call hands them back zeroed (each list's counter is reset by the kernel that
  // finishes the list), so a caller that reuses the workspace for the same shape may skip the init launch.
  if (!(cfg->flags & SSDBOX_DETECT_WS_CLEAN)) {
    rc = launch_init(nullptr, 0, cnt, (size_t)B * C + 4, nullptr, 0, nullptr, 0, st);
    if (rc) return rc;
  }

  if (P > 0) {
    DetStreamArgs sa{};
    rc = plan_ring(&sa.ring, scores, (long long)B * P, C, dev.sm_count, dev.max_smem_optin);
    if (rc) return rc;
    sa.keep = score_keep;
    sa.rf = rf;
    sa.cnt = cnt;
    sa.cand = cand;
    sa.P = P;
    sa.cap = cap;
    sa.thr = cfg->conf_thresh;
    sa.row_m = row_m;
    sa.row_s = row_s;
    void (*kern)(DetStreamArgs) = logits ? detect_stream_kernel<0, true> : detect_stream_kernel<0, false>;
    if (C == 81) kern = logits ? detect_stream_kernel<81, true> : detect_stream_kernel<81, false>;
    else if (C == 21) kern = logits ? detect_stream_kernel<21, true> : detect_stream_kernel<21, false>;
    SSDBOX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sa.ring.smem_bytes));
    SSDBOX_CARVE(kern);
{
    TimerScope ts__(KID_DET_STREAM, st);
    kern<<<sa.ring.grid, kRingThreads, sa.ring.smem_bytes, st>>>(sa);
  }
    SSDBOX_LAUNCH_OK("detect_stream_kernel");
  }

  DetSegArgs g{};
  g.B = B; g.P = P; g.C = C; g.top_k = top_k; g.cap = cap;
  g.nms_thr = cfg->nms_thresh; g.conf_thr = cfg->conf_thresh; g.var0 = cfg->var0; g.var1 = cfg->var1;
  g.prior_stride = (long long)cfg->prior_batch_stride;
  g.loc = loc; g.scores = scores; g.priors = priors; g.keep = score_keep; g.rf = rf;
  g.cnt = cnt; g.cand = cand; g.ovf_count = cnt + (size_t)B * C; g.ovf_list = ovf_list; g.big_count = cnt + (size_t)B * C + 1;
  g.tail_ticket = cnt + (size_t)B * C + 2; g.big_list = big_list; g.scratch = scratch; g.out = out; g.counts = counts;
  g.row_m = logits ? row_m : nullptr; g.row_s = logits ? row_s : nullptr;
  if (peers) {
    g.has_fin = 1;
    g.fin.rank = peers->rank;
    g.fin.world = peers->world;
    g.fin.timeout_ns = peer_timeout_ns(peers);
    for (int r = 0; r < peers->world; ++r) g.fin.bufs[r] = peers->bufs[r];
    g.fin.sums = loss_sums;
    g.fin.losses = losses;
  }
  const size_t seg_smem = (size_t)cap * 8 + nms_smem_bytes(top_k);
  SSDBOX_CUDA(cudaFuncSetAttribute(detect_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seg_smem));
  SSDBOX_CARVE(detect_segments_kernel);
  {
    TimerScope ts__(KID_DET_SEGMENT, st);
    detect_segments_kernel<<<(B * C + kSmallSegs - 1) / kSmallSegs, kSmallThreads, seg_smem, st>>>(g);
  }
  SSDBOX_LAUNCH_OK("detect_segments_kernel");
  // overflowed lists: the chunked kernel only SELECTS (it rewrites the list with <= 1024 keys and queues the
  // segment for the CTA-wide kernel that follows)
  int ovf_grid = dev.sm_count < kOverflowSlots ? dev.sm_count : kOverflowSlots;
  const size_t chunk_smem = 131072 + 288 + 192 + nms_smem_bytes(top_k);
  const bool chunked = top_k <= 256 && cap == 1024 && chunk_smem <= (size_t)dev.max_smem_optin - 1024;
  void (*ovf_kern)(DetSegArgs) = chunked ? detect_overflow_chunk_kernel : detect_overflow_kernel;
  const size_t ovf_smem = chunked ? chunk_smem : (size_t)16384 + 288 + nms_smem_bytes(top_k);
  const int ovf_threads = kOvfThreads;
  if (chunked) {
    const int items = B * ((C - 1 + kOvfClasses - 1) / kOvfClasses);
    if (ovf_grid > items) ovf_grid = items;
  } else if (ovf_grid > B * C) {
    ovf_grid = B * C;
  }
  if (ovf_grid < 1) ovf_grid = 1;
  SSDBOX_CUDA(cudaFuncSetAttribute(ovf_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ovf_smem));
  SSDBOX_CARVE(ovf_kern);
  SSDBOX_CUDA(cudaFuncSetAttribute(detect_segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seg_smem));
  SSDBOX_CARVE(detect_segment_kernel);
  const int big_grid = dev.sm_count * 8 < B * C ? dev.sm_count * 8 : B * C;
  {
    TimerScope ts__(KID_DET_OVERFLOW, st);
    ovf_kern<<<ovf_grid, ovf_threads, ovf_smem, st>>>(g);
  }
  SSDBOX_LAUNCH_OK("detect_overflow_kernel");
  {
    TimerScope ts__(KID_DET_SEGMENT_BIG, st);
    detect_segment_kernel<<<big_grid, kSegThreads, seg_smem, st>>>(g);
  }
  SSDBOX_LAUNCH_OK("detect_segment_kernel");
  return SSDBOX_OK;
}

extern "C" int ssdbox_nms(const float* boxes, const float* scores, int32_t n, float overlap, int32_t top_k,
                          int64_t* keep, int32_t* count, void* ws, size_t ws_bytes, ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(n >= 0, SSDBOX_EINVAL, "nms: negative n");
  SSDBOX_REQUIRE(top_k >= 1, SSDBOX_ESHAPE, "nms: top_k %d < 1", top_k);
  SSDBOX_REQUIRE(top_k <= kLargeTopKLimit || n <= kLargeTopKLimit, SSDBOX_ESHAPE, "nms: more than %d boxes to visit", kLargeTopKLimit);
  SSDBOX_REQUIRE(count, SSDBOX_EINVAL, "nms: null count");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {                                            // box_utils.py:292-293
    SSDBOX_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
    return SSDBOX_OK;
  }
  SSDBOX_REQUIRE(boxes && scores && keep && ws, SSDBOX_EINVAL, "nms: null pointer");
  SSDBOX_REQUIRE(ws_bytes >= nms_ws_bytes(n, top_k), SSDBOX_EWORKSPACE, "nms: workspace too small");
  if (top_k > kTopKLimit) {              // any top_k (box_utils.py:299-301): keys sorted in global memory, matrix-free sweep
    unsigned char* w = static_cast<unsigned char*>(ws);
    const size_t k = (size_t)(top_k < n ? top_k : n);
    const size_t off_keys = align_up((size_t)n * 4), off_box = off_keys + align_up(next_pow2((size_t)n) * 8);
    const size_t off_area = off_box + align_up(k * 16), off_pos = off_area + align_up(k * 4);
    nms_large_kernel<<<1, kLargeThreads, 0, st>>>(boxes, scores, n, overlap, top_k, keep, count, reinterpret_cast<uint32_t*>(w),
                                                  reinterpret_cast<unsigned long long*>(w + off_keys), reinterpret_cast<float4*>(w + off_box),
                                                  reinterpret_cast<float*>(w + off_area), reinterpret_cast<int32_t*>(w + off_pos));
    SSDBOX_LAUNCH_OK("nms_large_kernel");
    return SSDBOX_OK;
  }
  size_t smem = 16384 + 288 + nms_smem_bytes(top_k);
  SSDBOX_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_kernel<<<1, kOvfThreads, smem, st>>>(boxes, scores, n, overlap, top_k, keep, count, static_cast<uint32_t*>(ws));
  SSDBOX_LAUNCH_OK("nms_kernel");
  return SSDBOX_OK;
}
