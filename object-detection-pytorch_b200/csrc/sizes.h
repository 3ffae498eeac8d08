// Workspace layouts (shared between the sizing query in abi.cu and the op implementations).
#pragma once
#include <cstddef>
#include <cstdint>

#include "common.h"

namespace ssdbox {

constexpr int kGmaxLimit = 4096;       // truths per image (smem + int16 truth index)
constexpr int kHistBins = 2048;        // level-1 mining histogram: mine_bin() of the ordered key (128 bins per octave)
constexpr int kTopKLimit = 1024;       // NMS sweep keeps one removed-word per lane (the fast paths; larger top_k: the *_large kernels)
constexpr int kLargeTopKLimit = 65536; // removed-bit words of one list in shared memory (8 KB)
constexpr int kLargeCtas = 160;        // workspace slices of detect_large_kernel (>= SMs of the device that runs it)
constexpr int kClassLimit = 32766;     // labels travel as int16 (label + 1)

static inline int gt_pad(int gmax) { return gmax < 2 ? 2 : (gmax + 1) / 2 * 2; }

// ---- match ---------------------------------------------------------------------------------
struct MatchWs {
  unsigned long long* gt_best;  // [(B+1), gt_pad] packed (iou_bits << 32 | ~prior)
  uint32_t* done;               // [B] CTA tickets per image
  int16_t* lab;                 // [B,P] class target (0 = background)
  int16_t* tidx;                // [B,P] matched truth index
};
static inline size_t match_core_bytes(int B, int gmax) {
  return align_up((size_t)(B + 1) * gt_pad(gmax) * 8) + align_up((size_t)(B + 1) * 4);
}
static inline size_t match_ws_bytes(int B, int P, int gmax) {
  return match_core_bytes(B, gmax) + 2 * align_up((size_t)B * P * 2);
}
static inline void carve_match_core(Carver& c, int B, int gmax, MatchWs* w) {
  w->gt_best = c.take<unsigned long long>((size_t)(B + 1) * gt_pad(gmax));
  w->done = c.take<uint32_t>((size_t)(B + 1));
}

// ---- loss ----------------------------------------------------------------------------------
struct LossWs {
  MatchWs m;
  float* keys;        // [B,P]  lse - x[0]
  uint32_t* hist;     // [B, kHistBins]
  uint32_t* ukey;     // [B,P]  ordered mining keys (only used when they do not fit in smem)
  double* partial;    // [2B,3]
  uint32_t* ticket;   // [1]
};
static inline size_t loss_ws_bytes(int B, int P, int C, int gmax) {
  (void)C;
  return match_core_bytes(B, gmax) + align_up((size_t)B * P * 2) + align_up((size_t)B * P * 4) * 2 +
         align_up((size_t)B * kHistBins * 4) + align_up((size_t)B * 6 * 8) + 256;
}

// ---- mining in isolation ---------------------------------------------------------------------
static inline size_t mine_ws_bytes(int B, int P) { return align_up((size_t)B * P * 4) + 256; }

// ---- detect --------------------------------------------------------------------------------
constexpr int kOverflowSlots = 160;    // >= SM count: one ordered-score scratch row per resident CTA
static inline int detect_cand_cap(int top_k) { return top_k <= 512 ? 1024 : 2048; }
static inline size_t next_pow2(size_t v) {
  size_t p = 32;
  while (p < v) p <<= 1;
  return p;
}
// one list of the any-top_k path: ordered keys [P], sort keys [pow2 >= P], box / area / keep of the top_k
static inline size_t large_list_bytes(int P, int top_k) {
  const size_t k = (size_t)(top_k < P ? top_k : P);
  return align_up((size_t)P * 4) + align_up(next_pow2((size_t)P) * 8) + align_up(k * 16) + 2 * align_up(k * 4);
}
static inline size_t detect_ws_bytes(int B, int P, int C, int top_k) {
  if (top_k > kTopKLimit) return (size_t)kLargeCtas * large_list_bytes(P, top_k);
  return align_up((size_t)B * C * 4 + 16) + 2 * align_up((size_t)B * C * 4) +
         align_up((size_t)B * C * detect_cand_cap(top_k) * 8) +
         align_up((size_t)kOverflowSlots * P * 4) + 2 * align_up((size_t)B * P * 4);   // + softmax row max / sum (logits mode)
}

// ---- eval post-processing (detections -> flat list) -----------------------------------------
static inline size_t compact_ws_bytes(int B, int C) {
  return align_up((size_t)B * C * 4) + align_up(((size_t)B * C + 1) * 4) + 256;
}

// ---- VOC evaluation (rows = detections, M = truths) -------------------------------------------
static inline size_t voc_eval_ws_bytes(int rows, int M, int C) {
  const size_t n = (size_t)rows + 1;
  const size_t nblk = ((size_t)rows + 4095) / 4096;
  return align_up(((size_t)M + 1) * 8) + 6 * align_up(n * 4) + align_up((size_t)C * 4) +
         align_up(256 * (nblk ? nblk : 1) * 4) + 256;
}

// ---- nms -----------------------------------------------------------------------------------
static inline size_t nms_ws_bytes(int n, int top_k) {
  if (top_k > kTopKLimit) return large_list_bytes(n, top_k) + 256;
  return align_up((size_t)n * 4) + 256;
}

}  // namespace ssdbox
