// CTA-wide exact selection of the K largest order-preserving 32-bit keys (radix select,
// 11 + 11 + 10 bits).  Shared by hard-negative mining (ties -> LOWER index first: a stable
// descending sort, multibox_loss.py:99-103) and by top-k candidate selection before NMS
// (ties -> HIGHER index first: an ascending stable sort consumed from the tail,
// box_utils.py:299-312).
#pragma once
#include "sizes.h"
#include "ssdbox_dev.cuh"

namespace ssdbox {

// visits every ordered key once; VEC = 4 reads them as 16-byte vectors (P % 4 == 0, 16-byte aligned)
template <int VEC, class F>
__device__ __forceinline__ void for_each_key(const uint32_t* uk, int P, F f) {
  if (VEC == 4) {
    const uint4* u4 = reinterpret_cast<const uint4*>(uk);
    const int n4 = P >> 2;
    for (int q = threadIdx.x; q < n4; q += blockDim.x) {
      uint4 u = u4[q];
      f(u.x); f(u.y); f(u.z); f(u.w);
    }
  } else {
    for (int p = threadIdx.x; p < P; p += blockDim.x) f(uk[p]);
  }
}

// After the call an element is selected iff uk[p] != 0 && uk[p] >= T (returned); equal-to-T
// elements that lose the tie are demoted to T-1 in place.  K >= 1.  uk[p] == 0 marks elements
// outside the ranking.  hist1 (nullable, global) = precomputed level-1 histogram.
// s_hist: 2048 uint32, s_iscr: >= 64 ints, s_res: >= 2 ints (all shared memory).
template <bool kPreferHighIndex, int VEC = 1>
__device__ uint32_t cta_select_threshold(uint32_t* uk, int P, int K, const uint32_t* hist1, uint32_t* s_hist,
                                         int* s_iscr, int* s_res) {
  const int tid = threadIdx.x, T = blockDim.x;
  // level 1: bits 31..21
  for (int i = tid; i < kHistBins; i += T) s_hist[i] = hist1 ? hist1[i] : 0u;
  __syncthreads();
  if (!hist1) {
    for_each_key<VEC>(uk, P, [&](uint32_t u) {
      if (u) atomicAdd(&s_hist[u >> 21], 1u);
    });
    __syncthreads();
  }
  find_digit(s_hist, kHistBins, K, s_iscr, s_res);
  int d1 = s_res[0];
  if (d1 < 0) return 1u;   // fewer than K ranked elements: take them all
  int K2 = K - s_res[1];
  int n1 = (int)s_hist[d1];
  __syncthreads();
  if (K2 == n1) return ((uint32_t)d1 << 21) ? ((uint32_t)d1 << 21) : 1u;
  // level 2: bits 20..10 of the elements in bin d1
  for (int i = tid; i < 2048; i += T) s_hist[i] = 0u;
  __syncthreads();
  for_each_key<VEC>(uk, P, [&](uint32_t u) {
    if (u && (int)(u >> 21) == d1) atomicAdd(&s_hist[(u >> 10) & 2047u], 1u);
  });
  __syncthreads();
  find_digit(s_hist, 2048, K2, s_iscr, s_res);
  int d2 = s_res[0];
  int K3 = K2 - s_res[1];
  int n2 = (int)s_hist[d2];
  __syncthreads();
  uint32_t pre2 = ((uint32_t)d1 << 11) | (uint32_t)d2;
  if (K3 == n2) return (pre2 << 10) ? (pre2 << 10) : 1u;
  // level 3: bits 9..0
  for (int i = tid; i < 1024; i += T) s_hist[i] = 0u;
  __syncthreads();
  for_each_key<VEC>(uk, P, [&](uint32_t u) {
    if (u && (u >> 10) == pre2) atomicAdd(&s_hist[u & 1023u], 1u);
  });
  __syncthreads();
  find_digit(s_hist, 1024, K3, s_iscr, s_res);
  int d3 = s_res[0];
  int need = K3 - s_res[1];
  int n3 = (int)s_hist[d3];
  __syncthreads();
  uint32_t Tu = (pre2 << 10) | (uint32_t)d3;
  if (need == n3) return Tu;
  // ties straddle the cut: walk the equal keys in preference order, the first `need` win
  int running = 0;
  int nchunks = (P + T - 1) / T;
  for (int ch = 0; ch < nchunks; ++ch) {
    int base = (kPreferHighIndex ? (nchunks - 1 - ch) : ch) * T;
    int p = base + tid;
    int flag = (p < P && uk[p] == Tu) ? 1 : 0;
    int total;
    int ex = block_exclusive_scan(flag, s_iscr, &total);
    int rank = running + (kPreferHighIndex ? (total - ex - 1) : ex);
    if (flag && rank >= need) uk[p] = Tu - 1u;
    running += total;
    __syncthreads();
  }
  return Tu;
}

}  // namespace ssdbox
