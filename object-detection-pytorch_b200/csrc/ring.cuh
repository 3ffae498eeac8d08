// The row-tile ring shared by the two HBM-bound streaming kernels (MultiBoxLoss key pass and
// Detect candidate pass): a [rows, C] fp32 matrix is cut into tiles of R rows; each CTA owns a
// contiguous run of tiles and pulls them through NS shared-memory stages with 1-D TMA bulk
// copies (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) issued by one producer lane.
// kRingGroups consumer groups of kRingGroupWarps warps take the stages in turn; one thread owns one
// row of a stage (row stride C floats: conflict-free for odd C).  Measured on B200: 128-row tiles (41 KB
// bulk copies at C = 81) with 2 groups reach 94 % of the copy roofline; 64-row tiles with 4 groups and twice
// the stages -- the same bytes in flight -- are 40 % slower, the bulk-copy engine wants few large copies.
//
// Barriers: iteration `it` of a CTA uses stage it % NS but barrier pair j = it % (2*NS)
// (full[j]: count 1 + tx bytes, empty[j]: one arrival per warp of a group), phase it / (2*NS).
// With 2*NS pairs (NS even, so 2*NS is a multiple of the group count) every barrier is always
// consumed by the same consumer group, so each waiter
// follows its barrier phase by phase and the 1-bit parity can never alias (with one pair per
// stage and NS odd, a group would skip every other phase of a stage and could pass a wait two
// phases early).
#pragma once
#include <cstdlib>

#include "common.h"
#include "ssdbox_dev.cuh"

namespace ssdbox {

constexpr int kRingConsumerWarps = 8;
constexpr int kRingGroups = 2;                                      // consumer groups (4 x 64-row tiles measured 40 % slower)
constexpr int kRingGroupWarps = kRingConsumerWarps / kRingGroups;   // warps (x32 rows) per group = per tile
constexpr int kRingThreads = (kRingConsumerWarps + 1) * 32;
constexpr int kRingMaxStages = 12;
constexpr int kRingHeaderBytes = 384;   // 2 * kRingMaxStages pairs of 8-byte mbarriers

struct RingPlan {
  const float* src;
  long long rows;
  long long tiles;
  int C;
  int R;         // rows per tile = 32 * kRingGroupWarps * KR
  int KR;        // rows per consumer thread per tile (> 1 for narrow rows: keeps the bulk copies ~40 KB)
  int NS;
  int tiles_per_cta;
  int bulk_ok;
  int interleave;  // g > 0: groups of g consecutive CTAs share one contiguous run of tiles and take its tiles in turn
                   // (member m takes tiles m, m + g, m + 2g ... of the run); 0: every CTA has a run of its own
  int grid;
  size_t smem_bytes;
};

// host: choose R / NS / grid for `rows` x C given the device limits
static inline int plan_ring(RingPlan* p, const float* src, long long rows, int C, int sm_count, int max_smem) {
  size_t budget = (size_t)max_smem - 1024 - kRingHeaderBytes;
  // narrow rows (C = 21: 84 bytes): several rows per thread so that one bulk copy stays around 40 KB --
  // the per-SM bulk-copy engine wants few large copies (10 KB tiles measured 3x off the roofline)
  const int base = 32 * kRingGroupWarps;
  int KR = 1;
  while (KR < 32 && (size_t)base * (KR * 2) * C * 4 <= 49152) KR *= 2;
  int R = base * KR;
  while (R >= 32 && (size_t)2 * R * C * 4 > budget) {
    R >>= 1;
    if (KR > 1) KR >>= 1;
  }
  if (R < 32) return fail(SSDBOX_ESHAPE, "%d classes do not fit the shared-memory ring", C);
  size_t stage_bytes = (size_t)R * C * 4;
  int NS = (int)(budget / stage_bytes);
  if (NS > kRingMaxStages) NS = kRingMaxStages;
#ifdef SSDBOX_EXPERIMENTS      // tools/exp_*.sh builds only: the release library reads no environment
  if (const char* e = getenv("SSDBOX_RING_MAX_STAGES")) {
    int cap = atoi(e);
    if (cap >= 2 && NS > cap) NS = cap;
  }
#endif
  while (NS > 1 && (2 * NS) % kRingGroups) --NS;     // see the header comment
  p->src = src;
  p->rows = rows;
  p->C = C;
  p->R = R;
  p->KR = KR;
  p->NS = NS;
  p->tiles = (rows + R - 1) / R;
#ifdef SSDBOX_EXPERIMENTS
  if (const char* e = getenv("SSDBOX_RING_GRID")) {            // stream on fewer SMs
    int cap = atoi(e);
    if (cap >= 1 && cap < sm_count) sm_count = cap;
  }
#endif
  int grid = (int)(p->tiles < sm_count ? p->tiles : sm_count);
  if (grid < 1) grid = 1;
  p->tiles_per_cta = (int)((p->tiles + grid - 1) / grid);
  if (p->tiles_per_cta < 1) p->tiles_per_cta = 1;
  p->grid = (int)((p->tiles + p->tiles_per_cta - 1) / p->tiles_per_cta);
  if (p->grid < 1) p->grid = 1;
  p->bulk_ok = aligned16(src) ? 1 : 0;
  p->interleave = 0;
#ifdef SSDBOX_EXPERIMENTS
  if (const char* e = getenv("SSDBOX_RING_INTERLEAVE")) {
    int g = atoi(e);
    p->interleave = g < 1 ? 0 : (g > p->grid ? p->grid : g);
  }
#endif
  p->smem_bytes = kRingHeaderBytes + (size_t)NS * stage_bytes;
  return SSDBOX_OK;
}

struct RingCtx {
  uint64_t* full;
  uint64_t* empty;
  float* stages;
  size_t stage_floats;
  long long t0;
  long long tstep;   // tile of iteration it = t0 + it * tstep
  int n_local;
};

__device__ __forceinline__ RingCtx ring_setup(const RingPlan& p, unsigned char* smem_raw) {
  RingCtx r;
  r.full = reinterpret_cast<uint64_t*>(smem_raw);
  r.empty = r.full + 2 * kRingMaxStages;
  r.stages = reinterpret_cast<float*>(smem_raw + kRingHeaderBytes);
  r.stage_floats = (size_t)p.R * p.C;
  if (p.interleave) {
    const int g = p.interleave;
    const int group = blockIdx.x / g, member = blockIdx.x % g;
    const int members = min(g, (int)gridDim.x - group * g);                       // the last group may be short
    const long long lo = (long long)group * g * p.tiles_per_cta;                  // the members' own runs, pooled
    long long hi = lo + (long long)members * p.tiles_per_cta;
    if (hi > p.tiles) hi = p.tiles;
    r.t0 = lo + member;
    r.tstep = members;
    r.n_local = r.t0 < hi ? (int)((hi - r.t0 + members - 1) / members) : 0;
  } else {
    r.t0 = (long long)blockIdx.x * p.tiles_per_cta;
    r.tstep = 1;
    long long t1 = r.t0 + p.tiles_per_cta;
    if (t1 > p.tiles) t1 = p.tiles;
    r.n_local = t1 > r.t0 ? (int)(t1 - r.t0) : 0;
  }
  if (threadIdx.x == 0) {
    for (int j = 0; j < 2 * p.NS; ++j) {
      mbar_init(&r.full[j], 1);
      mbar_init(&r.empty[j], kRingGroupWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();
  return r;
}

// whole producer warp calls this and then exits
__device__ __forceinline__ void ring_produce(const RingPlan& p, const RingCtx& r) {
  const int lane = threadIdx.x & 31;
  uint64_t policy = l2_evict_first_policy();
  for (int it = 0; it < r.n_local; ++it) {
    const int s = it % p.NS;
    const int j = it % (2 * p.NS);
    if (it >= p.NS && lane == 0) {   // stage s was last used by iteration it - NS: wait for its release
      int prev = it - p.NS;
      mbar_wait(&r.empty[prev % (2 * p.NS)], (uint32_t)((prev / (2 * p.NS)) & 1));
    }
    __syncwarp();
    long long r0 = (r.t0 + it * r.tstep) * p.R;
    long long left = p.rows - r0;
    int nrows = left < p.R ? (int)left : p.R;
    uint32_t bytes = (uint32_t)nrows * (uint32_t)p.C * 4u;
    const float* src = p.src + r0 * p.C;
    float* dst = r.stages + (size_t)s * r.stage_floats;
    if (p.bulk_ok && (bytes & 15u) == 0u) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&r.full[j], bytes);
        bulk_g2s(dst, src, bytes, &r.full[j], policy);
      }
    } else {  // unaligned base or ragged tail: plain loads through the generic proxy
      int nf = nrows * p.C;
      for (int i = lane; i < nf; i += 32) dst[i] = src[i];
      __syncwarp();
      if (lane == 0) mbar_arrive(&r.full[j]);
    }
  }
}

}  // namespace ssdbox
