// MultiBoxLoss forward / backward: lib/layers/modules/multibox_loss.py:48-117 (+ autograd).
//
// Forward = 2 launches on the caller's stream (CUDA-graph capturable, no host sync), 3 the first time a workspace
// is used:
//   init_kernel        per-truth best-prior keys := (0, prior 0); tickets / histograms := 0.  Skipped with
//                      SSDBOX_LOSS_WS_CLEAN: every call hands this state back initialised
//   loss_stream_kernel THE HBM-bound kernel, warp-specialised, one persistent CTA per SM:
//                        1 producer warp   streams conf [B*P, C] through a ring of TMA bulk-copy
//                                          stages (cp.async.bulk + mbarrier)
//                        8 consumer warps  one thread per prior row: lse and the background key
//                                          lse - x[0]; level-1 mining histogram (top 11 bits of the
//                                          order-preserving key).  Independent of the class targets.
//                        1 scheduler warp + 6 or 8 match warps: box_utils.match while the consumers wait
//                                          on memory.  Work units (1024 priors of one image) come
//                                          from a global counter, the unit's truths and priors are
//                                          staged in shared memory (priors by TMA); 4 consecutive
//                                          priors per thread, warp bounding-box pruning, per-truth
//                                          argmax by REDUX + atomicMax.  The IoU matrix never exists.
//   mine_reduce_*      per image (one CTA, or a cluster of two for many priors; keys register-resident
//                      when P % 4 == 0 and P <= 24576): replays the forced assignment ("every truth
//                      keeps its best prior, last truth wins"), positives get CE = lse - x[target]
//                      (one gathered logit each) and move to the zero bin; radix-select of the
//                      num_neg-th largest mining key (level 1 from the streamed histogram, levels
//                      2/3 touch only the winning bin), canonical tie order, fixed-order fp64
//                      reduction of smooth-L1 / CE; the last CTA folds the partials and, multi-GPU,
//                      exchanges {sum_l, sum_c, N} with the other ranks over NVLink peer memory.
// (With SSDBOX_LOSS_SEPARATE_MATCH, or when two unit buffers do not fit beside the ring,
// match_kernel of match.cu runs as a fourth launch before the stream kernel instead.)
// The final CE over pos U neg needs no second pass over conf: CE of a negative is its mining key.
// Backward zero-fills grad_conf and touches conf only on the selected rows.
#include <cstdlib>

#include "ops.h"
#include "peer.cuh"
#include "ring.cuh"
#include "select.cuh"
#include "ssdbox_dev.cuh"

namespace ssdbox {

// ------------------------------------------------------------------------------------------------
// streaming pass
// ------------------------------------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
#ifdef SSDBOX_PHASE_TIMING
__device__ long long g_sphase[8 * 160];
__device__ long long g_mstat[8 * 160];
__device__ unsigned long long g_sgt[2 * 160];
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define SMARK(k) do { if (blockIdx.x < 160 && lane == 0) { g_sphase[blockIdx.x * 8 + (k)] = clock64(); \
    if ((k) == 0) g_sgt[blockIdx.x * 2] = gtimer_ns(); if ((k) == 7) g_sgt[blockIdx.x * 2 + 1] = gtimer_ns(); } } while (0)
#else
#define SMARK(k) do { } while (0)
#endif

constexpr int kMatchWarpsMax = 10;                                 // the kernel is launched with 6 or 10 match warps:
// 10 when the matching is the longer half of the kernel (narrow conf rows, C = 21: RFB300-VOC B=256 102 us with 6,
// 87 us with 8, 84 us with 10 -- 640 threads x 96 registers still fit the register file; 12 warps under a 704-thread
// launch bound: 78.5 us there but slower on the small batches, not kept), 6 when the conf stream is (C = 81: 16 warps
// = 4 per scheduler, 90 us against 93 us with 8; 4 / 5 / 7 warps measured 91.4 / - / 95.0 us).
// One work unit = match_warps * 128 priors.
static inline int match_warps_for(int C) { return C >= 48 ? 6 : 10; }
constexpr int kSchedWarp = kRingConsumerWarps + 1;                 // warp 8 = conf producer, 9 = unit scheduler
constexpr int kFirstMatchWarp = kRingConsumerWarps + 2;
constexpr int kStreamThreads = (kRingConsumerWarps + 2 + kMatchWarpsMax) * 32;      // launch bound; launched with match_warps
constexpr int kMatchBarBytes = 64;                                 // mfull[2], mempty[2]

struct StreamArgs {
  RingPlan ring;       // conf [B*P, C]
  const uint8_t* pool;
  float* key0;         // [B*P]  lse - x[0]   (mining key of a non-positive prior, multibox_loss.py:94)
  uint32_t* hist;
  int P;
  // fused matching (box_utils.py:92-130 on dedicated warps)
  int fuse;
  int match_warps, match_tile;   // warps doing the matching, priors per work unit (= 128 * match_warps)
  int B;
  const float* gt;
  const int32_t* gt_offsets;
  int gmax, gpad;
  const float* priors;
  long long prior_stride;
  const float* anchors_xyxy;
  RefineArgs rf;            // RefineDet fused: anchors decoded from arm_loc in the match warps, pool = ARM objectness
  float threshold;
  int binarize;
  unsigned long long* gt_best;
  int16_t* lab_out;
  int16_t* tidx_out;
  uint32_t* unit_counter;   // next matching work unit (zeroed by init_kernel)
  int unit_bytes;           // one unit buffer in shared memory
  int dbg;                  // SSDBOX_PHASE_TIMING builds: 1 = do not stream conf (matching alone)
  const double* shift;      // SSDBOX_LOSS_LSE_SHIFT: the value the log-sum-exp subtracts (nullptr: each row's maximum)
};

// log-sum-exp of one row with the row maximum (box_utils.py:273 uses one global maximum; same value
// up to fp32 rounding).  Compile-time class counts: the row sits in registers, the maximum is a tree of
// three-input max (FMNMX3) and the exponentials are fed / summed two at a time with packed fp32x2 FMA / ADD
// (FFMA2 / FADD2) on four independent accumulator pairs: ~300 issue slots per row of 81 classes instead of
// ~430 with scalar instructions (the consumer warps share their SM sub-partitions with the match warps).
// ex2.approx on (x - m) * log2(e).
template <int CT>
__device__ __forceinline__ float row_lse(const float* __restrict__ rp, int C) {
  float m, s;
  if (CT > 0) {
    float v[CT > 0 ? CT : 1];
#pragma unroll
    for (int c = 0; c < CT; ++c) v[c] = rp[c];
    float m4[4] = {v[0], v[0], v[0], v[0]};
#pragma unroll
    for (int c = 1; c + 1 < CT; c += 2) m4[(c >> 1) & 3] = fmax3(m4[(c >> 1) & 3], v[c], v[c + 1]);
    if ((CT & 1) == 0) m4[0] = fmaxf(m4[0], v[CT - 1]);          // elements 1 .. CT-1: an odd count leaves one over
    m = fmax3(m4[0], m4[1], fmaxf(m4[2], m4[3]));
    const float nml = -m * kLog2e;
    const unsigned long long l2 = pack_f32x2(kLog2e, kLog2e), n2 = pack_f32x2(nml, nml);
    unsigned long long s2[4] = {0ull, 0ull, 0ull, 0ull};         // four pairs of +0.0f
#pragma unroll
    for (int c = 0; c + 1 < CT; c += 2) {
      float a, b;
      unpack_f32x2(fma_f32x2(pack_f32x2(v[c], v[c + 1]), l2, n2), a, b);
      s2[(c >> 1) & 3] = add_f32x2(s2[(c >> 1) & 3], pack_f32x2(ex2_approx(a), ex2_approx(b)));
    }
    float a0, b0, a1, b1;
    unpack_f32x2(add_f32x2(add_f32x2(s2[0], s2[1]), add_f32x2(s2[2], s2[3])), a0, b0);
    s = a0 + b0;
    if (CT & 1) s += ex2_approx(fmaf(v[CT - 1], kLog2e, nml));
    (void)a1; (void)b1;
  } else {
    m = rp[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, rp[c]);
    float nml = -m * kLog2e;
    s = 0.f;
    for (int c = 0; c < C; ++c) s += ex2_approx(fmaf(rp[c], kLog2e, nml));
  }
  return logf(s) + m;
}

// box_utils.py:272-273 as written: log(sum(exp(x - shift))) + shift with the caller's shift (the batch maximum)
__device__ __forceinline__ float row_lse_shift(const float* __restrict__ rp, int C, float shift) {
  float s = 0.f;
  for (int c = 0; c < C; ++c) s = __fadd_rn(s, expf(__fsub_rn(rp[c], shift)));
  return __fadd_rn(logf(s), shift);
}

// ---- matching on dedicated warps -----------------------------------------------------------------
// Work unit = match_tile (128 per match warp) consecutive priors of one image.  Units are handed out dynamically (one
// global counter, heavy coarse-layer tiles first) so the matching load is balanced over the SMs
// whatever the truth counts of the images a CTA happens to stream.  The scheduler warp stages a
// unit in one of two shared-memory buffers: the unit's truths (+area, +label, +CTA-local best-prior
// keys) with plain loads and its priors with one TMA bulk copy; the match warps therefore never wait
// on a global load while the consumers saturate HBM.
struct MatchUnit {
  int b, p0, nrows, G;      // b < 0: no more work
};
struct UnitBuf {
  MatchUnit* hd;
  float4* pri;                // [match_tile] centre-form priors (or xyxy anchors)
  float4* box;                // [gpad] truth xyxy
  unsigned long long* best;   // [gpad] per-truth best prior over this unit
  float* area;                // [gpad]
  int* lab;                   // [gpad] class target (label + 1)
  float4* arm;                // [match_tile] RefineDet fused: the unit's ARM offsets (behind the rest; only then allocated)
};
__device__ __forceinline__ UnitBuf unit_buf(unsigned char* mbase, int buf, int unit_bytes, int gpad, int match_tile) {
  unsigned char* ub = mbase + kMatchBarBytes + (size_t)buf * unit_bytes;
  UnitBuf u;
  u.hd = reinterpret_cast<MatchUnit*>(ub);
  u.pri = reinterpret_cast<float4*>(ub + 16);
  u.box = u.pri + match_tile;
  u.best = reinterpret_cast<unsigned long long*>(u.box + gpad);
  u.area = reinterpret_cast<float*>(u.best + gpad);
  u.lab = reinterpret_cast<int*>(u.area + gpad);
  u.arm = reinterpret_cast<float4*>(u.lab + gpad);       // 16-byte aligned: gpad is even
  return u;
}
static inline int unit_buf_bytes(int gpad, int match_tile, bool refine) {
  return (int)align_up((size_t)16 + (size_t)match_tile * 16 + (size_t)gpad * 32 + (refine ? (size_t)match_tile * 16 : 0), 128);
}

// max IoU, lowest prior index on ties (box_utils.py:116); shared-memory copy first: the global
// atomic happens once per (unit, truth), never inside the truth loop
__device__ __forceinline__ void best_prior_update(unsigned long long* dst, uint32_t iou_bits, uint32_t p) {
  unsigned long long key = ((unsigned long long)iou_bits << 32) | (unsigned long long)(uint32_t)(~p);
  if (key > *reinterpret_cast<volatile unsigned long long*>(dst)) atomicMax(dst, key);
}

// whole scheduler warp
__device__ void match_sched_loop(const StreamArgs& a, unsigned char* mbase, int lane) {
  uint64_t* mfull = reinterpret_cast<uint64_t*>(mbase);
  uint64_t* mempty = mfull + 2;
  const int ntile = (a.P + a.match_tile - 1) / a.match_tile;
  const unsigned nunits = (unsigned)ntile * (unsigned)a.B;
  const float* src_base = a.anchors_xyxy ? a.anchors_xyxy : a.priors;
  for (int k = 0;; ++k) {
    const int buf = k & 1;
    UnitBuf ub = unit_buf(mbase, buf, a.unit_bytes, a.gpad, a.match_tile);
    if (k >= 2) {    // the match warps have released the unit staged two rounds ago
      if (lane == 0) mbar_wait(&mempty[buf], (uint32_t)(((k >> 1) - 1) & 1));
      __syncwarp();
    }
    unsigned u = 0;
    if (lane == 0) u = atomicAdd(a.unit_counter, 1u);
    u = __shfl_sync(SSDBOX_FULL_MASK, u, 0);
    if (u >= nunits) {
      if (lane == 0) {
        ub.hd->b = -1;
        mbar_arrive(&mfull[buf]);
      }
      return;
    }
    const int tile = ntile - 1 - (int)(u / (unsigned)a.B);     // coarse layers (most truths per warp) first
    const int b = (int)(u % (unsigned)a.B);
    const int p0 = tile * a.match_tile;
    const int nrows = a.P - p0 < a.match_tile ? a.P - p0 : a.match_tile;
    const int g0 = a.gt_offsets[b];
    int G = a.gt_offsets[b + 1] - g0;
    G = G < 0 ? 0 : (G > a.gmax ? a.gmax : G);
    for (int g = lane; g < G; g += 32) {
      const float* r = a.gt + (size_t)(g0 + g) * 5;
      Box t;
      t.x1 = r[0]; t.y1 = r[1]; t.x2 = r[2]; t.y2 = r[3];
      ub.box[g] = make_float4(t.x1, t.y1, t.x2, t.y2);
      ub.area[g] = box_area(t);
      ub.lab[g] = a.binarize ? 1 : (int)(r[4] + 1.0f);   // box_utils.py:129
      ub.best[g] = kBestInit;
    }
    if (lane == 0) *ub.hd = MatchUnit{b, p0, nrows, G};
    __syncwarp();
    if (lane == 0) {
      if (a.rf.arm_loc) {      // the shared priors of the unit and the image's ARM offsets: decoded by the match warps
        mbar_arrive_expect_tx(&mfull[buf], (uint32_t)nrows * 32u);
        bulk_g2s_plain(ub.pri, a.priors + (size_t)p0 * 4, (uint32_t)nrows * 16u, &mfull[buf]);
        bulk_g2s_plain(ub.arm, a.rf.arm_loc + ((size_t)b * (size_t)a.P + (size_t)p0) * 4, (uint32_t)nrows * 16u, &mfull[buf]);
      } else {
        const float* src = src_base + (size_t)b * (size_t)a.prior_stride + (size_t)p0 * 4;
        mbar_arrive_expect_tx(&mfull[buf], (uint32_t)nrows * 16u);
        bulk_g2s_plain(ub.pri, src, (uint32_t)nrows * 16u, &mfull[buf]);
      }
    }
  }
}

// One match warp owns 128 consecutive priors of the unit (4 consecutive priors per thread).
//  * truths are pruned 32 at a time: lane g tests truth g against the warp's bounding box, the
//    ballot is the list of truths to compute (a disjoint truth has IoU 0 with all 128 priors and
//    cannot move either running maximum, both update on strict >),
//  * the four IoUs of a thread are branch-free so their IEEE divisions overlap.
__device__ void match_unit_loop(const StreamArgs& a, unsigned char* mbase, int mw, int lane) {
  constexpr int K = 4;
  uint64_t* mfull = reinterpret_cast<uint64_t*>(mbase);
  uint64_t* mempty = mfull + 2;
#ifdef SSDBOX_PHASE_TIMING
  long long t_wait = 0, t_busy = 0, n_units = 0, n_truths = 0;
#endif
  for (int k = 0;; ++k) {
    const int buf = k & 1;
    UnitBuf ub = unit_buf(mbase, buf, a.unit_bytes, a.gpad, a.match_tile);
#ifdef SSDBOX_PHASE_TIMING
    long long c0 = clock64();
#endif
    mbar_wait(&mfull[buf], (uint32_t)((k >> 1) & 1));
#ifdef SSDBOX_PHASE_TIMING
    long long c1 = clock64();
    t_wait += c1 - c0;
#endif
    const MatchUnit hd = *ub.hd;
    if (hd.b < 0) break;
    const int G = hd.G;
    const int r0 = mw * (32 * K) + lane * K;      // first prior of this thread inside the unit
    Box box[K];
    float area[K], bt_ov[K];
    int bt_idx[K];
    bool valid[K];
    float wx1 = INFINITY, wy1 = INFINITY, wx2 = -INFINITY, wy2 = -INFINITY;
#pragma unroll
    for (int q = 0; q < K; ++q) {
      valid[q] = r0 + q < hd.nrows;
      box[q].x1 = box[q].y1 = box[q].x2 = box[q].y2 = 0.f;
      if (valid[q]) {
        float4 v = ub.pri[r0 + q];
        if (a.rf.arm_loc) {
          box[q] = decode_box(ub.arm[r0 + q], v, a.rf.var0, a.rf.var1);     // refined anchor, xyxy
        } else if (a.anchors_xyxy) {
          box[q].x1 = v.x; box[q].y1 = v.y; box[q].x2 = v.z; box[q].y2 = v.w;
        } else {
          box[q] = point_form(v);
        }
        wx1 = fminf(wx1, box[q].x1); wy1 = fminf(wy1, box[q].y1);
        wx2 = fmaxf(wx2, box[q].x2); wy2 = fmaxf(wy2, box[q].y2);
      }
      area[q] = box_area(box[q]);
      bt_ov[q] = 0.0f;     // all-zero IoU column -> truth 0 (first index), overlap 0
      bt_idx[q] = 0;
    }
    if (__ballot_sync(SSDBOX_FULL_MASK, valid[0]) != 0u) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        wx1 = fminf(wx1, __shfl_xor_sync(SSDBOX_FULL_MASK, wx1, d));
        wy1 = fminf(wy1, __shfl_xor_sync(SSDBOX_FULL_MASK, wy1, d));
        wx2 = fmaxf(wx2, __shfl_xor_sync(SSDBOX_FULL_MASK, wx2, d));
        wy2 = fmaxf(wy2, __shfl_xor_sync(SSDBOX_FULL_MASK, wy2, d));
      }
      const uint32_t pbase = (uint32_t)(hd.p0 + r0);
      for (int gbase = 0; gbase < G; gbase += 32) {
        bool touch = false;
        if (gbase + lane < G) {
          float4 tv = ub.box[gbase + lane];
          touch = tv.x < wx2 && tv.z > wx1 && tv.y < wy2 && tv.w > wy1;
        }
        uint32_t todo = __ballot_sync(SSDBOX_FULL_MASK, touch);
#ifdef SSDBOX_PHASE_TIMING
        n_truths += __popc(todo);
#endif
        while (todo) {                    // ascending truth index: first truth wins ties (box_utils.py:118)
          const int g = gbase + __ffs(todo) - 1;
          todo &= todo - 1;
          float4 tv = ub.box[g];
          Box t;
          t.x1 = tv.x; t.y1 = tv.y; t.x2 = tv.z; t.y2 = tv.w;
          const float ta = ub.area[g];
          float iou[K];
          iou_jaccard_multi<K>(t, ta, box, area, iou);
          float lm = 0.0f;
          uint32_t lp = 0xffffffffu;
#pragma unroll
          for (int q = 0; q < K; ++q) {
            float v = valid[q] ? iou[q] : 0.0f;
            if (v > bt_ov[q]) {           // strict
              bt_ov[q] = v;
              bt_idx[q] = g;
            }
            if (v > lm) {                 // strict + ascending p: lowest prior wins ties (:116)
              lm = v;
              lp = pbase + q;
            }
          }
          uint32_t mb = __reduce_max_sync(SSDBOX_FULL_MASK, __float_as_uint(lm));
          if (mb != 0u) {
            uint32_t pm = __reduce_min_sync(SSDBOX_FULL_MASK, (__float_as_uint(lm) == mb) ? lp : 0xffffffffu);
            if (lane == 0) best_prior_update(&ub.best[g], mb, pm);
          }
        }
      }
      short4 lo, to;
      int16_t* lop = reinterpret_cast<int16_t*>(&lo);
      int16_t* top = reinterpret_cast<int16_t*>(&to);
#pragma unroll
      for (int q = 0; q < K; ++q) {
        lop[q] = (int16_t)((G > 0 && !(bt_ov[q] < a.threshold)) ? ub.lab[bt_idx[q]] : 0);   // :130
        top[q] = (int16_t)bt_idx[q];
      }
      const size_t row0 = (size_t)hd.b * (size_t)a.P + (size_t)(hd.p0 + r0);
      if (valid[K - 1] && (row0 & 3u) == 0) {        // 8-byte stores
        *reinterpret_cast<short4*>(a.lab_out + row0) = lo;
        *reinterpret_cast<short4*>(a.tidx_out + row0) = to;
      } else {
#pragma unroll
        for (int q = 0; q < K; ++q)
          if (valid[q]) {
            a.lab_out[row0 + q] = lop[q];
            a.tidx_out[row0 + q] = top[q];
          }
      }
    }
    // every match warp is done with the unit: publish its per-truth candidates, release the buffer
    asm volatile("bar.sync 1, %0;" ::"r"(a.match_warps * 32) : "memory");
    for (int g = mw * 32 + lane; g < G; g += a.match_warps * 32) {
      unsigned long long v = ub.best[g];
      if (v > kBestInit) atomicMax(&a.gt_best[(size_t)hd.b * a.gpad + g], v);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&mempty[buf]);
#ifdef SSDBOX_PHASE_TIMING
    t_busy += clock64() - c1;
    ++n_units;
#endif
  }
#ifdef SSDBOX_PHASE_TIMING
  if (mw == 0 && lane == 0 && blockIdx.x < 160) {
    g_mstat[blockIdx.x * 8 + 0] = t_wait;
    g_mstat[blockIdx.x * 8 + 1] = t_busy;
    g_mstat[blockIdx.x * 8 + 2] = n_units;
    g_mstat[blockIdx.x * 8 + 3] = n_truths;
  }
#endif
}

template <int CT, bool SHIFT = false>
__global__ void __launch_bounds__(kStreamThreads, 1) loss_stream_kernel(StreamArgs a) {
  extern __shared__ __align__(128) unsigned char smem_ring[];
  const int C = CT > 0 ? CT : a.ring.C;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = a.ring.R, NS = a.ring.NS;

  unsigned char* mbase = smem_ring + kRingHeaderBytes + (size_t)NS * R * C * 4;   // matching area behind the ring
  if (a.fuse && tid == 0) {
    uint64_t* mfull = reinterpret_cast<uint64_t*>(mbase);
    mbar_init(&mfull[0], 1);
    mbar_init(&mfull[1], 1);
    mbar_init(&mfull[2], a.match_warps);
    mbar_init(&mfull[3], a.match_warps);
  }
  if (warp == 0) SMARK(0);
  RingCtx rc = ring_setup(a.ring, smem_ring);   // fences the barrier inits, __syncthreads()
  if (warp == 0) SMARK(1);
  if (warp == kRingConsumerWarps) {
    ring_produce(a.ring, rc);
    SMARK(2);
    return;
  }
  if (warp == kSchedWarp) {
    if (a.fuse) match_sched_loop(a, mbase, lane);
    return;
  }
  if (warp >= kFirstMatchWarp) {
    if (warp == kFirstMatchWarp) SMARK(3);
    if (a.fuse) match_unit_loop(a, mbase, warp - kFirstMatchWarp, lane);
    if (warp == kFirstMatchWarp) SMARK(4);
    if (warp == kFirstMatchWarp + a.match_warps - 1) SMARK(5);
    return;
  }
  const int wg = warp / kRingGroupWarps;
  const int rbase = tid % (32 * kRingGroupWarps);
  const int KR = a.ring.KR;
  const float shift = SHIFT ? (float)*a.shift : 0.f;
  for (int it = wg; it < rc.n_local; it += kRingGroups) {
    const int s = it % NS, j = it % (2 * NS), ph = it / (2 * NS);
    mbar_wait(&rc.full[j], (uint32_t)(ph & 1));
    for (int k = 0; k < KR; ++k) {
      const int r = k * (32 * kRingGroupWarps) + rbase;
      const long long row = (rc.t0 + it * rc.tstep) * R + r;
      if ((r < R) && (row < a.ring.rows)) {
        const float* rp = rc.stages + (size_t)s * rc.stage_floats + (size_t)r * C;
        const float lse = SHIFT ? row_lse_shift(rp, C, shift) : row_lse<CT>(rp, C);
        const float k0 = lse - rp[0];                        // multibox_loss.py:94 with conf_t = 0
        a.key0[row] = k0;                                    // (a positive's lse is recovered as key0 + x[0]: no second array)
        if (refine_member(a.rf, a.pool, (size_t)row)) {
          uint32_t b = (uint32_t)row / (uint32_t)a.P;
          atomicAdd(&a.hist[(size_t)b * kHistBins + mine_bin(f2ord(k0))], 1u);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&rc.empty[j]);
  }
  if (warp == 0) SMARK(6);
  if (warp == 7) SMARK(7);
}

// Decides whether the matching is fused (its two unit buffers must fit beside >= 3 ring stages).
static int plan_stream(StreamArgs* a, const float* conf, long long rows, int C, int sm_count, int max_smem,
                       bool want_fuse, size_t* smem_out) {
  a->fuse = 0;
  a->unit_bytes = 0;
  a->match_warps = match_warps_for(C);
  a->match_tile = a->match_warps * 128;
  if (want_fuse) {
    int ub = unit_buf_bytes(a->gpad, a->match_tile, a->rf.arm_loc != nullptr);
    size_t need = (size_t)kMatchBarBytes + 2 * (size_t)ub;
    if (need + 65536 < (size_t)max_smem) {
      RingPlan rp;
      int rc = plan_ring(&rp, conf, rows, C, sm_count, max_smem - (int)need);
      if (rc == SSDBOX_OK && rp.NS >= 3) {
        a->ring = rp;
        a->fuse = 1;
        a->unit_bytes = ub;
        *smem_out = rp.smem_bytes + need;
        return SSDBOX_OK;
      }
    }
  }
  int rc = plan_ring(&a->ring, conf, rows, C, sm_count, max_smem);
  if (rc) return rc;
  *smem_out = a->ring.smem_bytes;
  return SSDBOX_OK;
}

static int launch_stream(const StreamArgs& a, size_t smem, cudaStream_t st) {
  const int C = a.ring.C;
  void (*kern)(StreamArgs) = loss_stream_kernel<0>;
  if (C == 81) kern = loss_stream_kernel<81>;
  else if (C == 21) kern = loss_stream_kernel<21>;
  else if (C == 2) kern = loss_stream_kernel<2>;
  if (a.shift) kern = loss_stream_kernel<0, true>;        // fidelity mode: any C through the generic instantiation
  SSDBOX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SSDBOX_CARVE(kern);
  {
    TimerScope ts__(KID_LOSS_STREAM, st);
    kern<<<a.ring.grid, (kRingConsumerWarps + 2 + (a.fuse ? a.match_warps : 0)) * 32, smem, st>>>(a);
  }
  SSDBOX_LAUNCH_OK("loss_stream_kernel");
  return SSDBOX_OK;
}

// ------------------------------------------------------------------------------------------------
// per-image selection of the K largest ordered keys (descending, ties by ascending index)
// ------------------------------------------------------------------------------------------------
constexpr int kMineThreads = 1024;
constexpr int kMineFixedSmem = 8192 + 800 + 288;   // level histogram + reduction scratch (99 doubles) + int scratch

struct MineArgs {
  int B, P, C;
  int negpos_ratio;
  float var0, var1;
  int finalize;
  long long prior_stride;
  const float* loc;
  const float* priors;
  const float* gt;
  const int32_t* gt_offsets;
  const uint8_t* pool;
  RefineArgs rf;           // RefineDet fused: pool = ARM objectness, anchors decoded from arm_loc for the positives
  const float* keys;       // key0 = lse - x[0] from the stream kernel
  const float* conf;       // gathered per positive: x[target] and x[0] (lse = key0 + x[0])
  const int16_t* lab;
  const int16_t* tidx;
  uint32_t* hist;          // level-1 histogram; positives are moved to the zero bin here
  // fused matching: the forced assignment (box_utils.py:123-127) is replayed here
  int fuse;
  int gmax, gpad, binarize;
  const unsigned long long* gt_best;
  unsigned long long* gt_best_w;   // same array: the mining CTA of an image hands it back initialised
  int16_t* lab_w;
  int16_t* tidx_w;
  uint32_t* ukey_global;   // used when the ordered keys do not fit in shared memory
  int uk_in_smem;
  double* partial;
  uint32_t* ticket;
  double* sums;
  float* losses;
  int16_t* sel;
  uint8_t* dbg_neg;
  float* dbg_keys;
  // multi-GPU: exchange buffers of every rank (world == 0: single GPU)
  int peer_rank, peer_world, peer_defer;
  long long peer_timeout_ns;
  void* peer_bufs[SSDBOX_MAX_PEERS];
};

#ifdef SSDBOX_PHASE_TIMING
__device__ long long g_phase[16 * 64];
#define PHASE_MARK(k) do { __syncthreads(); if (blockIdx.x < 64 && threadIdx.x == 0) { g_phase[blockIdx.x * 16 + (k)] = clock64(); \
    if ((k) == 0) g_phase[blockIdx.x * 16 + 14] = (long long)gtimer_ns(); if ((k) == 7) g_phase[blockIdx.x * 16 + 15] = (long long)gtimer_ns(); } } while (0)
#else
#define PHASE_MARK(k) do { } while (0)
#endif

// one warp; s[0..2] in shared memory holds this rank's sums on entry and (unless the wait is
// deferred to ssdbox_multibox_loss_peer_finish) the global sums on exit
__device__ void peer_exchange(const MineArgs& a, double* s, int lane) {
  const double v0 = s[0], v1 = s[1], v2 = s[2];
  __syncwarp();
  unsigned long long epoch = peer_post(a.peer_bufs, a.peer_rank, a.peer_world, v0, v1, v2, lane);
  if (a.peer_defer) return;
  double g[3];
  peer_collect(a.peer_bufs, a.peer_rank, a.peer_world, epoch, lane, g, a.peer_timeout_ns);
  if (lane == 0) {
    s[0] = g[0];
    s[1] = g[1];
    s[2] = g[2];
  }
}

__global__ void peer_finish_kernel(PeerFinishArgs a) { peer_finish_warp(a, threadIdx.x & 31); }

// a rank whose local shard is empty still takes part in the exchange: it posts {0, 0, 0} for this call
// (and, unless the wait is deferred, collects and finalises) so that the epochs of all ranks stay in step
__global__ void peer_empty_kernel(PeerFinishArgs a, int defer, int finalize) {
  const int lane = threadIdx.x & 31;
  const unsigned long long epoch = peer_post(a.bufs, a.rank, a.world, 0.0, 0.0, 0.0, lane);
  double g[3] = {0.0, 0.0, 0.0};
  if (!defer) peer_collect(a.bufs, a.rank, a.world, epoch, lane, g, a.timeout_ns);
  if (lane == 0) {
    a.sums[0] = g[0];
    a.sums[1] = g[1];
    a.sums[2] = g[2];
    if (finalize && a.losses) {
      a.losses[0] = g[2] == 0.0 ? 0.0f : (float)(g[0] / g[2]);
      a.losses[1] = g[2] == 0.0 ? 0.0f : (float)(g[1] / g[2]);
    }
  }
}

// last CTA: fold the per-image partials in image order (bit-reproducible run to run); the loads
// are spread over the threads (one L2 round trip), the fp64 adds stay sequential in image order
__device__ void fold_partials(const MineArgs& a, double* s_dscr, double* s_big, int nparts) {
  // s_big: the (dead by now) level-histogram area, 1024 doubles.  One round trip brings up to kChunk
  // partials in; three threads then add one quantity each in CTA order (bit-reproducible, the same
  // association as a sequential loop over the partials).
  constexpr int kChunk = 336;
  const int tid = threadIdx.x;
  double acc = 0.0;
  for (int base = 0; base < nparts; base += kChunk) {
    const int n = nparts - base < kChunk ? nparts - base : kChunk;
    __syncthreads();
    for (int t = tid; t < 3 * n; t += blockDim.x) s_big[t] = __ldcg(&a.partial[(size_t)base * 3 + t]);
    __syncthreads();
    if (tid < 3)
      for (int i = 0; i < n; ++i) acc += s_big[i * 3 + tid];
  }
  __syncthreads();
  if (tid < 3) s_dscr[tid] = acc;
  __syncthreads();
  double sl = s_dscr[0], sc = s_dscr[1], sn = s_dscr[2];
  if (a.peer_world > 0) {     // uniform over the CTA
    __syncthreads();          // every thread has read s_dscr[0..2]
    if (tid < 32) peer_exchange(a, s_dscr, tid);
    __syncthreads();
    sl = s_dscr[0];
    sc = s_dscr[1];
    sn = s_dscr[2];
  }
  if (tid == 0) {
    a.ticket[0] = 0u;      // mining ticket and matching work-unit counter: clean for the next call
    a.ticket[1] = 0u;
    a.sums[0] = sl;
    a.sums[1] = sc;
    a.sums[2] = sn;
    if (a.finalize && a.losses) {     // multibox_loss.py:114-116 (N == 0 -> 0 instead of inf/nan)
      a.losses[0] = sn == 0.0 ? 0.0f : (float)(sl / sn);
      a.losses[1] = sn == 0.0 ? 0.0f : (float)(sc / sn);
    }
  }
}

// per-prior work of pass A: returns the ordered mining key (0 = outside the ranking)
__device__ __forceinline__ uint32_t mine_visit(const MineArgs& a, size_t i, int p, int b, int g0, const float* pri,
                                               float key, int lb, int inpool, int& npos, double& ce, double& l1) {
  if (a.dbg_keys) a.dbg_keys[i] = key;
  if (!inpool) return 0u;
  if (lb > 0) {
    ++npos;
    // CE of a positive = lse - x[target] (multibox_loss.py:94,110); it ranks as 0 (multibox_loss.py:97)
    float cep = (key + a.conf[i * (size_t)a.C]) - a.conf[i * (size_t)a.C + lb];
    if (a.dbg_keys) a.dbg_keys[i] = cep;
    ce += (double)cep;
    const float* row = a.gt + (size_t)(g0 + a.tidx[i]) * 5;
    Box m;
    m.x1 = row[0]; m.y1 = row[1]; m.x2 = row[2]; m.y2 = row[3];
    float4 t = encode_box(m, refine_center(a.rf, *reinterpret_cast<const float4*>(pri + (size_t)p * 4), i), a.var0, a.var1);
    float4 l = *reinterpret_cast<const float4*>(a.loc + i * 4);
    l1 += (double)(smooth_l1(l.x, t.x) + smooth_l1(l.y, t.y) + smooth_l1(l.z, t.z) + smooth_l1(l.w, t.w));
    return f2ord(0.0f);
  }
  return f2ord(key);
}

// VEC = 4: four consecutive priors per thread and 16/8/4-byte vector accesses (needs P % 4 == 0)
template <int VEC>
__global__ void __launch_bounds__(kMineThreads, 1) mine_reduce_kernel(MineArgs a) {
  extern __shared__ __align__(16) unsigned char smem_mine[];
  unsigned char* smem_raw = smem_mine;
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw);             // 2048
  double* s_dscr = reinterpret_cast<double*>(smem_raw + 8192);          // 100
  int* s_iscr = reinterpret_cast<int*>(smem_raw + 8192 + 800);          // 64
  int* s_res = s_iscr + 64;                                             // 8
  // smem mode: ordered keys + class targets of the whole image stay on chip between the passes
  uint32_t* uk = a.uk_in_smem ? reinterpret_cast<uint32_t*>(smem_raw + kMineFixedSmem)
                              : a.ukey_global + (size_t)blockIdx.x * a.P;
  int16_t* s_lab = reinterpret_cast<int16_t*>(smem_raw + kMineFixedSmem + (size_t)a.P * 4);
  __shared__ int s_last;

  const int b = blockIdx.x, tid = threadIdx.x, P = a.P;
  const size_t off = (size_t)b * P;
  const int g0 = a.gt_offsets[b];
  const float* pri = a.priors + (size_t)b * (size_t)a.prior_stride;
  const bool in_smem = a.uk_in_smem != 0;

  // phase F (fused matching only): every truth claims its best prior, sequentially, last truth
  // wins; a claimed prior becomes positive with that truth's label (box_utils.py:123-130)
  if (a.fuse) {
    int G = a.gt_offsets[b + 1] - g0;
    G = G < 0 ? 0 : (G > a.gmax ? a.gmax : G);
    const unsigned long long* best = a.gt_best + (size_t)b * a.gpad;
    unsigned long long* s_best = reinterpret_cast<unsigned long long*>(s_hist);   // 1024 entries
    const bool cached = G <= 1024;
    if (cached) {
      for (int j = tid; j < G; j += kMineThreads) s_best[j] = __ldcg(&best[j]);
      __syncthreads();
    }
    for (int j = tid; j < G; j += kMineThreads) {
      const uint32_t pj = ~(uint32_t)((cached ? s_best[j] : __ldcg(&best[j])) & 0xffffffffull);
      bool winner = pj < (uint32_t)P;
      for (int j2 = j + 1; winner && j2 < G; ++j2)
        if (~(uint32_t)((cached ? s_best[j2] : __ldcg(&best[j2])) & 0xffffffffull) == pj) winner = false;
      if (winner) {
        a.lab_w[off + pj] = (int16_t)(a.binarize ? 1 : (int)(a.gt[(size_t)(g0 + j) * 5 + 4] + 1.0f));
        a.tidx_w[off + pj] = (int16_t)j;
      }
    }
    __syncthreads();
    for (int j = tid; j < G; j += kMineThreads) a.gt_best_w[(size_t)b * a.gpad + j] = kBestInit;   // state handed back clean
  }

  // pass A: positives (count, CE, smooth-L1) and the ordered mining keys
  PHASE_MARK(0);
  int npos = 0;
  double ce = 0.0, l1 = 0.0;
  if (VEC == 4) {
    const float4* k4 = reinterpret_cast<const float4*>(a.keys + off);
    const short4* l4 = reinterpret_cast<const short4*>(a.lab + off);
    const uchar4* p4 = a.pool ? reinterpret_cast<const uchar4*>(a.pool + off) : nullptr;
    const int n4 = P >> 2;
    for (int q0 = tid; q0 < n4; q0 += 2 * kMineThreads) {
      float4 kk[2];
      short4 ll[2];
      uchar4 pp[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        int q = q0 + u * kMineThreads;
        kk[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        ll[u] = make_short4(0, 0, 0, 0);
        pp[u] = make_uchar4(1, 1, 1, 1);
        if (q < n4) {
          kk[u] = k4[q];
          ll[u] = l4[q];
          if (p4) pp[u] = p4[q];
          if (a.rf.arm_conf) {
            const size_t r0 = off + (size_t)q * 4;
            pp[u] = make_uchar4(refine_keeps(a.rf.arm_conf, r0, a.rf.theta), refine_keeps(a.rf.arm_conf, r0 + 1, a.rf.theta),
                                refine_keeps(a.rf.arm_conf, r0 + 2, a.rf.theta), refine_keeps(a.rf.arm_conf, r0 + 3, a.rf.theta));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        int q = q0 + u * kMineThreads;
        if (q >= n4) continue;
        int p = q * 4;
        size_t i = off + p;
        uint4 uo;
        uo.x = mine_visit(a, i, p, b, g0, pri, kk[u].x, ll[u].x, pp[u].x, npos, ce, l1);
        uo.y = mine_visit(a, i + 1, p + 1, b, g0, pri, kk[u].y, ll[u].y, pp[u].y, npos, ce, l1);
        uo.z = mine_visit(a, i + 2, p + 2, b, g0, pri, kk[u].z, ll[u].z, pp[u].z, npos, ce, l1);
        uo.w = mine_visit(a, i + 3, p + 3, b, g0, pri, kk[u].w, ll[u].w, pp[u].w, npos, ce, l1);
        *reinterpret_cast<uint4*>(uk + p) = uo;
        if (in_smem) *reinterpret_cast<short4*>(s_lab + p) = ll[u];
      }
    }
  } else {
    for (int p = tid; p < P; p += kMineThreads) {
      size_t i = off + p;
      float key = a.keys[i];
      int lb = a.lab[i];
      int inpool = refine_member(a.rf, a.pool, i) ? 1 : 0;
      uk[p] = mine_visit(a, i, p, b, g0, pri, key, lb, inpool, npos, ce, l1);
      if (in_smem) s_lab[p] = (int16_t)lb;
    }
  }
  PHASE_MARK(1);
  double npos_d = (double)npos;
  block_sum3(npos_d, ce, l1, s_dscr);
  int npos_blk = (int)(npos_d + 0.5);
  PHASE_MARK(2);

  // multibox_loss.py:101-102  num_neg = clamp(ratio * num_pos, max = P - 1)
  long long kk = (long long)a.negpos_ratio * npos_blk;
  if (kk > P - 1) kk = P - 1;
  const int K = (int)kk;
  uint32_t Tu = 0xffffffffu;
  // (the generic kernel builds its own radix histograms; the streamed one -- fine bins, see mine_bin -- is only
  // handed back clean below)
  if (K > 0) Tu = cta_select_threshold<false, VEC>(uk, P, K, nullptr, s_hist, s_iscr, s_res);
  __syncthreads();
  for (int i = tid; i < kHistBins; i += kMineThreads) a.hist[(size_t)b * kHistBins + i] = 0u;     // state handed back clean
  PHASE_MARK(3);

  // final pass: neg = rank < num_neg (:103); CE over pos U neg (:106-110).  The CE of a selected
  // negative is its mining key, recovered exactly from the ordered key (no second read of keys).
  double ce_neg = 0.0;
  auto decide = [&](uint32_t u, int lb, bool& negsel) -> int16_t {
    bool inpool = u != 0u;
    bool is_pos = inpool && lb > 0;
    negsel = K > 0 && inpool && u >= Tu;
    if (negsel && !is_pos) ce_neg += (double)ord2f(u);
    return is_pos ? (int16_t)lb : (negsel ? (int16_t)0 : (int16_t)-1);
  };
  if (VEC == 4) {
    const int n4 = P >> 2;
    for (int q = tid; q < n4; q += kMineThreads) {
      int p = q * 4;
      size_t i = off + p;
      uint4 u = *reinterpret_cast<const uint4*>(uk + p);
      short4 lb = in_smem ? *reinterpret_cast<const short4*>(s_lab + p) : *reinterpret_cast<const short4*>(a.lab + i);
      bool n0, n1, n2, n3;
      short4 so;
      so.x = decide(u.x, lb.x, n0);
      so.y = decide(u.y, lb.y, n1);
      so.z = decide(u.z, lb.z, n2);
      so.w = decide(u.w, lb.w, n3);
      *reinterpret_cast<short4*>(a.sel + i) = so;
      if (a.dbg_neg) *reinterpret_cast<uchar4*>(a.dbg_neg + i) = make_uchar4(n0, n1, n2, n3);
    }
  } else {
    for (int p = tid; p < P; p += kMineThreads) {
      size_t i = off + p;
      bool ns;
      a.sel[i] = decide(uk[p], in_smem ? (int)s_lab[p] : (int)a.lab[i], ns);
      if (a.dbg_neg) a.dbg_neg[i] = ns ? 1 : 0;
    }
  }
  PHASE_MARK(4);
  ce_neg = block_sum(ce_neg, s_dscr);
  PHASE_MARK(5);

  if (tid == 0) {
    a.partial[(size_t)b * 3 + 0] = l1;
    a.partial[(size_t)b * 3 + 1] = ce + ce_neg;
    a.partial[(size_t)b * 3 + 2] = (double)npos_blk;
    __threadfence();
    unsigned t = atomicAdd(a.ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  fold_partials(a, s_dscr, reinterpret_cast<double*>(smem_raw), a.B);
}

// ------------------------------------------------------------------------------------------------
// Register-resident variant for P % 4 == 0 and P <= 4 * kMineQ * 1024 (every SSD / RFB / FSSD head):
// each thread keeps its 24 mining keys and class targets in registers from the single load to the
// final store, so the kernel is one chain of a few memory round trips instead of several sweeps:
//   (1) keys, class targets, level-1 histogram, per-truth best priors and the truth rows are all
//       requested at once; (2) the forced assignment is replayed on a small shared list and patched
//       into the owners' registers; (3) positives are compacted (deterministic block scan) and
//       handled one per thread, so their dependent gathers (lse, the target logit from DRAM, loc,
//       prior) cost one round trip for the whole image; (4) radix select on the registers;
//   (5) sel / ce are produced from the registers.  Same results as mine_reduce_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kMineQ = 6;
constexpr int kForceListMax = 128;
constexpr int kRankCap = 1024;      // members of the threshold bin ranked by counting (more: radix fallback)

__device__ __forceinline__ int reg_find(const uint32_t* s_hist, int nbins, int K, int* s_iscr, int* s_res, int* above) {
  find_digit(s_hist, nbins, K, s_iscr, s_res);
  *above = s_res[1];
  return s_res[0];
}

// T = threads per CTA: 1024 (one CTA per SM), or 512 with two CTAs per SM for batches of more images than SMs -- the
// 256 images of RFB300-VOC then mine in ONE wave of 256 resident CTAs instead of two waves of 148.
template <int CL, int T = kMineThreads>
__global__ void __launch_bounds__(T, kMineThreads / T) mine_reduce_reg_kernel(MineArgs a) {
  constexpr int Q = kMineQ / CL;     // quads (4 priors) per thread
  constexpr int HB = kHistBins / T;  // histogram bins per thread
  extern __shared__ __align__(16) unsigned char smem_mine[];
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_mine);                        // 2048
  double* s_dscr = reinterpret_cast<double*>(smem_mine + 8192);                     // 100
  int* s_iscr = reinterpret_cast<int*>(smem_mine + 8192 + 800);                     // 64
  int* s_res = s_iscr + 64;                                                         // 8
  unsigned long long* s_best = reinterpret_cast<unsigned long long*>(smem_mine + kMineFixedSmem);   // kForceListMax
  float* s_gt = reinterpret_cast<float*>(s_best + 2 * kForceListMax);               // kForceListMax x 5 (+3 pad)
  uint32_t* s_list = reinterpret_cast<uint32_t*>(s_gt + 5 * kForceListMax + 8);     // P / CL + 4: prior | class << 16
  int16_t* s_ovr = reinterpret_cast<int16_t*>(s_list + ((a.P >> 2) + CL - 1) / CL * 4 + 4);   // forced label of a prior of this CTA, -1 = none
  __shared__ __align__(8) unsigned long long s_cand[kRankCap];      // members of the bin that holds the K-th key
  __shared__ unsigned long long s_x[1];
  __shared__ uint32_t s_ncand[1];
  __shared__ int s_last;
  __shared__ uint32_t s_xch[4];      // [0..1] positives per CTA of the cluster, [2..3] tie counts

  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int b = blockIdx.x / CL, tid = threadIdx.x, P = a.P;
  const size_t off = (size_t)b * P;
  const int g0 = a.gt_offsets[b];
  int G = a.gt_offsets[b + 1] - g0;
  G = G < 0 ? 0 : (G > a.gmax ? a.gmax : G);
  const float* pri = a.priors + (size_t)b * (size_t)a.prior_stride;
  const int n4 = P >> 2;
  const int n4h = (n4 + CL - 1) / CL;          // quads owned by one CTA of the cluster (contiguous range)
  const int q0 = rank * n4h;
  auto sync_all = [&]() {
    if (CL > 1) cluster_sync_all(); else __syncthreads();
  };
  const bool small_g = a.fuse && G <= kForceListMax;

  PHASE_MARK(0);
  // (1) everything this CTA needs from memory, requested together.  The loads that do not depend on the image's
  // truth count go first: a warp issues in order, and the truth offsets take a round trip of their own.
  uint32_t hreg[HB];
#pragma unroll
  for (int i = 0; i < HB; ++i) hreg[i] = a.hist[(size_t)b * kHistBins + i * T + tid];
  uint32_t uk[Q][4];     // ordered mining keys (0 = outside the ranking)
  short4 ll[Q];
  uint32_t poolmask = 0u;     // bit j*4+e: prior is ranked (inside P and inside the caller's pool)
  const float4* k4 = reinterpret_cast<const float4*>(a.keys + off);
  const short4* l4 = reinterpret_cast<const short4*>(a.lab + off);
  const uchar4* p4 = a.pool ? reinterpret_cast<const uchar4*>(a.pool + off) : nullptr;
#pragma unroll
  for (int j = 0; j < Q; ++j) {
    const int ql = tid + j * T, q = q0 + ql;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ql < n4h && q < n4) {
      kv = k4[q];
      uint32_t m = 0xfu;
      if (p4) {
        uchar4 pv = p4[q];
        m = (pv.x ? 1u : 0u) | (pv.y ? 2u : 0u) | (pv.z ? 4u : 0u) | (pv.w ? 8u : 0u);
      }
      if (a.rf.arm_conf) {       // RefineDet fused: the ARM objectness of the quad's four anchors (two 16-byte loads)
        const float4* c4 = reinterpret_cast<const float4*>(a.rf.arm_conf + (off + (size_t)q * 4) * 2);
        const float4 ca = c4[0], cb = c4[1];
        const float o0 = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(__fsub_rn(ca.x, ca.y))));
        const float o1 = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(__fsub_rn(ca.z, ca.w))));
        const float o2 = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(__fsub_rn(cb.x, cb.y))));
        const float o3 = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(__fsub_rn(cb.z, cb.w))));
        m = (o0 > a.rf.theta ? 1u : 0u) | (o1 > a.rf.theta ? 2u : 0u) | (o2 > a.rf.theta ? 4u : 0u) | (o3 > a.rf.theta ? 8u : 0u);
      }
      poolmask |= m << (4 * j);
    }
    uk[j][0] = f2ord(kv.x); uk[j][1] = f2ord(kv.y); uk[j][2] = f2ord(kv.z); uk[j][3] = f2ord(kv.w);
  }
  unsigned long long my_best = 0ull;
  if (small_g && tid < G) my_best = __ldcg(&a.gt_best[(size_t)b * a.gpad + tid]);
  const bool gt_cached = G <= kForceListMax;
  float my_gt = 0.f, my_gt2 = 0.f;      // 5 * G <= 640 floats: one per thread, a second one for the 512-thread CTA
  if (gt_cached && tid < 5 * G) my_gt = a.gt[(size_t)g0 * 5 + tid];
  if (T < 5 * kForceListMax && gt_cached && tid + T < 5 * G) my_gt2 = a.gt[(size_t)g0 * 5 + tid + T];
  if (a.fuse && !small_g) {
    // many truths: replay the forced assignment through global memory first (box_utils.py:123-130)
    const unsigned long long* best = a.gt_best + (size_t)b * a.gpad;
    for (int j = tid; j < G && rank == 0; j += T) {
      const uint32_t pj = ~(uint32_t)(__ldcg(&best[j]) & 0xffffffffull);
      bool winner = pj < (uint32_t)P;
      for (int j2 = j + 1; winner && j2 < G; ++j2)
        if (~(uint32_t)(__ldcg(&best[j2]) & 0xffffffffull) == pj) winner = false;
      if (winner) {
        a.lab_w[off + pj] = (int16_t)(a.binarize ? 1 : (int)(a.gt[(size_t)(g0 + j) * 5 + 4] + 1.0f));
        a.tidx_w[off + pj] = (int16_t)j;
      }
    }
    if (CL > 1) __threadfence();
    sync_all();
  }
#pragma unroll
  for (int j = 0; j < Q; ++j) {
    const int ql = tid + j * T, q = q0 + ql;
    ll[j] = make_short4(0, 0, 0, 0);
    if (ql < n4h && q < n4) ll[j] = l4[q];
  }
#pragma unroll
  for (int i = 0; i < HB; ++i) s_hist[i * T + tid] = hreg[i];
  if (small_g) {
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      const int ql = tid + j * T;
      if (ql < n4h) *reinterpret_cast<short4*>(s_ovr + ql * 4) = make_short4(-1, -1, -1, -1);
    }
  }
  if (small_g && tid < G) s_best[tid] = my_best;
  if (gt_cached && tid < 5 * G) s_gt[tid] = my_gt;
  if (T < 5 * kForceListMax && gt_cached && tid + T < 5 * G) s_gt[tid + T] = my_gt2;
  sync_all();          // cluster: the peer's histogram copy is in place before it receives remote updates
  // every CTA of the image has read its streamed histogram and best-prior keys: hand them back initialised
  // (the workspace state is clean again after the call, SSDBOX_LOSS_WS_CLEAN)
  if (rank == 0) {
#pragma unroll
    for (int i = 0; i < HB; ++i) a.hist[(size_t)b * kHistBins + i * T + tid] = 0u;
    if (a.fuse)
      for (int j = tid; j < G; j += T) a.gt_best_w[(size_t)b * a.gpad + j] = kBestInit;
  }

  PHASE_MARK(1);
  // (2) forced assignment: truth j keeps its best prior unless a later truth claims the same prior
  // (last truth wins); winners drop their label into a per-prior override array in shared memory
  // that every thread merges into its registers (one 8-byte shared load per quad)
  if (small_g) {
    const uint32_t gmask = __ballot_sync(SSDBOX_FULL_MASK, tid < G);     // whole warps reach this point together
    if (tid < G) {
      const uint32_t pj = ~(uint32_t)(s_best[tid] & 0xffffffffull);
      bool winner = pj < (uint32_t)P;
      if (G <= 32) {                 // one warp holds every truth: the highest lane with this prior wins
        const uint32_t same = __match_any_sync(gmask, pj);
        winner = winner && (31 - __clz(same)) == (int)(tid & 31);
      } else {
        for (int j2 = tid + 1; winner && j2 < G; ++j2)
          if (~(uint32_t)(s_best[j2] & 0xffffffffull) == pj) winner = false;
      }
      if (winner) {
        const int lb = a.binarize ? 1 : (int)(s_gt[tid * 5 + 4] + 1.0f);
        if (rank == 0) a.lab_w[off + pj] = (int16_t)lb;
        a.tidx_w[off + pj] = (int16_t)tid;      // every CTA of the cluster: its positives read it back below
        const int ql = (int)(pj >> 2) - q0;
        if (ql >= 0 && ql < n4h) s_ovr[ql * 4 + (int)(pj & 3u)] = (int16_t)lb;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      const int ql = tid + j * T;
      if (ql < n4h) {
        const short4 o = *reinterpret_cast<const short4*>(s_ovr + ql * 4);
        if (o.x >= 0) ll[j].x = o.x;
        if (o.y >= 0) ll[j].y = o.y;
        if (o.z >= 0) ll[j].z = o.z;
        if (o.w >= 0) ll[j].w = o.w;
      }
    }
  }

  PHASE_MARK(2);
  // (3) ordered keys in registers; positives leave their streamed bin for the zero bin and are
  // compacted into s_list in a fixed (thread-major) order
  int mypos = 0;
  const uint32_t zero_ord = f2ord(0.0f);
  const uint32_t rhist = CL > 1 ? peer_smem(s_hist, (uint32_t)(rank ^ 1)) : 0u;
#pragma unroll
  for (int j = 0; j < Q; ++j) {
    const int ql = tid + j * T, q = q0 + ql;
    if (a.dbg_keys && ql < n4h && q < n4)
      *reinterpret_cast<float4*>(a.dbg_keys + off + (size_t)q * 4) =
          make_float4(ord2f(uk[j][0]), ord2f(uk[j][1]), ord2f(uk[j][2]), ord2f(uk[j][3]));
    const int lb[4] = {ll[j].x, ll[j].y, ll[j].z, ll[j].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uint32_t u = 0u;
      if ((poolmask >> (4 * j + e)) & 1u) {
        u = uk[j][e];
        if (lb[e] > 0) {
          ++mypos;
          atomicSub(&s_hist[mine_bin(u)], 1u);
          atomicAdd(&s_hist[0], 1u);                       // mine_bin(zero_ord) == 0
          if (CL > 1) {
            peer_red_add(rhist + mine_bin(u) * 4u, 0xffffffffu);
            peer_red_add(rhist, 1u);
          }
          u = zero_ord;
        }
      }
      uk[j][e] = u;
    }
  }
  int npos_blk;
  int slot = block_exclusive_scan(mypos, s_iscr, &npos_blk);
  if (mypos) {
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      const int lb[4] = {ll[j].x, ll[j].y, ll[j].z, ll[j].w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (((poolmask >> (4 * j + e)) & 1u) && lb[e] > 0)
          s_list[slot++] = (uint32_t)((q0 + tid + j * T) * 4 + e) | ((uint32_t)lb[e] << 16);
    }
  }
  if (CL > 1 && tid == 0) {       // positives of this CTA -> both CTAs of the cluster
    s_xch[rank] = (uint32_t)npos_blk;
    peer_st_u32(peer_smem(&s_xch[rank], (uint32_t)(rank ^ 1)), (uint32_t)npos_blk);
  }
  __syncthreads();

  PHASE_MARK(3);
  int npos_img = npos_blk;
  if (CL > 1) {
    cluster_sync_all();           // remote histogram updates and positive counts have landed
    npos_img = (int)(s_xch[0] + s_xch[1]);
  }
  // one positive per thread: CE = lse - x[target] (multibox_loss.py:94,110) and smooth-L1 against
  // the encoded truth (:87-90, box_utils.py:215-222).  The gathers of the first 1024 positives (all of
  // them in practice) are only REQUESTED here -- the target logit comes from DRAM -- and consumed after
  // the selection, which does not depend on them.
  const bool havepos = tid < npos_blk;
  float p_lse = 0.f, p_xt = 0.f;
  int p_t = 0, p_p = 0;
  float4 p_l = make_float4(0.f, 0.f, 0.f, 0.f), p_pr = make_float4(1.f, 1.f, 1.f, 1.f);
  if (havepos) {
    const uint32_t ent = s_list[tid];
    p_p = (int)(ent & 0xffffu);
    const int lb = (int)(ent >> 16);
    const size_t i = off + p_p;
    p_lse = a.keys[i] + a.conf[i * (size_t)a.C];       // lse = key0 + x[0]
    p_xt = a.conf[i * (size_t)a.C + lb];
    p_t = a.tidx[i];
    p_l = *reinterpret_cast<const float4*>(a.loc + i * 4);
    p_pr = refine_center(a.rf, *reinterpret_cast<const float4*>(pri + (size_t)p_p * 4), i);
  }

  PHASE_MARK(4);
  // multibox_loss.py:101-102  num_neg = clamp(ratio * num_pos, max = P - 1)
  long long kk64 = (long long)a.negpos_ratio * npos_img;
  if (kk64 > P - 1) kk64 = P - 1;
  const int K = (int)kk64;

  // (4) the K-th largest ranked key.  X = (ordered key << 32 | ~prior) of the last selected element in the order
  // "key descending, prior ascending" (a stable descending sort, multibox_loss.py:99-103): an element is selected
  // iff its own packed value is >= X.  The streamed histogram has fine bins (mine_bin): the bin of the K-th key
  // usually has a handful of members, which are gathered into a shared list (both CTAs of a cluster push to both
  // lists) and ranked by counting -- two barriers instead of two more radix levels.  A crowded bin (tied keys)
  // falls back to the exact 11 + 11 + 10 bit radix select on the registers.
  unsigned long long X = ~0ull;
  if (K > 0) {
    int above;
    const int d1 = reg_find(s_hist, kHistBins, K, s_iscr, s_res, &above);
    if (d1 < 0) {
      X = 1ull << 32;      // fewer than K ranked elements: take them all (ranked keys are >= 1)
    } else {
      const int need = K - above;              // 1 <= need <= n1
      const int n1 = (int)s_hist[d1];
      __syncthreads();
      if (n1 <= kRankCap) {
        if (tid == 0) *s_ncand = 0u;
        sync_all();
        const uint32_t rcnt = CL > 1 ? peer_smem(s_ncand, (uint32_t)(rank ^ 1)) : 0u;
        const uint32_t rcand = CL > 1 ? peer_smem(s_cand, (uint32_t)(rank ^ 1)) : 0u;
#pragma unroll
        for (int j = 0; j < Q; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t u = uk[j][e];
            if (u && (int)mine_bin(u) == d1) {
              const uint32_t p = (uint32_t)((q0 + tid + j * T) * 4 + e);
              const unsigned long long v = ((unsigned long long)u << 32) | (unsigned long long)(uint32_t)(~p);
              s_cand[atomicAdd(s_ncand, 1u)] = v;
              if (CL > 1) peer_st_u64(rcand + peer_atom_add(rcnt, 1u) * 8u, v);
            }
          }
        sync_all();
        for (int t = tid; t < n1; t += T) {     // distinct values: exactly one has `need - 1` larger ones
          const unsigned long long x = s_cand[t];
          int larger = 0;
          for (int i = 0; i < n1; ++i) larger += s_cand[i] > x ? 1 : 0;
          if (larger == need - 1) *s_x = x;
        }
        __syncthreads();
        X = *s_x;
      } else {
        // crowded bin: exact radix select, level 1 rebuilt from the registers (top 11 bits of the ordered key)
        uint32_t Tu;
        for (int i = tid; i < kHistBins; i += T) s_hist[i] = 0u;
        sync_all();
#pragma unroll
        for (int j = 0; j < Q; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint32_t u = uk[j][e];
            if (u) {
              atomicAdd(&s_hist[u >> 21], 1u);
              if (CL > 1) peer_red_add(rhist + (u >> 21) * 4u, 1u);
            }
          }
        sync_all();
        const int r1 = reg_find(s_hist, kHistBins, K, s_iscr, s_res, &above);      // >= K ranked elements exist: r1 >= 0
        int K2 = K - above;
        int n1r = (int)s_hist[r1];
        __syncthreads();
        if (K2 == n1r) {
          Tu = ((uint32_t)r1 << 21) ? ((uint32_t)r1 << 21) : 1u;
        } else {
          for (int i = tid; i < kHistBins; i += T) s_hist[i] = 0u;
          sync_all();
#pragma unroll
          for (int j = 0; j < Q; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              uint32_t u = uk[j][e];
              if (u && (int)(u >> 21) == r1) {
                atomicAdd(&s_hist[(u >> 10) & 2047u], 1u);
                if (CL > 1) peer_red_add(rhist + ((u >> 10) & 2047u) * 4u, 1u);
              }
            }
          sync_all();
          int d2 = reg_find(s_hist, 2048, K2, s_iscr, s_res, &above);
          int K3 = K2 - above;
          int n2 = (int)s_hist[d2];
          __syncthreads();
          const uint32_t pre2 = ((uint32_t)r1 << 11) | (uint32_t)d2;
          if (K3 == n2) {
            Tu = (pre2 << 10) ? (pre2 << 10) : 1u;
          } else {
            for (int i = tid; i < 1024; i += T) s_hist[i] = 0u;
            sync_all();
#pragma unroll
            for (int j = 0; j < Q; ++j)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                uint32_t u = uk[j][e];
                if (u && (u >> 10) == pre2) {
                  atomicAdd(&s_hist[u & 1023u], 1u);
                  if (CL > 1) peer_red_add(rhist + (u & 1023u) * 4u, 1u);
                }
              }
            sync_all();
            int d3 = reg_find(s_hist, 1024, K3, s_iscr, s_res, &above);
            int need3 = K3 - above;
            int n3 = (int)s_hist[d3];
            __syncthreads();
            Tu = (pre2 << 10) | (uint32_t)d3;
            if (need3 != n3) {
              // ties straddle the cut: equal keys win in ascending prior order (stable descending sort)
              int running = 0;
              if (CL > 1) {         // the lower half of the priors (CTA 0) ranks first
                int mine_ties = 0;
#pragma unroll
                for (int j = 0; j < Q; ++j)
#pragma unroll
                  for (int e = 0; e < 4; ++e) mine_ties += uk[j][e] == Tu ? 1 : 0;
                int total_ties;
                block_exclusive_scan(mine_ties, s_iscr, &total_ties);
                if (tid == 0) {
                  s_xch[2 + rank] = (uint32_t)total_ties;
                  peer_st_u32(peer_smem(&s_xch[2 + rank], (uint32_t)(rank ^ 1)), (uint32_t)total_ties);
                }
                cluster_sync_all();
                running = rank == 0 ? 0 : (int)s_xch[2];
              }
#pragma unroll
              for (int j = 0; j < Q; ++j) {
                int cnt = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) cnt += uk[j][e] == Tu ? 1 : 0;
                int total;
                int ex = block_exclusive_scan(cnt, s_iscr, &total);
                int rk = running + ex;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (uk[j][e] == Tu) {
                    if (rk >= need3) uk[j][e] = Tu - 1u;
                    ++rk;
                  }
                running += total;
                __syncthreads();
              }
            }
          }
        }
        X = (unsigned long long)Tu << 32;      // tie losers were demoted below Tu: selected <=> key >= Tu
      }
    }
  }

  PHASE_MARK(5);
  double ce = 0.0, l1 = 0.0;
  auto positive = [&](size_t i, int p, float lse, float xt, int t, float4 l, float4 pr) {
    const float* row = gt_cached ? s_gt + t * 5 : a.gt + (size_t)(g0 + t) * 5;
    Box m;
    m.x1 = row[0]; m.y1 = row[1]; m.x2 = row[2]; m.y2 = row[3];
    const float cep = lse - xt;
    if (a.dbg_keys) a.dbg_keys[i] = cep;
    ce += (double)cep;
    float4 tt = encode_box(m, pr, a.var0, a.var1);
    l1 += (double)(smooth_l1(l.x, tt.x) + smooth_l1(l.y, tt.y) + smooth_l1(l.z, tt.z) + smooth_l1(l.w, tt.w));
  };
  if (havepos) positive(off + p_p, p_p, p_lse, p_xt, p_t, p_l, p_pr);
  for (int sidx = tid + T; sidx < npos_blk; sidx += T) {
    const uint32_t ent = s_list[sidx];
    const int p = (int)(ent & 0xffffu), lb = (int)(ent >> 16);
    const size_t i = off + p;
    positive(i, p, a.keys[i] + a.conf[i * (size_t)a.C], a.conf[i * (size_t)a.C + lb], a.tidx[i], *reinterpret_cast<const float4*>(a.loc + i * 4),
             refine_center(a.rf, *reinterpret_cast<const float4*>(pri + (size_t)p * 4), i));
  }
  // (5) neg = rank < num_neg (:103); CE over pos U neg (:106-110); the CE of a selected negative is
  // its mining key, recovered exactly from the ordered key
  double ce_neg = 0.0;
#pragma unroll
  for (int j = 0; j < Q; ++j) {
    const int ql = tid + j * T, q = q0 + ql;
    const int lb[4] = {ll[j].x, ll[j].y, ll[j].z, ll[j].w};
    int16_t so[4];
    unsigned char ng[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t u = uk[j][e];
      const bool inpool = u != 0u;
      const bool is_pos = inpool && lb[e] > 0;
      const uint32_t pidx = (uint32_t)(q * 4 + e);
      const bool negsel = K > 0 && inpool && (((unsigned long long)u << 32) | (unsigned long long)(uint32_t)(~pidx)) >= X;
      if (negsel && !is_pos) ce_neg += (double)ord2f(u);
      so[e] = is_pos ? (int16_t)lb[e] : (negsel ? (int16_t)0 : (int16_t)-1);
      ng[e] = negsel ? 1 : 0;
    }
    if (ql < n4h && q < n4) {
      *reinterpret_cast<short4*>(a.sel + off + (size_t)q * 4) = make_short4(so[0], so[1], so[2], so[3]);
      if (a.dbg_neg) *reinterpret_cast<uchar4*>(a.dbg_neg + off + (size_t)q * 4) = make_uchar4(ng[0], ng[1], ng[2], ng[3]);
    }
  }
  PHASE_MARK(6);
  block_sum3(ce_neg, ce, l1, s_dscr);       // three independent fixed-shape trees, one set of barriers
  PHASE_MARK(7);

  if (tid == 0) {
    a.partial[(size_t)blockIdx.x * 3 + 0] = l1;
    a.partial[(size_t)blockIdx.x * 3 + 1] = ce + ce_neg;
    a.partial[(size_t)blockIdx.x * 3 + 2] = (double)npos_blk;
    __threadfence();
    unsigned t = atomicAdd(a.ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  fold_partials(a, s_dscr, reinterpret_cast<double*>(smem_mine), (int)gridDim.x);
  PHASE_MARK(8);
}

__global__ void finalize_kernel(const double* __restrict__ sums, float* __restrict__ losses) {
  double n = sums[2];
  losses[0] = n == 0.0 ? 0.0f : (float)(sums[0] / n);      // NaN sums (a peer timed out) stay NaN
  losses[1] = n == 0.0 ? 0.0f : (float)(sums[1] / n);
}

// stand-alone mining on caller-supplied keys (multibox_loss.py:97-103 in isolation)
__global__ void __launch_bounds__(kMineThreads, 1)
mine_only_kernel(const float* __restrict__ keys, const uint8_t* __restrict__ pos, const uint8_t* __restrict__ pool, int P,
                 int ratio, uint8_t* __restrict__ neg, uint32_t* ukey_global, int uk_in_smem) {
  extern __shared__ __align__(16) unsigned char smem_mine[];
  unsigned char* smem_raw = smem_mine;
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw);
  double* s_dscr = reinterpret_cast<double*>(smem_raw + 8192);
  int* s_iscr = reinterpret_cast<int*>(smem_raw + 8192 + 800);
  int* s_res = s_iscr + 64;
  uint32_t* uk = uk_in_smem ? reinterpret_cast<uint32_t*>(smem_raw + kMineFixedSmem)
                            : ukey_global + (size_t)blockIdx.x * P;
  const int tid = threadIdx.x;
  const size_t off = (size_t)blockIdx.x * P;
  int npos = 0;
  for (int p = tid; p < P; p += kMineThreads) {
    int inpool = pool ? pool[off + p] : 1;
    int ps = pos[off + p] != 0;
    uint32_t u = 0u;
    if (inpool) {
      if (ps) { ++npos; u = f2ord(0.0f); } else u = f2ord(keys[off + p]);
    }
    uk[p] = u;
  }
  int npos_blk = (int)(block_sum((double)npos, s_dscr) + 0.5);
  long long kk = (long long)ratio * npos_blk;
  if (kk > P - 1) kk = P - 1;
  const int K = (int)kk;
  uint32_t Tu = 0xffffffffu;
  if (K > 0) Tu = cta_select_threshold<false>(uk, P, K, nullptr, s_hist, s_iscr, s_res);
  __syncthreads();
  for (int p = tid; p < P; p += kMineThreads) {
    uint32_t u = uk[p];
    neg[off + p] = (K > 0 && u != 0u && u >= Tu) ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------
// backward: d(loss_l)/d(loc) and d(loss_c)/d(conf)   (autograd of multibox_loss.py:87-116)
// ------------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 256;
#ifdef SSDBOX_EXPERIMENTS
#define SSDBOX_ABLATE(bit) ((a.ablate & (bit)) != 0)
#else
#define SSDBOX_ABLATE(bit) false
#endif

struct BwdArgs {
  int B, P, C;
  float var0, var1;
  long long prior_stride;
  const float* loc;
  const float* conf;
  const float* priors;
  const float* gt;
  const int32_t* gt_offsets;
  const int16_t* sel;
  const int16_t* tidx;
  const double* sums;
  const float* grad_out;
  float* grad_loc;
  float* grad_conf;
  int conf_aligned;
  RefineArgs rf;           // RefineDet fused: the positives' anchors are decoded from arm_loc
  int ablate;              // experiment builds only (SSDBOX_BWD_ABLATE): 1 no grad_loc, 2 no gradient rows, 4 no positives, 8 no sel,
                           // 16 no phase 0, 32 bulk stores without the evict_first hint, 64 phase 0 keeps the flags only
};

// (1) grad_conf := 0 at full store bandwidth (only ~4*num_pos rows per image are ever non-zero)
__global__ void __launch_bounds__(kBwdThreads) zero_fill_kernel(float* __restrict__ dst, size_t n, int aligned) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (aligned) {
    float4* d4 = reinterpret_cast<float4*>(dst);
    const size_t n4 = n >> 2;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    size_t i = tid;
    for (; i + 3 * stride < n4; i += 4 * stride) {
      __stcs(d4 + i, z);
      __stcs(d4 + i + stride, z);
      __stcs(d4 + i + 2 * stride, z);
      __stcs(d4 + i + 3 * stride, z);
    }
    for (; i < n4; i += stride) __stcs(d4 + i, z);
    for (size_t k = (n4 << 2) + tid; k < n; k += stride) dst[k] = 0.f;
  } else {
    for (size_t k = tid; k < n; k += stride) dst[k] = 0.f;
  }
}

// (2) one warp per 32 consecutive rows: grad_loc for every row (coalesced float4), then
// (softmax - onehot) * grad / N for the selected rows of the group.
__global__ void __launch_bounds__(kBwdThreads) loss_bwd_kernel(BwdArgs a) {
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)a.B * a.P;
  const long long warp_global = ((long long)blockIdx.x * kBwdThreads + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * kBwdThreads) >> 5;
  const double n = a.sums[2];
  const float scale_l = n > 0.0 ? (float)((double)a.grad_out[0] / n) : 0.0f;
  const float scale_c = n > 0.0 ? (float)((double)a.grad_out[1] / n) : 0.0f;
  const int C = a.C;
  for (long long grp = warp_global; grp * 32 < rows; grp += nwarps) {
    const long long row = grp * 32 + lane;
    int lb = -1;
    if (row < rows) {
      lb = a.sel[row];
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lb > 0) {
        int b = (int)(row / a.P);
        int p = (int)(row - (long long)b * a.P);
        const float* tr = a.gt + (size_t)(a.gt_offsets[b] + a.tidx[row]) * 5;
        Box m;
        m.x1 = tr[0]; m.y1 = tr[1]; m.x2 = tr[2]; m.y2 = tr[3];
        float4 t = encode_box(m, refine_center(a.rf, *reinterpret_cast<const float4*>(a.priors + (size_t)b * (size_t)a.prior_stride + (size_t)p * 4), (size_t)row),
                              a.var0, a.var1);
        float4 l = *reinterpret_cast<const float4*>(a.loc + row * 4);
        g.x = scale_l * fminf(fmaxf(l.x - t.x, -1.f), 1.f);     // smooth-L1': d for |d|<1, sign(d) otherwise
        g.y = scale_l * fminf(fmaxf(l.y - t.y, -1.f), 1.f);
        g.z = scale_l * fminf(fmaxf(l.z - t.z, -1.f), 1.f);
        g.w = scale_l * fminf(fmaxf(l.w - t.w, -1.f), 1.f);
      }
      *reinterpret_cast<float4*>(a.grad_loc + row * 4) = g;
    }
    uint32_t selmask = __ballot_sync(SSDBOX_FULL_MASK, lb >= 0);
    while (selmask) {
      int src = __ffs(selmask) - 1;
      selmask &= selmask - 1;
      int tl = __shfl_sync(SSDBOX_FULL_MASK, lb, src);
      const float* x = a.conf + (grp * 32 + src) * C;
      float* g = a.grad_conf + (grp * 32 + src) * C;
      float m = -INFINITY;
      for (int c = lane; c < C; c += 32) m = fmaxf(m, x[c]);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(SSDBOX_FULL_MASK, m, d));
      float sum = 0.f;
      for (int c = lane; c < C; c += 32) sum += expf(x[c] - m);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(SSDBOX_FULL_MASK, sum, d);
      float inv = 1.0f / sum;
      for (int c = lane; c < C; c += 32) g[c] = scale_c * (expf(x[c] - m) * inv - (c == tl ? 1.0f : 0.0f));
    }
  }
}

// ---- backward as ONE write-bound streaming kernel (aligned grad_conf, C <= 128) -------------------
// Persistent CTAs; every warp owns one shared-memory tile of kBwdTileRows x C floats (41 KB at C = 81)
// that is all zeros
// except while it carries the gradient rows of the selected priors of the tile it is working on:
//   1. `sel` of the tile (4 rows per lane; requested two tiles ahead), ballot of the selected rows
//   2. for each selected row: the logits row -> (softmax - onehot) * grad / N written into the tile
//   3. fence.proxy.async, ONE TMA bulk store of the whole tile (zeros + rows) to grad_conf
//   4. grad_loc rows of the tile with plain 16-byte stores (zeros, or smooth-L1' for positives)
//   5. when the bulk store has read the tile, the patched rows are zeroed again
// The zero fill therefore costs no store instructions, the few reads (3 MB of sel, 20 MB of logits rows)
// run tiles ahead of the write stream in the other warps, and nothing is written twice.
constexpr int kBwdRPL = 1;                          // rows of a tile per lane (1, 2 or 4)
constexpr int kBwdTileRows = 32 * kBwdRPL;          // 32 rows = 10 KB bulk stores at C = 81
constexpr int kBwdStreamWarps = 20;                 // measured: 5 warps x 128 rows 217 us, 10 x 64 134 us
#ifndef SSDBOX_BWD_SPLIT
#define SSDBOX_BWD_SPLIT 1
#endif
constexpr int kBwdSplit = SSDBOX_BWD_SPLIT;         // bulk stores per tile (1, 2 or 4; every part a multiple of 16 bytes)
static_assert(kBwdTileRows % kBwdSplit == 0 && (kBwdTileRows / kBwdSplit) % 4 == 0, "parts of whole rows, 16-byte multiples");

// per-tile read state of a warp: the class targets of its rows (requested two tiles ahead) and the
// logits of the first four selected rows (requested one tile ahead)
struct BwdPre {
  int lb[kBwdRPL];
  int rl[4], tl[4];
  float xv[4][4];
  uint32_t rest[kBwdRPL];      // selected rows beyond the first four, per row slot (bit = lane)
};

template <int CT>
__global__ void __launch_bounds__(kBwdStreamWarps * 32, 1) loss_bwd_stream_kernel(BwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem_bwd[];
  constexpr int RPL = kBwdRPL;
  const int C = CT > 0 ? CT : a.C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* tile = reinterpret_cast<float*>(smem_bwd) + (size_t)warp * kBwdTileRows * C;
  const long long rows = (long long)a.B * a.P;
  const long long tiles = (rows + kBwdTileRows - 1) / kBwdTileRows;
  const long long per_cta = (tiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = (long long)blockIdx.x * per_cta;
  const long long t_end = t_begin + per_cta < tiles ? t_begin + per_cta : tiles;
  const double n = a.sums[2];
  const float scale_l = n > 0.0 ? (float)((double)a.grad_out[0] / n) : 0.0f;
  const float scale_c = n > 0.0 ? (float)((double)a.grad_out[1] / n) : 0.0f;
  // Phase 0, before this CTA writes anything: pull everything the loop below will read -- the selection flags of the
  // CTA's rows, the logits rows they select, loc / tidx of the positives -- into L2 with evict_last priority (the
  // bulk stores below go out evict_first, so the 509 MB they write do not push it out again).
  // Measured (tools/micro/fill_patterns.cu, profiles/r04_micro_fill_patterns.txt): a small DRAM read that arrives
  // alone inside a saturated write stream costs the stream far more than its bytes (write->read->write
  // turn-arounds; 49 k flag reads: +6 us, one logits row per tile: +8-14 us, prefetched or not); as L2 hits the
  // same loads cost +2 us.  The rows are fetched with 4-byte cp.async into the (not yet used) tile: no registers,
  // every sector touched (prefetch.global.L2 turned out to be dropped: ncu showed the same misses with and without).
  const long long rbeg = t_begin * kBwdTileRows;
  const long long rend = t_end * kBwdTileRows < rows ? t_end * kBwdTileRows : rows;
#ifdef SSDBOX_EXPERIMENTS
  if (!(a.ablate & 16))
#endif
  {
    const uint64_t keep = l2_evict_last_policy();
    int slot = 0;
    constexpr int kGroupsAtOnce = 6;       // flag loads of a warp requested together (one round trip for the usual range)
    for (long long rb = rbeg + (long long)warp * 128; rb < rend; rb += (long long)kBwdStreamWarps * 128 * kGroupsAtOnce) {
      unsigned long long fl[kGroupsAtOnce];
#pragma unroll
      for (int g = 0; g < kGroupsAtOnce; ++g) {
        const long long rl = rb + (long long)g * kBwdStreamWarps * 128 + lane * 4;
        fl[g] = ~0ull;                       // four times -1
        if (rl + 3 < rend) {                 // rbeg is a multiple of 32 and sel is 8-byte aligned
          fl[g] = ld_u64_l2_keep(a.sel + rl, keep);
        } else {
          for (int k = 0; k < 4 && rl + k < rend; ++k)
            fl[g] = (fl[g] & ~(0xffffull << (16 * k))) | ((unsigned long long)(uint16_t)a.sel[rl + k] << (16 * k));
        }
      }
#ifdef SSDBOX_EXPERIMENTS
      if (a.ablate & 64) continue;
#endif
#pragma unroll
      for (int g = 0; g < kGroupsAtOnce; ++g) {
        const long long r = rb + (long long)g * kBwdStreamWarps * 128;
        if (r >= rend) break;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int lbk = (int)(int16_t)((fl[g] >> (16 * k)) & 0xffffull);
          uint32_t m = __ballot_sync(SSDBOX_FULL_MASK, lbk >= 0);
          const uint32_t pos = __ballot_sync(SSDBOX_FULL_MASK, lbk > 0);
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const long long row = r + src * 4 + k;
            const float* x = a.conf + row * C;
            float* dump = tile + (size_t)(slot & (kBwdTileRows - 1)) * C;
            for (int c = lane; c < C; c += 32) cp_async_4_hint(dump + c, x + c, keep);
            if (((pos >> src) & 1u) && lane == 0) {
              cp_async_4_hint(dump, a.loc + row * 4, keep);
              cp_async_4_hint(dump + 1, reinterpret_cast<const char*>(a.tidx) + ((row * 2) & ~3ll), keep);
              if (a.rf.arm_loc) cp_async_4_hint(dump + 2, a.rf.arm_loc + row * 4, keep);
            }
            ++slot;
          }
        }
      }
    }
    cp_async_wait_all();
    __syncwarp();
  }
  for (int i = lane; i < kBwdTileRows * C; i += 32) tile[i] = 0.f;
  __syncwarp();

  // sel of a tile (RPL consecutive rows per lane; row r of the tile = lane * RPL + k)
  auto load_sel = [&](long long t, int (&lb)[RPL]) {
#pragma unroll
    for (int k = 0; k < RPL; ++k) lb[k] = -1;
    if (t < t_end && !SSDBOX_ABLATE(8)) {
      const long long r0 = t * kBwdTileRows + lane * RPL;
      if (RPL == 4 && r0 + 3 < rows) {
        const short4 v = *reinterpret_cast<const short4*>(a.sel + r0);     // 8-byte aligned array, r0 % 4 == 0
        lb[0] = v.x; lb[1 % RPL] = v.y; lb[2 % RPL] = v.z; lb[3 % RPL] = v.w;
      } else if (RPL == 2 && r0 + 1 < rows) {
        const short2 v = *reinterpret_cast<const short2*>(a.sel + r0);
        lb[0] = v.x; lb[1 % RPL] = v.y;
      } else {
#pragma unroll
        for (int k = 0; k < RPL; ++k)
          if (r0 + k < rows) lb[k] = a.sel[r0 + k];
      }
    }
  };
  // the logits rows of the first four selected rows of the tile, requested together
  auto load_rows = [&](long long t, BwdPre& p) {
    const long long row0 = t * kBwdTileRows;
    int nfound = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) p.rl[j] = -1;
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      uint32_t m = __ballot_sync(SSDBOX_FULL_MASK, p.lb[k] >= 0);
      if (SSDBOX_ABLATE(2)) m = 0u;
      while (m && nfound < 4) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int tl = __shfl_sync(SSDBOX_FULL_MASK, p.lb[k], src);
        const int rl = src * RPL + k;
        const float* x = a.conf + (row0 + rl) * C;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j == nfound) {
            p.rl[j] = rl;
            p.tl[j] = tl;
#pragma unroll
            for (int u = 0; u < 4; ++u) p.xv[j][u] = lane + 32 * u < C ? x[lane + 32 * u] : -INFINITY;
          }
        ++nfound;
      }
      p.rest[k] = m;
    }
  };
  auto put_row = [&](int rl, int tl, const float (&xv)[4]) {
    float mx = fmaxf(fmaxf(xv[0], xv[1]), fmaxf(xv[2], xv[3]));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(SSDBOX_FULL_MASK, mx, d));
    float e[4], sum = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      e[u] = lane + 32 * u < C ? expf(xv[u] - mx) : 0.f;
      sum += e[u];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(SSDBOX_FULL_MASK, sum, d);
    const float inv = 1.0f / sum;
    float* g = tile + (size_t)rl * C;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = lane + 32 * u;
      if (c < C) g[c] = scale_c * (e[u] * inv - (c == tl ? 1.0f : 0.0f));
    }
  };

  const uint64_t wr_policy = l2_evict_first_policy();     // the 509 MB written here must not evict what phase 0 pulled in
  BwdPre cur, nxt;
  long long t = t_begin + warp;
  load_sel(t, cur.lb);
  load_sel(t + kBwdStreamWarps, nxt.lb);
  load_rows(t, cur);
  uint32_t patched[RPL];                       // rows of my tile that currently hold non-zero data (bit = lane, per slot)
#pragma unroll
  for (int k = 0; k < RPL; ++k) patched[k] = 0u;
  for (; t < t_end; t += kBwdStreamWarps) {
    const long long row0 = t * kBwdTileRows;
    const int nrows = (int)(rows - row0 < kBwdTileRows ? rows - row0 : kBwdTileRows);
    // read pipeline first (ahead of this iteration's stores): the selected logits rows of the NEXT tile
    // of this warp (its sel arrived an iteration ago), sel of the one after
    load_rows(t + kBwdStreamWarps, nxt);
    int lb_nn[RPL];
    load_sel(t + 2 * kBwdStreamWarps, lb_nn);
    // The tile goes out as kBwdSplit bulk stores of kBwdTileRows / kBwdSplit rows, one commit group each: while one
    // part is being patched the other parts are in flight (a store under load takes ~5 us to drain; what is in
    // flight per SM, not the issue rate, bounds the write stream).
    uint32_t selmask = 0u;
    if (RPL == 1) selmask = __ballot_sync(SSDBOX_FULL_MASK, cur.lb[0] >= 0);
#ifdef SSDBOX_EXPERIMENTS
    if (a.ablate & 2) selmask = 0u;
#endif
#pragma unroll
    for (int h = 0; h < kBwdSplit; ++h) {
      constexpr int RH = kBwdTileRows / kBwdSplit;                   // rows per part
      const uint32_t hm = RH * RPL >= 32 ? 0xffffffffu : (((1u << (RH / RPL)) - 1u) << (h * (RH / RPL)));   // its lanes
      // the store that used this part last (kBwdSplit commits ago) has read it: clear what it carried
      bulk_wait_read<kBwdSplit - 1>();
      __syncwarp();          // (lane 0 owns the bulk groups: nobody touches the part before its wait returns)
#pragma unroll
      for (int k = 0; k < RPL; ++k) {
        uint32_t m = patched[k] & hm;
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          float* g = tile + (size_t)(src * RPL + k) * C;
          for (int c = lane; c < C; c += 32) g[c] = 0.f;
        }
        const uint32_t now = RPL == 1 ? selmask : __ballot_sync(SSDBOX_FULL_MASK, cur.lb[k] >= 0);
        patched[k] = (patched[k] & ~hm) | (now & hm);
      }
      // gradient rows of the selected priors: (softmax - onehot) * grad / N
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (cur.rl[j] >= 0 && ((hm >> (cur.rl[j] / RPL)) & 1u)) put_row(cur.rl[j], cur.tl[j], cur.xv[j]);
#pragma unroll
      for (int k = 0; k < RPL; ++k) {
        uint32_t m = cur.rest[k] & hm;             // more than four selected rows in the tile
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const int tl = __shfl_sync(SSDBOX_FULL_MASK, cur.lb[k], src);
          const int rl = src * RPL + k;
          const float* x = a.conf + (row0 + rl) * C;
          float xv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) xv[u] = lane + 32 * u < C ? x[lane + 32 * u] : -INFINITY;
          put_row(rl, tl, xv);
        }
      }
      __syncwarp();
      const int nr = nrows - h * RH < RH ? nrows - h * RH : RH;        // rows of this part that exist
      const uint32_t bytes = nr > 0 ? (uint32_t)nr * (uint32_t)C * 4u : 0u;
      float* dst = a.grad_conf + (row0 + h * RH) * C;
      const float* part = tile + (size_t)h * RH * C;
      if (bytes != 0u && (bytes & 15u) == 0u) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
#ifdef SSDBOX_EXPERIMENTS
          if (a.ablate & 32) bulk_s2g(dst, part, bytes); else
#endif
          bulk_s2g_hint(dst, part, bytes, wr_policy);
        }
      } else {                                   // ragged last tile: plain stores (and an empty group: one commit per part)
        for (int i = lane; i < nr * C; i += 32) dst[i] = part[i];
        if (lane == 0) bulk_commit();
      }
    }
    // grad_loc of the tile: zeros, or smooth-L1' for positives
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      const long long row = row0 + lane * RPL + k;
      if (row >= rows || SSDBOX_ABLATE(1)) continue;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cur.lb[k] > 0 && !SSDBOX_ABLATE(4)) {
        const int b = (int)((uint32_t)row / (uint32_t)a.P);
        const int pi = (int)((uint32_t)row - (uint32_t)b * (uint32_t)a.P);
        const float4 l = *reinterpret_cast<const float4*>(a.loc + row * 4);
        const float4 pr = refine_center(a.rf, *reinterpret_cast<const float4*>(a.priors + (size_t)b * (size_t)a.prior_stride + (size_t)pi * 4), (size_t)row);
        const float* tr = a.gt + (size_t)(a.gt_offsets[b] + a.tidx[row]) * 5;
        Box mbox;
        mbox.x1 = tr[0]; mbox.y1 = tr[1]; mbox.x2 = tr[2]; mbox.y2 = tr[3];
        float4 tt = encode_box(mbox, pr, a.var0, a.var1);
        g.x = scale_l * fminf(fmaxf(l.x - tt.x, -1.f), 1.f);     // smooth-L1': d for |d|<1, sign(d) otherwise
        g.y = scale_l * fminf(fmaxf(l.y - tt.y, -1.f), 1.f);
        g.z = scale_l * fminf(fmaxf(l.z - tt.z, -1.f), 1.f);
        g.w = scale_l * fminf(fmaxf(l.w - tt.w, -1.f), 1.f);
      }
      __stcs(reinterpret_cast<float4*>(a.grad_loc + row * 4), g);
    }
    cur = nxt;
#pragma unroll
    for (int k = 0; k < RPL; ++k) nxt.lb[k] = lb_nn[k];
  }
  bulk_wait_all();          // the tile must outlive its last bulk store
}


}  // namespace ssdbox

using namespace ssdbox;

static int check_loss_cfg(const ssdbox_loss_cfg* c) {
  SSDBOX_REQUIRE(c, SSDBOX_EINVAL, "loss: null cfg");
  SSDBOX_REQUIRE(c->B >= 0 && c->P >= 0 && c->C >= 1 && c->gmax >= 0 && c->negpos_ratio >= 0, SSDBOX_EINVAL,
                 "loss: negative size");
  SSDBOX_REQUIRE(c->gmax <= kGmaxLimit, SSDBOX_ESHAPE, "loss: gmax %d > %d", c->gmax, kGmaxLimit);
  SSDBOX_REQUIRE(c->C <= kClassLimit, SSDBOX_ESHAPE, "loss: %d classes > %d", c->C, kClassLimit);
  SSDBOX_REQUIRE(c->B <= 65535, SSDBOX_ESHAPE, "loss: batch %d > 65535", c->B);
  SSDBOX_REQUIRE((long long)c->B * c->P < (1ll << 31), SSDBOX_ESHAPE, "loss: B*P must be < 2^31");
  SSDBOX_REQUIRE(c->prior_batch_stride == 0 || c->prior_batch_stride == (int64_t)c->P * 4, SSDBOX_EINVAL,
                 "loss: prior_batch_stride must be 0 or 4*P");
  return SSDBOX_OK;
}

extern "C" size_t ssdbox_peer_buffer_bytes(void) {
  return (size_t)kPeerHeaderBytes + 2 * (size_t)SSDBOX_MAX_PEERS * kPeerSlotBytes;
}

extern "C" int ssdbox_multibox_loss_fwd(const ssdbox_loss_cfg* cfg, const float* loc, const float* conf,
                                        const float* priors, const float* anchors_xyxy, const uint8_t* pool,
                                        const float* gt, const int32_t* gt_offsets, double* sums, float* losses,
                                        int16_t* sel, int16_t* tidx, int64_t* dbg_conf_t, float* dbg_loc_t,
                                        uint8_t* dbg_neg, float* dbg_keys, void* ws, size_t ws_bytes,
                                        ssdbox_stream_t stream) {
  return ssdbox_multibox_loss_fwd_peers(cfg, loc, conf, priors, anchors_xyxy, pool, gt, gt_offsets, sums, losses, sel,
                                        tidx, dbg_conf_t, dbg_loc_t, dbg_neg, dbg_keys, nullptr, ws, ws_bytes, stream);
}

static int loss_fwd_impl(const ssdbox_loss_cfg* cfg, const float* loc, const float* conf, const float* priors,
                         const float* anchors_xyxy, const uint8_t* pool, const ssdbox_refine* refine, const float* gt,
                         const int32_t* gt_offsets, double* sums, float* losses, int16_t* sel, int16_t* tidx,
                         int64_t* dbg_conf_t, float* dbg_loc_t, uint8_t* dbg_neg, float* dbg_keys,
                         const ssdbox_peer_group* peers, void* ws, size_t ws_bytes, ssdbox_stream_t stream);

extern "C" int ssdbox_multibox_loss_fwd_peers(const ssdbox_loss_cfg* cfg, const float* loc, const float* conf,
                                              const float* priors, const float* anchors_xyxy, const uint8_t* pool,
                                              const float* gt, const int32_t* gt_offsets, double* sums, float* losses,
                                              int16_t* sel, int16_t* tidx, int64_t* dbg_conf_t, float* dbg_loc_t,
                                              uint8_t* dbg_neg, float* dbg_keys, const ssdbox_peer_group* peers,
                                              void* ws, size_t ws_bytes, ssdbox_stream_t stream) {
  return loss_fwd_impl(cfg, loc, conf, priors, anchors_xyxy, pool, nullptr, gt, gt_offsets, sums, losses, sel, tidx, dbg_conf_t,
                       dbg_loc_t, dbg_neg, dbg_keys, peers, ws, ws_bytes, stream);
}

extern "C" int ssdbox_multibox_loss_fwd_refine(const ssdbox_loss_cfg* cfg, const float* loc, const float* conf,
                                               const float* priors, const ssdbox_refine* refine, const float* gt,
                                               const int32_t* gt_offsets, double* sums, float* losses, int16_t* sel,
                                               int16_t* tidx, int64_t* dbg_conf_t, float* dbg_loc_t, uint8_t* dbg_neg,
                                               float* dbg_keys, const ssdbox_peer_group* peers, void* ws, size_t ws_bytes,
                                               ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(refine, SSDBOX_EINVAL, "loss: null refine descriptor");
  return loss_fwd_impl(cfg, loc, conf, priors, nullptr, nullptr, refine, gt, gt_offsets, sums, losses, sel, tidx, dbg_conf_t,
                       dbg_loc_t, dbg_neg, dbg_keys, peers, ws, ws_bytes, stream);
}

// fills the kernels' RefineArgs from the ABI descriptor (validated); nullptr -> all zero
static int make_refine(const ssdbox_refine* refine, float var0, float var1, long long prior_batch_stride, int nonempty, RefineArgs* out) {
  RefineArgs r{};
  if (refine) {
    SSDBOX_REQUIRE(prior_batch_stride == 0, SSDBOX_EINVAL, "refine: priors must be the shared [P,4] tensor (prior_batch_stride 0)");
    SSDBOX_REQUIRE(!nonempty || refine->arm_loc, SSDBOX_EINVAL, "refine: null arm_loc");
    SSDBOX_REQUIRE(aligned16(refine->arm_loc) && aligned16(refine->arm_conf), SSDBOX_EALIGN, "refine: arm_loc / arm_conf must be 16-byte aligned");
    r.arm_loc = refine->arm_loc;
    r.arm_conf = refine->arm_conf;
    r.theta = refine->theta;
    r.var0 = var0;
    r.var1 = var1;
  }
  *out = r;
  return SSDBOX_OK;
}

static int loss_fwd_impl(const ssdbox_loss_cfg* cfg, const float* loc, const float* conf, const float* priors,
                         const float* anchors_xyxy, const uint8_t* pool, const ssdbox_refine* refine, const float* gt,
                         const int32_t* gt_offsets, double* sums, float* losses, int16_t* sel, int16_t* tidx,
                         int64_t* dbg_conf_t, float* dbg_loc_t, uint8_t* dbg_neg, float* dbg_keys,
                         const ssdbox_peer_group* peers, void* ws, size_t ws_bytes, ssdbox_stream_t stream) {
  int rc = check_loss_cfg(cfg);
  if (rc) return rc;
  RefineArgs rf;
  rc = make_refine(refine, cfg->var0, cfg->var1, cfg->prior_batch_stride, cfg->B > 0 && cfg->P > 0, &rf);
  if (rc) return rc;
  if (peers) {
    SSDBOX_REQUIRE(peers->world >= 1 && peers->world <= SSDBOX_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world,
                   SSDBOX_EINVAL, "loss: bad peer group (rank %d of %d)", peers->rank, peers->world);
    for (int r = 0; r < peers->world; ++r)
      SSDBOX_REQUIRE(peers->bufs[r] && (reinterpret_cast<uintptr_t>(peers->bufs[r]) & 15u) == 0, SSDBOX_EINVAL,
                     "loss: peer buffer %d is null or misaligned", r);
  }
  const int B = cfg->B, P = cfg->P, C = cfg->C;
  SSDBOX_REQUIRE(sums && ws && gt_offsets, SSDBOX_EINVAL, "loss: null pointer");
  SSDBOX_REQUIRE(!cfg->finalize || losses, SSDBOX_EINVAL, "loss: finalize needs `losses`");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((B == 0 || P == 0) && peers) {
    // empty local shard (global batch smaller than the world, last batch of an epoch): zeros are posted
    PeerFinishArgs pa{};
    pa.rank = peers->rank;
    pa.world = peers->world;
    pa.timeout_ns = peer_timeout_ns(peers);
    for (int r = 0; r < peers->world; ++r) pa.bufs[r] = peers->bufs[r];
    pa.sums = sums;
    pa.losses = losses;
    peer_empty_kernel<<<1, 32, 0, st>>>(pa, (cfg->flags & SSDBOX_LOSS_DEFER_PEER_WAIT) ? 1 : 0, cfg->finalize);
    SSDBOX_LAUNCH_OK("peer_empty_kernel");
    return SSDBOX_OK;
  }
  if (B == 0 || P == 0) {
    SSDBOX_CUDA(cudaMemsetAsync(sums, 0, 3 * sizeof(double), st));
    if (losses) SSDBOX_CUDA(cudaMemsetAsync(losses, 0, 2 * sizeof(float), st));
    return SSDBOX_OK;
  }
  SSDBOX_REQUIRE(loc && conf && priors && sel && tidx && (gt || cfg->gmax == 0), SSDBOX_EINVAL, "loss: null pointer");
  SSDBOX_REQUIRE(aligned16(loc) && aligned16(priors) && (!anchors_xyxy || aligned16(anchors_xyxy)) &&
                     (!dbg_loc_t || aligned16(dbg_loc_t)),
                 SSDBOX_EALIGN, "loss: box pointers must be 16-byte aligned");
  SSDBOX_REQUIRE((reinterpret_cast<uintptr_t>(conf) & 3u) == 0, SSDBOX_EALIGN, "loss: conf must be 4-byte aligned");
  SSDBOX_REQUIRE((reinterpret_cast<uintptr_t>(tidx) & 7u) == 0 && (reinterpret_cast<uintptr_t>(sel) & 7u) == 0,
                 SSDBOX_EALIGN, "loss: sel / tidx must be 8-byte aligned");
  SSDBOX_REQUIRE(ws_bytes >= loss_ws_bytes(B, P, C, cfg->gmax), SSDBOX_EWORKSPACE, "loss: workspace too small");
  DevInfo dev;
  rc = get_dev_info(&dev);
  if (rc) return rc;

  Carver c(ws);
  LossWs w;
  carve_match_core(c, B, cfg->gmax, &w.m);
  w.m.lab = c.take<int16_t>((size_t)B * P);
  w.m.tidx = tidx;
  w.keys = c.take<float>((size_t)B * P);
  w.ukey = c.take<uint32_t>((size_t)B * P);
  w.hist = c.take<uint32_t>((size_t)B * kHistBins);
  w.partial = c.take<double>((size_t)B * 6);     // {l1, ce, npos} per mining CTA (two per image when clustered)
  w.ticket = c.take<uint32_t>(2);      // [0] mine_reduce ticket, [1] next matching unit

  // State = per-truth best-prior keys, image tickets, mining histograms, work tickets.  Every call hands it back
  // initialised, so a caller that reuses the workspace for the same shape may skip this launch.
  if (!(cfg->flags & SSDBOX_LOSS_WS_CLEAN)) {
    rc = launch_init(w.m.gt_best, (size_t)(B + 1) * gt_pad(cfg->gmax), w.m.done, (size_t)B + 1, w.hist,
                     (size_t)B * kHistBins, w.ticket, 2, st);
    if (rc) return rc;
  }

  // Matching runs on dedicated warps of the streaming kernel unless the caller asked for the
  // separate kernel or the truths of a CTA's images do not fit in shared memory beside the ring.
  MatchArgs ma{gt, gt_offsets, cfg->gmax, priors, (long long)cfg->prior_batch_stride, anchors_xyxy, B, P,
               cfg->threshold, cfg->binarize_labels, rf};
  StreamArgs sa{};
  sa.rf = rf;
  sa.pool = pool;
  sa.key0 = w.keys;
  sa.hist = w.hist;
  sa.P = P;
  sa.B = B;
  sa.gt = gt;
  sa.gt_offsets = gt_offsets;
  sa.gmax = cfg->gmax;
  sa.gpad = gt_pad(cfg->gmax);
  sa.priors = priors;
  sa.prior_stride = (long long)cfg->prior_batch_stride;
  sa.anchors_xyxy = anchors_xyxy;
  sa.threshold = cfg->threshold;
  sa.binarize = cfg->binarize_labels;
  sa.gt_best = w.m.gt_best;
  sa.lab_out = w.m.lab;
  sa.tidx_out = tidx;
  sa.unit_counter = w.ticket + 1;
  sa.dbg = (cfg->flags >> 8) & 0xff;
  sa.shift = (cfg->flags & SSDBOX_LOSS_LSE_SHIFT) ? sums : nullptr;     // read by the streaming kernel; sums is written by the mining kernel after it
  size_t stream_smem = 0;
  rc = plan_stream(&sa, conf, (long long)B * P, C, dev.sm_count, dev.max_smem_optin,
                   !(cfg->flags & SSDBOX_LOSS_SEPARATE_MATCH), &stream_smem);
  if (rc) return rc;
#ifdef SSDBOX_PHASE_TIMING
  if (sa.dbg & 1) sa.ring.tiles = 0, sa.ring.tiles_per_cta = 0;     // matching alone
#endif
  if (!sa.fuse) {
    rc = launch_match(ma, w.m, w.m.lab, tidx, nullptr, st);
    if (rc) return rc;
  }
  rc = launch_stream(sa, stream_smem, st);
  if (rc) return rc;

  MineArgs m{};
  m.B = B; m.P = P; m.C = C;
  m.negpos_ratio = cfg->negpos_ratio;
  m.var0 = cfg->var0; m.var1 = cfg->var1;
  m.finalize = cfg->finalize;
  m.prior_stride = (long long)cfg->prior_batch_stride;
  m.loc = loc; m.priors = priors; m.gt = gt; m.gt_offsets = gt_offsets; m.pool = pool; m.rf = rf;
  m.keys = w.keys; m.conf = conf; m.lab = w.m.lab; m.tidx = tidx; m.hist = w.hist;
  m.fuse = sa.fuse; m.gmax = cfg->gmax; m.gpad = sa.gpad; m.binarize = cfg->binarize_labels;
  m.gt_best = w.m.gt_best; m.gt_best_w = w.m.gt_best; m.lab_w = w.m.lab; m.tidx_w = tidx;
  m.ukey_global = w.ukey;
  size_t fixed = kMineFixedSmem;
  m.uk_in_smem = (fixed + (size_t)P * 6 + 16 <= (size_t)dev.max_smem_optin - 1024) ? 1 : 0;
  m.partial = w.partial; m.ticket = w.ticket; m.sums = sums; m.losses = losses; m.sel = sel;
  m.dbg_neg = dbg_neg; m.dbg_keys = dbg_keys;
  m.peer_world = 0;
  m.peer_defer = (cfg->flags & SSDBOX_LOSS_DEFER_PEER_WAIT) ? 1 : 0;
  if (peers) {
    m.peer_rank = peers->rank;
    m.peer_world = peers->world;
    m.peer_timeout_ns = peer_timeout_ns(peers);
    for (int r = 0; r < peers->world; ++r) m.peer_bufs[r] = peers->bufs[r];
  }
  size_t smem = fixed + (m.uk_in_smem ? (size_t)P * 6 + 16 : 0);
  // vector mode needs every per-image row of keys / lab / pool / sel / debug arrays 16-byte friendly
  const bool vec4 = (P % 4 == 0) && aligned16(sel) && (!pool || (reinterpret_cast<uintptr_t>(pool) & 3u) == 0) &&
                    (!dbg_neg || (reinterpret_cast<uintptr_t>(dbg_neg) & 3u) == 0) && (!dbg_keys || aligned16(dbg_keys));
  void (*mkern)(MineArgs) = vec4 ? mine_reduce_kernel<4> : mine_reduce_kernel<1>;
  bool launched = false;
  int mine_threads = kMineThreads;
  if (vec4 && P <= 4 * kMineQ * kMineThreads && !(cfg->flags & SSDBOX_LOSS_GENERIC_MINE)) {
    // keys and class targets stay in registers; images with many priors are split over a cluster of
    // two CTAs (two SMs pull the image's keys, histograms merged through distributed shared memory)
    // (only while the 2*B CTAs still fit in one wave: one 1024-thread CTA per SM)
    const bool pair = P >= 8192 && 2 * B <= dev.sm_count && !(cfg->flags & SSDBOX_LOSS_NO_CLUSTER);
    // more images than SMs (RFB300-VOC B = 256): 512-thread CTAs, two per SM, so that the batch mines in one wave
    const bool half = !pair && P <= 4 * kMineQ * (kMineThreads / 2) &&
                      (B > dev.sm_count || (cfg->flags & SSDBOX_LOSS_MINE_HALF_CTA));
    mine_threads = half ? kMineThreads / 2 : kMineThreads;
    mkern = pair ? mine_reduce_reg_kernel<2> : (half ? mine_reduce_reg_kernel<1, kMineThreads / 2> : mine_reduce_reg_kernel<1>);
    smem = kMineFixedSmem + (size_t)kForceListMax * 16 + (size_t)(5 * kForceListMax + 8) * 4 + (size_t)P * 4 + 16 +
           (size_t)P * 2 + 16;      // ... + positives list + forced-label override array
    if (pair) {
      SSDBOX_CUDA(cudaFuncSetAttribute(mkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      SSDBOX_CARVE(mkern);
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(2 * B);
      lc.blockDim = dim3(kMineThreads);
      lc.dynamicSmemBytes = smem;
      lc.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      lc.attrs = at;
      lc.numAttrs = 1;
      {
        TimerScope ts__(KID_MINE, st);
        SSDBOX_CUDA(cudaLaunchKernelEx(&lc, mkern, m));
      }
      launched = true;
    }
  }
  if (!launched) {
    SSDBOX_CUDA(cudaFuncSetAttribute(mkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SSDBOX_CARVE(mkern);
    TimerScope ts__(KID_MINE, st);
    mkern<<<B, mine_threads, smem, st>>>(m);
  }
  SSDBOX_LAUNCH_OK("mine_reduce_kernel");

  if (dbg_conf_t || dbg_loc_t) {
    // encode against the centre-form priors (for RefineDet: the refined anchors' centre form)
    rc = launch_materialize(ma, cfg->var0, cfg->var1, w.m.lab, tidx, dbg_loc_t, dbg_conf_t, nullptr, st);
    if (rc) return rc;
  }
  return SSDBOX_OK;
}

#ifdef SSDBOX_PHASE_TIMING
extern "C" __attribute__((visibility("default"))) int ssdbox_debug_mstat(long long* out) {
  return cudaMemcpyFromSymbol(out, ssdbox::g_mstat, sizeof(long long) * 8 * 160) == cudaSuccess ? 0 : -5;
}
extern "C" __attribute__((visibility("default"))) int ssdbox_debug_sgt(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, ssdbox::g_sgt, sizeof(unsigned long long) * 2 * 160) == cudaSuccess ? 0 : -5;
}
extern "C" __attribute__((visibility("default"))) int ssdbox_debug_sphases(long long* out16) {
  return cudaMemcpyFromSymbol(out16, ssdbox::g_sphase, sizeof(long long) * 8 * 160) == cudaSuccess ? 0 : -5;
}
extern "C" __attribute__((visibility("default"))) int ssdbox_debug_phases(long long* out16) {
  return cudaMemcpyFromSymbol(out16, ssdbox::g_phase, sizeof(long long) * 16 * 64) == cudaSuccess ? 0 : -5;
}
#endif

extern "C" int ssdbox_multibox_loss_peer_finish(const ssdbox_peer_group* peers, double* sums, float* losses,
                                               ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(peers && sums, SSDBOX_EINVAL, "peer_finish: null pointer");
  SSDBOX_REQUIRE(peers->world >= 1 && peers->world <= SSDBOX_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world,
                 SSDBOX_EINVAL, "peer_finish: bad peer group");
  PeerFinishArgs a{};
  a.rank = peers->rank;
  a.world = peers->world;
  a.timeout_ns = peer_timeout_ns(peers);
  for (int r = 0; r < peers->world; ++r) a.bufs[r] = peers->bufs[r];
  a.sums = sums;
  a.losses = losses;
  peer_finish_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
  SSDBOX_LAUNCH_OK("peer_finish_kernel");
  return SSDBOX_OK;
}

extern "C" int ssdbox_multibox_loss_finalize(const double* sums, float* losses, ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(sums && losses, SSDBOX_EINVAL, "finalize: null pointer");
  finalize_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(sums, losses);
  SSDBOX_LAUNCH_OK("finalize_kernel");
  return SSDBOX_OK;
}

extern "C" int ssdbox_multibox_loss_bwd(const ssdbox_loss_cfg* cfg, const float* loc, const float* conf,
                                        const float* priors, const float* gt, const int32_t* gt_offsets,
                                        const int16_t* sel, const int16_t* tidx, const double* sums,
                                        const float* grad_out, float* grad_loc, float* grad_conf,
                                        ssdbox_stream_t stream) {
  return ssdbox_multibox_loss_bwd_refine(cfg, loc, conf, priors, nullptr, gt, gt_offsets, sel, tidx, sums, grad_out, grad_loc,
                                         grad_conf, stream);
}

extern "C" int ssdbox_multibox_loss_bwd_refine(const ssdbox_loss_cfg* cfg, const float* loc, const float* conf,
                                               const float* priors, const ssdbox_refine* refine, const float* gt,
                                               const int32_t* gt_offsets, const int16_t* sel, const int16_t* tidx,
                                               const double* sums, const float* grad_out, float* grad_loc,
                                               float* grad_conf, ssdbox_stream_t stream) {
  int rc = check_loss_cfg(cfg);
  if (rc) return rc;
  RefineArgs rf;
  rc = make_refine(refine, cfg->var0, cfg->var1, cfg->prior_batch_stride, cfg->B > 0 && cfg->P > 0, &rf);
  if (rc) return rc;
  if (cfg->B == 0 || cfg->P == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(loc && conf && priors && gt_offsets && sel && tidx && sums && grad_out && grad_loc && grad_conf,
                 SSDBOX_EINVAL, "loss_bwd: null pointer");
  SSDBOX_REQUIRE(aligned16(loc) && aligned16(priors) && aligned16(grad_loc), SSDBOX_EALIGN,
                 "loss_bwd: box pointers must be 16-byte aligned");
  SSDBOX_REQUIRE((reinterpret_cast<uintptr_t>(sel) & 7u) == 0, SSDBOX_EALIGN, "loss_bwd: sel must be 8-byte aligned");
  DevInfo dev;
  rc = get_dev_info(&dev);
  if (rc) return rc;
  BwdArgs a{};
  a.B = cfg->B; a.P = cfg->P; a.C = cfg->C;
  a.var0 = cfg->var0; a.var1 = cfg->var1;
  a.prior_stride = (long long)cfg->prior_batch_stride;
  a.loc = loc; a.conf = conf; a.priors = priors; a.gt = gt; a.gt_offsets = gt_offsets;
  a.sel = sel; a.tidx = tidx; a.sums = sums; a.grad_out = grad_out;
  a.grad_loc = grad_loc; a.grad_conf = grad_conf;
  a.rf = rf;
  a.conf_aligned = aligned16(grad_conf) ? 1 : 0;
#ifdef SSDBOX_EXPERIMENTS
  if (const char* e = getenv("SSDBOX_BWD_ABLATE")) a.ablate = atoi(e);
#endif
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long rows = (long long)a.B * a.P;
  const size_t stream_smem = (size_t)kBwdStreamWarps * kBwdTileRows * a.C * 4;
  if (a.conf_aligned && a.C <= 128 && stream_smem <= (size_t)dev.max_smem_optin - 1024
#ifdef SSDBOX_EXPERIMENTS
      && !getenv("SSDBOX_BWD_TWO_PASS")
#endif
  ) {
    long long tiles = (rows + kBwdTileRows - 1) / kBwdTileRows;
    int grid = (int)(tiles < dev.sm_count ? tiles : dev.sm_count);
    void (*kern)(BwdArgs) = a.C == 81 ? loss_bwd_stream_kernel<81> : (a.C == 21 ? loss_bwd_stream_kernel<21> : loss_bwd_stream_kernel<0>);
    SSDBOX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stream_smem));
    TimerScope ts__(KID_LOSS_BWD, st);
    kern<<<grid, kBwdStreamWarps * 32, stream_smem, st>>>(a);
  } else {
    TimerScope ts__(KID_LOSS_BWD, st);
    zero_fill_kernel<<<dev.sm_count * 16, kBwdThreads, 0, st>>>(grad_conf, (size_t)rows * a.C, a.conf_aligned);
    long long groups = (rows + 31) / 32;
    long long blocks = (groups + (kBwdThreads / 32) - 1) / (kBwdThreads / 32);
    if (blocks > (long long)dev.sm_count * 32) blocks = (long long)dev.sm_count * 32;
    loss_bwd_kernel<<<(int)blocks, kBwdThreads, 0, st>>>(a);
  }
  SSDBOX_LAUNCH_OK("loss_bwd_kernel");
  return SSDBOX_OK;
}

extern "C" int ssdbox_hard_negative_mine(const float* keys, const uint8_t* pos, const uint8_t* pool, int32_t B,
                                         int32_t P, int32_t negpos_ratio, uint8_t* neg, void* ws, size_t ws_bytes,
                                         ssdbox_stream_t stream) {
  SSDBOX_REQUIRE(B >= 0 && P >= 0 && negpos_ratio >= 0, SSDBOX_EINVAL, "mine: negative size");
  if (B == 0 || P == 0) return SSDBOX_OK;
  SSDBOX_REQUIRE(keys && pos && neg && ws, SSDBOX_EINVAL, "mine: null pointer");
  SSDBOX_REQUIRE(ws_bytes >= mine_ws_bytes(B, P), SSDBOX_EWORKSPACE, "mine: workspace too small");
  DevInfo dev;
  int rc = get_dev_info(&dev);
  if (rc) return rc;
  size_t fixed = kMineFixedSmem;
  int in_smem = (fixed + (size_t)P * 4 <= (size_t)dev.max_smem_optin - 1024) ? 1 : 0;
  size_t smem = fixed + (in_smem ? (size_t)P * 4 : 0);
  SSDBOX_CUDA(cudaFuncSetAttribute(mine_only_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mine_only_kernel<<<B, kMineThreads, smem, static_cast<cudaStream_t>(stream)>>>(keys, pos, pool, P, negpos_ratio, neg,
                                                                                static_cast<uint32_t*>(ws), in_smem);
  SSDBOX_LAUNCH_OK("mine_only_kernel");
  return SSDBOX_OK;
}
