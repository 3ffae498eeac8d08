/*
 * ssdbox.h -- C ABI of libssdbox.so: the SSD-series box hot path on B200 (sm_100a).
 *
 * The reference (arleyzhang/object-detection-pytorch) is pure Python and has no FFI of its
 * own; each entry point below names the reference function (file:line under the reference
 * root) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a maintainer
 * adds to lib/layers/ to route the reference's PriorBoxSSD / MultiBoxLoss / DetectOut
 * through these calls.
 *
 * Conventions
 *   - every pointer is a BORROWED DEVICE pointer (row-major, contiguous, fp32 unless typed
 *     otherwise) that the caller keeps alive until the stream work has completed; the only
 *     host pointers are the small *_cfg structs and ssdbox_last_error's buffer;
 *   - all work is enqueued on `stream` (a cudaStream_t); no call synchronises, allocates
 *     device memory or touches the default stream, so every call is CUDA-graph capturable;
 *   - scratch memory is caller-provided: ask ssdbox_workspace_bytes(), pass `ws`/`ws_bytes`
 *     (256-byte aligned).  The library keeps no mutable global state besides a cache of
 *     per-device attributes; calls are re-entrant and thread-safe;
 *   - every function returns SSDBOX_OK (0) or a negative SSDBOX_E* code, never throws across
 *     the ABI; ssdbox_last_error() returns the calling thread's last message;
 *   - there is NO CPU fallback: without a CUDA device the compute calls return SSDBOX_ECUDA.
 *
 * Ground truth layout (lib/datasets/det_dataset.py:63-85 collate -> train.py:128-130):
 *   gt          [gt_offsets[B], 5]  rows = (x1, y1, x2, y2, label0based) normalised xyxy
 *   gt_offsets  int32 [B+1]         image b owns rows gt_offsets[b] .. gt_offsets[b+1]-1
 * An image with zero rows is all background (multibox_loss_v1.py:70-71 skips such images).
 */
#ifndef SSDBOX_H_
#define SSDBOX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSDBOX_ABI_VERSION 1

#if defined(__GNUC__)
#define SSDBOX_API __attribute__((visibility("default")))
#else
#define SSDBOX_API
#endif

typedef void* ssdbox_stream_t; /* cudaStream_t */

enum {
  SSDBOX_OK = 0,
  SSDBOX_EINVAL = -1,     /* bad argument value (null pointer, negative size, nms_thresh<=0 ...) */
  SSDBOX_ESHAPE = -2,     /* unsupported shape (too many classes / truths / top_k ...)           */
  SSDBOX_EALIGN = -3,     /* pointer not aligned as required (boxes: 16 B)                       */
  SSDBOX_EWORKSPACE = -4, /* ws_bytes smaller than ssdbox_workspace_bytes()                      */
  SSDBOX_ECUDA = -5       /* CUDA runtime / launch error, message in ssdbox_last_error()         */
};

enum {
  SSDBOX_OP_MATCH = 1,
  SSDBOX_OP_LOSS_FWD = 2,
  SSDBOX_OP_DETECT = 3,
  SSDBOX_OP_NMS = 4,
  SSDBOX_OP_LSE = 5,
  SSDBOX_OP_MINE = 6,
  SSDBOX_OP_COMPACT = 7,      /* ssdbox_detections_compact: B, C used */
  SSDBOX_OP_VOC_EVAL = 8      /* ssdbox_voc_eval: P = detection rows, C = classes, gmax = truths (all images) */
};

SSDBOX_API int ssdbox_abi_version(void);
/* copies the calling thread's last error message (NUL terminated) into buf; returns its length */
SSDBOX_API int ssdbox_last_error(char* buf, size_t n);
/* bytes of scratch `op` needs for these sizes (P = priors, C = classes, gmax = max truths per
 * image, top_k as in DetectOut / nms, n = candidate count for SSDBOX_OP_NMS passed as P) */
SSDBOX_API size_t ssdbox_workspace_bytes(int op, int B, int P, int C, int gmax, int top_k);

/* ------------------------------------------------------------------------------------------
 * Prior generation -- PriorBoxSSD.forward / _create_prior (lib/layers/functions/prior_box.py:
 * 92-111, 122-143).  fp64 arithmetic, rounded once to fp32, optional clamp to [0,1].
 * ---------------------------------------------------------------------------------------- */
#define SSDBOX_MAX_LAYERS 16
#define SSDBOX_MAX_MIN_SIZES 4
#define SSDBOX_MAX_RATIOS 6

typedef struct {
  int32_t num_layers;
  int32_t clip;                 /* cfg.MODEL.CLIP  */
  int32_t flip;                 /* cfg.MODEL.FLIP  */
  int32_t has_max;              /* len(cfg.MODEL.MAX_SIZES) != 0 */
  double image_h, image_w;      /* cfg.MODEL.IMAGE_SIZE = (h, w) */
  int32_t feat_h[SSDBOX_MAX_LAYERS], feat_w[SSDBOX_MAX_LAYERS]; /* layer_dims */
  double step[SSDBOX_MAX_LAYERS];                                /* cfg.MODEL.STEPS */
  int32_t num_min[SSDBOX_MAX_LAYERS];
  double min_size[SSDBOX_MAX_LAYERS][SSDBOX_MAX_MIN_SIZES];      /* cfg.MODEL.MIN_SIZES */
  double max_size[SSDBOX_MAX_LAYERS];                            /* cfg.MODEL.MAX_SIZES */
  int32_t num_ratio[SSDBOX_MAX_LAYERS];
  double ratio[SSDBOX_MAX_LAYERS][SSDBOX_MAX_RATIOS];            /* cfg.MODEL.ASPECT_RATIOS */
} ssdbox_prior_cfg;

/* number of priors the configuration produces (host-side arithmetic only), <0 on error */
SSDBOX_API int64_t ssdbox_priorbox_count(const ssdbox_prior_cfg* cfg);
/* out: [count,4] (cx,cy,w,h) */
SSDBOX_API int ssdbox_priorbox(const ssdbox_prior_cfg* cfg, float* out, int64_t out_rows, ssdbox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Box algebra -- lib/layers/box_utils.py
 * ---------------------------------------------------------------------------------------- */
/* point_form :6-15   (cx,cy,w,h) -> (x1,y1,x2,y2) */
SSDBOX_API int ssdbox_point_form(const float* boxes, int64_t n, float* out, ssdbox_stream_t stream);
/* intent of center_size :18-27   (x1,y1,x2,y2) -> (cx,cy,w,h) */
SSDBOX_API int ssdbox_center_form(const float* boxes, int64_t n, float* out, ssdbox_stream_t stream);
/* jaccard :51-70   a[G,4] xyxy, b[P,4] xyxy -> out[G,P] */
SSDBOX_API int ssdbox_jaccard(const float* a, int32_t G, const float* b, int32_t P, float* out, ssdbox_stream_t stream);
/* encode :201-222  matched[n,4] xyxy, priors[n,4] centre form -> out[n,4] */
SSDBOX_API int ssdbox_encode(const float* matched, const float* priors, int64_t n, float var0, float var1,
                  float* out, ssdbox_stream_t stream);
/* decode :226-244  loc[n,4]; priors[prior_rows,4] reused cyclically (row i uses prior i % prior_rows)
 * so a whole [B,P,4] batch decodes in one call; out[n,4] xyxy.  out_center (nullable) receives
 * the centre form of the decoded box (RefineDet refined anchors). */
SSDBOX_API int ssdbox_decode(const float* loc, const float* priors, int64_t n, int64_t prior_rows, float var0,
                  float var1, float* out, float* out_center, ssdbox_stream_t stream);
/* log_sum_exp :265-273  x[rows,C] -> out[rows]; uses ONE global max over all of x like the
 * reference (two passes).  ws: SSDBOX_OP_LSE. */
/* out[0] (double, device) = max over x[0..n) -- the x.data.max() of box_utils.py:272; -inf for n = 0.  ws: SSDBOX_OP_LSE. */
SSDBOX_API int ssdbox_global_max(const float* x, int64_t n, double* out, void* ws, size_t ws_bytes, ssdbox_stream_t stream);
SSDBOX_API int ssdbox_log_sum_exp(const float* x, int64_t rows, int32_t C, float* out, void* ws, size_t ws_bytes,
                       ssdbox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * match -- box_utils.py:92-133 for a whole batch (the python loop of multibox_loss.py:69-74).
 *   priors            [P,4] centre form (prior_batch_stride = 0) or per-image [B,P,4]
 *                     (prior_batch_stride = 4*P floats; RefineDet refined anchors)
 *   anchors_xyxy      nullable; when given ([B,P,4] or [P,4], same stride rule) it replaces
 *                     point_form(priors) in the IoU (RefineDet: IoU against the decoded ARM boxes)
 *   loc_t  [B,P,4]    encoded regression targets           (nullable)
 *   conf_t [B,P]      int64 class targets, 0 = background  (nullable)
 *   match_idx [B,P]   int32 index of the matched truth within its image (nullable)
 *   overlap [B,P]     best-truth IoU, 2.0 for forced best priors (nullable)
 * IoU, argmax (first index on ties), the sequential "last truth wins" forced assignment and the
 * threshold compare are bit-exact with the reference; encode uses logf (1e-5 relative).
 * ws: SSDBOX_OP_MATCH.
 * ---------------------------------------------------------------------------------------- */
SSDBOX_API int ssdbox_match_encode(const float* gt, const int32_t* gt_offsets, int32_t gmax, const float* priors,
                        int64_t prior_batch_stride, const float* anchors_xyxy, int32_t B, int32_t P,
                        float threshold, float var0, float var1, int32_t binarize_labels,
                        float* loc_t, int64_t* conf_t, int32_t* match_idx, float* overlap, void* ws,
                        size_t ws_bytes, ssdbox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Hard-negative selection in isolation -- multibox_loss.py:97-103.
 *   keys [B,P]  mining loss (fp32), pos [B,P] uint8 (conf_t > 0), pool [B,P] uint8 nullable
 *   neg  [B,P]  uint8 out: rank(key zeroed at positives, descending, ties by ascending prior
 *               index) < min(negpos_ratio * num_pos, P-1)
 * Bit-exact with the reference's double sort on identical keys (canonical tie order).
 * ws: SSDBOX_OP_MINE.
 * ---------------------------------------------------------------------------------------- */
SSDBOX_API int ssdbox_hard_negative_mine(const float* keys, const uint8_t* pos, const uint8_t* pool, int32_t B,
                              int32_t P, int32_t negpos_ratio, uint8_t* neg, void* ws, size_t ws_bytes,
                              ssdbox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * MultiBoxLoss -- lib/layers/modules/multibox_loss.py:48-117 (+ autograd, train.py:143-144).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, P, C;
  int32_t gmax;               /* max truths per image (>= every gt_offsets difference) */
  float threshold;            /* overlap_thresh (0.5)  */
  int32_t negpos_ratio;       /* neg_pos (3)           */
  float var0, var1;           /* cfg.MODEL.VARIANCE    */
  int32_t binarize_labels;    /* RefineDet ARM: every truth label becomes class 1 */
  int32_t finalize;           /* 1: losses = sums / N on device; 0: leave sums for an all-reduce */
  int64_t prior_batch_stride; /* 0 or 4*P (see ssdbox_match_encode) */
  int32_t flags;              /* SSDBOX_LOSS_* bits */
  int32_t reserved;
} ssdbox_loss_cfg;

/* By default the matching runs on dedicated warps inside the streaming kernel (it overlaps the
 * HBM-bound pass over conf).  This flag runs it as its own kernel before the streaming pass
 * instead (same results; the library also falls back to it when a CTA's truths do not fit in
 * shared memory beside the ring). */
#define SSDBOX_LOSS_SEPARATE_MATCH 1
/* Use the generic (shared-memory) mining kernel even when the register-resident one applies
 * (P % 4 == 0 and P <= 24576); same results, for testing. */
#define SSDBOX_LOSS_GENERIC_MINE 2
/* Keep the register-resident mining kernel on one CTA per image (by default images with P >= 8192
 * are split over a thread-block cluster of two CTAs); same results, for testing. */
#define SSDBOX_LOSS_NO_CLUSTER 4
/* Peer-reduced calls only: the mining kernel posts this rank's sums to the peers but does not wait
 * for theirs; `sums` / `losses` hold LOCAL values until ssdbox_multibox_loss_peer_finish runs on
 * the same stream.  Work enqueued in between (e.g. DetectOut of the same step) overlaps the wait, so
 * rank skew no longer stalls the step. */
#define SSDBOX_LOSS_DEFER_PEER_WAIT 8
/* Workspace state is clean: skip the init launch.  The op keeps per-truth best-prior keys, per-image mining
 * histograms and work tickets in `ws`; every call hands them back initialised (the mining CTA of an image resets
 * what it has read, the last CTA the tickets).  A caller may set this flag iff the previous work on this `ws` was
 * a COMPLETED ssdbox_multibox_loss_fwd[_peers] call with the same (B, P, C, gmax) and the same flags.  Without the
 * flag nothing is assumed about the workspace contents.  The host mirror (MultiBoxLoss) sets it from the second
 * call on. */
#define SSDBOX_LOSS_WS_CLEAN 16
/* The reference's log_sum_exp subtracts ONE maximum -- x.data.max() over the whole batch_conf it sees
 * (box_utils.py:272-273) -- where this library by default subtracts each row's own maximum (same value up to fp32
 * rounding, and no underflow for rows far below the batch maximum).  With this flag the value to subtract is taken
 * from sums[0] ON ENTRY (a double; e.g. written by ssdbox_global_max on the same stream, all-reduced with MAX by the
 * caller when the batch is sharded over ranks) and the rows are evaluated as logf(sum expf(x - shift)) + shift like the
 * reference, underflow and all.  Slower (generic streaming kernel); a fidelity mode, not the default. */
#define SSDBOX_LOSS_LSE_SHIFT 32
/* mining with two 512-thread CTAs per SM instead of one 1024-thread CTA (chosen automatically when the batch has more
 * images than the device has SMs and the image fits: P <= 12288; this flag forces it wherever it fits -- tests).  The
 * selected sets are the same; the fp64 partial sums are added in a different (still fixed) order. */
#define SSDBOX_LOSS_MINE_HALF_CTA 64

/* forward.
 *   loc [B,P,4], conf [B,P,C] raw logits, priors, anchors_xyxy (nullable), gt/gt_offsets
 *   pool  [B,P] uint8 nullable  RefineDet: anchors with pool==0 are neither positive nor mined
 *   sums  double[3]  out: { sum smooth-L1 over positives, sum CE over pos U neg, N = #positives }
 *   losses float[2]  out: { loss_l, loss_c } = sums[0..1] / N when cfg->finalize (0 if N == 0)
 *   sel   [B,P] int16 out (kept for backward): -1 = row not in pos U neg, else its class target
 *   tidx  [B,P] int16 out (kept for backward): matched truth index within the image
 *   dbg_conf_t int64 [B,P], dbg_loc_t [B,P,4], dbg_neg uint8 [B,P], dbg_keys [B,P]: nullable
 *         materialisations of the reference's intermediates (tests / drop-in users of loc_t).
 * ws: SSDBOX_OP_LOSS_FWD. */
SSDBOX_API int ssdbox_multibox_loss_fwd(const ssdbox_loss_cfg* cfg, const float* loc, const float* conf,
                             const float* priors, const float* anchors_xyxy, const uint8_t* pool,
                             const float* gt, const int32_t* gt_offsets, double* sums, float* losses,
                             int16_t* sel, int16_t* tidx, int64_t* dbg_conf_t, float* dbg_loc_t,
                             uint8_t* dbg_neg, float* dbg_keys, void* ws, size_t ws_bytes,
                             ssdbox_stream_t stream);

/* ---- multi-GPU: loss sums reduced over NVLink peer memory inside the mining kernel ------------
 * The only cross-image coupling of MultiBoxLoss is { sum smooth-L1, sum CE, N } (multibox_loss.py:
 * 114-116).  With one process per GPU, every rank owns an exchange buffer of
 * ssdbox_peer_buffer_bytes() bytes, ZERO-FILLED ONCE at creation and mapped into every peer
 * (CUDA IPC / symmetric memory; the library never allocates).  The last CTA of the mining kernel
 * stores the rank's three sums into its slot of every peer's buffer (st.release.sys through
 * NVLink), waits for the slots of its own buffer (ld.acquire.sys), adds them in rank order
 * (bit-identical on every rank) and finalises: `sums` / `losses` then hold the GLOBAL values.
 * There is no separate collective launch.  Like any collective, every rank must issue the same
 * sequence of peer-reduced forwards; a call epoch kept in the buffer makes the call replayable
 * from a CUDA graph.  A rank whose local shard is empty (B == 0: global batch smaller than the world)
 * still calls: it posts zeros, so the epochs stay in step.  A peer that never arrives does NOT kill the
 * context: after wait_timeout_ms (wall clock; 0 = 30 s) the waiting rank gives up, its sums / losses of
 * that call become NaN and the event is counted in 64-bit word 1 of its own exchange buffer (word 0 is
 * the call epoch); the host mirror reads it (PeerExchange.timeouts()).  After a timeout the ranks are out
 * of step: rebuild the exchange (zero-fill + rendezvous) before the next call. */
#define SSDBOX_MAX_PEERS 16
typedef struct {
  int32_t rank, world;                 /* 1 <= world <= SSDBOX_MAX_PEERS */
  void* bufs[SSDBOX_MAX_PEERS];        /* bufs[r]: rank r's exchange buffer as addressable from THIS device */
  int64_t wait_timeout_ms;             /* bound of the wait for the peers' sums; 0 = default (30 000) */
} ssdbox_peer_group;
SSDBOX_API size_t ssdbox_peer_buffer_bytes(void);
/* completes a call made with SSDBOX_LOSS_DEFER_PEER_WAIT: waits for every rank's sums, adds them in
 * rank order, writes the global sums[3] and losses[2] (nullable). */
SSDBOX_API int ssdbox_multibox_loss_peer_finish(const ssdbox_peer_group* peers, double* sums, float* losses,
                                     ssdbox_stream_t stream);
/* ssdbox_multibox_loss_fwd with the reduction above; peers == NULL behaves like the plain call. */
SSDBOX_API int ssdbox_multibox_loss_fwd_peers(const ssdbox_loss_cfg* cfg, const float* loc, const float* conf,
                             const float* priors, const float* anchors_xyxy, const uint8_t* pool,
                             const float* gt, const int32_t* gt_offsets, double* sums, float* losses,
                             int16_t* sel, int16_t* tidx, int64_t* dbg_conf_t, float* dbg_loc_t,
                             uint8_t* dbg_neg, float* dbg_keys, const ssdbox_peer_group* peers, void* ws, size_t ws_bytes,
                             ssdbox_stream_t stream);
/* ---- RefineDet, fused two-step path (arXiv 1711.06897; SURVEY.md 8a-R -- no reference code, parity unpinned) ----
 * The ODM loss and the refined DetectOut take the ARM head's outputs as they are: an anchor of image b is
 * decode(arm_loc[b,p], priors[p]) (box_utils.py:238-243), recomputed inside the kernels that need it (the match
 * warps of the streaming kernel, the positives of the mining / backward kernels, the candidates of Detect), and
 * an anchor whose objectness softmax(arm_conf[b,p])[1] is <= theta leaves the positives, the hard-negative pool
 * and the detections.  No refined-anchor tensors ([B,P,4] xyxy + centre form) and no [B,P] mask are
 * materialised, and there is no extra launch.  Bit-identical to the materialised path (ssdbox_decode with
 * out_center + ssdbox_arm_filter feeding anchors_xyxy / pool / score_keep).  `priors` is the shared [P,4]
 * centre-form tensor (cfg->prior_batch_stride must be 0). */
typedef struct {
  const float* arm_loc;   /* [B,P,4] ARM regression offsets, 16-byte aligned */
  const float* arm_conf;  /* [B,P,2] ARM objectness logits, 16-byte aligned; NULL: no filtering */
  float theta;            /* objectness threshold (0.01) */
  int32_t reserved;
} ssdbox_refine;
SSDBOX_API int ssdbox_multibox_loss_fwd_refine(const ssdbox_loss_cfg* cfg, const float* loc, const float* conf,
                             const float* priors, const ssdbox_refine* refine,
                             const float* gt, const int32_t* gt_offsets, double* sums, float* losses,
                             int16_t* sel, int16_t* tidx, int64_t* dbg_conf_t, float* dbg_loc_t,
                             uint8_t* dbg_neg, float* dbg_keys, const ssdbox_peer_group* peers, void* ws, size_t ws_bytes,
                             ssdbox_stream_t stream);
SSDBOX_API int ssdbox_multibox_loss_bwd_refine(const ssdbox_loss_cfg* cfg, const float* loc, const float* conf,
                             const float* priors, const ssdbox_refine* refine, const float* gt, const int32_t* gt_offsets,
                             const int16_t* sel, const int16_t* tidx, const double* sums,
                             const float* grad_out, float* grad_loc, float* grad_conf,
                             ssdbox_stream_t stream);
/* losses = sums[0..1] / sums[2]  (after the caller all-reduced `sums` across ranks) */
SSDBOX_API int ssdbox_multibox_loss_finalize(const double* sums, float* losses, ssdbox_stream_t stream);
/* backward: grad_loc [B,P,4], grad_conf [B,P,C] (both fully written);
 * grad_out float[2] device = upstream gradients of (loss_l, loss_c); sums[2] = N (global). */
SSDBOX_API int ssdbox_multibox_loss_bwd(const ssdbox_loss_cfg* cfg, const float* loc, const float* conf,
                             const float* priors, const float* gt, const int32_t* gt_offsets,
                             const int16_t* sel, const int16_t* tidx, const double* sums,
                             const float* grad_out, float* grad_loc, float* grad_conf,
                             ssdbox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * nms -- box_utils.py:279-343.  boxes[n,4] xyxy, scores[n]; keep int64[n] zero padded (indices
 * into boxes, visiting order), count int32[1] (device).  Canonical tie order: equal scores are
 * visited higher index first.  Any top_k >= 1 like the reference (:299-301): up to 1024 the list is sorted and swept in
 * shared memory; beyond that a slower kernel sorts in the workspace and sweeps without a suppression matrix (at most
 * 65536 boxes visited).  ws: SSDBOX_OP_NMS (P = n; depends on top_k).
 * ---------------------------------------------------------------------------------------- */
SSDBOX_API int ssdbox_nms(const float* boxes, const float* scores, int32_t n, float overlap, int32_t top_k,
               int64_t* keep, int32_t* count, void* ws, size_t ws_bytes, ssdbox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * DetectOut.forward -- lib/layers/functions/detection.py:25-64.
 *   loc [B,P,4]; scores [B,P,C] (== [B*P,C], rfb_net.py:222-226) softmax probabilities
 *   priors [P,4] (or per image, prior_batch_stride = 4*P)
 *   score_keep [B,P] uint8 nullable: RefineDet, scores of anchors with 0 are treated as 0
 *   out [B,C,top_k,5] fully written: rows (score,x1,y1,x2,y2) in NMS order, zero padded,
 *       class-0 plane zero;  counts int32 [B,C] nullable
 * Errors like detection.py:19-20: nms_thresh <= 0 -> SSDBOX_EINVAL.  ws: SSDBOX_OP_DETECT.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, P, C;
  int32_t top_k;              /* 1..65536; > 1024: the slow any-top_k path (one CTA per list), not with SSDBOX_DETECT_LOGITS */
  float conf_thresh, nms_thresh;
  float var0, var1;
  int64_t prior_batch_stride;
  int32_t flags;              /* SSDBOX_DETECT_* bits */
  int32_t reserved;
} ssdbox_detect_cfg;

/* `scores` holds RAW LOGITS: the softmax every model applies right before DetectOut (ssd_v3.py:
 * 123-124, rfb_net.py:222-226) is fused into the streaming pass -- score = expf(x - rowmax) /
 * sum_c expf(x_c - rowmax), computed only for rows that can hold a candidate.  Saves the separate
 * softmax kernel's read + write of conf (2 x 4*B*P*C bytes).  Scores then agree with
 * torch.softmax to fp32 rounding (<= 1e-6 relative) instead of bit for bit. */
#define SSDBOX_DETECT_LOGITS 1
/* Workspace state is clean: skip the init launch.  The op keeps a few counters in `ws`; every call hands them
 * back zeroed (each list's counter is reset by the kernel that finishes the list).  A caller may therefore set
 * this flag iff the previous work on this `ws` was a COMPLETED ssdbox_detect call with the same (B, P, C, top_k)
 * on the same stream order (or the state region was just zero-filled).  Without the flag nothing is assumed
 * about the workspace contents.  The host mirror (DetectOut) sets it from the second call on. */
#define SSDBOX_DETECT_WS_CLEAN 2

SSDBOX_API int ssdbox_detect(const ssdbox_detect_cfg* cfg, const float* loc, const float* scores,
                  const float* priors, const uint8_t* score_keep, float* out, int32_t* counts, void* ws,
                  size_t ws_bytes, ssdbox_stream_t stream);
/* ssdbox_detect that also completes a loss forward made with SSDBOX_LOSS_DEFER_PEER_WAIT on the same stream (what
 * ssdbox_multibox_loss_peer_finish does, without a launch of its own): one warp of the last Detect kernel waits for
 * every rank's sums and writes the global loss_sums[3] / losses[2] (nullable).  In a step "loss forward, then
 * DetectOut" the Detect kernels run while the other ranks' sums arrive.  peers == NULL: plain ssdbox_detect. */
/* RefineDet inference: decode(odm_loc, refined anchors), scores of anchors with ARM objectness <= theta count as 0
 * (see ssdbox_refine above; `priors` shared [P,4], prior_batch_stride 0). */
SSDBOX_API int ssdbox_detect_refine(const ssdbox_detect_cfg* cfg, const float* loc, const float* scores,
                  const float* priors, const ssdbox_refine* refine, float* out, int32_t* counts, void* ws,
                  size_t ws_bytes, ssdbox_stream_t stream);
SSDBOX_API int ssdbox_detect_peers(const ssdbox_detect_cfg* cfg, const float* loc, const float* scores,
                  const float* priors, const uint8_t* score_keep, float* out, int32_t* counts,
                  const ssdbox_peer_group* peers, double* loss_sums, float* losses, void* ws, size_t ws_bytes,
                  ssdbox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Eval post-processing right after DetectOut -- replaces
 * lib/utils/evaluate_utils.py:63-70 (rescale), :127-139 / :175-190 (convert_ssd_result) and
 * :193-203 (EvalCOCO.post_proc).
 *   det        [B,C,K,5] rows (score, x1, y1, x2, y2), e.g. the output of ssdbox_detect
 *   extra      [B,2] (h, w) of each image, nullable (no rescale)          evaluate_utils.py:63-68
 *   image_ids  [B] fp32 nullable: dataset ids of the images (COCO modes)   :180
 *   mode 0  VOC  rows [xmin, ymin, xmax, ymax, score, image, cls]          :127-139
 *        1  COCO rows [xmin, ymin, xmax, ymax, score, image, cls, cocoid]  :175-190
 *        2  COCO result rows [cocoid, x1, y1, w, h, score, cls]            :193-199
 *   out        [capacity_rows, 7 | 8]; rows with score > 0 in (image, class, k) order (the order
 *              masked_select yields); rows beyond capacity_rows are dropped, *total still counts them
 *   total      int32[1] number of rows;  seg_offsets int32 [B*C+1] nullable: first row of every
 *              (image, class) segment (what EvalVOC.post_proc re-derives with numpy masks)
 * ws: SSDBOX_OP_COMPACT. */
SSDBOX_API int ssdbox_detections_compact(const float* det, int32_t B, int32_t C, int32_t K, const float* extra,
                              const float* image_ids, int32_t mode, float* out, int64_t capacity_rows,
                              int32_t* total, int32_t* seg_offsets, void* ws, size_t ws_bytes,
                              ssdbox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * PASCAL VOC evaluation of the accumulated detections -- replaces the chain
 * evaluate_detections -> write_voc_results_file -> do_python_eval -> voc_eval -> voc_ap of
 * lib/datasets/voc_eval.py:58-75, 78-106, 109-242, 244-262 for all classes in one call.
 *   rows        [num_rows, row_stride] fp32, columns 0..4 = (xmin, ymin, xmax, ymax, score) in pixels:
 *               the rows ssdbox_detections_compact writes (mode 0 / 1), all images concatenated,
 *               grouped by (image, class) segments
 *   seg_offsets int32 [num_images*num_classes + 1] first row of every (image, class) segment
 *   gt_boxes    [num_gt,4] fp32 pixel boxes as parse_rec yields them (:26-30), 16-byte aligned;
 *   gt_labels   int32 [num_gt] class index as in the rows' class column (1-based, 0 = background);
 *   gt_difficult uint8 [num_gt];  gt_offsets int32 [num_images+1] truths of image i
 * The reference prints every detection to a text file ('{:.3f}' score, '{:.1f}' coordinate + 1) and
 * parses it back; the same quantisation is applied arithmetically (exact, see voceval.cu).  Scores
 * must quantise into [0, 1]; *status counts the rows that do not (their bin is clamped), plus 2^30 when
 * seg_offsets does not cover the rows exactly (seg_offsets[num_images*num_classes] != num_rows).
 * Outputs (sorted order = by class, then descending quantised score, equal scores in row order --
 * the stable form of np.argsort(-confidence) :178):
 *   order       int32 [num_rows]  sorted position -> row index
 *   cls_offsets int32 [num_classes+1] class c owns sorted positions cls_offsets[c] .. cls_offsets[c+1]-1
 *   tpfp        uint8 [num_rows]  1 true positive, 2 false positive, 0 neither (matched a difficult truth :208)
 *   rec, prec   fp64 [num_rows]   :218-223, bit-exact
 *   ap          fp64 [num_classes] ap[c] for c >= 1; -1 for a class without detections (:238-241) and for
 *               c = 0 (rows of background segments sort in front with tpfp = 0 and are not evaluated);
 *               11-point metric bit-exact, area metric summed in a fixed tree order (1e-12 relative)
 *   npos        int32 [num_classes] non-difficult truths per class (:163)
 * ws: SSDBOX_OP_VOC_EVAL. */
typedef struct {
  int32_t num_images, num_classes;
  int32_t num_rows, row_stride;
  int32_t num_gt;
  int32_t use_07_metric;
  double ovthresh;
} ssdbox_voc_eval_cfg;
SSDBOX_API int ssdbox_voc_eval(const ssdbox_voc_eval_cfg* cfg, const float* rows, const int32_t* seg_offsets,
                    const float* gt_boxes, const int32_t* gt_labels, const uint8_t* gt_difficult,
                    const int32_t* gt_offsets, int32_t* order, int32_t* cls_offsets, uint8_t* tpfp, double* rec,
                    double* prec, double* ap, int32_t* npos, int32_t* status, void* ws, size_t ws_bytes,
                    ssdbox_stream_t stream);

/* The data-parallel part of RandomSampleCrop trials (lib/utils/augmentations.py:13-37 jaccard_numpy,
 * :250-268), batched over B images x T candidate rects, fp64 like the numpy pipeline (bit-exact).
 *   boxes       fp64 [box_offsets[B], 4] absolute xyxy truths;  box_offsets int32 [B+1]
 *   rects       int64 [B,T,4] candidate crops (x1, y1, x2, y2) (:248)
 *   overlap     fp64, image b holds [T, G_b] at element box_offsets[b]*T (nullable)
 *   minmax      fp64 [B,T,2] overlap.min(), overlap.max() of the trial (:254); (+inf, -inf) when G_b = 0
 *   center_mask uint8, same layout as overlap: truth centre strictly inside the rect (:257-268) (nullable) */
SSDBOX_API int ssdbox_crop_overlaps(const double* boxes, const int32_t* box_offsets, const int64_t* rects, int32_t B,
                         int32_t T, double* overlap, double* minmax, uint8_t* center_mask, ssdbox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Head-output layout -- replaces lib/models/ssd_v3.py:114-121
 * (rfb_net.py:213-220): permute(0,2,3,1).contiguous() of every multibox head output, view(B,-1)
 * and cat(dim 1), i.e. the producer of loc [B,P,4] / conf [B,P,C].
 *   src[k]   [B, channels_k, H_k, W_k] contiguous NCHW, channels_k = anchors_k * (4 | C), hw_k = H_k*W_k
 *   out      [B, sum_k hw_k * channels_k]: per image the layers in order, each in (h, w, channel) order
 * Every element read once and written once.  Layers with H*W % 4 == 0, at least 64 positions, 32..576 channels and a
 * 16-byte aligned pointer go through a tensor-map TMA ring (one launch for all of them), every other layer through a tile
 * kernel (a second launch); which path a layer takes changes nothing in the result (pure data movement, bit for bit). */
#define SSDBOX_MAX_HEADS 16
typedef struct {
  int32_t num_layers, B;
  int32_t channels[SSDBOX_MAX_HEADS];
  int32_t hw[SSDBOX_MAX_HEADS];
  const float* src[SSDBOX_MAX_HEADS];
} ssdbox_heads_cfg;
SSDBOX_API int ssdbox_heads_to_rows(const ssdbox_heads_cfg* cfg, float* out, ssdbox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * RefineDet glue (not in the reference snapshot; arXiv 1711.06897).
 *   arm_conf [n,2] logits -> keep[n] uint8 = softmax(arm_conf)[:,1] > theta
 * (refined anchors come from ssdbox_decode with out_center).
 * ---------------------------------------------------------------------------------------- */
SSDBOX_API int ssdbox_arm_filter(const float* arm_conf, int64_t n, float theta, uint8_t* keep, ssdbox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Opt-in per-kernel timing for bench.py / profiling (the only process-global state; off by
 * default).  While enabled every launch of the library is bracketed by CUDA events on the
 * caller's stream (do NOT enable during CUDA-graph capture).  ssdbox_timers_read synchronises
 * the pending events and returns the accumulated device time and launch count of one kernel:
 *   0 init, 1 match, 2 loss_stream, 3 mine_reduce, 4 loss_bwd, 5 detect_stream,
 *   6 detect_segment (warp path), 7 detect_overflow, 8 materialize, 9 detect_segment_big
 * ---------------------------------------------------------------------------------------- */
#define SSDBOX_KERNEL_COUNT 10
SSDBOX_API int ssdbox_timers_enable(int on);
SSDBOX_API int ssdbox_timers_read(int kernel_id, double* total_ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* SSDBOX_H_ */
