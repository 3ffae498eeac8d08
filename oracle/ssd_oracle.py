"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the SSD-series box hot path.

A self-contained restatement (torch-CPU fp32 tensor ops + plain Python float64) of the
reference algorithms, written so that every floating-point operation happens in the same
order and precision as in arleyzhang/object-detection-pytorch.  It exists because the
reference tree (``/root/reference``) does not travel to the GPU box.

PARITY STATUS
  * SSD path (priors, IoU, match, encode, decode, log-sum-exp, MultiBoxLoss, NMS, Detect):
    PINNED -- ``tests/test_oracle_vs_reference.py`` proves bit-identity against the imported
    reference in the build container, and ``tests/golden/*.npz`` (made by
    ``oracle/make_golden.py`` from the *reference*, not from this file) pin it on the GPU box.
  * RefineDet two-step path (``refine_*``): PARITY UNPINNED -- the reference snapshot has no
    RefineDet code (only README.md:6 mentions it); the functions below restate arXiv
    1711.06897 in the idiom of the SSD functions.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.  The product (``object-detection-pytorch_b200/``)
never does, and has no CPU fallback.

Citations are ``file:line`` relative to the reference root.

Tie-breaking: torch CPU ``max`` returns the first index among equal values (verified for both
reduction axes).  torch CPU ``sort`` is not stable for n > 16; every function that sorts takes
``stable`` (default True = the canonical order the CUDA path implements; ``stable=False``
reproduces the reference's literal ``sort`` call and is only meaningful on tie-free inputs).
"""
from math import sqrt

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# a1  prior generation                      lib/layers/functions/prior_box.py:25-50,92-143
# --------------------------------------------------------------------------------------
def _cell_anchors(cx, cy, k, m):
    """One feature-map cell: list of [cx, cy, w, h] in Python float64 (prior_box.py:122-143)."""
    img_h, img_w = m["IMAGE_SIZE"][0], m["IMAGE_SIZE"][1]
    out = []
    mins = m["MIN_SIZES"][k]
    if not isinstance(mins, (list, tuple)):
        mins = [mins]
    for ms in mins:
        sh = ms / img_h
        sw = ms / img_w
        out += [cx, cy, sw, sh]                                   # :129-131
        if len(m["MAX_SIZES"]) != 0:                              # :133-137
            mx = m["MAX_SIZES"][k]
            out += [cx, cy, sqrt(sw * (mx / img_w)), sqrt(sh * (mx / img_h))]
        for ar in m["ASPECT_RATIOS"][k]:                          # :139-142
            out += [cx, cy, sw * sqrt(ar), sh / sqrt(ar)]
            if m["FLIP"]:
                out += [cx, cy, sw / sqrt(ar), sh * sqrt(ar)]
    return out


def num_priors_per_cell(m):
    """prior_box.py:46-50."""
    return [len(_cell_anchors(0, 0, k, m)) // 4 for k in range(len(m["STEPS"]))]


def prior_boxes(m, layer_dims):
    """prior_box.py:92-111.  ``m`` is a dict with the cfg.MODEL fields (config.py:116-124)."""
    for v in m.get("VARIANCE", [0.1, 0.2]):
        if v <= 0:
            raise ValueError("Variances must be greater than 0")   # :33-35
    vals = []
    for k, (fh, fw) in enumerate(layer_dims):
        for i in range(fh):
            for j in range(fw):
                sx = m["IMAGE_SIZE"][1] / m["STEPS"][k]           # :99-102
                sy = m["IMAGE_SIZE"][0] / m["STEPS"][k]
                vals += _cell_anchors((j + 0.5) / sx, (i + 0.5) / sy, k, m)
    out = torch.tensor(vals, dtype=torch.float64).to(torch.float32).view(-1, 4)   # :107
    if m["CLIP"]:
        out.clamp_(min=0, max=1)                                  # :108-110
    return out


# --------------------------------------------------------------------------------------
# a2/a3  box algebra                                   lib/layers/box_utils.py:6-15,30-70
# --------------------------------------------------------------------------------------
def point_form(b):
    """(cx,cy,w,h) -> (x1,y1,x2,y2); half-extent first, then -/+ (box_utils.py:14-15)."""
    half = b[:, 2:] / 2
    return torch.cat((b[:, :2] - half, b[:, :2] + half), 1)


def iou_matrix(a, b):
    """IoU of xyxy boxes a[G,4] vs b[P,4] -> [G,P] (box_utils.py:43-48,63-70)."""
    hi = torch.min(a[:, None, 2:], b[None, :, 2:])
    lo = torch.max(a[:, None, :2], b[None, :, :2])
    ext = torch.clamp(hi - lo, min=0)
    inter = ext[..., 0] * ext[..., 1]
    area_a = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]))[:, None]
    area_b = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))[None, :]
    return inter / (area_a + area_b - inter)


# --------------------------------------------------------------------------------------
# a5/a8  encode / decode                                      box_utils.py:201-222,226-244
# --------------------------------------------------------------------------------------
def encode_boxes(matched, priors, variances):
    ctr = (matched[:, :2] + matched[:, 2:]) / 2 - priors[:, :2]
    ctr = ctr / (variances[0] * priors[:, 2:])
    ext = (matched[:, 2:] - matched[:, :2]) / priors[:, 2:]
    ext = torch.log(ext + 1e-10) / variances[1]
    return torch.cat([ctr, ext], 1)


def decode_boxes(loc, priors, variances):
    ctr = priors[:, :2] + loc[:, :2] * variances[0] * priors[:, 2:]
    ext = priors[:, 2:] * torch.exp(loc[:, 2:] * variances[1])
    lo = ctr - ext / 2            # :242
    hi = ext + lo                 # :243 (max corner is wh + min corner)
    return torch.cat((lo, hi), 1)


def center_form(b):
    """xyxy -> (cx,cy,w,h).  (box_utils.py:18-27 is broken in the snapshot; this is the intent.)"""
    return torch.cat(((b[:, 2:] + b[:, :2]) / 2, b[:, 2:] - b[:, :2]), 1)


# --------------------------------------------------------------------------------------
# a4  matching                                                       box_utils.py:92-133
# --------------------------------------------------------------------------------------
def match_image(threshold, truths, priors, variances, labels, anchors_xyxy=None):
    """Returns dict(loc[P,4] f32, conf[P] i64, truth_idx[P] i64, overlap[P] f32, best_prior[G] i64).

    ``anchors_xyxy`` (RefineDet only) replaces point_form(priors) in the IoU; ``priors`` is then
    the centre form of the same refined anchors.
    """
    P = priors.size(0)
    G = truths.size(0)
    if G == 0:
        # Our defined behaviour (the reference crashes; multibox_loss_v1.py:70-71 skips the image
        # and leaves the rows uninitialised): no truth -> every prior is background, zero targets.
        return dict(loc=torch.zeros(P, 4), conf=torch.zeros(P, dtype=torch.int64),
                    truth_idx=torch.zeros(P, dtype=torch.int64), overlap=torch.zeros(P),
                    best_prior=torch.zeros(0, dtype=torch.int64))
    ov = iou_matrix(truths, point_form(priors) if anchors_xyxy is None else anchors_xyxy)
    _, best_prior = ov.max(1)                 # :116 per truth, first index on ties
    best_ov, best_truth = ov.max(0)           # :118 per prior, first index on ties
    best_ov = best_ov.clone()
    best_truth = best_truth.clone()
    best_ov.index_fill_(0, best_prior, 2)     # :123
    for j in range(G):                        # :126-127 sequential, last truth wins
        best_truth[best_prior[j]] = j
    conf = (labels[best_truth] + 1)           # :129 (float)
    conf[best_ov < threshold] = 0             # :130
    loc = encode_boxes(truths[best_truth], priors, variances)
    return dict(loc=loc, conf=conf.to(torch.int64), truth_idx=best_truth, overlap=best_ov,
                best_prior=best_prior)


# --------------------------------------------------------------------------------------
# a7  log-sum-exp with ONE global max                                box_utils.py:265-273
# --------------------------------------------------------------------------------------
def log_sum_exp(x):
    gmax = x.max()
    return torch.log(torch.sum(torch.exp(x - gmax), 1, keepdim=True)) + gmax


# --------------------------------------------------------------------------------------
# hard-negative selection                                      multibox_loss.py:97-103
# --------------------------------------------------------------------------------------
def hard_negative_select(keys, pos, negpos_ratio, stable=True, pool=None):
    """keys[B,P] f32 mining loss, pos[B,P] bool -> neg[B,P] bool.

    The reference zeroes the keys at positives, sorts descending, sorts the permutation to get
    ranks, and keeps rank < min(ratio*num_pos, P-1).  ``pool`` (RefineDet only): priors outside
    the pool are removed from the ranking altogether.
    """
    k = keys.clone()
    k[pos] = 0
    if pool is not None:
        k[~pool] = float("-inf")
    _, order = k.sort(dim=1, descending=True, stable=stable)
    _, rank = order.sort(1)
    num_pos = pos.long().sum(1, keepdim=True)
    num_neg = torch.clamp(negpos_ratio * num_pos, max=pos.size(1) - 1)
    neg = rank < num_neg
    if pool is not None:
        neg = neg & pool
    return neg


# --------------------------------------------------------------------------------------
# a6  MultiBoxLoss.forward                                       multibox_loss.py:48-117
# --------------------------------------------------------------------------------------
def build_targets(loc_shape, priors, targets, threshold, variances):
    B, P = loc_shape[0], loc_shape[1]
    loc_t = torch.zeros(B, P, 4)
    conf_t = torch.zeros(B, P, dtype=torch.int64)
    truth_idx = torch.zeros(B, P, dtype=torch.int64)
    for b in range(B):
        t = targets[b]
        m = match_image(threshold, t[:, :4], priors, variances, t[:, 4])
        loc_t[b] = m["loc"]
        conf_t[b] = m["conf"]
        truth_idx[b] = m["truth_idx"]
    return loc_t, conf_t, truth_idx


def multibox_loss(loc_data, conf_data, priors, targets, num_classes, threshold=0.5,
                  negpos_ratio=3, variances=(0.1, 0.2), stable=True, detail=False):
    """Returns (loss_l, loss_c) or, with detail=True, a dict with every intermediate."""
    B, P = loc_data.size(0), loc_data.size(1)
    priors = priors[:P, :]                                              # :62
    loc_t, conf_t, truth_idx = build_targets(loc_data.shape, priors, targets, threshold, variances)
    pos = conf_t > 0                                                    # :82
    sel4 = pos.unsqueeze(2).expand_as(loc_data)
    loss_l = F.smooth_l1_loss(loc_data[sel4].view(-1, 4), loc_t[sel4].view(-1, 4),
                              reduction="sum")                          # :87-90
    flat = conf_data.view(-1, num_classes)
    keys = (log_sum_exp(flat) - flat.gather(1, conf_t.view(-1, 1))).view(B, P)   # :93-94
    neg = hard_negative_select(keys, pos, negpos_ratio, stable=stable)   # :97-103
    chosen = pos | neg
    loss_c = F.cross_entropy(conf_data[chosen.unsqueeze(2).expand_as(conf_data)]
                             .view(-1, num_classes), conf_t[chosen], reduction="sum")  # :106-110
    n = pos.long().sum()                                                # :114
    if detail:
        mk = keys.detach().clone()
        mk[pos] = 0
        return dict(loss_l=loss_l / n, loss_c=loss_c / n, sum_l=loss_l, sum_c=loss_c, n=n,
                    loc_t=loc_t, conf_t=conf_t, truth_idx=truth_idx, pos=pos, neg=neg,
                    mining_keys=mk)
    return loss_l / n, loss_c / n


def multibox_loss_grads(loc_data, conf_data, priors, targets, num_classes, **kw):
    """a11: autograd of a6 (train.py:143-144): returns (dloss/dloc, dloss/dconf) of loss_l+loss_c."""
    loc = loc_data.detach().clone().requires_grad_(True)
    conf = conf_data.detach().clone().requires_grad_(True)
    ll, lc = multibox_loss(loc, conf, priors, targets, num_classes, **kw)
    (ll + lc).backward()
    return loc.grad, conf.grad


# --------------------------------------------------------------------------------------
# a10  greedy NMS                                                   box_utils.py:279-343
# --------------------------------------------------------------------------------------
def greedy_nms(boxes, scores, overlap=0.5, top_k=200, stable=True):
    """Returns (keep[n] int64 zero-padded, count).  Visiting order: ascending sort consumed from
    the tail (box_utils.py:299-301,312), i.e. score descending and, with stable=True, the
    *higher* index first among equal scores."""
    n = scores.size(0)
    keep = torch.zeros(n, dtype=torch.int64)
    if boxes.numel() == 0:
        return keep, 0
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])     # :298
    _, asc = scores.sort(dim=0, stable=stable)
    order = asc[-top_k:].flip(0)
    bx = boxes[order]
    ar = area[order]
    m = order.numel()
    alive = torch.ones(m, dtype=torch.bool)
    count = 0
    for i in range(m):
        if not alive[i]:
            continue
        keep[count] = order[i]
        count += 1
        if i + 1 == m:
            break
        rest = bx[i + 1:]
        w = torch.clamp(torch.clamp(rest[:, 2], max=bx[i, 2]) - torch.clamp(rest[:, 0], min=bx[i, 0]),
                        min=0.0)                                       # :325-335
        h = torch.clamp(torch.clamp(rest[:, 3], max=bx[i, 3]) - torch.clamp(rest[:, 1], min=bx[i, 1]),
                        min=0.0)
        inter = w * h
        iou = inter / ((ar[i + 1:] - inter) + ar[i])                   # :339-340
        alive[i + 1:] &= iou.le(overlap)                               # :342
    return keep, count


# --------------------------------------------------------------------------------------
# a9  DetectOut.forward                                              detection.py:25-64
# --------------------------------------------------------------------------------------
def detect(loc_data, conf_data, priors, num_classes, top_k=200, conf_thresh=0.01,
           nms_thresh=0.45, variances=(0.1, 0.2), stable=True, score_mask=None,
           anchors_center=None):
    """conf_data: softmax scores [B,P,C] or [B*P,C].  Returns zeros[B,C,top_k,5] filled with
    (score, x1, y1, x2, y2) in NMS order.  (detection.py:60-63 is a no-op and is omitted.)

    RefineDet only: ``anchors_center`` [B,P,4] per-image refined anchors replace ``priors``;
    ``score_mask`` [B,P] bool zeroes the scores of filtered anchors before thresholding.
    """
    if nms_thresh <= 0:
        raise ValueError("nms_threshold must be non negative.")       # detection.py:19-20
    B = loc_data.size(0)
    P = priors.size(0) if anchors_center is None else anchors_center.size(1)
    out = torch.zeros(B, num_classes, top_k, 5)
    scores_all = conf_data.view(B, P, num_classes)
    for b in range(B):
        pri = priors if anchors_center is None else anchors_center[b]
        boxes = decode_boxes(loc_data[b].view(-1, 4), pri, variances)     # :43
        for c in range(1, num_classes):                               # :47
            s = scores_all[b, :, c]
            if score_mask is not None:
                s = torch.where(score_mask[b], s, torch.zeros_like(s))
            m = s > conf_thresh                                       # :48 strict
            if not bool(m.any()):
                continue
            sc = s[m]
            bx = boxes[m]
            ids, cnt = greedy_nms(bx, sc, nms_thresh, top_k, stable=stable)   # :56
            sel = ids[:cnt]
            out[b, c, :cnt] = torch.cat((sc[sel].unsqueeze(1), bx[sel]), 1)   # :57-59
    return out


# --------------------------------------------------------------------------------------
# a-R  RefineDet two-step path (PARITY UNPINNED: no reference code; arXiv 1711.06897)
# --------------------------------------------------------------------------------------
def refine_anchors(arm_loc, priors, variances=(0.1, 0.2)):
    """decode(arm_loc[b], priors) per image -> (xyxy[B,P,4], centre-form[B,P,4])."""
    xy = torch.stack([decode_boxes(arm_loc[b], priors, variances) for b in range(arm_loc.size(0))])
    cf = torch.stack([center_form(xy[b]) for b in range(xy.size(0))])
    return xy, cf


def arm_objectness(arm_conf):
    """softmax(arm_conf)[..., 1] written as 1 / (1 + exp(x0 - x1)) (the form the kernel uses)."""
    return 1.0 / (1.0 + torch.exp(arm_conf[..., 0] - arm_conf[..., 1]))


def refine_multibox_loss(arm_loc, arm_conf, odm_loc, odm_conf, priors, targets, num_classes,
                         threshold=0.5, negpos_ratio=3, variances=(0.1, 0.2), theta=0.01,
                         use_arm=False, stable=True, detail=False, anchors=None, pool=None):
    """use_arm=False: ARM loss  = MultiBoxLoss with binarised labels (C=2) on the raw priors.
    use_arm=True : ODM loss  = match against per-image refined anchors, anchors whose ARM
                   objectness <= theta leave the positives and the mining pool, then a6."""
    B, P = arm_loc.size(0), arm_loc.size(1)
    priors = priors[:P]
    if not use_arm:
        bt = [torch.cat([t[:, :4], torch.zeros_like(t[:, 4:5])], 1) for t in targets]
        return multibox_loss(arm_loc, arm_conf, priors, bt, 2, threshold, negpos_ratio, variances,
                             stable=stable, detail=detail)
    # `anchors` = (xyxy, centre form) and `pool` override the refinement / the ARM filter (tests feed the same
    # refined anchors to the CUDA path and to this function, so that only the ODM logic is compared)
    xy, cf = anchors if anchors is not None else refine_anchors(arm_loc.detach(), priors, variances)
    loc_t = torch.zeros(B, P, 4)
    conf_t = torch.zeros(B, P, dtype=torch.int64)
    for b in range(B):
        t = targets[b]
        m = match_image(threshold, t[:, :4], cf[b], variances, t[:, 4], anchors_xyxy=xy[b])
        loc_t[b] = m["loc"]
        conf_t[b] = m["conf"]
    if pool is None:
        pool = arm_objectness(arm_conf.detach()) > theta
    pos = (conf_t > 0) & pool
    sel4 = pos.unsqueeze(2).expand_as(odm_loc)
    loss_l = F.smooth_l1_loss(odm_loc[sel4].view(-1, 4), loc_t[sel4].view(-1, 4), reduction="sum")
    flat = odm_conf.view(-1, num_classes)
    conf_eff = torch.where(pos, conf_t, torch.zeros_like(conf_t))
    keys = (log_sum_exp(flat) - flat.gather(1, conf_eff.view(-1, 1))).view(B, P)
    neg = hard_negative_select(keys, pos, negpos_ratio, stable=stable, pool=pool)
    chosen = pos | neg
    loss_c = F.cross_entropy(odm_conf[chosen.unsqueeze(2).expand_as(odm_conf)].view(-1, num_classes),
                             conf_eff[chosen], reduction="sum")
    n = pos.long().sum()
    if detail:
        mk = keys.detach().clone()
        mk[pos] = 0
        return dict(loss_l=loss_l / n, loss_c=loss_c / n, sum_l=loss_l, sum_c=loss_c, n=n,
                    loc_t=loc_t, conf_t=conf_eff, conf_t_raw=conf_t, pos=pos, neg=neg, pool=pool, mining_keys=mk)
    return loss_l / n, loss_c / n


def refine_detect(arm_loc, arm_conf, odm_loc, odm_scores, priors, num_classes, top_k=200,
                  conf_thresh=0.01, nms_thresh=0.45, variances=(0.1, 0.2), theta=0.01, stable=True):
    """decode(odm_loc, refined anchors); scores of anchors with ARM objectness <= theta are zeroed."""
    _, cf = refine_anchors(arm_loc, priors, variances)
    keep = arm_objectness(arm_conf) > theta
    return detect(odm_loc, odm_scores, priors, num_classes, top_k, conf_thresh, nms_thresh,
                  variances, stable=stable, score_mask=keep, anchors_center=cf)


# --------------------------------------------------------------------------------------
# 8f rank 1: eval post-processing after Detect (lib/utils/evaluate_utils.py)
# --------------------------------------------------------------------------------------
def rescale_detections(det, extra):
    """evaluate_utils.py:63-68 on a copy: x *= w, y *= h with extra[:,0]=h, extra[:,1]=w."""
    det = det.clone()
    h = extra[:, 0].unsqueeze(-1).unsqueeze(-1)
    w = extra[:, 1].unsqueeze(-1).unsqueeze(-1)
    det[:, :, :, 1] *= w
    det[:, :, :, 3] *= w
    det[:, :, :, 2] *= h
    det[:, :, :, 4] *= h
    return det


def convert_ssd_result(det, coco_ids=None):
    """EvalVOC.convert_ssd_result (:127-139) / EvalCOCO.convert_ssd_result (:175-190): append the image
    index, the class index (and the coco id), keep rows with score > 0, reorder the columns to
    xmin, ymin, xmax, ymax, score, image, cls(, cocoid)."""
    B, C, K = det.shape[:3]
    idx = torch.arange(0, B).view(B, 1, 1, 1).expand(B, C, K, 1).to(det.dtype)
    cls = torch.arange(0, C).view(1, C, 1, 1).expand(B, C, K, 1).to(det.dtype)
    cols = [det, idx, cls]
    if coco_ids is not None:
        cid = torch.Tensor(list(coco_ids)).view(B, 1, 1, 1).expand(B, C, K, 1)
        cols.append(cid)
    full = torch.cat(cols, 3)
    n = full.size(3)
    mask = full[:, :, :, 0].gt(0.).unsqueeze(-1).expand(full.size())
    flat = torch.masked_select(full, mask).view(-1, n)
    order = [1, 2, 3, 4, 0, 5, 6] + ([7] if coco_ids is not None else [])
    return flat[:, order]


def coco_post_proc(rows):
    """EvalCOCO.post_proc (:193-199): x2,y2 -> w,h then cocoid, x1, y1, w, h, score, cls."""
    rows = rows.clone()
    rows[:, 2] -= rows[:, 0]
    rows[:, 3] -= rows[:, 1]
    return rows[:, [7, 0, 1, 2, 3, 4, 6]]


# --------------------------------------------------------------------------------------
# 8f rank 3: head-output layout (lib/models/ssd_v3.py:114-121, rfb_net.py:213-220)
# --------------------------------------------------------------------------------------
def heads_to_rows(outputs, k):
    """outputs: list of multibox head outputs [B, A*k, H, W] -> [B, P, k]:
    permute(0,2,3,1).contiguous() per layer (:115-116), view(B,-1) + cat(dim 1) (:118-119), view (:120-121)."""
    rows = [o.permute(0, 2, 3, 1).contiguous() for o in outputs]
    flat = torch.cat([o.view(o.size(0), -1) for o in rows], 1)
    return flat.view(flat.size(0), -1, k)
