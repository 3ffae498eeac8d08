"""TEST INFRASTRUCTURE ONLY -- CPU oracle (numpy, float64) for SURVEY.md 8f rank 4: the PASCAL VOC
detection-to-truth matching / AP of lib/datasets/voc_eval.py and the crop-sampling IoU of
lib/utils/augmentations.py.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import
this module; the product never does.

PARITY STATUS: PINNED.  tests/test_oracle_vs_reference.py runs the reference's own voc_eval() /
voc_ap() / jaccard_numpy() (imported from /root/reference) on result files and an annotation cache
written to a temporary directory and proves this restatement bit-identical (rec, prec, ap);
oracle/make_golden.py records tests/golden/voceval.npz from the REFERENCE for the GPU box.

What the reference does, in order (citations relative to the reference root):
  1. EvalVOC.post_proc (lib/utils/evaluate_utils.py:141-151) stores, per (class, image), the rows
     [xmin, ymin, xmax, ymax, score] of that image as float32.
  2. write_voc_results_file (lib/datasets/voc_eval.py:58-75) prints one line per detection into a
     per-class TEXT file: score as '{:.3f}', every coordinate + 1 (float32 add) as '{:.1f}'.
     The text round trip QUANTISES: voc_eval sees score = k/1000 and coordinates = m/10.
  3. voc_eval (:109-242) parses the file back to float64, sorts all detections of the class by
     -confidence (np.argsort, not stable), walks them in that order: IoU in float64 against the
     truths of the detection's image and class (:190-205), true positive when ovmax > ovthresh and the
     arg-max truth is neither difficult nor already claimed (:207-215), cumulative sums,
     rec = tp / npos, prec = tp / max(tp + fp, eps) (:218-223), voc_ap (:78-106).

Ties: quantised scores tie all the time and np.argsort's default kind is not stable, so the literal
reference order among equal scores is an accident of numpy's introsort.  `stable=True` (default) is
the canonical order the CUDA path implements -- equal scores keep their file order, i.e. (image, row)
order; `stable=False` makes the literal np.argsort(-confidence) call and is what is compared with the
reference (same numpy => same permutation).
"""
import numpy as np


# --------------------------------------------------------------------------------------
# text round trip                                      lib/datasets/voc_eval.py:58-75, :170-175
# --------------------------------------------------------------------------------------
def result_line(image_name, row):
    """One line of a per-class results file (:70-74).  `row` = float32 [xmin,ymin,xmax,ymax,score]."""
    row = np.asarray(row, dtype=np.float32)
    return '{:s} {:.3f} {:.1f} {:.1f} {:.1f} {:.1f}\n'.format(
        image_name, row[-1], row[0] + 1, row[1] + 1, row[2] + 1, row[3] + 1)


def parse_lines(lines):
    """:170-175 -> (image names, confidence float64 [n], BB float64 [n,4])"""
    split = [x.strip().split(' ') for x in lines]
    ids = [x[0] for x in split]
    conf = np.array([float(x[1]) for x in split])
    bb = np.array([[float(z) for z in x[2:]] for x in split]).reshape(-1, 4)
    return ids, conf, bb


def quantise_rows(rows):
    """What the text round trip does to float32 rows [n, >=5], without going through text:
    score -> float('%.3f' % score), coordinate -> float('%.1f' % (float32(coordinate) + 1)).
    (Kept literal -- the arithmetic shortcut rint(x*1000)/1000 is what the CUDA kernel uses and
    the tests compare the two.)"""
    rows = np.asarray(rows, dtype=np.float32)
    conf = np.array([float('{:.3f}'.format(s)) for s in rows[:, 4]], dtype=np.float64)
    bb = np.array([[float('{:.1f}'.format(v + 1)) for v in r[:4]] for r in rows], dtype=np.float64).reshape(-1, 4)
    return conf, bb


# --------------------------------------------------------------------------------------
# voc_ap                                                       lib/datasets/voc_eval.py:78-106
# --------------------------------------------------------------------------------------
def voc_ap(rec, prec, use_07_metric=True):
    """11-point interpolated AP (:85-93: thresholds np.arange(0, 1.1, 0.1), p/11 accumulated in that
    order) or the area under the precision envelope (:95-105)."""
    if use_07_metric:
        total = 0.
        for t in np.arange(0., 1.1, 0.1):
            sel = rec >= t
            best = np.max(prec[sel]) if sel.any() else 0
            total = total + best / 11.
        return total
    r = np.concatenate(([0.], rec, [1.]))
    p = np.concatenate(([0.], prec, [0.]))
    p = np.maximum.accumulate(p[::-1])[::-1]        # the backward running max of :97-98 (max is exact)
    step = np.where(r[1:] != r[:-1])[0]
    return np.sum((r[step + 1] - r[step]) * p[step + 1])


# --------------------------------------------------------------------------------------
# voc_eval for one class                                       lib/datasets/voc_eval.py:155-242
# --------------------------------------------------------------------------------------
def _overlaps(box, truths):
    """:191-203 -- float64, operation order kept: union = (area(box) + area(truth)) - inter."""
    x_lo = np.maximum(truths[:, 0], box[0])
    y_lo = np.maximum(truths[:, 1], box[1])
    x_hi = np.minimum(truths[:, 2], box[2])
    y_hi = np.minimum(truths[:, 3], box[3])
    inter = np.maximum(x_hi - x_lo, 0.) * np.maximum(y_hi - y_lo, 0.)
    union = ((box[2] - box[0]) * (box[3] - box[1]) +
             (truths[:, 2] - truths[:, 0]) * (truths[:, 3] - truths[:, 1]) - inter)
    with np.errstate(divide='ignore', invalid='ignore'):
        return inter / union


def voc_eval_class(det_image, conf, bb, gt_boxes, gt_difficult, ovthresh=0.5, use_07_metric=True,
                   stable=True):
    """det_image int [n]: image index of every detection (file order); conf / bb as parsed from the
    file; gt_boxes[i] float [G_i,4] / gt_difficult[i] bool [G_i]: the truths OF THIS CLASS in image i.
    Returns dict(rec, prec, ap, tp, fp, order, npos); rec = prec = ap = -1.0 when there is no
    detection (:238-241)."""
    npos = int(sum(int(np.sum(~np.asarray(d, dtype=bool))) for d in gt_difficult))      # :163
    n = len(det_image)
    if n == 0:
        return dict(rec=-1., prec=-1., ap=-1., tp=np.zeros(0), fp=np.zeros(0), order=np.zeros(0, np.int64), npos=npos)
    conf = np.asarray(conf, dtype=np.float64)
    order = np.argsort(-conf, kind='stable') if stable else np.argsort(-conf)       # :178
    bb = np.asarray(bb, dtype=np.float64)[order, :]
    img = [int(det_image[x]) for x in order]
    claimed = [[False] * len(g) for g in gt_boxes]
    tp = np.zeros(n)
    fp = np.zeros(n)
    for d in range(n):                                                                 # :187-216
        i = img[d]
        best, j = -np.inf, -1
        truths = np.asarray(gt_boxes[i]).astype(float).reshape(-1, 4)
        if truths.size > 0:
            ov = _overlaps(bb[d, :].astype(float), truths)
            best, j = np.max(ov), int(np.argmax(ov))       # first index on ties; NaN propagates (:204-205)
        if not best > ovthresh:
            fp[d] = 1.
        elif not gt_difficult[i][j]:                        # a difficult truth: neither tp nor fp (:208)
            if claimed[i][j]:
                fp[d] = 1.
            else:
                tp[d] = 1.
                claimed[i][j] = True
    fpc = np.cumsum(fp)                                                                # :218-223
    tpc = np.cumsum(tp)
    with np.errstate(divide='ignore', invalid='ignore'):
        rec = tpc / float(npos)
    prec = tpc / np.maximum(tpc + fpc, np.finfo(np.float64).eps)
    ap = voc_ap(rec, prec, use_07_metric)
    return dict(rec=rec, prec=prec, ap=float(ap), tp=tp, fp=fp, order=order, npos=npos)


def voc_eval_rows(rows, seg, num_images, num_classes, gt_boxes, gt_labels, gt_difficult, gt_offsets,
                  ovthresh=0.5, use_07_metric=True, stable=True):
    """The whole evaluate_detections chain (:308-311 -> :58-75 -> :244-262 -> :109-242) on the flat
    layout the CUDA path takes:
      rows  float32 [N, >=5] (xmin, ymin, xmax, ymax, score, ...) grouped by (image, class) segments,
      seg   int [num_images*num_classes + 1] first row of every segment (class 0 = background: empty),
      gt_boxes [M,4], gt_labels int [M] (1-based class = column `cls` of the rows), gt_difficult [M],
      gt_offsets int [num_images+1] (truths of image i are rows gt_offsets[i] .. gt_offsets[i+1]-1).
    Returns one voc_eval_class dict per class 1..num_classes-1 (+ 'rows': the row index of every
    detection of the class in file order) and the mean AP (:262)."""
    rows = np.asarray(rows, dtype=np.float32)
    seg = np.asarray(seg).astype(np.int64)
    gt_boxes = np.asarray(gt_boxes).reshape(-1, 4)
    gt_labels = np.asarray(gt_labels).astype(np.int64)
    gt_difficult = np.asarray(gt_difficult).astype(bool)
    out = []
    for c in range(1, num_classes):
        ridx = np.concatenate([np.arange(seg[i * num_classes + c], seg[i * num_classes + c + 1])
                               for i in range(num_images)] + [np.zeros(0, np.int64)]).astype(np.int64)
        dimg = np.concatenate([np.full(int(seg[i * num_classes + c + 1] - seg[i * num_classes + c]), i, dtype=np.int64)
                               for i in range(num_images)] + [np.zeros(0, np.int64)])
        conf, bb = quantise_rows(rows[ridx]) if len(ridx) else (np.zeros(0), np.zeros((0, 4)))
        gb, gd = [], []
        for i in range(num_images):
            sl = slice(int(gt_offsets[i]), int(gt_offsets[i + 1]))
            m = gt_labels[sl] == c
            gb.append(gt_boxes[sl][m])
            gd.append(gt_difficult[sl][m])
        r = voc_eval_class(dimg, conf, bb, gb, gd, ovthresh, use_07_metric, stable)
        r['rows'] = ridx
        out.append(r)
    return out, float(np.mean([r['ap'] for r in out]))                                 # :262


# --------------------------------------------------------------------------------------
# crop-sampling IoU                                        lib/utils/augmentations.py:13-37,250-268
# --------------------------------------------------------------------------------------
def jaccard_numpy(box_a, box_b):
    """:13-37 -- box_a [G,4], box_b [4]; arithmetic in the promoted dtype (float64 for the
    int64 `rect` of RandomSampleCrop :248)."""
    max_xy = np.minimum(box_a[:, 2:], box_b[2:])
    min_xy = np.maximum(box_a[:, :2], box_b[:2])
    inter = np.clip((max_xy - min_xy), a_min=0, a_max=np.inf)
    inter = inter[:, 0] * inter[:, 1]
    area_a = ((box_a[:, 2] - box_a[:, 0]) * (box_a[:, 3] - box_a[:, 1]))
    area_b = ((box_b[2] - box_b[0]) * (box_b[3] - box_b[1]))
    union = area_a + area_b - inter
    return inter / union


def crop_trial(boxes, rect):
    """The data-parallel part of one RandomSampleCrop trial (:250-268): overlap of every truth with
    the candidate rect, its min / max (the IoU constraint test :254), and the centre-in-rect mask
    (:261-268)."""
    overlap = jaccard_numpy(boxes, rect)
    centers = (boxes[:, :2] + boxes[:, 2:]) / 2.0
    m1 = (rect[0] < centers[:, 0]) * (rect[1] < centers[:, 1])
    m2 = (rect[2] > centers[:, 0]) * (rect[3] > centers[:, 1])
    return overlap, overlap.min(), overlap.max(), m1 * m2
