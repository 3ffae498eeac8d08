"""PASCAL VOC evaluation on the GPU -- host mirror of lib/datasets/voc_eval.py (evaluate_detections
:308-311, write_voc_results_file :58-75, do_python_eval :244-262, voc_eval :109-242, voc_ap :78-106)
and of the crop-sampling IoU of lib/utils/augmentations.py (:13-37, :250-268), backed by
ssdbox_voc_eval / ssdbox_crop_overlaps.

The reference accumulates `results[cls][img] = rows[:, 0:5]` batch by batch (EvalVOC.post_proc,
lib/utils/evaluate_utils.py:141-151), prints them to one text file per class and re-reads the files
class by class.  Here the flat rows of ssdbox.evaluate_utils.convert_ssd_result stay on the device:

    acc = VOCDetections(num_classes)
    for images, targets, extra in loader:                        # evaluate_utils.py:52-72
        rows, seg = convert_ssd_result(detector(loc, conf, priors), extra)
        acc.add(rows, seg)
    gt = VOCGroundTruth.from_recs(recs, imagenames, VOC_CLASSES)  # recs: the parse_rec dicts (:15-33)
    res, mAP = evaluate_detections(acc, gt, VOC_CLASSES)          # [(cls, ap, prec, rec)], mean AP

Equal quantised scores are ordered canonically (file order), see voceval.cu.  No CPU path.
"""
import numpy as np
import torch

from . import _abi

_ws = {}


class VOCGroundTruth(object):
    """The truths of an image set on the device: boxes fp32 [M,4], labels int32 [M] (1-based class
    index, 0 = background), difficult uint8 [M], offsets int32 [num_images+1]."""

    def __init__(self, boxes, labels, difficult, offsets, device):
        dev = torch.device(device)
        self.boxes = torch.as_tensor(np.asarray(boxes, dtype=np.float32).reshape(-1, 4)).to(dev).contiguous()
        self.labels = torch.as_tensor(np.asarray(labels, dtype=np.int32)).to(dev).contiguous()
        self.difficult = torch.as_tensor(np.asarray(difficult, dtype=np.uint8)).to(dev).contiguous()
        self.offsets = torch.as_tensor(np.asarray(offsets, dtype=np.int32)).to(dev).contiguous()
        if self.offsets.numel() < 1 or self.labels.numel() != self.boxes.size(0) or \
                self.difficult.numel() != self.boxes.size(0):
            raise ValueError("inconsistent ground-truth arrays")
        self.num_images = self.offsets.numel() - 1

    @classmethod
    def from_recs(cls, recs, imagenames, classes, device="cuda"):
        """recs[name] = list of {'name', 'difficult', 'bbox': [xmin, ymin, xmax, ymax]} as parse_rec
        (:15-33) returns them / annots.pkl caches them (:143-153); classes = the labelmap (class c of
        the detection rows is classes[c-1])."""
        index = {n: i + 1 for i, n in enumerate(classes)}
        boxes, labels, diff, offs = [], [], [], [0]
        for name in imagenames:
            for obj in recs[name]:
                if obj["name"] not in index:
                    continue
                boxes.append([float(v) for v in obj["bbox"]])
                labels.append(index[obj["name"]])
                diff.append(1 if obj["difficult"] else 0)
            offs.append(len(labels))
        return cls(boxes, labels, diff, offs, device)


class VOCDetections(object):
    """Accumulates convert_ssd_result outputs over the batches of an evaluation run
    (the role of EvalVOC.results, evaluate_utils.py:122-124,141-151)."""

    def __init__(self, num_classes):
        self.num_classes = int(num_classes)
        self._rows, self._seg = [], []
        self._nrows = 0

    def add(self, rows, seg):
        """rows [n, 7|8] and seg int32 [B*C+1] of one batch (ssdbox.evaluate_utils.convert_ssd_result)."""
        if seg.numel() < 1 or (seg.numel() - 1) % self.num_classes:
            raise ValueError("seg must hold B*num_classes+1 offsets")
        self._rows.append(rows)
        self._seg.append(seg[:-1].to(torch.int32) + self._nrows)
        self._nrows += int(rows.size(0))

    def flat(self):
        if not self._rows:
            raise ValueError("no detections were added")
        dev = self._rows[0].device
        rows = torch.cat(self._rows, 0).contiguous()
        seg = torch.cat(self._seg + [torch.tensor([self._nrows], dtype=torch.int32, device=dev)]).contiguous()
        return rows, seg


class VOCEvalResult(object):
    """Device outputs of ssdbox_voc_eval.  rec(c) / prec(c) are views of the class's sorted range."""

    def __init__(self, order, cls_offsets, tpfp, rec, prec, ap, npos):
        self.order, self.tpfp, self._rec, self._prec = order, tpfp, rec, prec
        self.cls_offsets, self.ap, self.npos = cls_offsets, ap, npos      # host numpy arrays

    def _range(self, c):
        return int(self.cls_offsets[c]), int(self.cls_offsets[c + 1])

    def rec(self, c):
        a, b = self._range(c)
        return self._rec[a:b]

    def prec(self, c):
        a, b = self._range(c)
        return self._prec[a:b]

    def rows_of(self, c):
        a, b = self._range(c)
        return self.order[a:b]

    @property
    def mean_ap(self):
        return float(np.mean(self.ap[1:])) if self.ap.size > 1 else float("nan")      # voc_eval.py:262


def voc_eval(rows, seg, gt, num_classes, ovthresh=0.5, use_07_metric=True):
    """voc_eval (:109-242) for every class at once.  rows [N, >=5] fp32 CUDA (xmin, ymin, xmax, ymax,
    score, ...), seg int32 [num_images*num_classes+1], gt a VOCGroundTruth."""
    if not rows.is_cuda:
        raise RuntimeError("ssdbox: voc_eval runs on CUDA tensors only (no CPU path)")
    dev = rows.device
    rows = _abi.as_f32(rows)
    if rows.dim() != 2 or rows.size(1) < 5:
        raise ValueError("rows must be [N, >=5]")
    C, I, N, M = int(num_classes), gt.num_images, rows.size(0), gt.boxes.size(0)
    seg = seg.to(dev, torch.int32).contiguous()
    if seg.numel() != I * C + 1:
        raise ValueError("seg must hold num_images*num_classes+1 offsets (%d), got %d" % (I * C + 1, seg.numel()))
    if gt.boxes.device != dev:
        raise ValueError("ground truth lives on %s, rows on %s" % (gt.boxes.device, dev))
    order = torch.empty(N, dtype=torch.int32, device=dev)
    tpfp = torch.empty(N, dtype=torch.uint8, device=dev)
    rec = torch.empty(N, dtype=torch.float64, device=dev)
    prec = torch.empty(N, dtype=torch.float64, device=dev)
    # the small outputs share one buffer so that the host needs a single D2H copy:
    # ap fp64 [C] | cls_offsets int32 [C+1] | npos int32 [C] | status int32 [1]
    small = torch.empty(8 * C + 4 * (2 * C + 2), dtype=torch.uint8, device=dev)
    ap = small[:8 * C].view(torch.float64)
    ints = small[8 * C:].view(torch.int32)
    cls_offsets, npos, status = ints[:C + 1], ints[C + 1:2 * C + 1], ints[2 * C + 1:]
    cfg = _abi.VocEvalCfg(I, C, N, rows.size(1), M, 1 if use_07_metric else 0, float(ovthresh))
    ws, n = _ws.setdefault(dev, _abi.Workspace()).get(_abi.workspace_bytes(_abi.OP_VOC_EVAL, 0, N, C, M), dev)
    _abi.check(_abi.lib().ssdbox_voc_eval(
        cfg, _abi.ptr(rows, torch.float32, "rows"), _abi.ptr(seg, torch.int32, "seg"),
        _abi.ptr(gt.boxes, torch.float32, "gt boxes"), _abi.ptr(gt.labels, torch.int32, "gt labels"),
        _abi.ptr(gt.difficult, torch.uint8, "gt difficult"), _abi.ptr(gt.offsets, torch.int32, "gt offsets"),
        _abi.ptr(order), _abi.ptr(cls_offsets), _abi.ptr(tpfp), _abi.ptr(rec), _abi.ptr(prec), _abi.ptr(ap),
        _abi.ptr(npos), _abi.ptr(status), ws, n, _abi.stream_ptr(dev)))
    host = small.cpu()
    h_ints = host[8 * C:].view(torch.int32).numpy()
    bad = int(h_ints[2 * C + 1])
    if bad & (1 << 30):
        raise ValueError("voc_eval: seg does not cover the %d rows exactly (seg[-1] must equal the row count)" % N)
    if bad:
        raise ValueError("voc_eval: %d detection scores fall outside [0, 1] after '%%.3f' rounding" % bad)
    return VOCEvalResult(order, h_ints[:C + 1].copy(), tpfp, rec, prec, host[:8 * C].view(torch.float64).numpy().copy(),
                         h_ints[C + 1:2 * C + 1].copy())


def evaluate_detections(detections, gt, classes, ovthresh=0.5, use_07_metric=True):
    """evaluate_detections / do_python_eval (:244-262, :308-311): returns ([(cls, ap, prec, rec)], mAP)
    with prec / rec as numpy float64 arrays (-1. for a class without detections, :238-241)."""
    rows, seg = detections.flat() if isinstance(detections, VOCDetections) else detections
    r = voc_eval(rows, seg, gt, len(classes) + 1, ovthresh, use_07_metric)
    res = []
    for c, name in enumerate(classes, start=1):
        a, b = r._range(c)
        if b == a:
            res.append((name, -1., -1., -1.))
        else:
            res.append((name, float(r.ap[c]), r.prec(c).cpu().numpy(), r.rec(c).cpu().numpy()))
    return res, float(np.mean([x[1] for x in res]))


def crop_overlaps(boxes, rects, want_overlap=True, want_mask=True):
    """jaccard_numpy + the centre test of RandomSampleCrop (augmentations.py:13-37, 250-268) for a
    batch: boxes = list of B float64 arrays [G_b,4] (absolute xyxy), rects int64 [B,T,4].
    Returns (overlap list of [T,G_b] | None, minmax [B,T,2], mask list of [T,G_b] bool | None)."""
    if not isinstance(rects, torch.Tensor) or not rects.is_cuda:
        raise RuntimeError("ssdbox: crop_overlaps runs on CUDA tensors only (no CPU path)")
    dev = rects.device
    rects = rects.to(torch.int64).contiguous()
    B, T = rects.size(0), rects.size(1)
    if len(boxes) != B:
        raise ValueError("need one box array per image")
    offs = np.zeros(B + 1, dtype=np.int32)
    offs[1:] = np.cumsum([len(b) for b in boxes])
    flat = [torch.as_tensor(np.asarray(b, dtype=np.float64).reshape(-1, 4)) for b in boxes]
    flat = torch.cat(flat, 0).to(dev).contiguous() if flat else torch.zeros(0, 4, dtype=torch.float64, device=dev)
    off_d = torch.as_tensor(offs).to(dev)
    total = int(offs[-1]) * T
    ov = torch.empty(total, dtype=torch.float64, device=dev) if want_overlap else None
    mk = torch.empty(total, dtype=torch.uint8, device=dev) if want_mask else None
    mm = torch.empty(B, T, 2, dtype=torch.float64, device=dev)
    _abi.check(_abi.lib().ssdbox_crop_overlaps(
        _abi.ptr(flat, torch.float64, "boxes"), _abi.ptr(off_d, torch.int32, "offsets"), _abi.ptr(rects, torch.int64, "rects"),
        B, T, _abi.ptr(ov, None, "overlap", True), _abi.ptr(mm), _abi.ptr(mk, None, "mask", True), _abi.stream_ptr(dev)))

    def split(t, cast=None):
        if t is None:
            return None
        out = []
        for b in range(B):
            g = int(offs[b + 1] - offs[b])
            v = t[int(offs[b]) * T:int(offs[b + 1]) * T].view(T, g)
            out.append(v.bool() if cast else v)
        return out

    return split(ov), mm, split(mk, True)
