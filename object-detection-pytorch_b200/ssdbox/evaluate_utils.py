"""Eval post-processing right after DetectOut -- host mirror of lib/utils/evaluate_utils.py:63-70
(rescale by the image size), EvalVOC / EvalCOCO.convert_ssd_result (:127-139, :175-190) and
EvalCOCO.post_proc (:193-203), backed by ssdbox_detections_compact.

    rows, seg = convert_ssd_result(detections, extra)                      # VOC: [n,7]
    rows, seg = convert_ssd_result(detections, extra, coco_ids=ids)        # COCO: [n,8]
    rows, seg = coco_result_rows(detections, extra, coco_ids=ids)          # COCO results: [n,7]

`detections` is DetectOut's [B,C,top_k,5] tensor (left untouched: the reference scales it in
place), `extra` [B,2] = (h, w) per image as the data loader yields it.  Rows come out in the order
the reference's masked_select produces: (image, class, k).  `seg` int32 [B*C+1] holds the first row
of every (image, class) segment, i.e. the slices EvalVOC.post_proc (:141-151) cuts with numpy masks:
results[cls][img_idx + b] = rows[seg[b*C+cls] : seg[b*C+cls+1], 0:5]."""
import torch

from . import _abi

_ws = {}


def _compact(detections, extra, image_ids, mode, capacity=None):
    if not detections.is_cuda:
        raise RuntimeError("ssdbox: convert_ssd_result runs on CUDA tensors only (no CPU path)")
    dev = detections.device
    det = _abi.as_f32(detections)
    if det.dim() != 4 or det.size(3) != 5:
        raise ValueError("detections must be [B, C, top_k, 5]")
    B, Cn, K = det.size(0), det.size(1), det.size(2)
    ex = None
    if extra is not None:
        ex = _abi.as_f32(extra, dev).reshape(B, -1)[:, :2].contiguous()
    ids = None
    if image_ids is not None:
        ids = torch.as_tensor(image_ids, dtype=torch.float32).to(dev).reshape(-1).contiguous()   # torch.Tensor(ids): fp32
        if ids.numel() != B:
            raise ValueError("need one image id per image")
    ncol = 8 if mode == 1 else 7
    cap = int(capacity) if capacity is not None else B * Cn * K
    out = torch.empty(cap, ncol, dtype=torch.float32, device=dev)
    total = torch.empty(1, dtype=torch.int32, device=dev)
    seg = torch.empty(B * Cn + 1, dtype=torch.int32, device=dev)
    ws_owner = _ws.setdefault(dev, _abi.Workspace())
    ws, n = ws_owner.get(_abi.workspace_bytes(_abi.OP_COMPACT, B, 0, Cn), dev)
    _abi.check(_abi.lib().ssdbox_detections_compact(
        _abi.ptr(det, torch.float32, "detections"), B, Cn, K, _abi.ptr(ex, torch.float32, "extra", True),
        _abi.ptr(ids, torch.float32, "image_ids", True), mode, _abi.ptr(out), cap, _abi.ptr(total), _abi.ptr(seg),
        ws, n, _abi.stream_ptr(dev)))
    if capacity is not None:
        seg.clamp_(max=cap)        # rows beyond the capacity were dropped: the segments end with the buffer
    return out, total, seg


def convert_ssd_result(detections, extra=None, coco_ids=None, capacity=None, sync=True):
    """EvalVOC.convert_ssd_result (coco_ids None) / EvalCOCO.convert_ssd_result, including the rescale
    of evaluate_utils.py:63-68.  sync=True slices the rows to their count (one 4-byte D2H read, the
    reference's masked_select synchronises as well); sync=False returns (buffer, count tensor, seg)."""
    out, total, seg = _compact(detections, extra, coco_ids, 0 if coco_ids is None else 1, capacity)
    if not sync:
        return out, total, seg
    return out[:min(int(total.item()), out.size(0))], seg


def coco_result_rows(detections, extra, coco_ids, capacity=None, sync=True):
    """convert_ssd_result followed by EvalCOCO.post_proc (:193-199): [cocoid, x1, y1, w, h, score, cls]."""
    out, total, seg = _compact(detections, extra, coco_ids, 2, capacity)
    if not sync:
        return out, total, seg
    return out[:min(int(total.item()), out.size(0))], seg


# --------------------------------------------------------------------------------------------------
# The evaluation solvers -- host mirror of EvalBase / EvalVOC / EvalCOCO (lib/utils/evaluate_utils.py:
# 14-78, 114-162, 165-222): same constructor and validate(net, priors) contract, so that
# eval_solver_factory (lib/utils/__init__.py:8-11) can hand them out (ssdbox.compat.install(patch_eval=True)).
# Everything between the network and the metric stays on the device: DetectOut -> result rows ->
# accumulation -> ssdbox_voc_eval.  Nothing is pickled or written to text files unless asked.
# --------------------------------------------------------------------------------------------------
# the PASCAL VOC label map in its canonical order (class c of the detection rows is VOC_CLASSES[c-1])
VOC_CLASSES = ('aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair', 'cow', 'diningtable',
               'dog', 'horse', 'motorbike', 'person', 'pottedplant', 'sheep', 'sofa', 'train', 'tvmonitor')


def parse_rec(filename):
    """One PASCAL VOC annotation file -> [{'name', 'pose', 'truncated', 'difficult', 'bbox'}] with
    bbox = xml value - 1, as voc_eval.py:15-33 returns it."""
    import xml.etree.ElementTree as ET
    objects = []
    for obj in ET.parse(filename).findall('object'):
        bb = obj.find('bndbox')
        text = lambda tag, d=None: (obj.find(tag).text if obj.find(tag) is not None else d)
        objects.append({'name': obj.find('name').text, 'pose': text('pose', 'Unspecified'),
                        'truncated': int(text('truncated', 0)), 'difficult': int(text('difficult', 0)),
                        'bbox': [int(bb.find(k).text) - 1 for k in ('xmin', 'ymin', 'xmax', 'ymax')]})
    return objects


class EvalBase(object):
    """EvalBase (:14-78).  `detector` defaults to DetectOut(C, 0, 200, 0.01, 0.45, cfg.MODEL.VARIANCE) (:16-17);
    pass conf_is_logits=True when the network returns raw class scores (softmax fused into Detect)."""

    def __init__(self, data_loader, cfg, detector=None, conf_is_logits=False):
        from .detection import DetectOut
        self.detector = detector or DetectOut(cfg.MODEL.NUM_CLASSES, 0, 200, 0.01, 0.45, cfg.MODEL.VARIANCE,
                                              conf_is_logits=conf_is_logits)
        self.data_loader = data_loader
        self.dataset = data_loader.dataset
        self.name = getattr(self.dataset, 'name', None)
        self.cfg = cfg
        self.results = None
        self.image_sets = getattr(self.dataset, 'image_sets', None)

    def reset_results(self):
        raise NotImplementedError

    def consume(self, detections, extra, img_idx):
        """rescale + convert_ssd_result + post_proc (:62-72) of one batch; returns the next img_idx"""
        raise NotImplementedError

    def evaluate_stats(self, classes=None, tb_writer=None):
        raise NotImplementedError

    def validate(self, net, priors, use_cuda=True, tb_writer=None):
        """:41-78.  `net(images, phase='eval')` returns (loc, conf) like the reference's models."""
        if not use_cuda:
            raise RuntimeError("ssdbox: the evaluation path runs on CUDA only (no CPU path)")
        self.reset_results()
        img_idx = 0
        dev = priors.device if priors.is_cuda else torch.device("cuda")
        priors = priors.to(dev)
        with torch.no_grad():
            for images, targets, extra in self.data_loader:
                loc, conf = net(images.to(dev, non_blocking=True), phase='eval')
                detections = self.detector(loc, conf, priors)
                img_idx = self.consume(detections, torch.as_tensor(extra).to(dev), img_idx)
        return self.evaluate_stats(None, tb_writer)


class EvalVOC(EvalBase):
    """EvalVOC (:114-162): returns (res, [mAP]) with res = [(cls, ap, prec, rec)] like do_python_eval
    (voc_eval.py:244-262).  The truths come from the dataset's annotation files (dataset.ids +
    dataset._annopath, parsed like voc_eval.py:15-33) or from `recs` / `ground_truth` if given."""

    def __init__(self, data_loader, cfg, detector=None, conf_is_logits=False, classes=None, recs=None,
                 use_07_metric=True, ovthresh=0.5, output_dir=None):
        super(EvalVOC, self).__init__(data_loader, cfg, detector, conf_is_logits)
        if getattr(cfg, 'DATASET', None) is not None and getattr(cfg.DATASET, 'NUM_EVAL_PICS', 0) > 0:
            raise Exception("not support voc")                                   # :118-119
        self.classes = list(classes) if classes is not None else None
        self.recs, self.use_07_metric, self.ovthresh, self.output_dir = recs, use_07_metric, ovthresh, output_dir
        self.ground_truth = None

    def reset_results(self):
        from .voc_eval import VOCDetections
        self.results = VOCDetections(self.cfg.MODEL.NUM_CLASSES)

    def consume(self, detections, extra, img_idx):
        rows, seg = convert_ssd_result(detections, extra)
        self.results.add(rows, seg)
        return img_idx + detections.size(0)

    def _image_names(self):
        return [i[1] if isinstance(i, (tuple, list)) else i for i in self.dataset.ids]

    def _ground_truth(self, device):
        from .voc_eval import VOCGroundTruth
        if self.ground_truth is None:
            names = self._image_names()
            recs = self.recs
            if recs is None:
                recs = {n: parse_rec(self.dataset._annopath % tuple(i) if isinstance(i, (tuple, list)) else self.dataset._annopath % i)
                        for n, i in zip(names, self.dataset.ids)}
            self.ground_truth = VOCGroundTruth.from_recs(recs, names, self._classes(), device)
        return self.ground_truth

    def _classes(self):
        if self.classes is None:
            if self.cfg.MODEL.NUM_CLASSES != len(VOC_CLASSES) + 1:
                raise ValueError("EvalVOC: %d classes are not the PASCAL VOC label map; pass classes=[...]" % self.cfg.MODEL.NUM_CLASSES)
            self.classes = list(VOC_CLASSES)
        return self.classes

    def evaluate_stats(self, classes=None, tb_writer=None):
        from .voc_eval import evaluate_detections
        rows, seg = self.results.flat()
        res, mean_ap = evaluate_detections((rows, seg), self._ground_truth(rows.device), self._classes(),
                                           self.ovthresh, self.use_07_metric)
        if self.output_dir is not None:                                           # voc_eval.py:259-260
            import os
            import pickle
            os.makedirs(self.output_dir, exist_ok=True)
            for cls, ap, prec, rec in res:
                with open(os.path.join(self.output_dir, cls + '_pr.pkl'), 'wb') as f:
                    pickle.dump({'rec': rec, 'prec': prec, 'ap': ap}, f)
        return res, [mean_ap]


class EvalCOCO(EvalBase):
    """EvalCOCO (:165-222): accumulates the COCO result rows [cocoid, x1, y1, w, h, score, cls] on the
    device; evaluate_stats hands them to pycocotools exactly like :206-222 (needs pycocotools)."""

    def __init__(self, data_loader, cfg, detector=None, conf_is_logits=False):
        super(EvalCOCO, self).__init__(data_loader, cfg, detector, conf_is_logits)
        n = getattr(cfg.DATASET, 'NUM_EVAL_PICS', 0) if getattr(cfg, 'DATASET', None) is not None else 0
        if n > 0:
            self.dataset.ids = self.dataset.ids[:n]                               # :168-169

    def reset_results(self):
        self.results = []

    def consume(self, detections, extra, img_idx):
        B = detections.size(0)
        rows, _ = coco_result_rows(detections, extra, self.dataset.ids[img_idx:img_idx + B])
        self.results.append(rows)
        return img_idx + B

    def result_rows(self):
        return torch.cat(self.results, 0) if self.results else torch.zeros(0, 7)

    def evaluate_stats(self, classes=None, tb_writer=None):
        import numpy as np
        from pycocotools.cocoeval import COCOeval
        res = self.result_rows().cpu().numpy()
        for r in res:
            r[6] = self.dataset.target_transform.inver_map[r[6]]
        coco = self.dataset.cocos[0]['coco']
        coco_pred = coco.loadRes(res)
        ev = COCOeval(coco, coco_pred, 'bbox')
        ev.params.imgIds = self.dataset.ids
        ev.evaluate()
        ev.accumulate()
        ev.summarize()
        res = ev.eval
        ap05 = res['precision'][0, :, :, 0, 2]
        ap95 = res['precision'][:, :, :, 0, 2]
        return res, [np.mean(ap05[ap05 > -1]), np.mean(ap95[ap95 > -1])]
