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
