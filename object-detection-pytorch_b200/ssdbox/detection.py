"""DetectOut -- drop-in for lib/layers/functions/detection.py:6-64.  Same constructor; call it
as `detector(loc_data, conf_data, prior_data)` (the reference relies on the legacy
Function.__call__ -> forward dispatch, evaluate_utils.py:60).  Returns the [B, C, top_k, 5]
tensor of (score, x1, y1, x2, y2) rows in NMS order, zero padded, on the input's device.

`conf_is_logits=True` (extension): conf_data holds the raw class logits and the
softmax of ssd_v3.py:123-124 / rfb_net.py:222-226 is fused into the candidate pass."""
import ctypes as C

import torch

from . import _abi


class DetectOut(object):
    def __init__(self, num_classes, bkg_label, top_k, conf_thresh, nms_thresh, variance, conf_is_logits=False):
        self.num_classes = num_classes
        self.background_label = bkg_label
        self.top_k = top_k
        self.nms_thresh = nms_thresh
        if nms_thresh <= 0:                                   # detection.py:19-20
            raise ValueError('nms_threshold must be non negative.')
        self.conf_thresh = conf_thresh
        self.variance = variance
        self.conf_is_logits = bool(conf_is_logits)
        self._ws = _abi.Workspace()
        self.last_counts = None

    def forward(self, loc_data, conf_data, prior_data, score_keep=None, out=None, pending=None, refine=None):
        """`pending` (extension): a PendingLoss of MultiBoxLoss.forward_packed_deferred made on the same stream -- its
        cross-rank wait rides on the last Detect kernel (ssdbox_detect_peers) instead of a launch of its own; call
        pending.wait() afterwards as usual (it then only hands out the results)."""
        if not loc_data.is_cuda:
            raise RuntimeError("ssdbox: DetectOut runs on CUDA tensors only (no CPU path)")
        dev = loc_data.device
        num = loc_data.size(0)
        pri = _abi.as_f32(prior_data, dev)
        if pri.dim() == 3 and pri.size(0) == 1:               # the reference docstring's [1, num_priors, 4]
            pri = pri[0]
        per_image = pri.dim() == 3
        if per_image and pri.size(0) != num:
            raise ValueError("ssdbox: per-image priors have batch %d, loc_data has %d" % (pri.size(0), num))
        P = pri.size(-2)
        loc = _abi.as_f32(loc_data).view(num, P, 4)
        scores = _abi.as_f32(conf_data, dev).view(num, P, self.num_classes)      # detection.py:38
        logits = self.conf_is_logits
        if logits and int(self.top_k) > 1024:
            # the any-top_k path of the library has no fused softmax: do what the reference does (ssd_v3.py:123-124)
            scores = torch.softmax(scores, dim=-1)
            logits = False
        if out is None:
            out = torch.empty(num, self.num_classes, self.top_k, 5, dtype=torch.float32, device=dev)
        counts = torch.empty(num, self.num_classes, dtype=torch.int32, device=dev)
        ws, n, clean = self._ws.acquire(_abi.workspace_bytes(_abi.OP_DETECT, num, P, self.num_classes, 0, self.top_k), dev,
                                        ("detect", num, P, self.num_classes, int(self.top_k)))
        cfg = _abi.DetectCfg(num, P, self.num_classes, int(self.top_k), float(self.conf_thresh),
                             float(self.nms_thresh), float(self.variance[0]), float(self.variance[1]),
                             4 * P if per_image else 0,
                             (_abi.DETECT_LOGITS if logits else 0) | (_abi.DETECT_WS_CLEAN if clean and int(self.top_k) <= 1024 else 0), 0)
        keep = score_keep.to(dev).to(torch.uint8).contiguous() if score_keep is not None else None
        fin = pending._detect_tail_args() if pending is not None else None
        if refine is not None:        # RefineDet fused (ssdbox_detect_refine): anchors refined / filtered inside the kernels
            from .multibox_loss import make_refine
            if per_image or keep is not None:
                raise ValueError("ssdbox: refine= needs the shared [P,4] priors and no score_keep")
            rf, keepalive = make_refine(refine)
            _abi.check(_abi.lib().ssdbox_detect_refine(
                C.byref(cfg), _abi.ptr(loc, torch.float32, "loc_data"), _abi.ptr(scores, torch.float32, "conf_data"),
                _abi.ptr(pri, torch.float32, "prior_data"), C.byref(rf), _abi.ptr(out, torch.float32, "out"), _abi.ptr(counts),
                ws, n, _abi.stream_ptr(dev)))
        elif fin is None:
            _abi.check(_abi.lib().ssdbox_detect(
                C.byref(cfg), _abi.ptr(loc, torch.float32, "loc_data"), _abi.ptr(scores, torch.float32, "conf_data"),
                _abi.ptr(pri, torch.float32, "prior_data"), _abi.ptr(keep, torch.uint8, "score_keep", True),
                _abi.ptr(out, torch.float32, "out"), _abi.ptr(counts), ws, n, _abi.stream_ptr(dev)))
        else:
            peers, sums, losses = fin
            _abi.check(_abi.lib().ssdbox_detect_peers(
                C.byref(cfg), _abi.ptr(loc, torch.float32, "loc_data"), _abi.ptr(scores, torch.float32, "conf_data"),
                _abi.ptr(pri, torch.float32, "prior_data"), _abi.ptr(keep, torch.uint8, "score_keep", True),
                _abi.ptr(out, torch.float32, "out"), _abi.ptr(counts), C.byref(peers), _abi.ptr(sums), _abi.ptr(losses),
                ws, n, _abi.stream_ptr(dev)))
            pending._completed_by_detect()
        self._ws.commit()
        self.last_counts = counts
        return out

    __call__ = forward
