"""TEST INFRASTRUCTURE ONLY -- loader for the *actual* reference implementation.

Imports the unmodified box hot path of arleyzhang/object-detection-pytorch from
``/root/reference`` (read-only mount, present only in the build container, never on
the GPU box) so that

  * ``oracle/ssd_oracle.py`` (the in-repo CPU restatement) can be proven bit-identical
    to the reference (``tests/test_oracle_vs_reference.py``), and
  * ``oracle/make_golden.py`` can record golden vectors under ``tests/golden/``.

Nothing in the product path (``object-detection-pytorch_b200/``) may import this file.

The reference targets PyTorch 0.3.1; three compatibility shims are applied *around* it
(the reference sources are never edited or copied), see SURVEY.md section 8c:

  S1  stub modules for ``tensorboardX`` / ``pycocotools`` (import chain only).
  S2  ``multibox_loss.py:97`` indexes a ``[B*P,1]`` tensor with a ``[B,P]`` mask, which
      torch>=0.4 rejects.  ``log_sum_exp`` is wrapped so that its result is a Tensor
      subclass whose ``__setitem__`` views itself with the mask's shape first
      (identical semantics to the 0.3.1 behaviour: element-wise masked fill).
  S3  ``DetectOut`` is a legacy autograd ``Function``: ``.forward`` is called directly,
      and ``nms`` is wrapped to return ``(keep, 0)`` on an empty candidate set
      (``box_utils.py:292-293`` returns a bare tensor there, ``detection.py:50`` relied
      on the 0.3-era ``dim()==0`` idiom to skip such classes).
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("SSDBOX_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lib", "layers", "box_utils.py"))


def _install_stubs():
    if "tensorboardX" not in sys.modules:
        m = types.ModuleType("tensorboardX")

        class SummaryWriter(object):  # pragma: no cover - never used
            def __init__(self, *a, **k):
                pass

        m.SummaryWriter = SummaryWriter
        sys.modules["tensorboardX"] = m
    if "pycocotools" not in sys.modules:
        pkg = types.ModuleType("pycocotools")
        coco = types.ModuleType("pycocotools.coco")
        cocoeval = types.ModuleType("pycocotools.cocoeval")
        coco.COCO = type("COCO", (), {})
        cocoeval.COCOeval = type("COCOeval", (), {})
        pkg.coco = coco
        pkg.cocoeval = cocoeval
        sys.modules["pycocotools"] = pkg
        sys.modules["pycocotools.coco"] = coco
        sys.modules["pycocotools.cocoeval"] = cocoeval


class _MaskViewTensor(torch.Tensor):
    """S2: masked assignment with a differently-shaped (same numel) boolean mask."""

    def __setitem__(self, key, value):
        if isinstance(key, torch.Tensor) and key.dtype in (torch.bool, torch.uint8) \
                and key.shape != self.shape and key.numel() == self.numel():
            torch.Tensor.__setitem__(self.as_subclass(torch.Tensor).view(key.shape), key, value)
            return
        torch.Tensor.__setitem__(self, key, value)


_REF = None


def load():
    """Returns a namespace with the reference's hot-path symbols (shimmed)."""
    global _REF
    if _REF is not None:
        return _REF
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import lib.layers.box_utils as box_utils
    import lib.layers.modules.multibox_loss as mbl
    import lib.layers.functions.detection as det
    import lib.layers.functions.prior_box as pb
    from lib.utils.config import cfg as ref_cfg

    # S2
    _orig_lse = box_utils.log_sum_exp

    def _lse_shim(x):
        return _orig_lse(x).as_subclass(_MaskViewTensor)

    mbl.log_sum_exp = _lse_shim

    # S3
    _orig_nms = box_utils.nms

    def _nms_shim(boxes, scores, overlap=0.5, top_k=200):
        if boxes.numel() == 0:
            return scores.new_zeros(scores.size(0)).long(), 0
        return _orig_nms(boxes, scores, overlap, top_k)

    det.nms = _nms_shim

    ns = types.SimpleNamespace()
    ns.box_utils = box_utils
    ns.cfg = ref_cfg
    ns.PriorBoxSSD = pb.PriorBoxSSD
    ns.MultiBoxLoss = mbl.MultiBoxLoss
    ns.DetectOut = det.DetectOut
    ns.nms = _nms_shim
    ns.raw_nms = _orig_nms

    def multibox_loss(num_classes, predictions, targets, threshold=0.5, neg_pos=3,
                      variance=(0.1, 0.2)):
        """reference MultiBoxLoss.forward on CPU (train.py:99-100 ctor arguments)."""
        crit = mbl.MultiBoxLoss(num_classes, threshold, True, 0, True, neg_pos, 0.5, False,
                                use_gpu=False)
        crit.variance = list(variance)
        old = torch.get_default_dtype()
        return crit.forward(predictions, targets)

    def detect(num_classes, loc, conf, priors, top_k=200, conf_thresh=0.01, nms_thresh=0.45,
               variance=(0.1, 0.2)):
        """reference DetectOut.forward on CPU (evaluate_utils.py:16-17 ctor arguments)."""
        d = det.DetectOut.__new__(det.DetectOut)
        det.DetectOut.__init__(d, num_classes, 0, top_k, conf_thresh, nms_thresh, list(variance))
        return det.DetectOut.forward(d, loc, conf, priors)

    ns.multibox_loss = multibox_loss
    ns.detect = detect
    _REF = ns
    return ns
