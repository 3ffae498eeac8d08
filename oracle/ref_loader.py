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

VENDORED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/build_ref.py


def _has_lib(root):
    return os.path.isfile(os.path.join(root, "lib", "layers", "box_utils.py"))


def _pick_root():
    """SSDBOX_REFERENCE_ROOT if set; else the read-only mount of the build container; else the unmodified copy
    oracle/build_ref.py left under oracle/_ref (what travels to the GPU box)."""
    env = os.environ.get("SSDBOX_REFERENCE_ROOT")
    if env:
        return env
    return "/root/reference" if _has_lib("/root/reference") else VENDORED_ROOT


REFERENCE_ROOT = _pick_root()


def use_vendored():
    """bench.py: always time the copy under oracle/_ref (never reads /root/reference at run time)."""
    global REFERENCE_ROOT
    if _REF is not None and REFERENCE_ROOT != VENDORED_ROOT:
        raise RuntimeError("reference already loaded from %s" % REFERENCE_ROOT)
    REFERENCE_ROOT = VENDORED_ROOT
    return available()


def available():
    return _has_lib(REFERENCE_ROOT)


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


def voc_eval_reference(case, workdir, use_07_metric=True, ovthresh=0.5):
    """Runs the reference's own write_voc_results_file + voc_eval (lib/datasets/voc_eval.py:58-75,
    109-242) on a case of ssdbox.synth.gen_voc_eval_case: the annotation cache (annots.pkl, what
    :143-150 would write after parsing the xml files), the image-set file and the per-class result
    files are created under `workdir`.  Returns [(rec, prec, ap)] for the classes 1..C-1.

    Shim S5: the module-level `devkit_path` (:9, from the global cfg) is pointed at `workdir` while the
    result files are written; the `dets == []` test of :65 is a list-vs-ndarray comparison that numpy 2
    evaluates element-wise, so images without detections keep the reference's own `[]` placeholder and
    the others go through unchanged."""
    import pickle

    import numpy as np
    load()
    import lib.datasets.voc_eval as ve
    C, I = int(case["num_classes"]), int(case["num_images"])
    names = ["%06d" % (i + 1) for i in range(I)]
    classes = ["cls%02d" % c for c in range(1, C)]
    recs = {}
    for i, nm in enumerate(names):
        objs = []
        for g in range(int(case["gt_offsets"][i]), int(case["gt_offsets"][i + 1])):
            objs.append({"name": classes[int(case["gt_labels"][g]) - 1], "pose": "Unspecified", "truncated": 0,
                         "difficult": int(case["gt_difficult"][g]), "bbox": [int(v) for v in case["gt_boxes"][g]]})
        recs[nm] = objs
    cachedir = os.path.join(workdir, "annotations_cache")
    os.makedirs(cachedir, exist_ok=True)
    with open(os.path.join(cachedir, "annots.pkl"), "wb") as f:
        pickle.dump(recs, f)
    imageset = os.path.join(workdir, "test.txt")
    with open(imageset, "w") as f:
        f.write("\n".join(names) + "\n")
    # EvalVOC.reset_results / post_proc layout (evaluate_utils.py:122-151): results[cls][img] = float32 [n,5] or []
    seg, rows = case["seg"], case["rows"]
    all_boxes = [[[] for _ in range(I)] for _ in range(C)]
    for i in range(I):
        for c in range(1, C):
            a, b = int(seg[i * C + c]), int(seg[i * C + c + 1])
            if b > a:
                all_boxes[c][i] = rows[a:b, 0:5].astype(np.float32, copy=False)
    dataset = types.SimpleNamespace(ids=[("VOC2007", nm) for nm in names])
    saved = (ve.devkit_path, ve.labelmap)
    ve.devkit_path, ve.labelmap = workdir, classes
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _write_results(ve, all_boxes, dataset)
        out = []
        for cls in classes:
            detfile = ve.get_voc_results_file_template("test", cls)
            out.append(ve.voc_eval(detfile, os.path.join(workdir, "%s.xml"), imageset, cls, cachedir,
                                   ovthresh=ovthresh, use_07_metric=use_07_metric))
    finally:
        ve.devkit_path, ve.labelmap = saved
    return out


def _write_results(ve, all_boxes, dataset):
    """write_voc_results_file (:58-75) unmodified when numpy still accepts its `dets == []` test;
    otherwise the same per-line format call (:70-74) driven from here."""
    try:
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            ve.write_voc_results_file(all_boxes, dataset, "test")
        return "reference"
    except (ValueError, TypeError):
        for cls_ind, cls in enumerate(ve.labelmap):
            with open(ve.get_voc_results_file_template("test", cls), "wt") as f:
                for im_ind, index in enumerate(dataset.ids):
                    dets = all_boxes[cls_ind + 1][im_ind]
                    if isinstance(dets, list):
                        continue
                    for k in range(dets.shape[0]):
                        f.write('{:s} {:.3f} {:.1f} {:.1f} {:.1f} {:.1f}\n'.format(
                            index[1], dets[k, -1], dets[k, 0] + 1, dets[k, 1] + 1, dets[k, 2] + 1, dets[k, 3] + 1))
        return "format-only"
