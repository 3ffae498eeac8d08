"""Installs the ssdbox modules under the reference's import names so that its train.py /
eval.py run unchanged:  `import ssdbox.compat; ssdbox.compat.install()` before
`from lib.layers import *`.

After install(), `lib.layers`, `lib.layers.functions`, `lib.layers.modules` and
`lib.layers.box_utils` of an already-imported reference tree expose the CUDA-backed
PriorBoxSSD / DetectOut / MultiBoxLoss and box_utils functions."""
import sys

from . import box_utils as _bu
from .detection import DetectOut
from .multibox_loss import MultiBoxLoss
from .prior_box import PriorBoxSSD

_BOX_FUNCS = ["point_form", "center_size", "jaccard", "match", "encode", "decode", "log_sum_exp", "nms"]


def install(variance_from_cfg=True, patch_eval=False):
    """Patches the reference modules in sys.modules (imports them first if `lib` is importable).
    patch_eval=True also swaps the evaluation solvers (lib.utils.evaluate_utils.EvalVOC / EvalCOCO and
    lib.utils.eval_solver_map, what eval_solver_factory hands out) for the device-resident ones."""
    try:
        import lib.layers  # noqa: F401  (reference tree on sys.path)
    except Exception as e:  # pragma: no cover - reference not importable here
        raise RuntimeError("ssdbox.compat.install(): the reference package `lib` is not importable: %r" % (e,))

    loss_cls = MultiBoxLoss
    if variance_from_cfg:
        from lib.utils.config import cfg as ref_cfg

        class MultiBoxLossCfg(MultiBoxLoss):
            """reads cfg.MODEL.VARIANCE at construction like multibox_loss.py:46"""

            def __init__(self, *a, **k):
                k.setdefault("variance", list(ref_cfg.MODEL.VARIANCE))
                super(MultiBoxLossCfg, self).__init__(*a, **k)

        MultiBoxLossCfg.__name__ = "MultiBoxLoss"
        loss_cls = MultiBoxLossCfg

    for name in ("lib.layers", "lib.layers.functions", "lib.layers.modules"):
        mod = sys.modules.get(name)
        if mod is None:
            continue
        for sym, obj in (("PriorBoxSSD", PriorBoxSSD), ("DetectOut", DetectOut), ("MultiBoxLoss", loss_cls)):
            if hasattr(mod, sym):
                setattr(mod, sym, obj)
    for name, sym, obj in (("lib.layers.functions.prior_box", "PriorBoxSSD", PriorBoxSSD),
                           ("lib.layers.functions.detection", "DetectOut", DetectOut),
                           ("lib.layers.modules.multibox_loss", "MultiBoxLoss", loss_cls)):
        mod = sys.modules.get(name)
        if mod is not None:
            setattr(mod, sym, obj)
    bu = sys.modules.get("lib.layers.box_utils")
    if bu is not None:
        for f in _BOX_FUNCS:
            setattr(bu, f, getattr(_bu, f))
    if patch_eval:
        from . import evaluate_utils as _eu
        mod = sys.modules.get("lib.utils.evaluate_utils")
        if mod is not None:
            mod.EvalVOC, mod.EvalCOCO, mod.EvalBase = _eu.EvalVOC, _eu.EvalCOCO, _eu.EvalBase
        utils = sys.modules.get("lib.utils")
        if utils is not None:
            utils.EvalVOC, utils.EvalCOCO = _eu.EvalVOC, _eu.EvalCOCO
            if hasattr(utils, "eval_solver_map"):
                utils.eval_solver_map.update({'VOC0712': _eu.EvalVOC, 'COCO2014': _eu.EvalCOCO})
    return loss_cls
