"""ssdbox -- B200-native SSD-series box hot path (PriorBox -> match/encode -> MultiBoxLoss ->
Detect/NMS, plus RefineDet's two-step variant) behind the reference's module signatures.

Host code is this thin Python layer; all arithmetic lives in lib/libssdbox.so (sm_100a CUDA,
C ABI in include/ssdbox.h).  There is no CPU fallback."""
from . import _abi, box_utils, configs, synth          # noqa: F401
from .detection import DetectOut                        # noqa: F401
from .multibox_loss import MultiBoxLoss, pack_targets   # noqa: F401
from .prior_box import PriorBoxSSD                      # noqa: F401
from .refine import RefineDetectOut, RefineMultiBoxLoss, arm_filter, refine_anchors  # noqa: F401
from .overlap import MultiStreamStep, TwoStreamStep     # noqa: F401

__all__ = ["PriorBoxSSD", "DetectOut", "MultiBoxLoss", "RefineMultiBoxLoss", "RefineDetectOut", "TwoStreamStep", "MultiStreamStep",
           "box_utils", "configs", "synth"]
