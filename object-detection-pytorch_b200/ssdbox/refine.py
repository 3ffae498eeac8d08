"""RefineDet two-step ARM -> ODM path on the same kernels (arXiv 1711.06897).

NOT in the reference snapshot (README.md:6 only names it): the semantics below are this
repository's own restatement, mirrored by oracle/ssd_oracle.py
`refine_*`; parity is therefore *unpinned* against the reference.

  ARM loss   (use_ARM=False): MultiBoxLoss with every truth label binarised to class 1 (C = 2).
  ODM loss   (use_ARM=True) : anchors are refined per image, decode(arm_loc, priors); truths are
                              matched against the refined boxes and encoded against their centre
                              form; anchors whose ARM objectness softmax(arm_conf)[:,1] <= theta
                              leave the positives and the hard-negative pool.
  Inference                 : decode(odm_loc, refined anchors), scores of filtered anchors -> 0,
                              then DetectOut.
"""
import torch

from . import _abi
from .detection import DetectOut
from .multibox_loss import MultiBoxLoss, pack_targets


def refine_anchors(arm_loc, priors, variance=(0.1, 0.2)):
    """decode(arm_loc[b], priors) for the whole batch -> (xyxy [B,P,4], centre form [B,P,4])."""
    loc = _abi.as_f32(arm_loc.detach())
    pri = _abi.as_f32(priors, loc.device)
    xy = torch.empty_like(loc)
    cf = torch.empty_like(loc)
    _abi.check(_abi.lib().ssdbox_decode(_abi.ptr(loc), _abi.ptr(pri), loc.numel() // 4, pri.size(0),
                                        float(variance[0]), float(variance[1]), _abi.ptr(xy), _abi.ptr(cf),
                                        _abi.stream_ptr(loc.device)))
    return xy, cf


def arm_filter(arm_conf, theta=0.01):
    """uint8 [B,P]: 1 where softmax(arm_conf)[...,1] > theta."""
    x = _abi.as_f32(arm_conf.detach())
    keep = torch.empty(x.shape[:-1], dtype=torch.uint8, device=x.device)
    _abi.check(_abi.lib().ssdbox_arm_filter(_abi.ptr(x), x.numel() // 2, float(theta), _abi.ptr(keep),
                                            _abi.stream_ptr(x.device)))
    return keep


class RefineMultiBoxLoss(MultiBoxLoss):
    """forward((arm_loc, arm_conf, odm_loc, odm_conf, priors), targets) -> (loss_l, loss_c)."""

    def __init__(self, num_classes, overlap_thresh, prior_for_matching, bkg_label, neg_mining, neg_pos,
                 neg_overlap, encode_target, use_gpu=True, theta=0.01, use_ARM=False, variance=(0.1, 0.2),
                 distributed=None, process_group=None, fused=True):
        super(RefineMultiBoxLoss, self).__init__(num_classes, overlap_thresh, prior_for_matching, bkg_label,
                                                 neg_mining, neg_pos, neg_overlap, encode_target, use_gpu, variance,
                                                 distributed, process_group)
        self.theta = theta
        self.use_ARM = use_ARM
        self.binarize_labels = not use_ARM
        # fused (default): arm_loc / arm_conf go to the loss kernels as they are (ssdbox_multibox_loss_fwd_refine): no
        # refined-anchor tensors, no mask, no extra launch.  fused=False materialises them first (refine_anchors +
        # arm_filter, two small kernels) -- same results bit for bit, kept for comparison.
        self.fused = bool(fused)

    def forward(self, predictions, targets):
        arm_loc, arm_conf, odm_loc, odm_conf, priors = predictions
        dev = arm_loc.device
        P = arm_loc.size(1)
        pri = _abi.as_f32(priors, dev)[:P].contiguous()
        gt, offsets, gmax = pack_targets(targets, dev)
        return self.forward_packed_two_step(arm_loc, arm_conf, odm_loc, odm_conf, pri, gt, offsets, gmax)

    def forward_packed_two_step(self, arm_loc, arm_conf, odm_loc, odm_conf, pri, gt, offsets, gmax):
        """forward with the targets already in the C-ABI layout (no host work: CUDA-graph capturable)."""
        P = arm_loc.size(1)
        if not self.use_ARM:
            loc = _abi.as_f32(arm_loc)
            conf = _abi.as_f32(arm_conf).view(loc.size(0), P, 2)
            return self.forward_packed(loc, conf, pri, gt, offsets, gmax)
        loc = _abi.as_f32(odm_loc)
        conf = _abi.as_f32(odm_conf).view(loc.size(0), P, self.num_classes)
        if self.fused:
            return self.forward_packed(loc, conf, pri, gt, offsets, gmax, refine=(arm_loc, arm_conf, self.theta))
        xy, cf = refine_anchors(arm_loc, pri, self.variance)
        pool = arm_filter(arm_conf, self.theta)
        return self.forward_packed(loc, conf, cf, gt, offsets, gmax, anchors_xyxy=xy, pool=pool)


class RefineDetectOut(DetectOut):
    """detector(arm_loc, arm_conf, odm_loc, odm_scores, priors) -> [B,C,top_k,5]."""

    def __init__(self, num_classes, bkg_label, top_k, conf_thresh, nms_thresh, variance, theta=0.01, fused=True):
        super(RefineDetectOut, self).__init__(num_classes, bkg_label, top_k, conf_thresh, nms_thresh, variance)
        self.theta = theta
        self.fused = bool(fused)      # see RefineMultiBoxLoss

    def forward(self, arm_loc, arm_conf, odm_loc, odm_scores, prior_data, out=None):
        if self.fused:
            return DetectOut.forward(self, odm_loc, odm_scores, prior_data, out=out, refine=(arm_loc, arm_conf, self.theta))
        _, cf = refine_anchors(arm_loc, prior_data, self.variance)
        keep = arm_filter(arm_conf, self.theta)
        return DetectOut.forward(self, odm_loc, odm_scores, cf, score_keep=keep, out=out)

    __call__ = forward
