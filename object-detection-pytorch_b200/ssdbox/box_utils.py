"""Drop-in mirror of lib/layers/box_utils.py: same names, argument order and return values;
every function runs a CUDA kernel of libssdbox.so on CUDA tensors (no CPU path)."""
import torch

from . import _abi

_ws = _abi.Workspace()


def point_form(boxes):
    """box_utils.py:6-15  (cx,cy,w,h) -> (xmin,ymin,xmax,ymax)."""
    b = _abi.as_f32(boxes)
    out = torch.empty_like(b)
    _abi.check(_abi.lib().ssdbox_point_form(_abi.ptr(b, torch.float32, "boxes"), b.size(0), _abi.ptr(out),
                                            _abi.stream_ptr(b.device)))
    return out


def center_size(boxes):
    """box_utils.py:18-27 (the reference's version has a torch.cat arity bug; this is its intent)."""
    b = _abi.as_f32(boxes)
    out = torch.empty_like(b)
    _abi.check(_abi.lib().ssdbox_center_form(_abi.ptr(b, torch.float32, "boxes"), b.size(0), _abi.ptr(out),
                                             _abi.stream_ptr(b.device)))
    return out


def jaccard(box_a, box_b):
    """box_utils.py:51-70  IoU of xyxy boxes [A,4] x [B,4] -> [A,B]."""
    a = _abi.as_f32(box_a)
    b = _abi.as_f32(box_b, a.device)
    out = torch.empty(a.size(0), b.size(0), dtype=torch.float32, device=a.device)
    _abi.check(_abi.lib().ssdbox_jaccard(_abi.ptr(a, torch.float32, "box_a"), a.size(0),
                                         _abi.ptr(b, torch.float32, "box_b"), b.size(0), _abi.ptr(out),
                                         _abi.stream_ptr(a.device)))
    return out


def encode(matched, priors, variances):
    """box_utils.py:201-222."""
    m = _abi.as_f32(matched)
    p = _abi.as_f32(priors, m.device)
    out = torch.empty_like(m)
    _abi.check(_abi.lib().ssdbox_encode(_abi.ptr(m, torch.float32, "matched"), _abi.ptr(p, torch.float32, "priors"),
                                        m.size(0), float(variances[0]), float(variances[1]), _abi.ptr(out),
                                        _abi.stream_ptr(m.device)))
    return out


def decode(loc, priors, variances):
    """box_utils.py:226-244.  loc [n,4] or [B,P,4] with priors [P,4]."""
    l = _abi.as_f32(loc)
    p = _abi.as_f32(priors, l.device)
    out = torch.empty_like(l)
    n = l.numel() // 4
    _abi.check(_abi.lib().ssdbox_decode(_abi.ptr(l, torch.float32, "loc"), _abi.ptr(p, torch.float32, "priors"), n,
                                        p.numel() // 4, float(variances[0]), float(variances[1]), _abi.ptr(out),
                                        None, _abi.stream_ptr(l.device)))
    return out


def log_sum_exp(x):
    """box_utils.py:265-273  [N,C] -> [N,1] with the reference's single global max."""
    x = _abi.as_f32(x)
    out = torch.empty(x.size(0), 1, dtype=torch.float32, device=x.device)
    ws, n = _ws.get(_abi.workspace_bytes(_abi.OP_LSE), x.device)
    _abi.check(_abi.lib().ssdbox_log_sum_exp(_abi.ptr(x, torch.float32, "x"), x.size(0), x.size(1), _abi.ptr(out),
                                             ws, n, _abi.stream_ptr(x.device)))
    return out


def match_batch(threshold, gt, gt_offsets, gmax, priors, variances, anchors_xyxy=None, binarize=False,
                want_overlap=False):
    """Batched box_utils.match: returns (loc_t [B,P,4] f32, conf_t [B,P] i64, match_idx [B,P] i32[, overlap])."""
    pri = _abi.as_f32(priors)
    dev = pri.device
    B = gt_offsets.numel() - 1
    per_image = pri.dim() == 3
    P = pri.size(-2)
    loc_t = torch.empty(B, P, 4, dtype=torch.float32, device=dev)
    conf_t = torch.empty(B, P, dtype=torch.int64, device=dev)
    midx = torch.empty(B, P, dtype=torch.int32, device=dev)
    ov = torch.empty(B, P, dtype=torch.float32, device=dev) if want_overlap else None
    ws, n = _ws.get(_abi.workspace_bytes(_abi.OP_MATCH, B, P, 0, gmax), dev)
    anc = _abi.as_f32(anchors_xyxy, dev) if anchors_xyxy is not None else None
    _abi.check(_abi.lib().ssdbox_match_encode(
        _abi.ptr(gt, torch.float32, "gt"), _abi.ptr(gt_offsets, torch.int32, "gt_offsets"), int(gmax),
        _abi.ptr(pri, torch.float32, "priors"), 4 * P if per_image else 0,
        _abi.ptr(anc, torch.float32, "anchors_xyxy", allow_none=True), B, P, float(threshold),
        float(variances[0]), float(variances[1]), 1 if binarize else 0, _abi.ptr(loc_t), _abi.ptr(conf_t),
        _abi.ptr(midx), _abi.ptr(ov, allow_none=True), ws, n, _abi.stream_ptr(dev)))
    return (loc_t, conf_t, midx, ov) if want_overlap else (loc_t, conf_t, midx)


def match(threshold, truths, priors, variances, labels, loc_t, conf_t, idx):
    """box_utils.py:92-133: fills loc_t[idx] / conf_t[idx] in place for one image."""
    dev = priors.device
    gt = torch.cat([_abi.as_f32(truths, dev), _abi.as_f32(labels, dev).view(-1, 1)], 1).contiguous()
    offs = torch.tensor([0, gt.size(0)], dtype=torch.int32).to(dev)
    lt, ct, _ = match_batch(threshold, gt, offs, gt.size(0), priors, variances)
    loc_t[idx] = lt[0].to(loc_t.device)
    conf_t[idx] = ct[0].to(conf_t.device)


def hard_negative_mine(keys, pos, negpos_ratio, pool=None):
    """multibox_loss.py:97-103 in isolation: keys [B,P] f32, pos [B,P] bool -> neg [B,P] bool."""
    k = _abi.as_f32(keys)
    dev = k.device
    ps = pos.to(dev).to(torch.uint8).contiguous()
    pl = pool.to(dev).to(torch.uint8).contiguous() if pool is not None else None
    neg = torch.empty_like(ps)
    B, P = k.shape
    ws, n = _ws.get(_abi.workspace_bytes(_abi.OP_MINE, B, P), dev)
    _abi.check(_abi.lib().ssdbox_hard_negative_mine(_abi.ptr(k), _abi.ptr(ps), _abi.ptr(pl, allow_none=True), B, P,
                                                    int(negpos_ratio), _abi.ptr(neg), ws, n, _abi.stream_ptr(dev)))
    return neg.bool()


def nms(boxes, scores, overlap=0.5, top_k=200):
    """box_utils.py:279-343: returns (keep LongTensor[n] zero padded, count int).
    On empty input the reference returns the bare `keep` tensor (:292-293); so do we."""
    s = _abi.as_f32(scores)
    dev = s.device
    keep = torch.zeros(s.size(0), dtype=torch.int64, device=dev)
    if boxes.numel() == 0:
        return keep
    b = _abi.as_f32(boxes, dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    ws, n = _ws.get(_abi.workspace_bytes(_abi.OP_NMS, 0, s.size(0), 0, 0, top_k), dev)
    _abi.check(_abi.lib().ssdbox_nms(_abi.ptr(b, torch.float32, "boxes"), _abi.ptr(s, torch.float32, "scores"),
                                     s.size(0), float(overlap), int(top_k), _abi.ptr(keep), _abi.ptr(count), ws, n,
                                     _abi.stream_ptr(dev)))
    return keep, int(count.item())
