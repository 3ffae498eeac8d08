"""Head-output layout -- host mirror of lib/models/ssd_v3.py:114-121 (rfb_net.py:213-220):

    loc.append(l(x).permute(0, 2, 3, 1).contiguous()) ...            # per source layer
    loc = torch.cat([o.view(o.size(0), -1) for o in loc], 1).view(B, -1, 4)

`heads_to_rows(outputs, K)` takes the raw multibox head outputs (list of [B, A_k*K, H_k, W_k] CUDA
tensors) and returns the [B, P, K] tensor in one launch (ssdbox_heads_to_rows).
No autograd: inference / target-generation side only."""
import ctypes as C

import torch

from . import _abi


def heads_to_rows(outputs, K, out=None):
    if len(outputs) == 0:
        raise ValueError("need at least one head output")
    if len(outputs) > _abi.MAX_HEADS:
        raise ValueError("at most %d source layers" % _abi.MAX_HEADS)
    if not outputs[0].is_cuda:
        raise RuntimeError("ssdbox: heads_to_rows runs on CUDA tensors only (no CPU path)")
    dev = outputs[0].device
    B = outputs[0].size(0)
    cfg = _abi.HeadsCfg()
    cfg.num_layers = len(outputs)
    cfg.B = B
    keep = []
    total = 0
    for k, o in enumerate(outputs):
        if o.dim() != 4 or o.size(0) != B:
            raise ValueError("head output %d must be [B, A*K, H, W]" % k)
        if o.size(1) % K:
            raise ValueError("head output %d has %d channels, not a multiple of %d" % (k, o.size(1), K))
        t = _abi.as_f32(o.detach(), dev)
        keep.append(t)
        cfg.channels[k] = t.size(1)
        cfg.hw[k] = t.size(2) * t.size(3)
        cfg.src[k] = C.c_void_p(t.data_ptr())
        total += t.size(1) * t.size(2) * t.size(3)
    if out is None:
        out = torch.empty(B, total // K, K, dtype=torch.float32, device=dev)
    elif out.numel() != B * total or not out.is_contiguous():
        raise ValueError("out must be a contiguous tensor of %d elements" % (B * total))
    _abi.check(_abi.lib().ssdbox_heads_to_rows(C.byref(cfg), _abi.ptr(out, torch.float32, "out"), _abi.stream_ptr(dev)))
    return out
