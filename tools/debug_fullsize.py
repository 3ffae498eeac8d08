"""Debug / first-look timing at BASELINE sizes: runs loss fwd(+bwd) and detect for growing batch
sizes with a sync after every call and prints the library's per-kernel device times."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch_b200"))
import torch
import ssdbox
from ssdbox import _abi, configs, synth

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "ssd512_coco"
batches = [int(b) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 16, 64]
cfg, c = configs.get(name)
C = cfg.MODEL.NUM_CLASSES
pri = ssdbox.PriorBoxSSD(cfg).forward(c["layer_dims"], keep_on_device=True)
P = pri.size(0)
for B in batches:
    tg = [t.to(dev) for t in synth.gen_targets(B, C, c["gt_max"], 0)]
    loc = (torch.randn(B, P, 4, device=dev) * 0.5)
    conf = torch.randn(B, P, C, device=dev)
    conf[..., 0] += 4
    sc = torch.randn(B, P, C, device=dev)
    sc[..., 0] += 10
    sc = torch.softmax(sc, -1)
    torch.cuda.synchronize()
    crit = ssdbox.MultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False)
    det = ssdbox.DetectOut(C, 0, 200, 0.01, 0.45, [0.1, 0.2])
    print("B=%d P=%d C=%d" % (B, P, C), flush=True)
    for rep in range(3):
        _abi.timers_enable(True)
        l = loc.clone().requires_grad_(True)
        x = conf.clone().requires_grad_(True)
        ll, lc = crit((l, x, pri), tg)
        torch.cuda.synchronize()
        print("  fwd ok", float(ll), float(lc), flush=True)
        (ll + lc).backward()
        torch.cuda.synchronize()
        print("  bwd ok", float(l.grad.abs().sum()), float(x.grad.abs().sum()), flush=True)
        out = det(loc, sc, pri)
        torch.cuda.synchronize()
        print("  det ok", int((out[..., 0] > 0).sum()), int(det.last_counts.max()), flush=True)
        t = _abi.timers_read()
        _abi.timers_enable(False)
        print("  kernels(us):", {k: round(v[0] * 1e3 / max(v[1], 1), 1) for k, v in t.items() if v[1]}, flush=True)
    del conf, sc, loc
print("done")
