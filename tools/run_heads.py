"""profiling driver: ssdbox_heads_to_rows on the conf heads of SSD512-COCO, B=64 (ncu -k regex:heads ...)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch_b200"))
import torch
from ssdbox import _abi
if os.environ.get("SSDBOX_EXP"):
    _abi.LIB_PATH = os.path.join(ROOT, "tools", os.environ["SSDBOX_EXP"] if os.environ["SSDBOX_EXP"].endswith(".so") else "libssdbox_exp.so")
import ssdbox
from ssdbox import configs, heads as HD
name = sys.argv[1] if len(sys.argv) > 1 else "ssd512_coco"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
cfg, c = configs.get(name); C = cfg.MODEL.NUM_CLASSES; B = 64
per_cell = ssdbox.PriorBoxSSD(cfg).num_priors
outs = [torch.randn(B, a * C, h, w, device=dev) for a, (h, w) in zip(per_cell, c["layer_dims"])]
P = sum(a * h * w for a, (h, w) in zip(per_cell, c["layer_dims"]))
rows = torch.empty(B, P, C, device=dev)
HD.heads_to_rows(outs, C, out=rows)
ref = torch.cat([o.permute(0, 2, 3, 1).contiguous().view(B, -1) for o in outs], 1).view(B, P, C)
print("equal", bool(torch.equal(ref, rows)))
del ref
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for it in range(reps):
    flush.zero_()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); HD.heads_to_rows(outs, C, out=rows); e1.record(); torch.cuda.synchronize()
    ts.append(1e3 * e0.elapsed_time(e1))
ts.sort()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
flush.zero_(); e0.record()
for it in range(reps):
    HD.heads_to_rows(outs, C, out=rows)
e1.record(); torch.cuda.synchronize()
b2b = 1e3 * e0.elapsed_time(e1) / reps
nb = 2 * B * P * C * 4
print("%s heads_to_rows median %.1f us  min %.1f us  back-to-back %.1f us = %.2f TB/s  [%s]" % (
    name, ts[len(ts) // 2], ts[0], b2b, nb / b2b / 1e6, " ".join("%s=%s" % kv for kv in os.environ.items() if kv[0].startswith("SSDBOX_"))))
