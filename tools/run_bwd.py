"""profiling driver: MultiBoxLoss forward + backward at SSD512-COCO B=64 (ncu -k regex:loss_bwd ...); flags via argv[1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch_b200"))
import torch
from ssdbox import _abi
if os.environ.get("SSDBOX_EXP"):
    _abi.LIB_PATH = os.path.join(ROOT, "tools", os.environ["SSDBOX_EXP"] if os.environ["SSDBOX_EXP"].endswith(".so") else "libssdbox_exp.so")
import ssdbox
from ssdbox import configs, synth
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dev = torch.device("cuda:0")
cfg, c = configs.get("ssd512_coco"); C = 81; B = 64
pri = ssdbox.PriorBoxSSD(cfg).forward(c["layer_dims"], keep_on_device=True); P = pri.size(0)
tg = synth.gen_targets(B, C, 32, 0)
gt, offs = synth.pack_targets(tg); gt, offs = gt.to(dev), offs.to(dev)
loc = (torch.randn(B, P, 4, device=dev) * 0.5).requires_grad_(True)
conf = torch.randn(B, P, C, device=dev); conf[..., 0] += 4; conf.requires_grad_(True)
crit = ssdbox.MultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False)
crit.abi_flags = flags
from ssdbox import _abi as A
ts = []
for it in range(8):
    a, b = crit.forward_packed(loc, conf, pri, gt, offs, 32)
    torch.cuda.synchronize()
    A.timers_enable(True)
    (a + b).backward()
    torch.cuda.synchronize()
    k = A.timers_read(); A.timers_enable(False)
    ts.append(1e3 * k["loss_bwd"][0] / max(k["loss_bwd"][1], 1))
    loc.grad = None; conf.grad = None
ts = sorted(ts[2:])
print("loss_bwd kernel: median %.1f us  min %.1f us  [%s]" % (ts[len(ts) // 2], ts[0], " ".join("%s=%s" % kv for kv in os.environ.items() if kv[0].startswith("SSDBOX_"))))
