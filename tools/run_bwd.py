"""profiling driver: MultiBoxLoss forward + backward at SSD512-COCO B=64 (ncu -k regex:loss_bwd ...); flags via argv[1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch_b200"))
import torch
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
for it in range(4):
    a, b = crit.forward_packed(loc, conf, pri, gt, offs, 32)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); (a + b).backward(); e1.record(); torch.cuda.synchronize()
    print("backward %.1f us" % (1e3 * e0.elapsed_time(e1)))
    loc.grad = None; conf.grad = None
