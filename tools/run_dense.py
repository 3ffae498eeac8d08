"""profiling driver: DetectOut on dense scores at SSD512-COCO B=64 (ncu -k regex:detect_overflow|detect_segment_kernel ...)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch_b200"))
import torch
import ssdbox
from ssdbox import configs, synth
dev = torch.device("cuda:0")
cfg, c = configs.get("ssd512_coco"); C = 81; B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pri = ssdbox.PriorBoxSSD(cfg).forward(c["layer_dims"], keep_on_device=True); P = pri.size(0)
loc = torch.randn(B, P, 4, device=dev) * 0.5
g = torch.Generator(device=dev).manual_seed(3)
x = torch.randn(B, P, C, device=dev, generator=g); x[..., 0] += 4.0
sc = torch.softmax(x, -1); del x
det = ssdbox.DetectOut(C, 0, 200, 0.01, 0.45, [0.1, 0.2])
for it in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = det(loc, sc, pri); e1.record(); torch.cuda.synchronize()
    print("detect dense %.1f us, detections %d" % (1e3 * e0.elapsed_time(e1), int((out[..., 0] > 0).sum())))
