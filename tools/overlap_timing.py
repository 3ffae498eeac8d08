"""debug: do match_kernel and loss_stream_kernel overlap? (%globaltimer min start / max end of each)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch_b200"))
import torch
from ssdbox import _abi
_abi.LIB_PATH = os.path.join(ROOT, "tools", "libssdbox_dbg.so")
import ssdbox
from ssdbox import configs, synth
dev = torch.device("cuda:0")
cfg, c = configs.get("ssd512_coco"); Cn = 81; B = 64
pri = ssdbox.PriorBoxSSD(cfg).forward(c["layer_dims"], keep_on_device=True); P = pri.size(0)
tg = [t.to(dev) for t in synth.gen_targets(B, Cn, 32, 0)]
loc = torch.randn(B, P, 4, device=dev) * 0.5
conf = torch.randn(B, P, Cn, device=dev); conf[..., 0] += 4
crit = ssdbox.MultiBoxLoss(Cn, 0.5, True, 0, True, 3, 0.5, False)
lib = _abi.lib()
lib.ssdbox_debug_gtimes.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_ulonglong * 8)()
for it in range(4):
    lib.ssdbox_debug_gtimes(buf, 1)
    with torch.no_grad():
        crit((loc, conf, pri), tg)
    torch.cuda.synchronize()
    lib.ssdbox_debug_gtimes(buf, 0)
    m0, m1, s0, s1 = buf[0], buf[1], buf[2], buf[3]
    base = min(m0, s0)
    print("match  start %6.1f us end %6.1f us | stream start %6.1f us end %6.1f us" % ((m0 - base) / 1e3, (m1 - base) / 1e3, (s0 - base) / 1e3, (s1 - base) / 1e3))
