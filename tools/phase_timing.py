"""debug: per-phase clock64 stamps of mine_reduce_kernel (CTA 1) from a -DSSDBOX_PHASE_TIMING build"""
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
for _ in range(3):
    with torch.no_grad():
        crit((loc, conf, pri), tg)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 16)()
    _abi.lib().ssdbox_debug_phases.argtypes = [C.c_void_p]
    print(_abi.lib().ssdbox_debug_phases(buf), [buf[i + 1] - buf[i] for i in range(5)], "cycles: passA, sums, select, final, sum")
    sb = (C.c_longlong * (8 * 160))()
    _abi.lib().ssdbox_debug_sphases.argtypes = [C.c_void_p]
    _abi.lib().ssdbox_debug_sphases(sb)
    import numpy as np
    arr = np.array(list(sb), dtype=np.int64).reshape(160, 8)[:148]
    rel = arr - arr[:, :1]
    print(" stream per-CTA cycles: match_w0_end min/med/max %s | cons_w0_end min/med/max %s | producer_end max %d"
          % (np.percentile(rel[:, 4], [0, 50, 100]).astype(int).tolist(), np.percentile(rel[:, 6], [0, 50, 100]).astype(int).tolist(), rel[:, 2].max()))
    ms = (C.c_longlong * (8 * 160))()
    _abi.lib().ssdbox_debug_mstat.argtypes = [C.c_void_p]
    _abi.lib().ssdbox_debug_mstat(ms)
    st = np.array(list(ms), dtype=np.int64).reshape(160, 8)[:148]
    for i in (3, 138, 57, 70):
        print("  CTA %d warp0: load %d cyc, g-loop %d cyc over %d iters; truths visited %d, computed %d; match_end %d" % (i, st[i,0], st[i,1], st[i,2], st[i,4], st[i,3], rel[i,4]))
    worst = np.argsort(-rel[:, 4])[:6]
    print(" slowest match CTAs:", [(int(i), int(rel[i, 4]), int(rel[i, 6])) for i in worst])
    buf = (C.c_longlong * 32)()
    _abi.lib().ssdbox_debug_match_phases.argtypes = [C.c_void_p]
    _abi.lib().ssdbox_debug_match_phases(buf)
    print(" match heavy tile:", [buf[i + 1] - buf[i] for i in range(5)], " light tile:", [buf[16 + i + 1] - buf[16 + i] for i in range(5)],
          "(gt->smem, priors+bbox, g-loop, writes, publish+ticket); light starts", buf[16] - buf[0], "after heavy")
