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
NAME = sys.argv[1] if len(sys.argv) > 1 else "ssd512_coco"
cfg, c = configs.get(NAME); Cn = cfg.MODEL.NUM_CLASSES; B = int(sys.argv[2]) if len(sys.argv) > 2 else c["batch"]
pri = ssdbox.PriorBoxSSD(cfg).forward(c["layer_dims"], keep_on_device=True); P = pri.size(0)
tg = [t.to(dev) for t in synth.gen_targets(B, Cn, c["gt_max"], 0)]
loc = torch.randn(B, P, 4, device=dev) * 0.5
conf = torch.randn(B, P, Cn, device=dev); conf[..., 0] += 4
crit = ssdbox.MultiBoxLoss(Cn, 0.5, True, 0, True, 3, 0.5, False)
import time
for it in range(4):
    crit.abi_flags = {0: 0, 1: 0, 2: 256, 3: 1}[it]
    print('--- flags', crit.abi_flags, '(0 fused, 256 fused without conf streaming, 1 separate match kernel)')
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    with torch.no_grad():
        crit((loc, conf, pri), tg)
    e1.record(); torch.cuda.synchronize(); print('  forward total %.1f us' % (1e3 * e0.elapsed_time(e1)))
    buf = (C.c_longlong * (16 * 64))()
    _abi.lib().ssdbox_debug_phases.argtypes = [C.c_void_p]
    _abi.lib().ssdbox_debug_phases(buf)
    import numpy as np
    ph = np.array(list(buf), dtype=np.int64).reshape(64, 16)
    d = np.diff(ph[:, :8], axis=1)
    print("  mine phases (cycles, median over CTAs / max): load %d/%d  forced %d/%d  keys+scan %d/%d  positives %d/%d  select %d/%d  final %d/%d  sum %d/%d ; total med %d max %d; kernel span %d"
          % tuple([x for i in range(7) for x in (int(np.median(d[:, i])), int(d[:, i].max()))] + [int(np.median(ph[:, 7] - ph[:, 0])), int((ph[:, 7] - ph[:, 0]).max()), int(ph[:, :8].max() - ph[:, 0].min())]))
    gt0, gt1 = ph[:, 14], ph[:, 15]
    print("  mine CTAs 0..63 (globaltimer): start spread %.1f us, end spread %.1f us, first start -> last end %.1f us, CTA duration med %.1f max %.1f us"
          % ((gt0.max() - gt0.min()) / 1e3, (gt1.max() - gt1.min()) / 1e3, (gt1.max() - gt0.min()) / 1e3, np.median(gt1 - gt0) / 1e3, (gt1 - gt0).max() / 1e3))
    sb = (C.c_longlong * (8 * 160))()
    _abi.lib().ssdbox_debug_sphases.argtypes = [C.c_void_p]
    _abi.lib().ssdbox_debug_sphases(sb)
    import numpy as np
    arr = np.array(list(sb), dtype=np.int64).reshape(160, 8)[:148]
    rel = arr - arr[:, :1]
    print(" stream per-CTA cycles: match_w0_end min/med/max %s | cons_w0_end min/med/max %s | producer_end max %d"
          % (np.percentile(rel[:, 4], [0, 50, 100]).astype(int).tolist(), np.percentile(rel[:, 6], [0, 50, 100]).astype(int).tolist(), rel[:, 2].max()))
    gt_ = (C.c_ulonglong * (2 * 160))()
    _abi.lib().ssdbox_debug_sgt.argtypes = [C.c_void_p]
    _abi.lib().ssdbox_debug_sgt(gt_)
    g = np.array(list(gt_), dtype=np.int64).reshape(160, 2)[:148]
    span_ns = g[:, 1].max() - g[:, 0].min()
    print("  stream kernel span (globaltimer, first CTA start -> last consumer end): %.1f us; CTA start spread %.1f us; consumer-end spread %.1f us; cycles/ns of CTA 0: %.3f"
          % (span_ns / 1e3, (g[:, 0].max() - g[:, 0].min()) / 1e3, (g[:, 1].max() - g[:, 1].min()) / 1e3, rel[0, 7] / max(g[0, 1] - g[0, 0], 1)))
    ms = (C.c_longlong * (8 * 160))()
    _abi.lib().ssdbox_debug_mstat.argtypes = [C.c_void_p]
    _abi.lib().ssdbox_debug_mstat(ms)
    st = np.array(list(ms), dtype=np.int64).reshape(160, 8)[:148]
    print("  match warp 0 per CTA: wait cyc min/med/max %s | busy cyc min/med/max %s | units min/med/max %s | truths/unit %.1f | busy cyc/unit %.0f"
          % (np.percentile(st[:, 0], [0, 50, 100]).astype(int).tolist(), np.percentile(st[:, 1], [0, 50, 100]).astype(int).tolist(),
             np.percentile(st[:, 2], [0, 50, 100]).astype(int).tolist(), st[:, 3].sum() / max(st[:, 2].sum(), 1), st[:, 1].sum() / max(st[:, 2].sum(), 1)))
    buf = (C.c_longlong * 32)()
    _abi.lib().ssdbox_debug_match_phases.argtypes = [C.c_void_p]
    _abi.lib().ssdbox_debug_match_phases(buf)
    print(" match heavy tile:", [buf[i + 1] - buf[i] for i in range(5)], " light tile:", [buf[16 + i + 1] - buf[16 + i] for i in range(5)],
          "(gt->smem, priors+bbox, g-loop, writes, publish+ticket); light starts", buf[16] - buf[0], "after heavy")
