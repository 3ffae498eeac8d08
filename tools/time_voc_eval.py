"""Times ssdbox.voc_eval.voc_eval on a VOC2007-test sized synthetic result set (per call, CUDA events)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch_b200"))
import torch
from ssdbox import synth
from ssdbox import voc_eval as VE
dev = torch.device("cuda:0")
case = synth.gen_voc_eval_case(4952, 21, 11, fp_max=12)
gt = VE.VOCGroundTruth(case["gt_boxes"], case["gt_labels"], case["gt_difficult"], case["gt_offsets"], dev)
rows, seg = torch.as_tensor(case["rows"]).to(dev), torch.as_tensor(case["seg"]).to(dev)
for use07 in (True, False):
    VE.voc_eval(rows, seg, gt, 21, 0.5, use07)
    torch.cuda.synchronize()
    for rep in range(3):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        t0 = time.perf_counter()
        e[0].record()
        for _ in range(5):
            r = VE.voc_eval(rows, seg, gt, 21, 0.5, use07)
        e[1].record()
        torch.cuda.synchronize()
        print("use07=%s rep %d: %.1f us per call (events), %.1f us (host clock), mAP %.6f" % (use07, rep, 1e3 * e[0].elapsed_time(e[1]) / 5, 1e6 * (time.perf_counter() - t0) / 5, r.mean_ap))
