"""TEST INFRASTRUCTURE ONLY -- records tests/golden/voceval.npz from the *reference itself*
(SURVEY.md 8f rank 4).  Run in the build container (needs /root/reference):

    python oracle/make_golden_voc.py

  * voc_*: a synthetic evaluation set (ssdbox.synth.gen_voc_eval_case, 3-decimal scores pairwise
    distinct inside every class so that np.argsort's unstable order cannot matter) pushed through the
    reference's write_voc_results_file + voc_eval (lib/datasets/voc_eval.py:58-75, 109-242) by
    oracle/ref_loader.voc_eval_reference, for the 11-point and the area metric;
  * crop_*: the reference's jaccard_numpy (lib/utils/augmentations.py:20-37) on seeded truths / rects.
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch_b200"))

from oracle import ref_loader  # noqa: E402
from ssdbox import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "voceval.npz")
CASE = dict(num_images=60, num_classes=21, seed=7, distinct_scores=True)


def crop_inputs(seed=5, B=6, T=8):
    rs = np.random.RandomState(seed)
    boxes, rects = [], np.zeros((B, T, 4), dtype=np.int64)
    for b in range(B):
        G = rs.randint(1, 9)
        wh = rs.rand(G, 2) * 200 + 5
        xy = rs.rand(G, 2) * 250
        boxes.append(np.concatenate([xy, xy + wh], 1))          # float64 absolute coordinates
        for t in range(T):
            w, h = rs.uniform(0.3 * 500, 500), rs.uniform(0.3 * 375, 375)
            left, top = rs.uniform(500 - w), rs.uniform(375 - h)
            rects[b, t] = [int(left), int(top), int(left + w), int(top + h)]      # augmentations.py:248
    return boxes, rects


def main():
    ref_loader.load()
    case = synth.gen_voc_eval_case(**CASE)
    out = {k: v for k, v in case.items() if isinstance(v, np.ndarray)}
    for tag, use07 in (("07", True), ("area", False)):
        with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(io.StringIO()):
            res = ref_loader.voc_eval_reference(case, d, use07)
        out["ap_" + tag] = np.array([float(r[2]) for r in res])
        if use07:
            out["rec"] = np.concatenate([np.atleast_1d(r[0]) for r in res if np.ndim(r[0])])
            out["prec"] = np.concatenate([np.atleast_1d(r[1]) for r in res if np.ndim(r[1])])
            out["count"] = np.array([len(r[0]) if np.ndim(r[0]) else 0 for r in res], dtype=np.int32)
    import lib.utils.augmentations as aug
    boxes, rects = crop_inputs()
    ov = []
    for b, bx in enumerate(boxes):
        for t in range(rects.shape[1]):
            ov.append(aug.jaccard_numpy(bx, rects[b, t]))
    out["crop_overlap"] = np.concatenate(ov)
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "mAP07 %.4f" % out["ap_07"].mean(), "rows", case["rows"].shape[0])


if __name__ == "__main__":
    main()
