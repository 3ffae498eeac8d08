"""The in-repo oracle must be bit-identical to the imported reference (build container only).

Skipped where /root/reference is not mounted (the GPU box): there the committed golden
vectors (tests/test_oracle_golden.py) pin the oracle instead.
"""
import os

import pytest
import torch

from oracle import ref_loader
from oracle import ssd_oracle as O
from ssdbox import configs, synth

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def _priors(name):
    cfg, c = configs.get(name)
    return O.prior_boxes(cfg.MODEL, c["layer_dims"]), cfg, c


@pytest.mark.parametrize("name", list(configs.CONFIGS))
def test_priors_bit_exact(ref, name):
    pri, cfg, c = _priors(name)
    rp = ref.PriorBoxSSD(cfg)
    assert rp.num_priors == O.num_priors_per_cell(cfg.MODEL)
    assert torch.equal(rp.forward(c["layer_dims"]), pri)
    assert pri.size(0) == c["num_priors"]


def test_box_algebra_bit_exact(ref):
    pri, _, _ = _priors("ssd300_voc")
    bu = ref.box_utils
    t = synth.gen_targets(1, 21, 16, 5)[0]
    assert torch.equal(bu.point_form(pri), O.point_form(pri))
    assert torch.equal(bu.jaccard(t[:, :4], bu.point_form(pri)), O.iou_matrix(t[:, :4], O.point_form(pri)))
    m = t[torch.randint(0, t.size(0), (pri.size(0),)), :4]
    assert torch.equal(bu.encode(m, pri, [0.1, 0.2]), O.encode_boxes(m, pri, [0.1, 0.2]))
    loc = synth.gen_loc(1, pri.size(0), 3)[0]
    assert torch.equal(bu.decode(loc, pri, [0.1, 0.2]), O.decode_boxes(loc, pri, [0.1, 0.2]))
    x = synth.gen_train_logits(2, 500, 21, 1).view(-1, 21)
    assert torch.equal(bu.log_sum_exp(x), O.log_sum_exp(x))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_match_bit_exact(ref, seed):
    pri, _, _ = _priors("ssd300_voc")
    tg = synth.gen_targets(4, 21, 16, seed)
    P = pri.size(0)
    loc_t = torch.zeros(4, P, 4)
    conf_t = torch.zeros(4, P, dtype=torch.int64)
    for b, t in enumerate(tg):
        ref.box_utils.match(0.5, t[:, :4], pri, [0.1, 0.2], t[:, 4], loc_t, conf_t, b)
        m = O.match_image(0.5, t[:, :4], pri, [0.1, 0.2], t[:, 4])
        assert torch.equal(conf_t[b], m["conf"])
        assert torch.equal(loc_t[b], m["loc"])


def test_match_duplicate_truths_last_wins(ref):
    pri, _, _ = _priors("ssd300_voc")
    t = torch.tensor([[0.1, 0.1, 0.4, 0.5, 3.0], [0.1, 0.1, 0.4, 0.5, 7.0]])
    P = pri.size(0)
    loc_t = torch.zeros(1, P, 4)
    conf_t = torch.zeros(1, P, dtype=torch.int64)
    ref.box_utils.match(0.5, t[:, :4], pri, [0.1, 0.2], t[:, 4], loc_t, conf_t, 0)
    m = O.match_image(0.5, t[:, :4], pri, [0.1, 0.2], t[:, 4])
    assert torch.equal(conf_t[0], m["conf"])
    assert int(m["conf"][m["best_prior"][0]]) == 8


@pytest.mark.parametrize("name,B,seed", [("ssd300_voc", 4, 0), ("ssd300_voc", 3, 1),
                                         ("fssd300_coco", 2, 2), ("refinedet320_voc", 3, 3)])
def test_multibox_loss_bit_exact(ref, name, B, seed):
    pri, cfg, c = _priors(name)
    C = cfg.MODEL.NUM_CLASSES
    P = pri.size(0)
    tg = synth.gen_targets(B, C, c["gt_max"], seed)
    loc = synth.gen_loc(B, P, seed)
    conf = synth.gen_train_logits(B, P, C, seed)
    rl, rc = ref.multibox_loss(C, (loc, conf, pri), tg)
    # stable=False reproduces the literal reference sort; stable=True is the canonical order.
    for stable in (False, True):
        ol, oc = O.multibox_loss(loc, conf, pri, tg, C, stable=stable)
        assert float(rl) == float(ol)
        assert float(rc) == float(oc)


@pytest.mark.parametrize("bias,seed", [(10.0, 0), (10.0, 1), (6.0, 2)])
def test_detect_bit_exact(ref, bias, seed):
    pri, cfg, c = _priors("ssd300_voc")
    B, P, C = 2, pri.size(0), 21
    loc = synth.gen_loc(B, P, seed)
    sc = synth.gen_detect_scores(B, P, C, seed, bkg_bias=bias)
    r = ref.detect(C, loc, sc, pri)
    for stable in (False, True):
        o = O.detect(loc, sc, pri, C, stable=stable)
        assert torch.equal(r, o)
    # conf flattened to [B*P, C] (rfb_net.py:222-226) is accepted too
    assert torch.equal(r, O.detect(loc, sc.view(-1, C), pri, C))


def test_nms_bit_exact(ref):
    g = torch.Generator().manual_seed(7)
    for n, k in [(1, 200), (5, 200), (50, 10), (300, 200), (700, 200)]:
        xy = torch.rand(n, 2, generator=g) * 0.6
        wh = torch.rand(n, 2, generator=g) * 0.3 + 0.02
        boxes = torch.cat([xy, xy + wh], 1)
        scores = torch.rand(n, generator=g)
        rk, rc = ref.nms(boxes, scores, 0.45, k)
        ok, oc = O.greedy_nms(boxes, scores, 0.45, k)
        assert rc == oc and torch.equal(rk, ok)
    rk, rc = ref.nms(torch.zeros(0, 4), torch.zeros(0), 0.45, 200)
    ok, oc = O.greedy_nms(torch.zeros(0, 4), torch.zeros(0), 0.45, 200)
    assert rc == oc == 0


def test_compat_install_rebinds_reference_symbols():
    """ssdbox.compat.install() swaps the CUDA-backed classes into the reference's namespaces
    (run in a subprocess: it patches lib.layers globally)."""
    import subprocess
    import sys
    code = r'''
import sys, os
root = %r
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "object-detection-pytorch_b200"))
import warnings; warnings.filterwarnings("ignore")
from oracle import ref_loader
ref_loader._install_stubs()
sys.path.insert(0, ref_loader.REFERENCE_ROOT)
import lib.layers
import ssdbox, ssdbox.compat
cls = ssdbox.compat.install()
from lib.layers import *            # what train.py:16 / evaluate_utils.py:9 do
import lib.layers.box_utils as bu
assert PriorBoxSSD is ssdbox.PriorBoxSSD and DetectOut is ssdbox.DetectOut
assert MultiBoxLoss is cls and issubclass(cls, ssdbox.MultiBoxLoss)
assert bu.nms is ssdbox.box_utils.nms and bu.match is ssdbox.box_utils.match
from lib.utils.config import cfg
crit = MultiBoxLoss(21, 0.5, True, 0, True, 3, 0.5, False, True)     # train.py:99-100
assert crit.variance == list(cfg.MODEL.VARIANCE)
pb = PriorBoxSSD(cfg)                                                 # models/__init__.py:28
assert pb.num_priors == [4, 6, 6, 6, 4, 4]
det = DetectOut(21, 0, 200, 0.01, 0.45, cfg.MODEL.VARIANCE)           # evaluate_utils.py:16-17
# the evaluation solvers eval_solver_factory hands out (lib/utils/__init__.py:4-11)
import lib.utils
from ssdbox import evaluate_utils as EU
ssdbox.compat.install(patch_eval=True)
assert lib.utils.eval_solver_map["VOC0712"] is EU.EvalVOC and lib.utils.eval_solver_map["COCO2014"] is EU.EvalCOCO
import types
loader = types.SimpleNamespace(dataset=types.SimpleNamespace(name="VOC0712", ids=[], image_sets=[("2007", "test")]))
solver = lib.utils.eval_solver_factory(loader, cfg)                   # eval.py:100
assert isinstance(solver, EU.EvalVOC) and solver.detector.top_k == 200 and solver._classes() == list(lib.datasets.VOC_CLASSES)
print("ok")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_eval_post_processing_bit_exact(ref):
    """oracle convert_ssd_result / coco_post_proc == the reference's EvalVOC / EvalCOCO methods
    (lib/utils/evaluate_utils.py:63-68,127-139,175-203), called unbound on a stand-in object."""
    import importlib
    import types
    eu = importlib.import_module("lib.utils.evaluate_utils")
    g = torch.Generator().manual_seed(3)
    B, C, K = 3, 5, 7
    det = torch.zeros(B, C, K, 5)
    for b in range(B):
        for c in range(1, C):
            n = int(torch.randint(0, K + 1, (1,), generator=g))
            det[b, c, :n, 0] = torch.rand(n, generator=g).sort(descending=True).values * 0.98 + 0.01
            det[b, c, :n, 1:] = torch.rand(n, 4, generator=g)
    extra = torch.tensor([[375.0, 500.0], [333.0, 500.0], [480.0, 640.0]])
    scaled = O.rescale_detections(det, extra)
    want = det.clone()
    h = extra[:, 0].unsqueeze(-1).unsqueeze(-1)
    w = extra[:, 1].unsqueeze(-1).unsqueeze(-1)
    want[:, :, :, 1] *= w; want[:, :, :, 3] *= w; want[:, :, :, 2] *= h; want[:, :, :, 4] *= h
    assert torch.equal(scaled, want)
    rv, _ = eu.EvalVOC.convert_ssd_result(None, scaled.clone(), 0)
    assert torch.equal(O.convert_ssd_result(scaled), rv)
    ids = [139, 285, 632]
    holder = types.SimpleNamespace(dataset=types.SimpleNamespace(ids=ids), results=[])
    rc, idt = eu.EvalCOCO.convert_ssd_result(holder, scaled.clone(), 0)
    assert torch.equal(O.convert_ssd_result(scaled, coco_ids=ids), rc)
    eu.EvalCOCO.post_proc(holder, rc.clone(), 0, idt)
    assert (holder.results[0] == O.coco_post_proc(rc).numpy()).all()


def test_head_output_layout_bit_exact(ref):
    """oracle heads_to_rows == what the reference's SSD.forward (lib/models/ssd_v3.py:113-121) returns:
    the real SSD300-VGG16 model (random weights) is run on CPU, its multibox head outputs are captured
    with forward hooks and re-laid out by the oracle."""
    from lib.models import model_factory
    from lib.utils.config import cfg as ref_cfg
    torch.manual_seed(0)
    model, _, _ = model_factory(phase="train", cfg=ref_cfg)
    outs = {"loc": [], "conf": []}
    handles = []
    for name in ("loc", "conf"):
        for layer in getattr(model, name).children():
            handles.append(layer.register_forward_hook(lambda m, i, o, name=name: outs[name].append(o.detach().clone())))
    with torch.no_grad():
        loc, conf = model(torch.randn(2, 3, 300, 300), phase="train")
    for h in handles:
        h.remove()
    assert len(outs["loc"]) == 6 and outs["conf"][0].shape == (2, 84, 38, 38)
    assert torch.equal(O.heads_to_rows(outs["loc"], 4), loc)
    assert torch.equal(O.heads_to_rows(outs["conf"], ref_cfg.MODEL.NUM_CLASSES), conf)


@pytest.mark.parametrize("use07", [True, False])
def test_voc_eval_bit_exact(ref, use07, tmp_path):
    """oracle voc_eval_rows == the reference's write_voc_results_file + voc_eval + voc_ap
    (lib/datasets/voc_eval.py:58-75, 78-106, 109-242) on result files and an annotation cache written
    to a temporary directory: rec, prec and ap bit for bit, with the literal (unstable) argsort and --
    on a tie-free set -- with the canonical stable order the CUDA path implements."""
    import contextlib
    import io
    import numpy as np
    from oracle import voc_oracle as V
    for kw, stable in ((dict(seed=3), False), (dict(seed=4, distinct_scores=True), True)):
        case = synth.gen_voc_eval_case(50, 21, **kw)
        with contextlib.redirect_stdout(io.StringIO()):
            want = ref_loader.voc_eval_reference(case, str(tmp_path), use07)
        got, mean_ap = V.voc_eval_rows(case["rows"], case["seg"], 50, 21, case["gt_boxes"], case["gt_labels"],
                                       case["gt_difficult"], case["gt_offsets"], 0.5, use07, stable=stable)
        assert len(want) == len(got) == 20
        seen = 0
        for (rec, prec, ap), m in zip(want, got):
            if np.ndim(rec) == 0:
                assert rec == prec == ap == -1. and m["ap"] == -1.
                continue
            seen += 1
            assert np.array_equal(rec, m["rec"], equal_nan=True) and np.array_equal(prec, m["prec"])
            assert ap == m["ap"]
        assert seen >= 15
        assert mean_ap == float(np.mean([w[2] for w in want]))


def test_crop_overlaps_bit_exact(ref):
    """oracle jaccard_numpy == lib/utils/augmentations.py:20-37 (float64 truths, int64 rect)."""
    import numpy as np
    import lib.utils.augmentations as aug
    from oracle import voc_oracle as V
    rs = np.random.RandomState(0)
    for _ in range(20):
        G = rs.randint(1, 12)
        xy = rs.rand(G, 2) * 300
        boxes = np.concatenate([xy, xy + rs.rand(G, 2) * 150 + 1], 1)
        rect = np.array([int(rs.uniform(0, 200)), int(rs.uniform(0, 200)), int(rs.uniform(220, 500)), int(rs.uniform(220, 400))])
        assert np.array_equal(aug.jaccard_numpy(boxes, rect), V.jaccard_numpy(boxes, rect))


def test_parse_rec_bit_exact(ref, tmp_path):
    """ssdbox.evaluate_utils.parse_rec == lib/datasets/voc_eval.py:15-33 on a PASCAL VOC annotation file."""
    import lib.datasets.voc_eval as ve
    from ssdbox import evaluate_utils as EU
    xml = """<annotation><folder>VOC2007</folder><filename>000001.jpg</filename><size><width>353</width><height>500</height><depth>3</depth></size>
<object><name>dog</name><pose>Left</pose><truncated>1</truncated><difficult>0</difficult><bndbox><xmin>48</xmin><ymin>240</ymin><xmax>195</xmax><ymax>371</ymax></bndbox></object>
<object><name>person</name><pose>Unspecified</pose><truncated>0</truncated><difficult>1</difficult><bndbox><xmin>8</xmin><ymin>12</ymin><xmax>352</xmax><ymax>498</ymax></bndbox></object>
</annotation>"""
    f = tmp_path / "000001.xml"
    f.write_text(xml)
    assert EU.parse_rec(str(f)) == ve.parse_rec(str(f))
