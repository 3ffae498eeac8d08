"""GPU parity of SURVEY.md 8f rank 4 (through the C ABI): ssdbox_voc_eval against the CPU oracle
(oracle/voc_oracle.py, pinned bit-exact to the reference's lib/datasets/voc_eval.py) and against
rec / prec / ap recorded from the reference itself; ssdbox_crop_overlaps against jaccard_numpy.

Bars: sorted order, tp / fp flags, npos, rec, prec and the 11-point AP bit-exact (float64); the area
AP within 1e-12 relative (np.sum adds pairwise, the kernel in a fixed tree order)."""
import numpy as np
import pytest
import torch

from oracle import voc_oracle as V
from ssdbox import synth
from ssdbox import voc_eval as VE
from tests import _util as U

pytestmark = pytest.mark.gpu
AREA_REL = 1e-12


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _gt(case, dev):
    return VE.VOCGroundTruth(case["gt_boxes"], case["gt_labels"], case["gt_difficult"], case["gt_offsets"], dev)


def _check(case, dev, use07, ovthresh=0.5, rows=None):
    I, C = int(case["num_images"]), int(case["num_classes"])
    r = rows if rows is not None else case["rows"]
    want, want_map = V.voc_eval_rows(r, case["seg"], I, C, case["gt_boxes"], case["gt_labels"], case["gt_difficult"],
                                     case["gt_offsets"], ovthresh, use07, stable=True)
    got = VE.voc_eval(torch.as_tensor(r).to(dev), torch.as_tensor(case["seg"]).to(dev), _gt(case, dev), C, ovthresh, use07)
    assert got.cls_offsets[0] == 0 and got.cls_offsets[C] == r.shape[0]
    bg = int(got.cls_offsets[1])                                  # rows of background segments sort in front
    for c in range(1, C):
        w = want[c - 1]
        assert int(got.npos[c]) == w["npos"]
        n = len(w["tp"])
        assert got.cls_offsets[c + 1] - got.cls_offsets[c] == n
        if n == 0:
            assert got.ap[c] == -1.0 and w["ap"] == -1.0
            continue
        assert np.array_equal(got.rows_of(c).cpu().numpy(), w["rows"][w["order"]])       # the sorted order itself
        assert bg == sum(int(case["seg"][i * C + 1] - case["seg"][i * C]) for i in range(I))
        a, b = got._range(c)
        flags = got.tpfp[a:b].cpu().numpy()
        assert np.array_equal(flags == 1, w["tp"] == 1.0) and np.array_equal(flags == 2, w["fp"] == 1.0)
        assert np.array_equal(got.rec(c).cpu().numpy(), w["rec"], equal_nan=True)         # bit-exact float64
        assert np.array_equal(got.prec(c).cpu().numpy(), w["prec"])
        if use07:
            assert got.ap[c] == w["ap"]                                                   # bit-exact
        else:
            # 1e-12 relative; a class whose truths are all difficult has npos = 0 -> rec = 0/0 and the
            # area metric is NaN in the reference as well
            assert (np.isnan(got.ap[c]) and np.isnan(w["ap"])) or abs(got.ap[c] - w["ap"]) <= AREA_REL * abs(w["ap"])
    if use07:
        assert got.mean_ap == want_map
    return got


@pytest.mark.parametrize("use07", [True, False])
def test_voc_eval_golden_from_reference(dev, use07):
    """rec / prec / ap recorded from the reference's own voc_eval (tests/golden/voceval.npz)."""
    g = U.golden("voceval.npz")
    I, C = len(g["gt_offsets"]) - 1, 21
    gt = VE.VOCGroundTruth(g["gt_boxes"], g["gt_labels"], g["gt_difficult"], g["gt_offsets"], dev)
    got = VE.voc_eval(torch.as_tensor(g["rows"]).to(dev), torch.as_tensor(g["seg"]).to(dev), gt, C, 0.5, use07)
    ap = got.ap[1:]
    if use07:
        assert np.array_equal(ap, g["ap_07"])
        assert np.array_equal(got._rec.cpu().numpy(), g["rec"], equal_nan=True)
        assert np.array_equal(got._prec.cpu().numpy(), g["prec"])
        assert (got.cls_offsets[2:] - got.cls_offsets[1:-1]).tolist() == g["count"].tolist()
    else:
        assert np.all(np.abs(ap - g["ap_area"]) <= AREA_REL * np.abs(g["ap_area"]))


@pytest.mark.parametrize("use07", [True, False])
@pytest.mark.parametrize("seed", [0, 1])
def test_voc_eval_against_oracle_with_ties(dev, seed, use07):
    """400 images x 21 classes: thousands of detections per class, 3-decimal scores tie constantly
    (several radix tiles per class, claims contested between equal-score detections)."""
    case = synth.gen_voc_eval_case(400, 21, seed, fp_max=12)
    assert case["rows"].shape[0] > 4096 * 6
    _check(case, dev, use07)


def test_voc_eval_three_radix_passes_and_wide_rows(dev):
    """81 classes need 17 key bits (three 8-bit passes); rows with 8 columns (the COCO convert layout)."""
    case = synth.gen_voc_eval_case(120, 81, 5, gt_max=10)
    rows8 = np.concatenate([case["rows"], np.full((case["rows"].shape[0], 1), 7.0, np.float32)], 1)
    _check(case, dev, True, rows=rows8)
    _check(case, dev, False, ovthresh=0.3)


def test_voc_eval_edge_cases(dev):
    # images without truths or detections, classes without detections / truths, all-difficult classes
    case = synth.gen_voc_eval_case(30, 6, 9, gt_max=2, fp_max=1, difficult_p=0.6)
    _check(case, dev, True)
    _check(case, dev, False)
    # every score identical: the order is the file order, one claim per truth goes to the first row
    rows = case["rows"].copy()
    rows[:, 4] = 0.5
    _check(case, dev, True, rows=rows)
    # exact duplicates of a truth box and degenerate (zero-area) detections
    rows = case["rows"].copy()
    rows[::3, 2:4] = rows[::3, 0:2]
    _check(case, dev, True, rows=rows)
    # 0/0: a zero-area detection on a zero-area truth gives a NaN overlap, which np.max propagates and
    # `ovmax > ovthresh` rejects -> false positive (voc_eval.py:203-216)
    mini = dict(num_images=1, num_classes=2, rows=np.float32([[9, 9, 9, 9, 0.9, 0, 1], [5, 5, 20, 20, 0.8, 0, 1]]),
                seg=np.int32([0, 0, 2]), gt_boxes=np.float32([[10, 10, 10, 10]]), gt_labels=np.int32([1]),
                gt_difficult=np.uint8([0]), gt_offsets=np.int32([0, 1]))
    got = _check(mini, dev, True)
    assert got.tpfp.cpu().tolist() == [2, 2] and got.ap[1] == 0.0
    # rows in a background (class 0) segment are carried along but never evaluated
    seg0 = case["seg"].astype(np.int64).copy()
    extra_rows = np.float32([[1, 1, 9, 9, 0.7, 0, 0], [2, 2, 8, 8, 0.6, 0, 0]])
    seg0[1:] += 2                                                 # two rows in (image 0, class 0)
    with_bg = dict(case, rows=np.concatenate([extra_rows, case["rows"]], 0), seg=seg0.astype(np.int32))
    got = _check(with_bg, dev, True)
    assert got.cls_offsets[1] == 2 and got.tpfp[:2].cpu().tolist() == [0, 0] and got.ap[0] == -1.0
    # no detections at all: every class reports -1 (voc_eval.py:238-241)
    empty = dict(case, rows=np.zeros((0, 7), np.float32), seg=np.zeros_like(case["seg"]))
    got = _check(empty, dev, True)
    assert np.all(got.ap[1:] == -1.0)
    # scores that do not quantise into [0, 1] are reported, not silently mis-sorted
    bad = case["rows"].copy()
    bad[0, 4] = 1.7
    with pytest.raises(ValueError, match="outside"):
        VE.voc_eval(torch.as_tensor(bad).to(dev), torch.as_tensor(case["seg"]).to(dev), _gt(case, dev), 6)
    with pytest.raises(RuntimeError, match="no CPU path"):
        VE.voc_eval(torch.as_tensor(case["rows"]), torch.as_tensor(case["seg"]), _gt(case, dev), 6)
    # segments that do not cover the rows are reported, not evaluated on uninitialised keys
    with pytest.raises(ValueError, match="does not cover"):
        VE.voc_eval(torch.as_tensor(np.concatenate([case["rows"], case["rows"][:3]], 0)).to(dev), torch.as_tensor(case["seg"]).to(dev), _gt(case, dev), 6)


def test_voc_eval_after_detect_pipeline(dev):
    """DetectOut -> convert_ssd_result -> VOCDetections over two batches -> evaluate_detections, against
    the oracle fed with the same accumulated rows; truths through VOCGroundTruth.from_recs."""
    import ssdbox
    from ssdbox import evaluate_utils as EU
    names = ["cls%02d" % c for c in range(1, 21)]
    acc = VE.VOCDetections(21)
    recs, imagenames = {}, []
    rs = np.random.RandomState(2)
    for batch in range(2):
        x = U.seeded_inputs("ssd300_voc", 3, 10 + batch)
        det = ssdbox.DetectOut(21, 0, 200, 0.01, 0.45, [0.1, 0.2])(x["loc"].to(dev), x["scores"].to(dev), x["priors"].to(dev))
        extra = torch.tensor([[375.0, 500.0], [333.0, 500.0], [480.0, 640.0]], device=dev)
        rows, seg = EU.convert_ssd_result(det, extra)
        acc.add(rows, seg)
        for b in range(3):
            nm = "%06d" % (batch * 3 + b)
            imagenames.append(nm)
            h, w = float(extra[b, 0]), float(extra[b, 1])
            recs[nm] = [{"name": names[int(t[4])], "difficult": int(rs.rand() < 0.2),
                         "bbox": [int(t[0] * w), int(t[1] * h), int(t[2] * w), int(t[3] * h)]} for t in x["targets"][b].tolist()]
    gt = VE.VOCGroundTruth.from_recs(recs, imagenames, names, dev)
    res, mean_ap = VE.evaluate_detections(acc, gt, names)
    rows, seg = acc.flat()
    want, want_map = V.voc_eval_rows(rows.cpu().numpy(), seg.cpu().numpy(), 6, 21, gt.boxes.cpu().numpy(), gt.labels.cpu().numpy(),
                                     gt.difficult.cpu().numpy(), gt.offsets.cpu().numpy())
    assert mean_ap == want_map
    for (name, ap, prec, rec), w in zip(res, want):
        assert ap == w["ap"] and np.array_equal(prec, w["prec"]) and np.array_equal(rec, w["rec"], equal_nan=True)


def test_crop_overlaps(dev):
    """jaccard_numpy + the centre mask of RandomSampleCrop (augmentations.py:13-37, 250-268): golden
    overlaps recorded from the reference, and min / max / mask against the oracle."""
    from oracle import make_golden_voc as MG
    g = U.golden("voceval.npz")
    boxes, rects = MG.crop_inputs()
    boxes.append(np.zeros((0, 4)))                                       # an image without truths
    rects = np.concatenate([rects, rects[:1]], 0)
    ov, mm, mask = VE.crop_overlaps(boxes, torch.as_tensor(rects).to(dev))
    got = np.concatenate([o.cpu().numpy().reshape(-1) for o in ov])
    assert np.array_equal(got, g["crop_overlap"])                        # bit-exact float64
    mm = mm.cpu().numpy()
    for b, bx in enumerate(boxes[:-1]):
        for t in range(rects.shape[1]):
            o, lo, hi, m = V.crop_trial(bx, rects[b, t])
            assert mm[b, t, 0] == lo and mm[b, t, 1] == hi
            assert np.array_equal(mask[b][t].cpu().numpy(), m.astype(bool))
    assert np.isinf(mm[-1]).all()


def _write_voc_xml(path, objs):
    rows = "".join(
        "<object><name>%s</name><pose>Unspecified</pose><truncated>0</truncated><difficult>%d</difficult>"
        "<bndbox><xmin>%d</xmin><ymin>%d</ymin><xmax>%d</xmax><ymax>%d</ymax></bndbox></object>" % ((o[0], o[1]) + tuple(o[2]))
        for o in objs)
    with open(path, "w") as f:
        f.write("<annotation>%s</annotation>" % rows)


def test_eval_solvers_end_to_end(dev, tmp_path):
    """EvalVOC.validate (lib/utils/evaluate_utils.py:41-78, 114-162) with a stand-in network, data loader
    and PASCAL VOC annotation files: network outputs -> DetectOut -> result rows -> accumulation ->
    ssdbox_voc_eval, against the CPU oracle chain (detect, rescale, convert_ssd_result, voc_eval_rows).
    EvalCOCO's accumulated result rows against the oracle's post_proc rows."""
    import types
    import ssdbox
    from oracle import ssd_oracle as O
    from ssdbox import configs
    from ssdbox import evaluate_utils as EU
    pri = O.prior_boxes(U.SMALL_MODEL, U.SMALL_DIMS)
    P, C, B, nb = pri.size(0), 21, 3, 2
    cfg = configs.AttrDict(MODEL=U.SMALL_MODEL)
    rs = np.random.RandomState(4)
    batches, outputs, ids, all_det = [], [], [], []
    (tmp_path / "Annotations").mkdir()
    gtb, gtl, gtd, goff = [], [], [], [0]
    for k in range(nb):
        loc = synth.gen_loc(B, P, 40 + k)
        sc = synth.gen_detect_scores(B, P, C, 40 + k, bkg_bias=6.0)
        extra = torch.tensor([[375.0, 500.0], [333.0, 500.0], [480.0, 640.0]])
        batches.append((torch.zeros(B, 3, 8, 8), None, extra))
        outputs.append((loc, sc))
        det = O.detect(loc, sc, pri, C)
        all_det.append(O.convert_ssd_result(O.rescale_detections(det, extra)).numpy())
        for b in range(B):
            name = "%06d" % (k * B + b)
            ids.append((str(tmp_path), name))
            # truths: a few of the image's own detections (so that there are true positives) + a difficult one
            rows = all_det[-1][all_det[-1][:, 5] == b]
            pick = rows[rs.permutation(len(rows))[:4]] if len(rows) else np.zeros((0, 7))
            objs = [(EU.VOC_CLASSES[int(r[6]) - 1], int(j == 3), [int(r[0]) + 1, int(r[1]) + 1, int(r[2]) + 2, int(r[3]) + 2]) for j, r in enumerate(pick)]
            _write_voc_xml(str(tmp_path / "Annotations" / (name + ".xml")), objs)
            for o in objs:
                gtb.append([v - 1 for v in o[2]]); gtl.append(EU.VOC_CLASSES.index(o[0]) + 1); gtd.append(o[1])
            goff.append(len(gtl))
    dataset = types.SimpleNamespace(name="VOC0712", ids=ids, image_sets=[("2007", "test")],
                                    _annopath="%s/Annotations/%s.xml")          # voc0712.py:100
    calls = iter(outputs)

    class Loader(object):
        def __init__(self):
            self.dataset = dataset

        def __iter__(self):
            return iter(batches)

    def net(images, phase="eval"):
        loc, sc = next(calls)
        return loc.to(images.device), sc.to(images.device)

    solver = EU.EvalVOC(Loader(), cfg, output_dir=str(tmp_path / "out"))
    res, maps = solver.validate(net, pri.to(dev))
    # oracle chain on the same data
    seg, rows = [0], []
    for k in range(nb):
        d = all_det[k]
        for b in range(B):
            for c in range(C):
                r = d[(d[:, 5] == b) & (d[:, 6] == c)]
                rows.append(r)
                seg.append(seg[-1] + len(r))
    rows = np.concatenate(rows, 0)
    want, want_map = V.voc_eval_rows(rows, np.asarray(seg), nb * B, C, np.float32(gtb).reshape(-1, 4), np.int32(gtl), np.uint8(gtd), np.int32(goff))
    assert len(res) == 20 and maps == [want_map]
    hit = 0
    for (name, ap, prec, rec), w in zip(res, want):
        assert ap == w["ap"]
        if w["ap"] != -1.0:
            hit += 1
            assert np.array_equal(prec, w["prec"]) and np.array_equal(rec, w["rec"], equal_nan=True)
    assert hit >= 5 and max(x[1] for x in res) > 0.0
    import pickle
    with open(str(tmp_path / "out" / (res[0][0] + "_pr.pkl")), "rb") as f:
        assert pickle.load(f)["ap"] == res[0][1]
    # COCO: accumulated result rows
    coco_ids = [100 + 7 * i for i in range(nb * B)]
    calls = iter(outputs)
    cds = types.SimpleNamespace(name="COCO2014", ids=coco_ids, image_sets=None)

    class CLoader(Loader):
        def __init__(self):
            self.dataset = cds

    cs = EU.EvalCOCO(CLoader(), configs.AttrDict(MODEL=U.SMALL_MODEL, DATASET=configs.AttrDict(NUM_EVAL_PICS=0)))
    cs.reset_results()
    idx = 0
    for (images, _, extra), (loc, sc) in zip(batches, outputs):
        det = cs.detector(loc.to(dev), sc.to(dev), pri.to(dev))
        idx = cs.consume(det, extra.to(dev), idx)
    got = cs.result_rows().cpu()
    want_rows = torch.cat([O.coco_post_proc(O.convert_ssd_result(O.rescale_detections(O.detect(loc, sc, pri, C), batches[k][2]),
                                                                coco_ids=coco_ids[k * B:(k + 1) * B])) for k, (loc, sc) in enumerate(outputs)], 0)
    assert idx == nb * B and torch.equal(got[:, [0, 5, 6]], want_rows[:, [0, 5, 6]])
    assert float((got - want_rows).abs().max()) <= 1e-5 * 640
