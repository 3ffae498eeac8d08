"""CPU: the oracle restatement against the golden vectors recorded from the reference
(oracle/make_golden.py).  Runs anywhere -- this is what pins the oracle on the GPU box."""
import numpy as np
import pytest
import torch

from oracle import ssd_oracle as O
from ssdbox import configs
from tests import _util as U

VAR = [0.1, 0.2]


@pytest.fixture(scope="module")
def kat():
    return U.golden("kat.npz")


def test_nms_known_answer(kat):
    b, s = torch.tensor(kat["nms_boxes"]), torch.tensor(kat["nms_scores"])
    k, c = O.greedy_nms(b, s, 0.45, 200)
    assert c == int(kat["nms_count"]) == 3 and k.tolist() == kat["nms_keep"].tolist() == [4, 0, 2, 0, 0]
    k, c = O.greedy_nms(b, s, 0.45, 2)
    assert c == int(kat["nms_count_top2"]) == 2 and k.tolist() == kat["nms_keep_top2"].tolist()


def test_box_algebra_known_answers(kat):
    gt, pr = torch.tensor(kat["iou_gt"]), torch.tensor(kat["iou_prior"])
    assert np.array_equal(O.iou_matrix(gt, O.point_form(pr)).numpy(), kat["iou"])
    assert abs(float(kat["iou"].item()) - 0.592592657) < 1e-8
    enc = O.encode_boxes(gt, pr, VAR)
    assert np.array_equal(enc.numpy(), kat["encode"])
    assert np.array_equal(O.decode_boxes(enc, pr, VAR).numpy(), kat["decode_of_encode"])
    assert np.array_equal(O.decode_boxes(torch.tensor(kat["decode_loc"]), pr, VAR).numpy(), kat["decode"])
    assert np.array_equal(O.log_sum_exp(torch.tensor(kat["lse_x"])).numpy(), kat["lse"])


def test_match_known_answer(kat):
    pri = U.oracle_priors("ssd300_voc")
    m = O.match_image(0.5, torch.tensor(kat["match_truths"]), pri, VAR, torch.tensor(kat["match_labels"]))
    assert np.array_equal(m["conf"].numpy().astype(np.int16), kat["match_conf_t"])
    pos = m["conf"] > 0
    assert int(pos.sum()) == 22 and int(pos.nonzero().sum()) == 167321
    assert m["best_prior"].tolist() == [8134, 8279, 564]
    assert np.array_equal(m["loc"][pos].numpy(), kat["match_loc_t_pos"])
    d = O.match_image(0.5, torch.tensor(kat["dup_truths"]), pri, VAR, torch.tensor(kat["dup_labels"]))
    assert np.array_equal(d["conf"].numpy().astype(np.int16), kat["dup_conf_t"])
    assert int(d["conf"][8074]) == 8      # last truth wins the shared best prior
    assert set(d["conf"].unique().tolist()) == {0, 4, 8}


@pytest.mark.parametrize("name", list(configs.CONFIGS))
def test_priors_golden(kat, name):
    p = U.oracle_priors(name)
    assert U.digest(p) == str(kat["priors_sha_" + name])
    assert np.array_equal(p[:8].numpy(), kat["priors_head_" + name])
    assert np.array_equal(p[-8:].numpy(), kat["priors_tail_" + name])
    assert p.double().sum().item() == float(kat["priors_sum64_" + name])


def test_priors_flip_equivalence():
    """prior_box.py:161-175: flip=True with ar=[2] == flip=False with ar=[2, 1/2] (to 1e-8)."""
    _, c = configs.get("ssd300_voc")
    a = dict(configs.get("ssd300_voc")[0].MODEL)
    b = dict(a)
    a["ASPECT_RATIOS"] = [[2]] * 6
    b["ASPECT_RATIOS"] = [[2, 1 / 2]] * 6
    b["FLIP"] = False
    pa, pb = O.prior_boxes(a, c["layer_dims"]), O.prior_boxes(b, c["layer_dims"])
    assert pa.shape == pb.shape and float((pa - pb).abs().max()) < 1e-8
    assert int((U.oracle_priors("ssd300_voc") == 1.0).sum()) == 444


def test_small_fixture():
    g = U.golden("small790.npz")
    pri = torch.tensor(g["priors"])
    assert torch.equal(pri, O.prior_boxes(U.SMALL_MODEL, U.SMALL_DIMS))
    tg = U.unpack_targets(g["gt"], g["gt_offsets"])
    loc, conf, sc = torch.tensor(g["loc"]), torch.tensor(g["conf"]), torch.tensor(g["scores"])
    d = O.multibox_loss(loc, conf, pri, tg, 21, detail=True)
    assert np.array_equal(d["conf_t"].numpy().astype(np.int16), g["conf_t"])
    assert np.array_equal(d["loc_t"].numpy(), g["loc_t"])
    assert float(d["loss_l"]) == float(g["loss_l"]) and float(d["loss_c"]) == float(g["loss_c"])
    assert np.array_equal(O.detect(loc, sc, pri, 21, top_k=20).numpy(), g["detect_top20"])
    k, c = O.greedy_nms(O.decode_boxes(loc[0], pri, VAR), sc[0, :, 5].contiguous(), 0.45, 50)
    assert c == int(g["nms_count"]) and np.array_equal(k.numpy().astype(np.int32), g["nms_keep"])


@pytest.mark.parametrize("name,B,seed", [("ssd300_voc", 4, 0), ("fssd300_coco", 2, 1), ("ssd512_coco", 2, 2)])
def test_seeded_fixture(name, B, seed):
    g = U.golden("seeded.npz")
    key = "%s_b%d_s%d" % (name, B, seed)
    x = U.seeded_inputs(name, B, seed)
    if U.digest(x["priors"], x["loc"], x["conf"], x["scores"], *x["targets"]) != str(g[key + "_inputs_sha"]):
        pytest.skip("torch RNG stream differs from the one the fixture was recorded with")
    d = O.multibox_loss(x["loc"], x["conf"], x["priors"], x["targets"], x["C"], detail=True)
    assert np.array_equal(d["conf_t"].numpy().astype(np.int8), g[key + "_conf_t"])
    assert U.digest(d["loc_t"]) == str(g[key + "_loc_t_sha"])
    assert [float(d["loss_l"]), float(d["loss_c"])] == g[key + "_loss"].tolist()
    det = O.detect(x["loc"], x["scores"], x["priors"], x["C"])
    nz = det[..., 0] > 0
    assert np.array_equal(nz.sum(-1).numpy().astype(np.int16), g[key + "_det_counts"])
    assert np.array_equal(det[nz].numpy(), g[key + "_det_rows"])


def test_hard_negative_properties():
    x = U.seeded_inputs("ssd300_voc", 3, 5)
    d = O.multibox_loss(x["loc"], x["conf"], x["priors"], x["targets"], x["C"], detail=True)
    npos = d["pos"].sum(1)
    nneg = d["neg"].sum(1)
    assert torch.equal(nneg, torch.clamp(3 * npos, max=x["P"] - 1))
    assert not bool((d["pos"] & d["neg"]).any())
    for b, t in enumerate(x["targets"]):            # every truth keeps at least one positive
        assert int(npos[b]) >= 1
    # selected negatives dominate the rejected ones
    for b in range(3):
        k = d["mining_keys"][b]
        rej = ~(d["neg"][b] | d["pos"][b])
        assert float(k[d["neg"][b]].min()) >= float(k[rej].max())


def test_nms_idempotent_and_empty():
    g = torch.Generator().manual_seed(3)
    xy = torch.rand(300, 2, generator=g) * 0.6
    boxes = torch.cat([xy, xy + torch.rand(300, 2, generator=g) * 0.3 + 0.02], 1)
    scores = torch.rand(300, generator=g)
    k, c = O.greedy_nms(boxes, scores, 0.45, 200)
    kept = k[:c]
    k2, c2 = O.greedy_nms(boxes[kept], scores[kept], 0.45, 200)
    assert c2 == c and k2[:c2].tolist() == list(range(c))
    assert O.greedy_nms(torch.zeros(0, 4), torch.zeros(0), 0.45, 200)[1] == 0
    with pytest.raises(ValueError):
        O.detect(torch.zeros(1, 4, 4), torch.zeros(1, 4, 3), torch.ones(4, 4), 3, nms_thresh=0.0)


def test_empty_truth_image_is_all_background():
    pri = U.oracle_priors("refinedet320_voc")
    m = O.match_image(0.5, torch.zeros(0, 4), pri, VAR, torch.zeros(0))
    assert int(m["conf"].abs().sum()) == 0 and float(m["loc"].abs().sum()) == 0.0


def test_eval_post_processing_golden():
    """oracle convert_ssd_result / coco_post_proc against rows recorded from the reference's
    EvalVOC / EvalCOCO (lib/utils/evaluate_utils.py:63-70,127-139,175-203)."""
    g = U.golden("evalpost.npz")
    det, extra = torch.tensor(g["det"]), torch.tensor(g["extra"])
    scaled = O.rescale_detections(det, extra)
    assert np.array_equal(O.convert_ssd_result(scaled).numpy(), g["voc"])
    coco = O.convert_ssd_result(scaled, coco_ids=g["ids"].tolist())
    assert np.array_equal(coco.numpy(), g["coco"])
    assert np.array_equal(O.coco_post_proc(coco).numpy(), g["coco_rows"])


def test_head_output_layout_golden():
    """oracle heads_to_rows against the loc tensor the reference's SSD300 model returned for the head
    outputs recorded with it (lib/models/ssd_v3.py:113-121)."""
    g = U.golden("heads.npz")
    outs = [torch.tensor(g["loc_in%d" % k]) for k in range(6)]
    assert np.array_equal(O.heads_to_rows(outs, 4).numpy(), g["loc_out"])


def test_voc_eval_golden():
    """oracle voc_eval_rows against rec / prec / ap recorded from the reference's voc_eval
    (lib/datasets/voc_eval.py:109-242; oracle/make_golden_voc.py), both AP metrics."""
    from oracle import voc_oracle as V
    g = U.golden("voceval.npz")
    I, C = len(g["gt_offsets"]) - 1, 21
    for tag, use07 in (("07", True), ("area", False)):
        got, _ = V.voc_eval_rows(g["rows"], g["seg"], I, C, g["gt_boxes"], g["gt_labels"], g["gt_difficult"],
                                 g["gt_offsets"], 0.5, use07)
        assert np.array_equal(np.array([m["ap"] for m in got]), g["ap_" + tag])
    rec = np.concatenate([m["rec"] for m in got if np.ndim(m["rec"])])
    prec = np.concatenate([m["prec"] for m in got if np.ndim(m["prec"])])
    assert np.array_equal(rec, g["rec"], equal_nan=True) and np.array_equal(prec, g["prec"])
    assert [len(m["tp"]) for m in got] == g["count"].tolist()


def test_text_round_trip_is_arithmetic():
    """The CUDA path replaces '{:.3f}' / '{:.1f}' + float() (voc_eval.py:70-74, 170-175) by
    rint(x * 10^d) / 10^d in float64; the two agree bit for bit on float32 inputs, including exact
    halves (x * 10^d is exact in float64, rint and the formatter both round half to even)."""
    rs = np.random.RandomState(0)
    s = np.concatenate([rs.rand(20000).astype(np.float32), np.float32([0.0005, 0.0015, 0.0025, 0.5, 1.0, 0.9995, 0.0104999]),
                        (np.arange(0, 2000, dtype=np.float32) + 0.5) / np.float32(1000)])
    x = np.concatenate([(rs.rand(20000) * 600 - 20).astype(np.float32), np.float32([0.05, 0.15, 0.25, 12.25, 100.75, -0.05, -3.35]),
                        np.arange(0, 4000, dtype=np.float32) / np.float32(20)])
    want_s = np.array([float('{:.3f}'.format(v)) for v in s])
    want_x = np.array([float('{:.1f}'.format(v + 1)) for v in x])
    assert np.array_equal(np.rint(s.astype(np.float64) * 1000.0) / 1000.0, want_s)
    assert np.array_equal(np.rint((x + np.float32(1)).astype(np.float64) * 10.0) / 10.0, want_x)


def test_crop_overlaps_golden():
    from oracle import make_golden_voc as MG
    from oracle import voc_oracle as V
    g = U.golden("voceval.npz")
    boxes, rects = MG.crop_inputs()
    got = np.concatenate([V.jaccard_numpy(bx, rects[b, t]) for b, bx in enumerate(boxes) for t in range(rects.shape[1])])
    assert np.array_equal(got, g["crop_overlap"])
