"""GPU parity, second set: the shapes and semantics the first round left untested.

* backward at the headline shape (SSD512-COCO, B=64) and full batches of cfg4 / cfg5,
* RefineDet split into (a) the refinement arithmetic (1e-5) and (b) the ODM logic on the ORACLE's refined
  anchors / ARM mask (class targets, negative sets, keep-lists bit-exact),
* zero-area truths in the matching, 3-D priors, empty shards in the peer exchange, the deferred-wait guard,
  one module driven from two streams.
Bars as in test_gpu_parity.py: indices / labels / sets / keep-lists bit-exact, floats 1e-5 relative.
"""
import pytest
import torch

import ssdbox
from oracle import ssd_oracle as O
from ssdbox import box_utils as BU
from ssdbox import configs, synth
from tests import _util as U
from tests.test_gpu_parity import _check_neg_sets, _compare_detect, _gpu_targets

pytestmark = pytest.mark.gpu
VAR = [0.1, 0.2]
REL = 1e-5


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


# ------------------------------------------------------------------------------------------------
# a11 at the headline shape: loss_bwd_stream_kernel<81> with 148 persistent CTAs, many tiles per warp
# ------------------------------------------------------------------------------------------------
def test_full_size_ssd512_coco_backward(dev):
    cfg, c = configs.get("ssd512_coco")
    B, P, C = 64, 24564, 81
    pri = U.oracle_priors("ssd512_coco")
    tg = synth.gen_targets(B, C, 32, 0)
    loc_h = synth.gen_loc(B, P, 0)
    conf_h = synth.gen_train_logits(B, P, C, 0)
    loc = loc_h.to(dev).requires_grad_(True)
    conf = conf_h.to(dev).requires_grad_(True)
    crit = ssdbox.MultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False)
    ll, lc = crit((loc, conf, pri.to(dev)), _gpu_targets(tg, dev))
    (ll + lc).backward()
    sums, sel, _ = crit._last
    n_full = float(sums[2])
    chosen = sel >= 0
    # rows outside pos U neg carry exactly zero gradient (multibox_loss.py:106-110 never reads them)
    gc = conf.grad
    assert float(gc[~chosen].abs().max()) == 0.0
    assert float(loc.grad[sel <= 0].abs().max()) == 0.0
    # every selected row: softmax - onehot sums to ~0 and is non-zero
    rs = gc[chosen]
    assert float(rs.sum(-1).abs().max()) < 1e-6 and bool((rs.abs().sum(-1) > 0).all())
    # oracle autograd on a 3-image subset; per-image gradients only differ by the normaliser N
    sub = [0, 31, 63]
    r = O.multibox_loss(loc_h[sub], conf_h[sub], pri, [tg[i] for i in sub], C, detail=True)
    n_sub = float(r["n"])
    gl, gcr = O.multibox_loss_grads(loc_h[sub], conf_h[sub], pri, [tg[i] for i in sub], C)
    same = torch.equal((sel[sub] >= 0).cpu(), r["pos"] | r["neg"])
    scale = n_full / n_sub
    U.assert_close_rel(loc.grad[sub].cpu() * scale, gl, REL, 1e-8, "grad_loc subset")      # 1e-5 relative
    if same:
        U.assert_close_rel(gc[sub].cpu() * scale, gcr, REL, 1e-8, "grad_conf subset")      # 1e-5 relative
    else:       # a mining key within 2e-6 of the threshold flipped: compare the rows both selected
        both = ((sel[sub] >= 0).cpu() & (r["pos"] | r["neg"]))
        U.assert_close_rel((gc[sub].cpu() * scale)[both], gcr[both], REL, 1e-8, "grad_conf subset (common rows)")
    # the last image's last tile (ragged tail of the persistent partition) is covered by `sub`


@pytest.mark.parametrize("name,B,seed", [("fssd300_coco", 32, 40), ("refinedet320_voc", 32, 41), ("ssd300_voc", 32, 42)])
def test_full_batch_small_configs(dev, name, B, seed):
    """cfg1 / cfg4 / cfg5 at their full batch (the multi-tile / persistent paths change with B)."""
    x = U.seeded_inputs(name, B, seed)
    loc = x["loc"].to(dev).requires_grad_(True)
    conf = x["conf"].to(dev).requires_grad_(True)
    crit = ssdbox.MultiBoxLoss(x["C"], 0.5, True, 0, True, 3, 0.5, False)
    d = crit.intermediates((loc, conf, x["priors"].to(dev)), _gpu_targets(x["targets"], dev))
    r = O.multibox_loss(x["loc"], x["conf"], x["priors"], x["targets"], x["C"], detail=True)
    n_diff = _check_neg_sets(d, r, x["P"])
    U.assert_close_rel(d["loss_l"], r["loss_l"], REL, 0, "loss_l")
    U.assert_close_rel(d["loss_c"], r["loss_c"], REL, 0, "loss_c")
    U.assert_close_rel(d["loc_t"], r["loc_t"], REL, 1e-6, "loc_t")
    (d["loss_l"] + d["loss_c"]).backward()
    gl, gc = O.multibox_loss_grads(x["loc"], x["conf"], x["priors"], x["targets"], x["C"])
    U.assert_close_rel(loc.grad, gl, REL, 1e-8, "grad_loc")
    if n_diff == 0:
        U.assert_close_rel(conf.grad, gc, REL, 1e-8, "grad_conf")
    out = ssdbox.DetectOut(x["C"], 0, 200, 0.01, 0.45, VAR)(x["loc"].to(dev), x["scores"].to(dev), x["priors"].to(dev))
    sub = [0, B // 2, B - 1]
    _compare_detect(out[sub].cpu(), O.detect(x["loc"][sub], x["scores"][sub], x["priors"], x["C"]), name + " detect subset")


# ------------------------------------------------------------------------------------------------
# a-R RefineDet, split: refinement arithmetic (1e-5) | ODM logic on the oracle's anchors (bit-exact)
# ------------------------------------------------------------------------------------------------
def _refine_case(B, seed):
    pri = U.oracle_priors("refinedet320_voc")
    P, C = pri.size(0), 21
    tg = synth.gen_targets(B, C, 16, seed)
    arm_loc, arm_conf = synth.gen_arm_outputs(B, P, seed)
    odm_loc = synth.gen_loc(B, P, seed + 1)
    odm_conf = synth.gen_train_logits(B, P, C, seed + 2)
    sc = synth.gen_detect_scores(B, P, C, seed + 3, bkg_bias=8.0)
    return pri, P, C, tg, arm_loc, arm_conf, odm_loc, odm_conf, sc


def test_refine_anchor_arithmetic(dev):
    """The only floating-point step of the two-step path that is not shared with SSD: decode(arm_loc, priors)
    -> xyxy and centre form (1e-5 relative), objectness softmax(arm_conf)[:,1] > theta (flips only where the
    objectness is within 1e-6 relative of theta: expf ulp)."""
    pri, P, C, tg, arm_loc, arm_conf, *_ = _refine_case(8, 50)
    xy, cf = ssdbox.refine_anchors(arm_loc.to(dev), pri.to(dev))
    oxy, ocf = O.refine_anchors(arm_loc, pri)
    U.assert_close_rel(xy, oxy, REL, 1e-6, "refined xyxy")
    U.assert_close_rel(cf, ocf, REL, 1e-6, "refined centre form")
    keep = ssdbox.arm_filter(arm_conf.to(dev), 0.01).cpu().bool()
    obj = O.arm_objectness(arm_conf)
    flips = keep != (obj > 0.01)
    assert float(((obj[flips] - 0.01).abs() / 0.01).max()) < 1e-6 if bool(flips.any()) else True
    assert 0.05 < float((~keep).float().mean()) < 0.6          # the filter is exercised


@pytest.mark.parametrize("B,seed", [(6, 60), (32, 61)])
def test_refinedet_odm_logic_bit_exact_on_oracle_anchors(dev, B, seed):
    """ODM loss and RefineDetectOut fed with the ORACLE's refined anchors and ARM mask: what is left is the
    reference's own match / mining / NMS logic (box_utils.py:92-133, multibox_loss.py:97-103, box_utils.py:279-343
    with per-image anchors and a pool), so class targets, matched truths, negative sets and keep-lists must
    be bit-exact and the losses within 1e-5."""
    pri, P, C, tg, arm_loc, arm_conf, odm_loc, odm_conf, sc = _refine_case(B, seed)
    oxy, ocf = O.refine_anchors(arm_loc, pri)
    pool = O.arm_objectness(arm_conf) > 0.01
    r = O.refine_multibox_loss(arm_loc, arm_conf, odm_loc, odm_conf, pri, tg, C, use_arm=True, detail=True,
                               anchors=(oxy, ocf), pool=pool)
    crit = ssdbox.RefineMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, use_ARM=True)
    gt, offs, gmax = ssdbox.pack_targets(_gpu_targets(tg, dev), dev)
    crit._debug = dict(conf_t=torch.empty(B, P, dtype=torch.int64, device=dev),
                       loc_t=torch.empty(B, P, 4, dtype=torch.float32, device=dev),
                       neg=torch.empty(B, P, dtype=torch.uint8, device=dev),
                       keys=torch.empty(B, P, dtype=torch.float32, device=dev))
    ll, lc = crit.forward_packed(odm_loc.to(dev), odm_conf.to(dev), ocf.to(dev), gt, offs, gmax,
                                 anchors_xyxy=oxy.to(dev), pool=pool.to(torch.uint8).to(dev))
    d = crit._debug
    crit._debug = None
    sums, sel, tidx = crit._last
    assert torch.equal(d["conf_t"].cpu(), r["conf_t_raw"])                      # match labels: bit-exact
    pos_gpu = (sel > 0).cpu()
    assert torch.equal(pos_gpu, r["pos"])                                      # positives inside the pool: bit-exact
    assert torch.equal(sel.cpu().long()[pos_gpu], r["conf_t"][pos_gpu])
    n_diff = _check_neg_sets(dict(conf_t=r["conf_t_raw"], neg=d["neg"]), dict(conf_t=r["conf_t_raw"], neg=r["neg"], mining_keys=
                             torch.where(pool, r["mining_keys"], torch.full_like(r["mining_keys"], float("-inf")))), P)
    assert not bool((d["neg"].cpu().bool() & ~pool).any())                     # filtered anchors are never mined
    assert int(sums[2]) == int(r["n"])
    U.assert_close_rel(d["loc_t"].cpu()[r["pos"]], r["loc_t"][r["pos"]], REL, 1e-6, "ODM loc_t")   # 1e-5 relative
    U.assert_close_rel(ll, r["loss_l"], REL, 0, "ODM loss_l")                  # 1e-5 relative
    U.assert_close_rel(lc, r["loss_c"], REL, 0, "ODM loss_c")
    # inference on the same anchors / mask: keep-lists bit-exact
    det = ssdbox.DetectOut(C, 0, 200, 0.01, 0.45, VAR)
    out = det.forward(odm_loc.to(dev), sc.to(dev), ocf.to(dev), score_keep=pool.to(torch.uint8).to(dev)).cpu()
    ref = O.detect(odm_loc, sc, pri, C, score_mask=pool, anchors_center=ocf)
    _compare_detect(out, ref, "RefineDetectOut on oracle anchors")
    assert int((out[..., 0] > 0).sum()) > 0


def test_refinedet_end_to_end_modules(dev):
    """The public two-step modules (their own refinement on the GPU) against the oracle end to end: losses
    within 5e-5 (the refined anchors differ by expf ulps, which moves IoUs at the 1e-7 level), at most a handful
    of detections differ."""
    B = 8
    pri, P, C, tg, arm_loc, arm_conf, odm_loc, odm_conf, sc = _refine_case(B, 70)
    preds = tuple(t.to(dev) for t in (arm_loc, arm_conf, odm_loc, odm_conf, pri))
    gtg = _gpu_targets(tg, dev)
    arm = ssdbox.RefineMultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, use_ARM=False)
    ll, lc = arm(preds, gtg)
    rl, rc = O.refine_multibox_loss(arm_loc, arm_conf, odm_loc, odm_conf, pri, tg, 2, use_arm=False)
    U.assert_close_rel(ll, rl, REL, 0, "ARM loss_l")
    U.assert_close_rel(lc, rc, REL, 0, "ARM loss_c")
    odm = ssdbox.RefineMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, use_ARM=True)
    ll, lc = odm(preds, gtg)
    rl, rc = O.refine_multibox_loss(arm_loc, arm_conf, odm_loc, odm_conf, pri, tg, C, use_arm=True)
    U.assert_close_rel(ll, rl, 5e-5, 0, "ODM loss_l")
    U.assert_close_rel(lc, rc, 5e-5, 0, "ODM loss_c")
    det = ssdbox.RefineDetectOut(C, 0, 200, 0.01, 0.45, VAR, theta=0.01)
    out = det(arm_loc.to(dev), arm_conf.to(dev), odm_loc.to(dev), sc.to(dev), pri.to(dev)).cpu()
    ref = O.refine_detect(arm_loc, arm_conf, odm_loc, sc, pri, C)
    assert int((out[..., 0] != ref[..., 0]).sum()) <= 4
    # backward of the ODM loss reaches odm_loc / odm_conf only
    ol = odm_loc.to(dev).requires_grad_(True)
    oc = odm_conf.to(dev).requires_grad_(True)
    ll, lc = odm((arm_loc.to(dev), arm_conf.to(dev), ol, oc, pri.to(dev)), gtg)
    (ll + lc).backward()
    assert float(ol.grad.abs().sum()) > 0 and float(oc.grad.abs().sum()) > 0


# ------------------------------------------------------------------------------------------------
# semantics at the edges
# ------------------------------------------------------------------------------------------------
def test_match_zero_area_truths(dev):
    """A truth with zero width / height has IoU 0 with every (positive-area) prior (box_utils.py:63-70: 0 / area):
    its best prior is index 0 (first maximum, :116) and is force-matched (:123-127).  Bit-exact with the oracle.
    (0/0 = NaN needs a zero-area PRIOR under a zero-area truth; PriorBoxSSD never emits one -- the kernels define
    IoU := 0 there, documented in DESIGN.md.)"""
    pri = U.oracle_priors("ssd300_voc")
    tg = [torch.tensor([[0.30, 0.30, 0.30, 0.60, 4.0], [0.1, 0.1, 0.5, 0.6, 2.0]]),      # zero width + a normal truth
          torch.tensor([[0.5, 0.5, 0.5, 0.5, 9.0]]),                                      # a point
          torch.tensor([[0.2, 0.7, 0.6, 0.7, 1.0], [0.2, 0.7, 0.6, 0.7, 3.0], [0.25, 0.2, 0.7, 0.8, 5.0]])]
    gt, offs = synth.pack_targets(tg)
    loc_t, conf_t, midx, ov = BU.match_batch(0.5, gt.to(dev), offs.to(dev), 3, pri.to(dev), VAR, want_overlap=True)
    for b, t in enumerate(tg):
        m = O.match_image(0.5, t[:, :4], pri, VAR, t[:, 4])
        assert torch.equal(conf_t[b].cpu(), m["conf"]), b
        assert torch.equal(midx[b].cpu().long(), m["truth_idx"]), b
        assert torch.equal(ov[b].cpu(), m["overlap"]), b
    assert int(conf_t[1, 0]) == 10 and int((conf_t[1] > 0).sum()) == 1      # the point truth claims prior 0 only
    # the fused path (matching inside loss_stream) agrees
    loc = synth.gen_loc(3, pri.size(0), 1)
    conf = synth.gen_train_logits(3, pri.size(0), 21, 1)
    crit = ssdbox.MultiBoxLoss(21, 0.5, True, 0, True, 3, 0.5, False)
    d = crit.intermediates((loc.to(dev), conf.to(dev), pri.to(dev)), _gpu_targets(tg, dev))
    assert torch.equal(d["conf_t"], conf_t)


def test_priors_3d_shapes(dev):
    x = U.seeded_inputs("ssd300_voc", 2, 3)
    loc, conf, pri, sc = x["loc"].to(dev), x["conf"].to(dev), x["priors"].to(dev), x["scores"].to(dev)
    crit = ssdbox.MultiBoxLoss(x["C"], 0.5, True, 0, True, 3, 0.5, False)
    gt, offs, gmax = ssdbox.pack_targets(_gpu_targets(x["targets"], dev), dev)
    a = crit.forward_packed(loc, conf, pri, gt, offs, gmax)
    b = crit.forward_packed(loc, conf, pri.unsqueeze(0), gt, offs, gmax)                  # [1,P,4] as the reference docstring says
    c = crit.forward_packed(loc, conf, pri.unsqueeze(0).expand(2, -1, -1).contiguous(), gt, offs, gmax)   # per-image priors
    assert float(a[0]) == float(b[0]) == float(c[0]) and float(a[1]) == float(b[1]) == float(c[1])
    with pytest.raises(ValueError):
        crit.forward_packed(loc, conf, pri.unsqueeze(0).expand(3, -1, -1).contiguous(), gt, offs, gmax)
    det = ssdbox.DetectOut(x["C"], 0, 200, 0.01, 0.45, VAR)
    o1 = det(loc, sc, pri).clone()
    assert torch.equal(det(loc, sc, pri.unsqueeze(0)), o1)
    with pytest.raises(ValueError):
        det(loc, sc, pri.unsqueeze(0).expand(3, -1, -1).contiguous())


def test_pack_targets_host_and_device(dev):
    """host targets travel as one pinned staging copy, device targets as one gather: same packed layout."""
    tg = synth.gen_targets(5, 21, 9, 3)
    tg[2] = torch.tensor([-1.0])                           # multibox_loss_v1.py:70 sentinel
    g1, o1, m1 = ssdbox.pack_targets(tg, dev)
    g2, o2, m2 = ssdbox.pack_targets([t.to(dev) for t in tg], dev)
    assert m1 == m2 and torch.equal(o1, o2) and torch.equal(g1[:int(o1[-1])], g2[:int(o2[-1])])
    assert g1.data_ptr() % 16 == 0 and o1.dtype == torch.int32
    ref, roffs = synth.pack_targets([t if t.dim() == 2 else torch.zeros(0, 5) for t in tg])
    assert torch.equal(g1[:ref.size(0)].cpu(), ref) and torch.equal(o1.cpu(), roffs)


def test_peer_exchange_empty_shard_and_pending_guard(dev):
    """(1) a rank whose local shard is empty (global batch < world) still posts zeros and keeps its epoch in
    step; (2) a second deferred forward before PendingLoss.wait() raises instead of desynchronising the banks."""
    x = U.seeded_inputs("ssd300_voc", 2, 5)
    crit = ssdbox.MultiBoxLoss(x["C"], 0.5, True, 0, True, 3, 0.5, False)
    crit.use_local_peer_exchange(dev)
    pri = x["priors"].to(dev)
    P, C = x["P"], x["C"]
    e_loc = torch.zeros(0, P, 4, device=dev)
    e_conf = torch.zeros(0, P, C, device=dev)
    gt0 = torch.zeros(1, 5, device=dev)
    offs0 = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.no_grad():
        ll, lc = crit.forward_packed(e_loc, e_conf, pri, gt0, offs0, 0)
    assert float(ll) == 0.0 and float(lc) == 0.0 and int(crit._peers.buf[0]) == 1
    pend = crit.forward_packed_deferred(e_loc, e_conf, pri, gt0, offs0, 0)
    a, b = pend.wait()
    assert float(a) == 0.0 and float(b) == 0.0 and int(crit._peers.buf[0]) == 2
    # a non-empty call afterwards still matches the plain module (epochs in step)
    plain = ssdbox.MultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False)
    gt, offs, gmax = ssdbox.pack_targets(_gpu_targets(x["targets"], dev), dev)
    with torch.no_grad():
        want = plain.forward_packed(x["loc"].to(dev), x["conf"].to(dev), pri, gt, offs, gmax)
        got = crit.forward_packed(x["loc"].to(dev), x["conf"].to(dev), pri, gt, offs, gmax)
    assert float(want[0]) == float(got[0]) and float(want[1]) == float(got[1])
    # guard
    pend = crit.forward_packed_deferred(x["loc"].to(dev), x["conf"].to(dev), pri, gt, offs, gmax)
    with pytest.raises(RuntimeError):
        crit.forward_packed_deferred(x["loc"].to(dev), x["conf"].to(dev), pri, gt, offs, gmax)
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            crit.forward_packed(x["loc"].to(dev), x["conf"].to(dev), pri, gt, offs, gmax)
    dl, dc = pend.wait()
    assert float(dl) == float(want[0]) and float(dc) == float(want[1])
    assert crit._peers.timeouts() == 0 if hasattr(crit._peers, "timeouts") else True


def test_one_module_two_streams(dev):
    """Per-(device, stream) workspaces: the same DetectOut / MultiBoxLoss objects driven from two streams at once
    give the single-stream results."""
    xa = U.seeded_inputs("ssd300_voc", 4, 80)
    xb = U.seeded_inputs("ssd300_voc", 4, 81)
    det = ssdbox.DetectOut(xa["C"], 0, 200, 0.01, 0.45, VAR)
    crit = ssdbox.MultiBoxLoss(xa["C"], 0.5, True, 0, True, 3, 0.5, False)
    ins = []
    for x in (xa, xb):
        gt, offs, gmax = ssdbox.pack_targets(_gpu_targets(x["targets"], dev), dev)
        ins.append((x["loc"].to(dev), x["scores"].to(dev), x["conf"].to(dev), x["priors"].to(dev), gt, offs, gmax))
    want = []
    with torch.no_grad():
        for loc, sc, conf, pri, gt, offs, gmax in ins:
            want.append((det(loc, sc, pri).clone(), torch.stack(crit.forward_packed(loc, conf, pri, gt, offs, gmax)).clone()))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for rep in range(5):
        got = []
        for s, (loc, sc, conf, pri, gt, offs, gmax) in zip(streams, ins):
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s), torch.no_grad():
                got.append((det(loc, sc, pri), torch.stack(crit.forward_packed(loc, conf, pri, gt, offs, gmax))))
        torch.cuda.synchronize()
        for (o, l), (wo, wl) in zip(got, want):
            assert torch.equal(o, wo) and torch.equal(l, wl)


# ------------------------------------------------------------------------------------------------
# self-cleaning workspaces (SSDBOX_LOSS_WS_CLEAN / SSDBOX_DETECT_WS_CLEAN): from the second call on a module
# skips its init launch; every call must therefore leave the state exactly as the init kernel would
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flags", [0, 1, 2, 4, 3])       # fused / separate matching, generic mining, no cluster
def test_loss_workspace_clean_after_every_call(dev, flags):
    from ssdbox import _abi
    name, B = "ssd300_voc", 5
    pri = U.oracle_priors(name).to(dev)
    C = 21
    reused = ssdbox.MultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False)
    reused.abi_flags = flags
    for it, (seed, gtm) in enumerate([(1, 16), (2, 16), (3, 9), (4, 16), (5, 150), (6, 16)]):
        tg = synth.gen_targets(B, C, gtm, seed, gt_min=max(1, gtm // 2))
        if it == 3:
            tg[1] = torch.zeros(0, 5)                      # an image without truths
            tg[2] = torch.cat([tg[2], tg[2][:1]], 0)       # duplicate truth (shared best prior)
        loc = synth.gen_loc(B, 8732, seed).to(dev)
        conf = synth.gen_train_logits(B, 8732, C, seed).to(dev)
        fresh = ssdbox.MultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False)
        fresh.abi_flags = flags
        want = fresh.intermediates((loc, conf, pri), _gpu_targets(tg, dev))
        got = reused.intermediates((loc, conf, pri), _gpu_targets(tg, dev))
        for k in ("conf_t", "neg", "sel", "tidx", "keys", "sums", "loc_t"):
            assert torch.equal(want[k], got[k]), (it, k)
    tags = [t for t in reused._state.ws.tags.values() if t is not None]
    assert tags and tags[0][0] == "loss"


def test_detect_workspace_clean_after_every_call(dev):
    name = "ssd300_voc"
    pri = U.oracle_priors(name).to(dev)
    P, C, B = 8732, 21, 2
    reused = ssdbox.DetectOut(C, 0, 200, 0.01, 0.45, VAR)
    reused_lg = ssdbox.DetectOut(C, 0, 200, 0.01, 0.45, VAR, conf_is_logits=True)
    g = torch.Generator().manual_seed(5)
    for it, bias in enumerate([10.0, 1.0, 10.0, 6.0, 1.0, 9.0, 10.0]):     # sparse, dense (overflow), sparse, CTA-wide lists, ...
        logits = torch.randn(B, P, C, generator=g)
        logits[..., 0] += bias
        if it == 3:
            logits[..., 5] += 5.0                         # one class overflows while the others stay sparse
        sc = torch.softmax(logits, -1).to(dev)
        loc = synth.gen_loc(B, P, it).to(dev)
        want = ssdbox.DetectOut(C, 0, 200, 0.01, 0.45, VAR)(loc, sc, pri)
        got = reused(loc, sc, pri)
        assert torch.equal(want, got), it
        want = ssdbox.DetectOut(C, 0, 200, 0.01, 0.45, VAR, conf_is_logits=True)(loc, logits.to(dev), pri)
        got = reused_lg(loc, logits.to(dev), pri)
        assert torch.equal(want, got), ("logits", it)
    # the counters really are zero after a call (what SSDBOX_DETECT_WS_CLEAN promises)
    torch.cuda.synchronize()
    ws = reused._ws.buf
    assert int(ws[:(B * C + 4) * 4].view(torch.int32).abs().sum()) == 0


def test_deferred_loss_completed_by_detect(dev):
    """ssdbox_detect_peers: the cross-rank wait of a deferred loss forward rides on the last Detect kernel (world of
    one rank here; tests/_mgpu_worker.py runs it across GPUs).  Same losses as the plain module, same detections,
    over several steps (both slot banks) and under CUDA-graph replay."""
    x = U.seeded_inputs("ssd300_voc", 4, 90)
    loc, conf, pri, sc = x["loc"].to(dev), x["conf"].to(dev), x["priors"].to(dev), x["scores"].to(dev)
    gt, offs, gmax = ssdbox.pack_targets(_gpu_targets(x["targets"], dev), dev)
    plain = ssdbox.MultiBoxLoss(x["C"], 0.5, True, 0, True, 3, 0.5, False)
    with torch.no_grad():
        wl, wc = plain.forward_packed(loc, conf, pri, gt, offs, gmax)
    det = ssdbox.DetectOut(x["C"], 0, 200, 0.01, 0.45, VAR)
    want = det(loc, sc, pri).clone()
    crit = ssdbox.MultiBoxLoss(x["C"], 0.5, True, 0, True, 3, 0.5, False)
    crit.use_local_peer_exchange(dev)

    def step():
        with torch.no_grad():
            pend = crit.forward_packed_deferred(loc, conf, pri, gt, offs, gmax)
            out = det.forward(loc, sc, pri, pending=pend)
            ll, lc = pend.wait()
        return ll, lc, out
    for it in range(3):
        ll, lc, out = step()
        assert float(ll) == float(wl) and float(lc) == float(wc), it
        assert torch.equal(out, want)
    assert int(crit._peers.buf[0]) == 3
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
        step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        ll, lc, out = step()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert float(ll) == float(wl) and float(lc) == float(wc) and torch.equal(out, want)
    assert int(crit._peers.buf[0]) == 8          # 3 eager + 2 warm-up + 3 replays (the capture itself does not run)


def test_detect_graph_replay_sparse_dense_sparse(dev):
    """One captured DetectOut graph replayed on sparse scores (the overflow kernels exit on one load), on dense scores
    (streaming top-k select + rewritten lists) and on sparse scores again: always the eager results, and the
    workspace state is clean after every replay."""
    name, B = "ssd300_voc", 2
    pri = U.oracle_priors(name).to(dev)
    P, C = 8732, 21
    g = torch.Generator().manual_seed(11)
    loc = synth.gen_loc(B, P, 3).to(dev)
    cases = []
    for bias in (10.0, 1.0, 9.0, 1.5):
        logits = torch.randn(B, P, C, generator=g)
        logits[..., 0] += bias
        cases.append(torch.softmax(logits, -1).to(dev))
    eager = ssdbox.DetectOut(C, 0, 200, 0.01, 0.45, VAR)
    want = [eager(loc, sc, pri).clone() for sc in cases]
    assert int((want[1][..., 0] > 0).sum()) > 10 * int((want[0][..., 0] > 0).sum())      # the dense case really is dense
    det = ssdbox.DetectOut(C, 0, 200, 0.01, 0.45, VAR)
    buf = cases[0].clone()
    out = torch.empty(B, C, 200, 5, device=dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        det.forward(loc, buf, pri, out=out)
        det.forward(loc, buf, pri, out=out)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        det.forward(loc, buf, pri, out=out)
    for rep in range(2):
        for sc, w in zip(cases, want):
            buf.copy_(sc)
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(out, w), rep
    # the eager module (always-launched overflow kernels) and the replayed graph leave the same clean workspace state
    ws = det._ws.buf
    assert int(ws[:(B * C + 4) * 4].view(torch.int32).abs().sum()) == 0


@pytest.mark.parametrize("B,seed,flags", [(5, 100, 0), (32, 101, 0), (5, 102, 1), (5, 103, 2), (5, 104, 4)])
def test_refinedet_fused_equals_materialised(dev, B, seed, flags):
    """The fused two-step path (arm_loc / arm_conf handed to the kernels: ssdbox_multibox_loss_fwd_refine / _bwd_refine /
    ssdbox_detect_refine) against the materialised one (refine_anchors + arm_filter feeding anchors_xyxy / pool /
    score_keep): same arithmetic, so every target, set, sum, gradient and detection row must be bit-identical -- for the
    fused and the separate matching kernel, the register-resident and the generic mining kernel."""
    pri, P, C, tg, arm_loc, arm_conf, odm_loc, odm_conf, sc = _refine_case(B, seed)
    if B == 5:
        tg[1] = torch.zeros(0, 5)                      # an image without truths
    gtg = _gpu_targets(tg, dev)
    res = []
    for fused in (True, False):
        crit = ssdbox.RefineMultiBoxLoss(C, 0.5, True, 0, True, 3, 0.5, False, use_ARM=True, fused=fused)
        crit.abi_flags = flags
        ol = odm_loc.to(dev).requires_grad_(True)
        oc = odm_conf.to(dev).requires_grad_(True)
        d = crit.intermediates((arm_loc.to(dev), arm_conf.to(dev), ol, oc, pri.to(dev)), gtg)
        (d["loss_l"] + 2.0 * d["loss_c"]).backward()
        d["grad_loc"], d["grad_conf"] = ol.grad.clone(), oc.grad.clone()
        res.append(d)
    a, b = res
    for k in ("conf_t", "neg", "sel", "tidx", "keys", "sums", "loc_t", "grad_loc", "grad_conf"):
        assert torch.equal(a[k], b[k]), k
    assert float(a["loss_l"]) == float(b["loss_l"]) and float(a["loss_c"]) == float(b["loss_c"])
    assert int((a["sel"] > 0).sum()) > 0 and int(a["neg"].sum()) > 0
    args = tuple(t.to(dev) for t in (arm_loc, arm_conf, odm_loc, sc, pri))
    o_f = ssdbox.RefineDetectOut(C, 0, 200, 0.01, 0.45, VAR, theta=0.01, fused=True)(*args)
    o_m = ssdbox.RefineDetectOut(C, 0, 200, 0.01, 0.45, VAR, theta=0.01, fused=False)(*args)
    assert torch.equal(o_f, o_m) and int((o_f[..., 0] > 0).sum()) > 0
    # dense scores: the overflow kernels read the ARM objectness as well
    dense = synth.gen_detect_scores(min(B, 2), P, C, seed, bkg_bias=1.0).to(dev)
    args = tuple(t[:min(B, 2)].to(dev) for t in (arm_loc, arm_conf, odm_loc)) + (dense, pri.to(dev))
    o_f = ssdbox.RefineDetectOut(C, 0, 200, 0.01, 0.45, VAR, theta=0.01, fused=True)(*args)
    o_m = ssdbox.RefineDetectOut(C, 0, 200, 0.01, 0.45, VAR, theta=0.01, fused=False)(*args)
    assert torch.equal(o_f, o_m)


def test_two_stream_step_equals_serial(dev):
    """ssdbox.TwoStreamStep (DetectOut on a side stream, submitted first; MultiBoxLoss on the current stream): same
    losses and detections as the two ops back to back, eagerly and under CUDA-graph replay."""
    x = U.seeded_inputs("ssd300_voc", 6, 120)
    loc, conf, pri, sc = x["loc"].to(dev), x["conf"].to(dev), x["priors"].to(dev), x["scores"].to(dev)
    gt, offs, gmax = ssdbox.pack_targets(_gpu_targets(x["targets"], dev), dev)
    crit = ssdbox.MultiBoxLoss(x["C"], 0.5, True, 0, True, 3, 0.5, False)
    det = ssdbox.DetectOut(x["C"], 0, 200, 0.01, 0.45, VAR)
    with torch.no_grad():
        wl, wc = crit.forward_packed(loc, conf, pri, gt, offs, gmax)
        want = det(loc, sc, pri).clone()
    step = ssdbox.TwoStreamStep(dev)
    out_buf = torch.empty_like(want)

    def run():
        with torch.no_grad():
            (ll, lc), out = step(lambda: crit.forward_packed(loc, conf, pri, gt, offs, gmax),
                                 lambda: det.forward(loc, sc, pri, out=out_buf))
        return ll, lc, out
    for _ in range(3):
        ll, lc, out = run()
        torch.cuda.synchronize()
        assert float(ll) == float(wl) and float(lc) == float(wc) and torch.equal(out, want)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
        run()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        ll, lc, out = run()
    out_buf.zero_()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert float(ll) == float(wl) and float(lc) == float(wc) and torch.equal(out, want)


# ------------------------------------------------------------------------------------------------
# head layout through the TMA ring (layers with H*W % 4 == 0, >= 64 positions, 32..576 channels) and through the
# generic tile kernel (everything else), mixed in one call -- pure data movement, bit-exact against torch
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("B,shapes", [
    (2, [(324, 64, 64), (486, 32, 32), (486, 16, 16), (486, 8, 8), (486, 4, 4), (324, 2, 2), (324, 1, 1)]),   # SSD512-COCO conf heads
    (3, [(40, 10, 10), (35, 7, 5), (576, 8, 8), (577, 8, 8), (33, 9, 12), (256, 6, 12), (257, 12, 6)]),     # ragged position tiles, 1 / 2 / 3 boxes, odd layers
    (1, [(84, 38, 38), (126, 19, 19), (126, 10, 10), (126, 5, 5), (84, 3, 3), (84, 1, 1)]),                 # SSD300-VOC conf heads
    (5, [(64, 40, 40), (64, 20, 20)]),
])
def test_head_layout_tma_and_generic_layers(dev, B, shapes):
    from ssdbox import heads as H
    gen = torch.Generator().manual_seed(11)
    outs = [torch.randn(B, ch, h, w, generator=gen).to(dev) for ch, h, w in shapes]
    want = torch.cat([o.permute(0, 2, 3, 1).contiguous().view(B, -1) for o in outs], 1)
    got = H.heads_to_rows(outs, 1)
    assert torch.equal(got.view(B, -1), want)
    # a second call into a poisoned output: every element is written
    out = torch.full_like(got, float("nan"))
    H.heads_to_rows(outs, 1, out=out)
    assert torch.equal(out.view(B, -1), want)


@pytest.mark.gpu
def test_head_layout_unaligned_source_falls_back(dev):
    """a source that is not 16-byte aligned cannot be described by a tensor map: the generic kernel takes it"""
    from ssdbox import heads as H
    gen = torch.Generator().manual_seed(12)
    big = torch.randn(2 * 64 * 16 * 16 + 1, generator=gen).to(dev)
    o = big[1:].view(2, 64, 16, 16)
    assert o.data_ptr() % 16 != 0
    want = o.permute(0, 2, 3, 1).contiguous().view(2, -1)
    assert torch.equal(H.heads_to_rows([o], 64).view(2, -1), want)


# ------------------------------------------------------------------------------------------------
# any top_k (box_utils.py:299-301 takes whatever it is given): beyond 1024 the slow matrix-free kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("n,k", [(1500, 1025), (3000, 2000), (3000, 5000), (24564, 3000), (70, 4096), (30000, 30000)])
def test_nms_any_top_k_bit_exact(dev, n, k):
    g = torch.Generator().manual_seed(n + k)
    xy = torch.rand(n, 2, generator=g) * 0.8
    wh = torch.rand(n, 2, generator=g) * 0.12 + 0.01
    boxes = torch.cat([xy, xy + wh], 1)
    scores = torch.rand(n, generator=g)
    if n == 3000:
        scores = (scores * 50).floor() / 50                 # ties everywhere: equal scores are visited higher index first
        boxes[::9, 2:] = boxes[::9, :2]                     # zero-area boxes
    ok, oc = O.greedy_nms(boxes, scores, 0.45, k)
    keep, cnt = BU.nms(boxes.to(dev), scores.to(dev), 0.45, k)
    assert cnt == oc and torch.equal(keep.cpu(), ok)


@pytest.mark.gpu
def test_nms_large_path_equals_fast_path(dev):
    """same input through both kernels: top_k = 1024 (shared-memory matrix) and top_k = 1025 with only 1024 boxes alive"""
    g = torch.Generator().manual_seed(77)
    n = 1024
    xy = torch.rand(n, 2, generator=g) * 0.7
    boxes = torch.cat([xy, xy + torch.rand(n, 2, generator=g) * 0.2 + 0.01], 1).to(dev)
    scores = torch.rand(n, generator=g).to(dev)
    k1, c1 = BU.nms(boxes, scores, 0.45, 1024)
    k2, c2 = BU.nms(boxes, scores, 0.45, 1025)
    assert c1 == c2 and torch.equal(k1, k2)


@pytest.mark.gpu
@pytest.mark.parametrize("name,B,seed,bias,top_k", [("ssd300_voc", 2, 3, 4.0, 1500), ("ssd300_voc", 1, 5, 1.0, 2048),
                                                   ("fssd300_coco", 1, 6, 9.0, 1100)])
def test_detect_any_top_k(dev, name, B, seed, bias, top_k):
    x = U.seeded_inputs(name, B, seed, bkg_bias=bias)
    det = ssdbox.DetectOut(x["C"], 0, top_k, 0.01, 0.45, (0.1, 0.2))
    out = det(x["loc"].to(dev), x["scores"].to(dev), x["priors"].to(dev))
    assert out.shape == (B, x["C"], top_k, 5)
    ref = O.detect(x["loc"], x["scores"], x["priors"], x["C"], top_k=top_k)
    _compare_detect(out.cpu(), ref, "%s top_k=%d" % (name, top_k))
    assert torch.equal(det.last_counts.cpu().long(), (ref[..., 0] > 0).sum(-1))
    # raw logits with such a top_k: the host mirror applies the softmax itself (no fused path beyond 1024)
    if name == "fssd300_coco":
        logit = torch.log(x["scores"].clamp_min(1e-30))
        det2 = ssdbox.DetectOut(x["C"], 0, top_k, 0.01, 0.45, (0.1, 0.2), conf_is_logits=True)
        out2 = det2(x["loc"].to(dev), logit.to(dev), x["priors"].to(dev)).cpu()
        assert torch.equal(out2[..., 0] > 0, out.cpu()[..., 0] > 0)


# ------------------------------------------------------------------------------------------------
# log_sum_exp with ONE maximum over the batch (box_utils.py:272-273) as a fidelity switch
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_global_max_entry_point(dev):
    from ssdbox import _abi
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(5, 777, 21, generator=g) * 7).to(dev)
    out = torch.zeros(1, dtype=torch.float64, device=dev)
    ws = torch.empty(256, dtype=torch.uint8, device=dev)
    _abi.check(_abi.lib().ssdbox_global_max(_abi.ptr(x), x.numel(), _abi.ptr(out), _abi.ptr(ws), 256, _abi.stream_ptr(dev)))
    assert float(out) == float(x.max())
    _abi.check(_abi.lib().ssdbox_global_max(None, 0, _abi.ptr(out), _abi.ptr(ws), 256, _abi.stream_ptr(dev)))
    assert float(out) == float("-inf")


@pytest.mark.gpu
@pytest.mark.parametrize("name,B,seed,offset", [("ssd300_voc", 3, 0, 0.0), ("ssd300_voc", 2, 1, 40.0), ("fssd300_coco", 2, 2, 60.0)])
def test_loss_lse_global_max_mode(dev, name, B, seed, offset):
    """lse_global_max=True evaluates log(sum(exp(x - M))) + M with M = the batch maximum, like the reference: against the
    oracle (which does exactly that) the mining keys agree at least as tightly as in the default per-row mode, the mined
    sets agree up to keys tied within that tolerance, the losses to 1e-5."""
    x = U.seeded_inputs(name, B, seed)
    conf = x["conf"].clone()
    conf[0, :50] += offset                     # one image far above the others: its maximum is everybody's shift
    ref = O.multibox_loss(x["loc"], conf, x["priors"], x["targets"], x["C"], detail=True)
    crit = ssdbox.MultiBoxLoss(x["C"], 0.5, True, 0, True, 3, 0.5, False)
    crit.lse_global_max = True
    pred = (x["loc"].to(dev), conf.to(dev), x["priors"].to(dev))
    tg = [t.to(dev) for t in x["targets"]]
    d = crit.intermediates(pred, tg)
    assert torch.equal(d["conf_t"].cpu(), ref["conf_t"])
    keys = d["keys"].cpu().clone()
    keys[ref["pos"]] = 0                       # (the oracle's mining keys are zeroed at the positives, multibox_loss.py:96)
    # absolute agreement relative to the magnitude the arithmetic ran at (the shift), as in the reference
    scale = max(1.0, float(conf.max()))
    assert float((keys - ref["mining_keys"]).abs().max()) <= 4e-6 * scale
    diff = d["neg"].cpu().bool() != ref["neg"]            # mined sets: identical except at keys tied with the cut within the tolerance
    for b in diff.any(1).nonzero().flatten().tolist():
        k = int(ref["neg"][b].sum())
        kth = ref["mining_keys"][b].sort(descending=True).values[k - 1]
        assert float((ref["mining_keys"][b][diff[b]] - kth).abs().max()) <= 8e-6 * scale, "non-tie mining mismatch in image %d" % b
    U.assert_close_rel(d["loss_l"].cpu(), ref["loss_l"], 1e-5, 1e-7, "loss_l")
    U.assert_close_rel(d["loss_c"].cpu(), ref["loss_c"], 2e-5 * scale, 1e-6, "loss_c")
    # the default mode on the same inputs: same targets, same losses to the same tolerance
    crit2 = ssdbox.MultiBoxLoss(x["C"], 0.5, True, 0, True, 3, 0.5, False)
    ll2, lc2 = crit2(pred, tg)
    U.assert_close_rel(lc2.cpu(), ref["loss_c"], 2e-5 * scale, 1e-6, "loss_c (row max)")
    # and the flag survives backward (the gradient is the same softmax either way)
    loc_g = pred[0].clone().requires_grad_(True)
    conf_g = pred[1].clone().requires_grad_(True)
    ll, lc = crit((loc_g, conf_g, pred[2]), tg)
    (ll + lc).backward()
    gl, gc = O.multibox_loss_grads(x["loc"], conf, x["priors"], x["targets"], x["C"])
    U.assert_close_rel(conf_g.grad.cpu(), gc, 1e-4, 1e-6, "grad_conf")


@pytest.mark.gpu
def test_any_top_k_edge_cases(dev):
    # nms: nothing / one box / top_k far beyond n
    keep = BU.nms(torch.zeros(0, 4, device=dev), torch.zeros(0, device=dev), 0.45, 5000)
    assert isinstance(keep, torch.Tensor) and keep.numel() == 0
    k, c = BU.nms(torch.tensor([[0.1, 0.1, 0.4, 0.5]], device=dev), torch.tensor([0.7], device=dev), 0.45, 60000)
    assert c == 1 and int(k[0]) == 0
    # DetectOut: fewer priors than top_k, an image without any candidate, B = 0
    g = torch.Generator().manual_seed(21)
    P, C, top_k = 300, 4, 1200
    cxcy = torch.rand(P, 2, generator=g)
    pri = torch.cat([cxcy, torch.rand(P, 2, generator=g) * 0.3 + 0.05], 1)
    loc = torch.randn(2, P, 4, generator=g) * 0.3
    sc = torch.softmax(torch.randn(2, P, C, generator=g) * 2, -1)
    sc[1] = 0.0                                                   # image 1: nothing above conf_thresh
    det = ssdbox.DetectOut(C, 0, top_k, 0.01, 0.45, (0.1, 0.2))
    out = det(loc.to(dev), sc.to(dev), pri.to(dev)).cpu()
    ref = O.detect(loc, sc, pri, C, top_k=top_k)
    _compare_detect(out, ref, "P < top_k")
    assert float(out[1].abs().sum()) == 0.0
    out0 = det(loc[:0].to(dev), sc[:0].to(dev), pri.to(dev))
    assert out0.shape == (0, C, top_k, 5)


@pytest.mark.gpu
def test_head_layout_edge_cases(dev):
    from ssdbox import heads as H
    g = torch.Generator().manual_seed(31)
    # exactly at the limits of the TMA path (64 positions, 32 channels), next to layers just outside it, and B = 0
    shapes = [(32, 8, 8), (31, 8, 8), (32, 7, 9), (576, 8, 8), (64, 2, 32)]
    outs = [torch.randn(3, ch, h, w, generator=g).to(dev) for ch, h, w in shapes]
    want = torch.cat([o.permute(0, 2, 3, 1).contiguous().view(3, -1) for o in outs], 1)
    assert torch.equal(H.heads_to_rows(outs, 1).view(3, -1), want)
    # channels-last / sliced inputs are made contiguous by the host mirror
    big = torch.randn(3, 80, 8, 8, generator=g).to(dev)
    view = big[:, 8:72]                                            # non-contiguous NCHW slice
    want = view.permute(0, 2, 3, 1).contiguous().view(3, -1)
    assert torch.equal(H.heads_to_rows([view], 64).view(3, -1), want)
    empty = [torch.zeros(0, 64, 8, 8, device=dev)]
    assert H.heads_to_rows(empty, 64).shape[0] == 0


# ------------------------------------------------------------------------------------------------
# mining with two 512-thread CTAs per SM (batches of more images than SMs): same selection as the 1024-thread kernel
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,B,seed", [("ssd300_voc", 5, 0), ("rfb300_voc", 3, 1), ("refinedet320_voc", 4, 2)])
def test_mining_half_cta_equals_full_cta(dev, name, B, seed):
    from ssdbox import _abi
    x = U.seeded_inputs(name, B, seed)
    pred = (x["loc"].to(dev), x["conf"].to(dev), x["priors"].to(dev))
    tg = [t.to(dev) for t in x["targets"]]
    res = {}
    for flags in (0, _abi.LOSS_MINE_HALF_CTA):
        crit = ssdbox.MultiBoxLoss(x["C"], 0.5, True, 0, True, 3, 0.5, False)
        crit.abi_flags = flags
        res[flags] = crit.intermediates(pred, tg)
    a, b = res[0], res[_abi.LOSS_MINE_HALF_CTA]
    for k in ("conf_t", "neg", "sel", "tidx", "keys"):
        assert torch.equal(a[k], b[k]), k
    assert float(a["sums"][2]) == float(b["sums"][2])
    assert float(((a["sums"] - b["sums"]).abs() / a["sums"].abs().clamp_min(1e-30)).max()) <= 1e-12      # fp64 sums, another fixed order
    U.assert_close_rel(b["loss_l"].cpu(), a["loss_l"].cpu(), 2e-7, 0, "loss_l")
    U.assert_close_rel(b["loss_c"].cpu(), a["loss_c"].cpu(), 2e-7, 0, "loss_c")


@pytest.mark.gpu
def test_mining_more_images_than_sms(dev):
    """B = 150 > 148 SMs: the 512-thread kernel is picked by itself; the generic kernel (any shape) selects the same rows"""
    from ssdbox import _abi
    cfg, c = configs.get("ssd300_voc")
    x = U.seeded_inputs("ssd300_voc", 2, 7)
    B = 150
    g = torch.Generator().manual_seed(9)
    loc = torch.randn(B, x["P"], 4, generator=g).to(dev) * 0.5
    conf = torch.randn(B, x["P"], x["C"], generator=g).to(dev)
    conf[..., 0] += 3.0
    tg = [t.to(dev) for t in synth.gen_targets(B, x["C"], 8, 11)]
    pri = x["priors"].to(dev)
    out = {}
    for flags in (0, _abi.LOSS_GENERIC_MINE):
        crit = ssdbox.MultiBoxLoss(x["C"], 0.5, True, 0, True, 3, 0.5, False)
        crit.abi_flags = flags
        ll, lc = crit((loc, conf, pri), tg)
        out[flags] = (ll.cpu(), lc.cpu(), crit._last[1].clone(), crit._last[0].clone())
    a, b = out[0], out[_abi.LOSS_GENERIC_MINE]
    assert torch.equal(a[2], b[2])                                   # the selection flags of all 150 images
    assert float(a[3][2]) == float(b[3][2])
    U.assert_close_rel(a[0], b[0], 2e-7, 0, "loss_l")
    U.assert_close_rel(a[1], b[1], 2e-7, 0, "loss_c")
