"""TEST INFRASTRUCTURE ONLY -- records golden vectors from the *reference itself*.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
Writes small fixtures to tests/golden/.  Every output array below is produced by the
unmodified reference functions (through oracle/ref_loader.py's shims), never by the oracle;
tests/test_oracle_golden.py then pins oracle/ssd_oracle.py to them on any machine, and the
``-m gpu`` tests pin the CUDA path to them.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch_b200"))

from oracle import ref_loader  # noqa: E402
from ssdbox import configs, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# A reduced SSD300-style head (4 coarse maps, 790 priors) keeps the fixtures small.
SMALL_MODEL = configs.AttrDict(MODEL=configs.AttrDict(
    IMAGE_SIZE=(300, 300), STEPS=[32, 64, 100, 300], MIN_SIZES=[111, 162, 213, 264],
    MAX_SIZES=[162, 213, 264, 315], ASPECT_RATIOS=[[2, 3], [2, 3], [2], [2]],
    VARIANCE=[0.1, 0.2], CLIP=True, FLIP=True, NUM_CLASSES=21))
SMALL_DIMS = [[10, 10], [5, 5], [3, 3], [1, 1]]


def digest(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def np32(t):
    return t.detach().numpy()


def main():
    ref = ref_loader.load()
    bu = ref.box_utils
    os.makedirs(OUT, exist_ok=True)
    var = [0.1, 0.2]

    # ---- RNG-free known answers -------------------------------------
    cfg300, c300 = configs.get("ssd300_voc")
    pri300 = ref.PriorBoxSSD(cfg300).forward(c300["layer_dims"])
    kat = {}
    nb = torch.tensor([[0, 0, 1, 1], [0.1, 0, 1.1, 1], [0, 0, 0.5, 0.5], [2, 2, 3, 3], [2.05, 2, 3.05, 3]])
    ns = torch.tensor([0.9, 0.8, 0.7, 0.6, 0.95])
    k, c = bu.nms(nb, ns, 0.45, 200)
    kat.update(nms_boxes=np32(nb), nms_scores=np32(ns), nms_keep=k.numpy(), nms_count=np.int64(c))
    k2, c2 = bu.nms(nb, ns, 0.45, 2)
    kat.update(nms_keep_top2=k2.numpy(), nms_count_top2=np.int64(c2))
    gt1 = torch.tensor([[0.35, 0.25, 0.65, 0.70]])
    pr1 = torch.tensor([[0.5, 0.5, 0.2, 0.4]])
    kat.update(iou_gt=np32(gt1), iou_prior=np32(pr1), iou=np32(bu.jaccard(gt1, bu.point_form(pr1))))
    enc = bu.encode(gt1, pr1, var)
    kat.update(encode=np32(enc), decode_of_encode=np32(bu.decode(enc, pr1, var)))
    l2 = torch.tensor([[0.5, -0.25, 1.0, -2.0]])
    kat.update(decode_loc=np32(l2), decode=np32(bu.decode(l2, pr1, var)))
    tr = torch.tensor([[0.10, 0.15, 0.45, 0.60], [0.40, 0.30, 0.90, 0.95], [0.70, 0.05, 0.78, 0.12]])
    lb = torch.tensor([11., 14., 6.])
    lt = torch.zeros(1, pri300.size(0), 4)
    ct = torch.zeros(1, pri300.size(0), dtype=torch.int64)
    bu.match(0.5, tr, pri300, var, lb, lt, ct, 0)
    kat.update(match_truths=np32(tr), match_labels=np32(lb), match_conf_t=ct[0].numpy().astype(np.int16),
               match_loc_t_pos=np32(lt[0][ct[0] > 0]))
    dup = torch.tensor([[0.1, 0.1, 0.4, 0.5], [0.1, 0.1, 0.4, 0.5]])
    dl = torch.tensor([3., 7.])
    bu.match(0.5, dup, pri300, var, dl, lt, ct, 0)
    kat.update(dup_truths=np32(dup), dup_labels=np32(dl), dup_conf_t=ct[0].numpy().astype(np.int16))
    x = torch.tensor([[1., 2., 3.], [-50., -51., -52.]])
    kat.update(lse_x=np32(x), lse=np32(bu.log_sum_exp(x)))
    for name in configs.CONFIGS:
        cfg, c = configs.get(name)
        p = ref.PriorBoxSSD(cfg).forward(c["layer_dims"])
        kat["priors_sha_" + name] = np.array(digest(p))
        kat["priors_head_" + name] = np32(p[:8])
        kat["priors_tail_" + name] = np32(p[-8:])
        kat["priors_sum64_" + name] = np.float64(p.double().sum().item())
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **kat)

    # ---- small full-tensor fixture: inputs AND reference outputs ---------------------------
    pri = ref.PriorBoxSSD(SMALL_MODEL).forward(SMALL_DIMS)
    P, C, B = pri.size(0), 21, 3
    tg = synth.gen_targets(B, C, 6, 11)
    loc = synth.gen_loc(B, P, 11)
    conf = synth.gen_train_logits(B, P, C, 11)
    sc = synth.gen_detect_scores(B, P, C, 11, bkg_bias=5.0)
    loc_t = torch.zeros(B, P, 4)
    conf_t = torch.zeros(B, P, dtype=torch.int64)
    for b, t in enumerate(tg):
        bu.match(0.5, t[:, :4], pri, var, t[:, 4], loc_t, conf_t, b)
    ll, lc = ref.multibox_loss(C, (loc, conf, pri), tg)
    det = ref.detect(C, loc, sc, pri, top_k=20)
    flat, offs = synth.pack_targets(tg)
    nk, nc = bu.nms(bu.decode(loc[0], pri, var), sc[0, :, 5].contiguous(), 0.45, 50)
    np.savez_compressed(
        os.path.join(OUT, "small790.npz"), priors=np32(pri), gt=np32(flat), gt_offsets=offs.numpy(),
        loc=np32(loc), conf=np32(conf), scores=np32(sc), loc_t=np32(loc_t),
        conf_t=conf_t.numpy().astype(np.int16), loss_l=np32(ll), loss_c=np32(lc),
        detect_top20=np32(det), nms_keep=nk.numpy().astype(np.int32), nms_count=np.int64(nc))

    # ---- seeded full-size fixtures: inputs regenerated from the seed, outputs stored -------
    seeded = {}
    for name, B, seed in [("ssd300_voc", 4, 0), ("fssd300_coco", 2, 1), ("ssd512_coco", 2, 2)]:
        cfg, c = configs.get(name)
        pri = ref.PriorBoxSSD(cfg).forward(c["layer_dims"])
        P, C = pri.size(0), cfg.MODEL.NUM_CLASSES
        tg = synth.gen_targets(B, C, c["gt_max"], seed)
        loc = synth.gen_loc(B, P, seed)
        conf = synth.gen_train_logits(B, P, C, seed)
        sc = synth.gen_detect_scores(B, P, C, seed, bkg_bias=10.0)
        loc_t = torch.zeros(B, P, 4)
        conf_t = torch.zeros(B, P, dtype=torch.int64)
        for b, t in enumerate(tg):
            bu.match(0.5, t[:, :4], pri, var, t[:, 4], loc_t, conf_t, b)
        ll, lc = ref.multibox_loss(C, (loc, conf, pri), tg)
        det = ref.detect(C, loc, sc, pri)
        key = "%s_b%d_s%d" % (name, B, seed)
        nz = det[..., 0] > 0
        seeded[key + "_inputs_sha"] = np.array(digest(pri, loc, conf, sc, *tg))
        seeded[key + "_conf_t"] = conf_t.numpy().astype(np.int8)
        seeded[key + "_loc_t_sha"] = np.array(digest(loc_t))
        seeded[key + "_loc_t_pos_sum64"] = np.float64(loc_t[conf_t > 0].double().sum().item())
        seeded[key + "_loss"] = np.array([float(ll), float(lc)], dtype=np.float32)
        seeded[key + "_det_counts"] = nz.sum(-1).numpy().astype(np.int16)
        seeded[key + "_det_rows"] = np32(det[nz])
    np.savez_compressed(os.path.join(OUT, "seeded.npz"), **seeded)
    # ---- eval post-processing after Detect (evaluate_utils.py:63-70,127-139,175-203) -------
    import importlib
    import types
    eu = importlib.import_module("lib.utils.evaluate_utils")
    g = torch.Generator().manual_seed(17)
    B, Cn, K = 4, 6, 9
    det = torch.zeros(B, Cn, K, 5)
    for b in range(B):
        for c in range(1, Cn):
            n = int(torch.randint(0, K + 1, (1,), generator=g))
            det[b, c, :n, 0] = torch.rand(n, generator=g).sort(descending=True).values * 0.98 + 0.01
            det[b, c, :n, 1:] = torch.rand(n, 4, generator=g)
    extra = torch.tensor([[375.0, 500.0], [333.0, 500.0], [480.0, 640.0], [427.0, 640.0]])
    ids = [139, 285, 632, 724]
    scaled = det.clone()
    h = extra[:, 0].unsqueeze(-1).unsqueeze(-1)          # evaluate_utils.py:63-68, verbatim order
    w = extra[:, 1].unsqueeze(-1).unsqueeze(-1)
    scaled[:, :, :, 1] *= w
    scaled[:, :, :, 3] *= w
    scaled[:, :, :, 2] *= h
    scaled[:, :, :, 4] *= h
    voc, _ = eu.EvalVOC.convert_ssd_result(None, scaled.clone(), 0)
    holder = types.SimpleNamespace(dataset=types.SimpleNamespace(ids=ids), results=[])
    coco, idt = eu.EvalCOCO.convert_ssd_result(holder, scaled.clone(), 0)
    eu.EvalCOCO.post_proc(holder, coco.clone(), 0, idt)
    np.savez_compressed(os.path.join(OUT, "evalpost.npz"), det=np32(det), extra=np32(extra),
                        ids=np.array(ids, dtype=np.float32), voc=np32(voc), coco=np32(coco),
                        coco_rows=holder.results[0].astype(np.float32))
    # ---- head-output layout (ssd_v3.py:113-121): the real model's forward on captured head outputs ----
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
        mloc, mconf = model(torch.randn(1, 3, 300, 300), phase="train")
    for h in handles:
        h.remove()
    # the loc heads only (34928 floats): inputs and the reference's own output
    heads = {"loc_out": np32(mloc)}
    for k, o in enumerate(outs["loc"]):
        heads["loc_in%d" % k] = np32(o)
    np.savez_compressed(os.path.join(OUT, "heads.npz"), **heads)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
