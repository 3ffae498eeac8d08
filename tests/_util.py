"""Shared helpers for the test-suite (input regeneration, golden loading, digests)."""
import hashlib
import os

import numpy as np
import torch

from oracle import ssd_oracle as O
from ssdbox import configs, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SMALL_MODEL = configs.AttrDict(
    IMAGE_SIZE=(300, 300), STEPS=[32, 64, 100, 300], MIN_SIZES=[111, 162, 213, 264],
    MAX_SIZES=[162, 213, 264, 315], ASPECT_RATIOS=[[2, 3], [2, 3], [2], [2]],
    VARIANCE=[0.1, 0.2], CLIP=True, FLIP=True, NUM_CLASSES=21)
SMALL_DIMS = [[10, 10], [5, 5], [3, 3], [1, 1]]


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def digest(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


_PRI = {}


def oracle_priors(name):
    if name not in _PRI:
        cfg, c = configs.get(name)
        _PRI[name] = O.prior_boxes(cfg.MODEL, c["layer_dims"])
    return _PRI[name]


def seeded_inputs(name, B, seed, bkg_bias=10.0):
    cfg, c = configs.get(name)
    pri = oracle_priors(name)
    P, C = pri.size(0), cfg.MODEL.NUM_CLASSES
    tg = synth.gen_targets(B, C, c["gt_max"], seed)
    loc = synth.gen_loc(B, P, seed)
    conf = synth.gen_train_logits(B, P, C, seed)
    sc = synth.gen_detect_scores(B, P, C, seed, bkg_bias=bkg_bias)
    return dict(priors=pri, targets=tg, loc=loc, conf=conf, scores=sc, P=P, C=C, B=B)


def unpack_targets(flat, offs):
    flat = torch.as_tensor(flat)
    return [flat[int(offs[i]):int(offs[i + 1])] for i in range(len(offs) - 1)]


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float(((a - b).abs() / b.abs().clamp_min(1e-30)).max()) if a.numel() else 0.0


def assert_close_rel(a, b, tol=1e-5, atol=1e-6, what=""):
    """|a-b| <= atol + tol*|b| element-wise (the north-star float tolerance is 1e-5 relative;
    atol covers values that are themselves ~0, e.g. an encoded centre offset of exactly 0)."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = (a - b).abs() > (atol + tol * b.abs())
    assert not bool(bad.any()), "%s: %d/%d elements out of tolerance, max abs diff %g" % (
        what, int(bad.sum()), a.numel(), float((a - b).abs().max()))
