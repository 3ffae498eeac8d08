"""Synthetic inputs of the named shapes, shared by tests and bench.

Everything is drawn on the CPU from ``torch.Generator().manual_seed(seed)`` so that the CUDA
path and the CPU oracle see identical fp32 tensors.  Priors are supplied by the caller (the
real PriorBoxSSD output of the configuration, never random).
"""
import torch


def gen_targets(batch, num_classes, gt_max, seed, gt_min=1):
    """list of B tensors [G_b,5] = (x1,y1,x2,y2,label0) with G_b ~ U{gt_min..gt_max}."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(batch):
        n = int(torch.randint(gt_min, gt_max + 1, (1,), generator=g))
        wh = torch.rand(n, 2, generator=g) * 0.45 + 0.05
        xy = torch.rand(n, 2, generator=g) * (1 - wh)
        lab = torch.randint(0, num_classes - 1, (n, 1), generator=g).float()
        out.append(torch.cat([xy, xy + wh, lab], 1))
    return out


def gen_loc(batch, num_priors, seed):
    g = torch.Generator().manual_seed(seed + 1000)
    return torch.randn(batch, num_priors, 4, generator=g) * 0.5


def gen_train_logits(batch, num_priors, num_classes, seed, bkg_bias=4.0):
    """N(0,1) logits with a background bias (train-time conf_data)."""
    g = torch.Generator().manual_seed(seed + 2000)
    x = torch.randn(batch, num_priors, num_classes, generator=g)
    x[..., 0] += bkg_bias
    return x


def gen_detect_scores(batch, num_priors, num_classes, seed, bkg_bias=10.0):
    """softmax scores; bkg_bias=10 -> sparse/realistic, 4 -> dense/worst case."""
    g = torch.Generator().manual_seed(seed + 3000)
    x = torch.randn(batch, num_priors, num_classes, generator=g)
    x[..., 0] += bkg_bias
    return torch.softmax(x, -1)


def gen_arm_outputs(batch, num_priors, seed):
    """RefineDet ARM head outputs: arm_loc = 0.2 * N(0,1) (a refinement, smaller than
    a regression from scratch) and binary objectness logits 2.5 * N(0,1) with a background bias of 2 --
    roughly a quarter of the anchors fall below theta = 0.01 and are filtered."""
    g = torch.Generator().manual_seed(seed + 5000)
    arm_loc = torch.randn(batch, num_priors, 4, generator=g) * 0.2
    arm_conf = torch.randn(batch, num_priors, 2, generator=g) * 2.5
    arm_conf[..., 0] += 2.0
    return arm_loc, arm_conf


def pack_targets(targets):
    """list of [G_b,5] -> (flat [sum G,5] float32, offsets int32 [B+1]) -- the C-ABI layout."""
    offs = [0]
    for t in targets:
        offs.append(offs[-1] + (int(t.size(0)) if t.dim() == 2 else 0))
    rows = [t for t in targets if t.dim() == 2 and t.size(0) > 0]
    flat = torch.cat(rows, 0).float() if rows else torch.zeros(0, 5)
    return flat.contiguous(), torch.tensor(offs, dtype=torch.int32)


def gen_voc_eval_case(num_images, num_classes, seed, gt_max=6, fp_max=6, top_k=200, difficult_p=0.15,
                      distinct_scores=False):
    """A synthetic PASCAL-VOC evaluation set in the flat layout of ssdbox.voc_eval.

    Truths: per image G ~ U{1..gt_max} integer pixel boxes (as parse_rec yields them: xml value - 1),
    class ~ U{1..C-1}, `difficult` with probability difficult_p.  Detections per (image, class): up to
    two jittered copies of every truth of the class (true-positive / duplicate candidates) plus
    U{0..fp_max} random boxes, fractional float32 pixel coordinates, scores descending within the
    segment like DetectOut emits them.  distinct_scores=True keeps the 3-decimal quantised scores of a
    class pairwise different (<= 989 detections per class), which makes np.argsort's unstable order
    irrelevant.  Returns a dict of numpy arrays: rows [N,7] (xmin, ymin, xmax, ymax, score, image, cls),
    seg [I*C+1], gt_boxes [M,4] float32, gt_labels [M] int32, gt_difficult [M] uint8, gt_offsets [I+1]."""
    import numpy as np
    rs = np.random.RandomState(seed)
    C = num_classes
    gtb, gtl, gtd, goff = [], [], [], [0]
    segs = [[None] * C for _ in range(num_images)]
    for i in range(num_images):
        G = rs.randint(1, gt_max + 1)
        wh = rs.randint(20, 220, size=(G, 2))
        xy = rs.randint(0, 480 - 220, size=(G, 2))
        boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        labels = rs.randint(1, C, size=G).astype(np.int32)
        diff = (rs.rand(G) < difficult_p).astype(np.uint8)
        gtb.append(boxes); gtl.append(labels); gtd.append(diff); goff.append(goff[-1] + G)
        for c in range(1, C):
            cand, sc = [], []
            for g in np.where(labels == c)[0]:
                for _ in range(rs.randint(0, 3)):
                    size = np.tile(boxes[g, 2:] - boxes[g, :2], 2)
                    cand.append(boxes[g] + (rs.rand(4) - 0.5) * 0.3 * size)
                    sc.append(rs.rand() * 0.7 + 0.3)
            for _ in range(rs.randint(0, fp_max + 1)):
                w2 = rs.rand(2) * 200 + 10
                p = rs.rand(2) * 300
                cand.append(np.concatenate([p, p + w2]))
                sc.append(rs.rand() * 0.5 + 0.011)
            cand, sc = cand[:top_k], sc[:top_k]
            if cand:
                o = np.argsort(-np.asarray(sc, dtype=np.float32), kind="stable")
                segs[i][c] = np.concatenate([np.asarray(cand, dtype=np.float32)[o],
                                             np.asarray(sc, dtype=np.float32)[o][:, None]], 1)
    if distinct_scores:
        for c in range(1, C):
            n = sum(len(segs[i][c]) for i in range(num_images) if segs[i][c] is not None)
            assert n <= 989, "too many detections of one class for pairwise distinct 3-decimal scores"
            ks = (rs.permutation(989)[:n] + 11).astype(np.float32) / np.float32(1000.0)
            at = 0
            for i in range(num_images):
                if segs[i][c] is not None:
                    m = len(segs[i][c])
                    segs[i][c][:, 4] = np.sort(ks[at:at + m])[::-1]
                    at += m
    rows, seg = [], [0]
    for i in range(num_images):
        for c in range(C):
            s = segs[i][c]
            if s is not None:
                rows.append(np.concatenate([s, np.full((len(s), 1), i, np.float32), np.full((len(s), 1), c, np.float32)], 1))
            seg.append(seg[-1] + (0 if s is None else len(s)))
    rows = np.concatenate(rows, 0).astype(np.float32) if rows else np.zeros((0, 7), np.float32)
    return dict(rows=rows, seg=np.asarray(seg, np.int32), gt_boxes=np.concatenate(gtb, 0).astype(np.float32),
                gt_labels=np.concatenate(gtl).astype(np.int32), gt_difficult=np.concatenate(gtd).astype(np.uint8),
                gt_offsets=np.asarray(goff, np.int32), num_images=num_images, num_classes=C)
