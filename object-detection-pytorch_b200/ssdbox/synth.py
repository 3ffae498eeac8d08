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


def pack_targets(targets):
    """list of [G_b,5] -> (flat [sum G,5] float32, offsets int32 [B+1]) -- the C-ABI layout."""
    offs = [0]
    for t in targets:
        offs.append(offs[-1] + (int(t.size(0)) if t.dim() == 2 else 0))
    rows = [t for t in targets if t.dim() == 2 and t.size(0) > 0]
    flat = torch.cat(rows, 0).float() if rows else torch.zeros(0, 5)
    return flat.contiguous(), torch.tensor(offs, dtype=torch.int32)
