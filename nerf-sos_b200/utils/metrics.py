"""Evaluation metrics on the device the images live on: SSIM (utils/ssim.py:7-40 of the reference), adjusted
Rand index (sklearn.metrics.adjusted_rand_score as called by engines/eval.py:70-74), 2-means clustering of the
semantic logits (utils/misc.py:40-50, sklearn KMeans there) and foreground IoU.  The reference moves every map to the
host and runs sklearn per view; here a full 1008x756 view never leaves the GPU."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _window(size, channel, device, dtype):
    g = torch.tensor([math.exp(-(x - size // 2) ** 2 / (2 * 1.5 ** 2)) for x in range(size)], dtype=torch.float32)
    g = (g / g.sum()).unsqueeze(1)
    w2 = (g @ g.t()).unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, size, size).contiguous().to(device=device, dtype=dtype)


def ssim(img1, img2, window_size=11, size_average=True, format="NCHW"):
    """Gaussian-window SSIM (sigma 1.5, C1=0.01^2, C2=0.03^2).  format 'HWC' takes one image like eval.py:88."""
    if format == "HWC":
        img1, img2 = img1.permute(2, 0, 1).unsqueeze(0), img2.permute(2, 0, 1).unsqueeze(0)
    c = img1.shape[1]
    w = _window(window_size, c, img1.device, img1.dtype)
    pad = window_size // 2
    mu1, mu2 = F.conv2d(img1, w, padding=pad, groups=c), F.conv2d(img2, w, padding=pad, groups=c)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = F.conv2d(img1 * img1, w, padding=pad, groups=c) - mu1_sq
    s2 = F.conv2d(img2 * img2, w, padding=pad, groups=c) - mu2_sq
    s12 = F.conv2d(img1 * img2, w, padding=pad, groups=c) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean() if size_average else m.mean(1).mean(1).mean(1)


def adjusted_rand_score(labels_true, labels_pred):
    """ARI from the contingency table (same special cases as sklearn: identical trivial clusterings -> 1)."""
    a = torch.as_tensor(labels_true).reshape(-1).long()
    b = torch.as_tensor(labels_pred).reshape(-1).long().to(a.device)
    n = a.numel()
    if n == 0:
        return torch.tensor(1.0, device=a.device)
    _, ai = torch.unique(a, return_inverse=True)
    _, bi = torch.unique(b, return_inverse=True)
    na, nb = int(ai.max()) + 1, int(bi.max()) + 1
    cont = torch.bincount(ai * nb + bi, minlength=na * nb).reshape(na, nb).double()
    tn_pairs = n * (n - 1.0)
    sum_sq = (cont * cont).sum()
    ra, rb = cont.sum(1), cont.sum(0)
    # pair confusion matrix (sklearn.metrics.cluster.pair_confusion_matrix)
    tp = sum_sq - n
    fp = (cont * rb[None, :]).sum() - sum_sq
    fn = (cont * ra[:, None]).sum() - sum_sq
    tn = tn_pairs - tp - fp - fn
    if fn == 0 and fp == 0:
        return torch.tensor(1.0, device=a.device)
    return (2.0 * (tp * tn - fn * fp) / ((tp + fn) * (fn + tn) + (tp + fp) * (fp + tn))).float()


def kmeans_labels(x, n_clusters=2, iters=50):
    """Lloyd's algorithm on [..., C] features, deterministic farthest-point initialisation.  Returns int64 labels with
    the leading shape of x plus a trailing 1 (segmap_cluster returns [H, W, 1])."""
    lead = x.shape[:-1]
    p = x.reshape(-1, x.shape[-1]).float()
    cent = p.mean(0, keepdim=True)
    cs = [p[((p - cent) ** 2).sum(-1).argmax()]]
    for _ in range(1, n_clusters):
        d = torch.stack([((p - c) ** 2).sum(-1) for c in cs], 0).min(0).values
        cs.append(p[d.argmax()])
    c = torch.stack(cs, 0)
    lab = None
    for _ in range(iters):
        new = torch.cdist(p, c).argmin(1)
        if lab is not None and torch.equal(new, lab):
            break
        lab = new
        for k in range(n_clusters):
            m = lab == k
            if m.any():
                c[k] = p[m].mean(0)
    return lab.reshape(*lead, 1)


def binary_iou(pred, gt):
    """Foreground IoU of two {0,1} maps; a clustering has no polarity, so the better of the two assignments counts."""
    p, g = torch.as_tensor(pred).reshape(-1).bool(), torch.as_tensor(gt).reshape(-1).bool().to(torch.as_tensor(pred).device)
    def iou(a, b):
        u = (a | b).sum().float()
        return (a & b).sum().float() / u if u > 0 else torch.tensor(1.0, device=a.device)
    return torch.maximum(iou(p, g), iou(~p, g))
