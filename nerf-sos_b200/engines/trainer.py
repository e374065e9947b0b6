"""Drop-in for engines/trainer.py:train_one_step of the reference (patch_tune recipe of scripts/train_*.sh),
data-parallel aware.  Same signature, same loss assembly and the same returned dict; differences:

  * `model(...)` is the fused CUDA render (kernel A) and the correlation losses are kernel B;
  * with torch.distributed initialised, each rank holds B/G whole patches: per-patch tensors are all-gathered
    (parallel.gather_cat) so that negatives may live on another rank, every rank evaluates the identical
    global-batch loss, and parameter gradients are summed with one flat all-reduce before optimizer.step();
  * the DINO pass runs under no_grad (the reference builds a graph that carries no useful gradient);
  * reference crashes are guarded, not reproduced: `cls_` undefined when --use_dino is off (trainer.py:125),
    the unused `sacrebleu` import, CPU KMeans/ARI logging is skipped unless sklearn is importable.
Numerical quirks are kept: the FINE depth feeds both geometry-loss calls (trainer.py:159-160), double
ImageNet normalisation lives inside the injected `dino.get_vit_attn_feat`.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .. import parallel as P
from ..utils.image import get_similarity_matrix, img2mse, mse2psnr


def normalize_batch(batch):
    mean = batch.new_tensor([0.485, 0.456, 0.406]).view(-1, 1, 1)
    std = batch.new_tensor([0.229, 0.224, 0.225]).view(-1, 1, 1)
    return (batch - mean) / std


def train_one_step(batch, model, optimizer, scheduler, train_loader, global_step, losses, device, args, group=None,
                   randoms=None):
    model, dino = model
    seg_loss, contrast_loss, correlation_loss, geoCorrelation_loss = losses
    model.train()
    near, far = train_loader.dataset.near_far()
    radii = train_loader.dataset.radii() if hasattr(train_loader.dataset, "radii") else None
    batch = [b.to(device) for b in batch]
    batch_rays, gt = batch[0], batch[1]
    masks = batch[2] if len(batch) > 2 else None
    if not args.patch_tune:
        raise NotImplementedError("only the --patch_tune recipe (all shipped scripts) is implemented")
    Bl, Ps = batch_rays.shape[0], args.patch_size                      # local patches on this rank
    batch_rays = batch_rays.reshape(-1, *batch_rays.shape[2:]).permute(1, 0, 2)      # [2, Bl*P*P, 3]
    gt = gt.reshape(Bl, Ps, Ps, 3)

    kw = {"radii": radii}
    if randoms is not None:
        kw["randoms"] = randoms
    ret = model(batch_rays, (near, far), **kw)                       # kernel A (trainer.py:68)
    patch = lambda t: t.reshape(Bl, Ps, Ps, t.shape[-1])
    rgb, rgb0 = patch(ret["rgb"]), patch(ret["rgb0"])
    depth = patch(ret["depth"])
    ray_o, ray_d = patch(batch_rays[0]), patch(batch_rays[1])
    has_sem = "semantics" in ret
    if has_sem:
        sem, sem0 = patch(ret["semantics"]), patch(ret["semantics0"])

    cls_ = feat = None
    if args.use_dino:
        with torch.no_grad():
            dino_in = F.interpolate(rgb.detach().permute(0, 3, 1, 2), (Ps * args.patch_stride, Ps * args.patch_stride))
            d = dino.get_vit_attn_feat(normalize_batch(dino_in))
        cls_, feat = d["cls_"], d["feat"]

    optimizer.zero_grad()
    # ---- global batch: gather the per-patch tensors (identity on one GPU)
    rgb_g, rgb0_g, gt_g = P.gather_cat(rgb, group), P.gather_cat(rgb0, group), P.gather_cat(gt, group)
    img_loss = img2mse(rgb_g, gt_g)
    psnr = mse2psnr(img_loss)
    loss = args.rgb_w * img_loss
    img_loss0 = img2mse(rgb0_g, gt_g)
    loss = loss + args.rgb_w * img_loss0
    zero = torch.zeros((), device=device)
    corr0 = corr1 = geo0 = geo1 = contrast_l = zero
    sim = None
    if cls_ is not None:
        sim = get_similarity_matrix(P.gather_cat(cls_, group))
    if args.use_correlation and has_sem and feat is not None:
        side = int(math.sqrt(feat.shape[-2]))
        feat_g = P.gather_cat(feat, group)
        feat_g = feat_g.reshape(feat_g.shape[0], side, side, feat_g.shape[-1]).permute(0, 3, 1, 2)
        s0 = P.gather_cat(sem0, group).permute(0, 3, 1, 2)
        s1 = P.gather_cat(sem, group).permute(0, 3, 1, 2)
        corr0 = args.correlation_w * correlation_loss(feat_g, s0, sim)
        corr1 = args.correlation_w * correlation_loss(feat_g, s1, sim)
        loss = loss + corr0 + corr1
    if args.use_geoCorr and has_sem:
        s0 = P.gather_cat(sem0, group).permute(0, 3, 1, 2)
        s1 = P.gather_cat(sem, group).permute(0, 3, 1, 2)
        dep = P.gather_cat(depth.detach(), group).permute(0, 3, 1, 2).contiguous()
        ro = P.gather_cat(ray_o, group).permute(0, 3, 1, 2)
        rd = P.gather_cat(ray_d, group).permute(0, 3, 1, 2)
        geo0 = args.Gcorrelation_w * geoCorrelation_loss(dep, s0, [ro, rd, gt_g], sim)      # fine depth for both (:159)
        geo1 = args.Gcorrelation_w * geoCorrelation_loss(dep, s1, [ro, rd, gt_g], sim)
        loss = loss + geo0 + geo1
    if getattr(args, "use_contrast", False) and contrast_loss is not None and cls_ is not None:
        contrast_l = args.contrast_w * contrast_loss(cls_)
        loss = loss + contrast_l

    loss.backward()
    P.allreduce_gradients(model.parameters(), group)          # every rank back-propagated the same global loss: SUM
    optimizer.step()
    if scheduler is not None:
        scheduler.step(global_step)
    return dict(loss=loss, psnr=psnr, sem0=zero, sem1=zero, img0=img_loss0, img1=img_loss, contrast=contrast_l, corr0=corr0,
                corr1=corr1, geo_corr0=geo0, geo_corr1=geo1, clus_ari=0, clus_ari_fg=0, sem_ari=0, sem_ari_fg=0)
