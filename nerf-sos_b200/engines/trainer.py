"""Drop-in for engines/trainer.py:train_one_step of the reference (patch_tune recipe of scripts/train_*.sh),
data-parallel aware.  Same signature, same loss assembly and the same returned dict; differences:

  * `model(...)` is the fused CUDA render (kernel A) and the correlation losses are kernel B;
  * with torch.distributed initialised each rank holds B/G whole patches.  One packed all-gather moves the small
    per-patch tensors (semantic codes, depth, rays, DINO features: 0.5 MB per patch) so that negatives may live on
    another rank; every rank evaluates the loss rows of ITS OWN patches only (kernel B's sharded phases: the batch-wide
    `old_mean` of image.py:316-319 / :420-424 is one all-reduce of 8 scalars), the code gradients that land on remote
    negatives come home in one all-reduce, and the parameter gradients are summed with one flat all-reduce before
    optimizer.step().  Four collectives per step, no host synchronisation; loss and gradients equal the single-GPU
    step on the global batch (tests/dist_gpu_worker.py: 1e-6 / 1e-5).  The same code path runs on one GPU (the
    collectives are identities);
  * the DINO pass runs under no_grad (the reference builds a graph that carries no useful gradient);
  * the random sample coordinates of the appearance loss (image.py:343-344) come from a generator seeded with the
    step, so that all ranks evaluate the same stochastic loss;
  * reference crashes are guarded, not reproduced: `cls_` undefined when --use_dino is off (trainer.py:125), the
    unused `sacrebleu` import; the ARI logging of trainer.py:174-198 (every i_print steps, when masks are in the
    batch) runs on the device (utils/metrics.py) instead of CPU sklearn / KMeans.
Numerical quirks are kept: the FINE depth feeds both geometry-loss calls (trainer.py:159-160), double
ImageNet normalisation lives inside the injected `dino.get_vit_attn_feat`.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .. import parallel as P
from ..utils.image import get_similarity_matrix, img2mse, mse2psnr


_NORM = {}


def normalize_batch(batch):
    """trainer.py:24-29 (ImageNet mean / std); the two constants live on the device once instead of two pageable host copies per step."""
    key = (batch.device, batch.dtype)
    if key not in _NORM:
        _NORM[key] = (batch.new_tensor([0.485, 0.456, 0.406]).view(-1, 1, 1), batch.new_tensor([0.229, 0.224, 0.225]).view(-1, 1, 1))
    mean, std = _NORM[key]
    return (batch - mean) / std


def _ari_block(ret, masks, args, Bl, Ps):
    """trainer.py:174-198 on the device: ARI of the argmax / 2-means segmentation against the batch masks."""
    from ..utils.metrics import adjusted_rand_score, kmeans_labels
    logits = ret["semantics"].detach().float().reshape(Bl, Ps, Ps, -1)
    prob = logits if getattr(args, "clus_no_sfm", False) else logits.softmax(-1)
    sem_pred = logits.softmax(-1).argmax(-1)
    clus = torch.stack([kmeans_labels(prob[i], n_clusters=getattr(args, "N_cluster", 2)).reshape(Ps, Ps) for i in range(Bl)])
    gt = masks.reshape(Bl, Ps, Ps).long()
    fg = gt == 1
    return dict(clus_ari=adjusted_rand_score(gt, clus), clus_ari_fg=adjusted_rand_score(gt[fg], clus[fg]),
                sem_ari=adjusted_rand_score(gt, sem_pred), sem_ari_fg=adjusted_rand_score(gt[fg], sem_pred[fg]))


def train_one_step(batch, model, optimizer, scheduler, train_loader, global_step, losses, device, args, group=None,
                   randoms=None, coords=None):
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
    rank, ws = P.world(group)
    Bl, Ps = batch_rays.shape[0], args.patch_size                      # local patches on this rank
    Bg, q0 = Bl * ws, Bl * rank                                        # global batch (equal shares), my first patch
    M = Ps * Ps
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
    sd = ret["semantics"].shape[-1] if has_sem else 0
    if has_sem:
        sem, sem0 = patch(ret["semantics"]), patch(ret["semantics0"])

    cls_ = feat = None
    if args.use_dino:
        with torch.no_grad():
            dino_in = F.interpolate(rgb.detach().permute(0, 3, 1, 2), (Ps * args.patch_stride, Ps * args.patch_stride))
            d = dino.get_vit_attn_feat(normalize_batch(dino_in))
        cls_, feat = d["cls_"], d["feat"]

    optimizer.zero_grad()
    # ---- image loss: batch mean = sum of the ranks' local means / G (equal shares)
    img_loss_l = img2mse(rgb, gt)
    img_loss0_l = img2mse(rgb0, gt)
    local = args.rgb_w * (img_loss_l + img_loss0_l) / ws               # this rank's share of the global loss (autograd)
    zero = torch.zeros((), device=device)
    use_app = bool(args.use_correlation and has_sem and feat is not None)
    use_geo = bool(args.use_geoCorr and has_sem)
    use_con = bool(getattr(args, "use_contrast", False) and contrast_loss is not None and cls_ is not None)
    parts = [zero, zero, zero, zero]                                   # corr0, corr1, geo0, geo1: this rank's shares
    g0_all = g1_all = None
    contrast_l = zero
    if use_app or use_geo or use_con:
        # ---- ONE packed all-gather of the per-patch tensors the cross-patch negatives need
        cols = [sem0.detach().reshape(Bl, -1), sem.detach().reshape(Bl, -1)] if has_sem else []
        if use_geo:
            cols += [depth.detach().reshape(Bl, -1), ray_o.reshape(Bl, -1), ray_d.reshape(Bl, -1)]
        if feat is not None:
            cols += [feat.reshape(Bl, -1).float(), cls_.reshape(Bl, -1).float()]
        packed = P.all_gather_rows(torch.cat(cols, 1), group)          # [Bg, L]
        off = [0]

        def take(n, shape):
            t = packed[:, off[0]:off[0] + n].reshape(Bg, *shape)
            off[0] += n
            return t

        if has_sem:                                                    # leaf copies: d(loss rows of this rank)/d(code of ANY patch)
            s0_all = take(M * sd, (Ps, Ps, sd)).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
            s1_all = take(M * sd, (Ps, Ps, sd)).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        if use_geo:
            dep_all = take(M, (Ps, Ps, 1)).permute(0, 3, 1, 2).contiguous()
            ro_all = take(3 * M, (Ps, Ps, 3)).permute(0, 3, 1, 2)
            rd_all = take(3 * M, (Ps, Ps, 3)).permute(0, 3, 1, 2)
        sim = None
        if feat is not None:
            nf, cf = feat.shape[-2], feat.shape[-1]
            side = int(math.sqrt(nf))
            feat_all = take(nf * cf, (side, side, cf)).permute(0, 3, 1, 2)
            cls_all = take(cls_.shape[-1], (cls_.shape[-1],))
            sim = get_similarity_matrix(cls_all)
        # ---- phase 1 of every loss call (row means of my query patches), then ONE all-reduce of the old_mean sums
        pend = []
        # random negatives (rand_neg, or no similarity matrix) must be the same on every rank: drawn from a step-seeded host
        # generator, one permutation per loss call like the reference (image.py:349-356); the argmin choice needs nothing
        ngen = torch.Generator().manual_seed(1_000_003 * int(global_step) + 29)

        def shared_neg(mod):
            rand = bool(getattr(mod, "rand_neg", False))
            if not (rand or sim is None):
                return None
            perm = torch.randperm(Bg, generator=ngen)
            if not rand:                                               # super_perm: no patch is its own negative
                perm[perm == torch.arange(Bg)] += 1
                perm = perm % Bg
            return perm.to(device)
        if use_app:
            if coords is None:                                         # shared by all ranks: seeded with the step
                gen = torch.Generator().manual_seed(1_000_003 * int(global_step) + 17)
                shp = (2, 2, Bg, correlation_loss.feature_samples, correlation_loss.feature_samples, 2)
                pin = torch.device(device).type == "cuda"                  # pinned source: no staging copy (the CPU tests have no driver)
                coords = torch.rand(shp, generator=gen, pin_memory=pin).mul_(2).sub_(1).to(device, non_blocking=True)
            for i, s_all in enumerate((s0_all, s1_all)):
                pend.append((i, correlation_loss.begin(feat_all, s_all, sim, q0, Bl, coords=(coords[i][0], coords[i][1]),
                                                       neg=shared_neg(correlation_loss))))
        if use_geo:
            for i, s_all in enumerate((s0_all, s1_all)):                # the FINE depth for both (trainer.py:159-160)
                pend.append((2 + i, geoCorrelation_loss.begin(dep_all.clone(), s_all, [ro_all, rd_all, None], sim, q0, Bl,
                                                              neg=shared_neg(geoCorrelation_loss))))
        if pend:
            sums = torch.stack([p.sums for _, p in pend])
            P.all_reduce_sum_(sums, group)
            w = [args.correlation_w, args.correlation_w, args.Gcorrelation_w, args.Gcorrelation_w]
            for j, (i, p) in enumerate(pend):
                p.sums = sums[j]
                fn = correlation_loss if i < 2 else geoCorrelation_loss
                parts[i] = w[i] * fn.finish(p)                         # phase 2: my share of the loss, graph to s0_all / s1_all
            g0_all, g1_all = torch.autograd.grad(parts[0] + parts[1] + parts[2] + parts[3], [s0_all, s1_all], allow_unused=True)
        if use_con:
            contrast_l = args.contrast_w * contrast_loss(cls_all)       # no gradient path to the render (cls_ is detached)
    # ---- ONE all-reduce: code gradients for every patch (mine may sit on other ranks' negatives) + the logged scalars
    scal = torch.stack([p.detach() for p in parts] + [img_loss_l.detach() / ws, img_loss0_l.detach() / ws])
    if g0_all is not None or g1_all is not None:
        z = torch.zeros(Bg, sd, Ps, Ps, device=device)
        buf = torch.cat([(g0_all if g0_all is not None else z).reshape(-1), (g1_all if g1_all is not None else z).reshape(-1), scal])
        P.all_reduce_sum_(buf, group)
        n = Bg * sd * M
        g0 = buf[:n].reshape(Bg, sd, Ps, Ps)[q0:q0 + Bl].permute(0, 2, 3, 1)
        g1 = buf[n:2 * n].reshape(Bg, sd, Ps, Ps)[q0:q0 + Bl].permute(0, 2, 3, 1)
        scal = buf[2 * n:]
        local = local + (sem0 * g0).sum() + (sem * g1).sum()           # surrogate: d/d(sem) == the all-reduced code gradient
    else:
        P.all_reduce_sum_(scal, group)
    corr0, corr1, geo0, geo1, img_loss, img_loss0 = scal.unbind(0)
    loss = args.rgb_w * (img_loss + img_loss0) + corr0 + corr1 + geo0 + geo1 + contrast_l      # value of the global loss
    psnr = mse2psnr(img_loss)

    ari = dict(clus_ari=0, clus_ari_fg=0, sem_ari=0, sem_ari_fg=0)
    i_print = getattr(args, "i_print", 0)
    if has_sem and masks is not None and i_print and ((global_step % i_print == 0 and global_step > 0) or global_step == 1):
        ari = _ari_block(ret, masks, args, Bl, Ps)

    local.backward()
    P.allreduce_gradients(model.parameters(), group)          # sum of the ranks' shares
    optimizer.step()
    if scheduler is not None:
        scheduler.step(global_step)
    return dict(loss=loss, psnr=psnr, sem0=zero, sem1=zero, img0=img_loss0, img1=img_loss, contrast=contrast_l, corr0=corr0,
                corr1=corr1, geo_corr0=geo0, geo_corr1=geo1, **ari)


def save_checkpoint(path, global_step, model, optimizer):
    """trainer.py:216-222: same dictionary layout ('global_step', 'model', 'optimizer'); FusedAdam's state_dict uses
    torch.optim.Adam's names, so either side can resume from the other's file."""
    torch.save({"global_step": global_step, "model": model.state_dict(), "optimizer": optimizer.state_dict()}, path)
