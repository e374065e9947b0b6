"""Drop-in for engines/eval.py of the reference: eval_one_view (:31-98), evaluate (:101-212), render_video (:214-270) and
export_density (:279-304).

A view is rendered by ONE call into the fused kernel (no ray_chunk loop, no [H*W,192,6] `raw` tensor -- 3.5 GB per
1008x756 view in the reference), optionally ray-sharded over the ranks of a process group, and every metric is computed
on the device.  LPIPS and the DINO-attention polarity check need external weights that do not exist offline: `lpips_fn` /
`dino` are optional call-ins with the reference's interfaces; without them 'lpips' is NaN and the cluster polarity is
left as found (ARI is invariant to it)."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from .. import parallel as P
from ..utils.image import img2mse, mse2psnr
from ..utils.metrics import adjusted_rand_score, binary_iou, kmeans_labels, ssim


def eval_one_view(model, batch, near_far, radii, device, clus_no_sfm=False, N_cluster=2, group=None, lpips_fn=None,
                  **render_kwargs):
    model.eval()
    near, far = near_far
    with torch.no_grad():
        rays = batch["rays"].to(device)                                   # [2, H, W, 3]
        render_kwargs.setdefault("retraw", False)
        if P.world(group)[1] > 1:
            ret = P.render_sharded(model, rays, (near, far), group=group, keys=None, radii=radii, **render_kwargs)   # every per-ray key
        else:
            ret = model(rays, (near, far), radii=radii, **render_kwargs)
        ret = dict(ret)
        zero = torch.zeros(1, device=device)
        clus_ari = clus_ari_fg = sem_ari = sem_ari_fg = iou = zero
        if "semantics" in ret:
            logits = ret["semantics"].float()
            prob = logits if clus_no_sfm else logits.softmax(-1)
            sem_pred = logits.softmax(-1).argmax(-1, keepdim=True)
            clus = kmeans_labels(prob, n_clusters=N_cluster)
            ret["sem"], ret["clustering"] = sem_pred, clus
            if "masks" in batch:
                gt = batch["masks"].to(device).long().reshape(sem_pred.shape)
                fg = gt == 1
                clus_ari, sem_ari = adjusted_rand_score(gt, clus).reshape(1), adjusted_rand_score(gt, sem_pred).reshape(1)
                clus_ari_fg, sem_ari_fg = adjusted_rand_score(gt[fg], clus[fg]).reshape(1), adjusted_rand_score(gt[fg], sem_pred[fg]).reshape(1)
                iou = binary_iou(clus, gt).reshape(1)
        metrics = {}
        if "target_s" in batch:
            tgt = batch["target_s"].to(device)
            ret["target_s"] = tgt
            mse = img2mse(ret["rgb"], tgt)
            metrics = {"mse": mse, "psnr": mse2psnr(mse), "ssim": ssim(ret["rgb"], tgt, format="HWC"),
                       "lpips": lpips_fn(ret["rgb"], tgt, format="HWC") if lpips_fn else torch.tensor(float("nan"), device=device),
                       "clus_ari": clus_ari, "clus_ari_fg": clus_ari_fg, "sem_ari": sem_ari, "sem_ari_fg": sem_ari_fg, "seg_iou": iou}
        return ret, metrics


KEYS = ["mse", "psnr", "ssim", "lpips", "clus_ari", "clus_ari_fg", "sem_ari", "sem_ari_fg", "seg_iou"]


def evaluate(model, dataset, device, save_dir=None, fast_mode=False, ret_cluster=False, clus_no_sfm=False, N_cluster=2, find_fg=False,
             dino=None, group=None, lpips_fn=None, **render_kwargs):
    near, far = dataset.near_far()
    radii = dataset.radii()
    allm = {k: [] for k in KEYS}
    for i in range(len(dataset)):
        if fast_mode and i >= 1:
            break
        batch = dataset[i]
        ret, m = eval_one_view(model, batch, (near, far), radii, device, clus_no_sfm, N_cluster, group=group, lpips_fn=lpips_fn,
                               **render_kwargs)
        for k in KEYS:
            allm[k].append(float(m[k]))
        if find_fg and dino is not None and "clustering" in ret:          # eval.py:133-144: foreground = higher DINO attention
            from .trainer import normalize_batch
            x = normalize_batch(ret["rgb"].permute(2, 0, 1).unsqueeze(0))
            attn = dino.get_vit_attn_feat_noresize(x)["attn"].reshape(1, 1, x.shape[2] // 16, x.shape[3] // 16)
            attn = torch.nn.functional.interpolate(attn, x.shape[2:]).reshape(x.shape[2], x.shape[3], 1)
            c = ret["clustering"]
            if attn[c == 1].mean() < attn[c == 0].mean():
                ret["clustering"] = 1 - c
        if save_dir is not None:
            os.makedirs(save_dir, exist_ok=True)
            np.savez_compressed(os.path.join(save_dir, f"view_{i:03d}.npz"),
                                **{k: ret[k].cpu().numpy() for k in ("rgb", "depth", "acc", "sem", "clustering") if k in ret})
    tot = {f"total_{k}": float(np.mean(v)) for k, v in allm.items() if v}
    if allm["mse"]:
        tot["total_psnr"] = float(mse2psnr(torch.tensor(tot["total_mse"])))   # PSNR of the mean MSE (eval.py:182-183)
    allm.update(tot)
    if save_dir is not None:
        with open(os.path.join(save_dir, "log.json"), "w") as f:
            json.dump(allm, f)
    return {"mse": tot.get("total_mse"), "psnr": tot.get("total_psnr"), "ssim": tot.get("total_ssim"), "lpips": tot.get("total_lpips"),
            "clus_ari": tot.get("total_clus_ari"), "sem_ari": tot.get("total_sem_ari"), "seg_iou": tot.get("total_seg_iou"), "all": allm}


def _to8b(x):
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def render_video(model, dataset, device, save_dir, suffix="", fps=30, quality=8, ret_cluster=True, fast_mode=False, clus_no_sfm=False,
                 N_cluster=2, find_fg=True, dino=None, group=None, writer=None, **render_kwargs):
    """eval.py:214-270: renders every view of `dataset` (one fused launch per view) and writes rgb / disp / sem / clus videos.
    `writer(path, frames_uint8, fps=, quality=)` defaults to imageio.mimwrite when imageio is importable; offline the uint8
    frame stacks are saved as `<name>.npy` next to where the .mp4 would go.  The foreground polarity check (:231-245) needs a
    DINO extractor: pass `dino=` (the reference builds one from torch.hub here); without it the clustering is kept as found.
    Returns the frame stacks."""
    near, far = dataset.near_far()
    radii = dataset.radii()
    rgbs, disps, sems, clusters = [], [], [], []
    ret = {}
    for i in range(len(dataset)):
        if fast_mode and i >= 2:
            break
        ret, _ = eval_one_view(model, dataset[i], (near, far), radii, device, clus_no_sfm, N_cluster, group=group, **render_kwargs)
        if "sem" in ret:
            sems.append(ret["sem"].float().cpu().numpy())
        if ret_cluster and "clustering" in ret:
            c = ret["clustering"]
            if find_fg and dino is not None:
                from .trainer import normalize_batch
                x = normalize_batch(ret["rgb"].permute(2, 0, 1).unsqueeze(0))
                attn = dino.get_vit_attn_feat_noresize(x)["attn"].reshape(1, 1, x.shape[2] // 16, x.shape[3] // 16)
                attn = torch.nn.functional.interpolate(attn, x.shape[2:]).reshape(x.shape[2], x.shape[3], 1)
                if attn[c == 1].mean() < attn[c == 0].mean():
                    c = 1 - c
            clusters.append(c.cpu().numpy())
        rgbs.append(ret["rgb"].cpu().numpy())
        disps.append(ret["disp"].cpu().numpy())
    if writer is None:
        try:
            import imageio
            writer = getattr(imageio, "mimwrite", None)            # (a stub module without it counts as absent)
        except ImportError:
            writer = None
        if writer is None:
            writer = lambda path, frames, **kw: np.save(os.path.splitext(path)[0] + ".npy", frames)
    os.makedirs(save_dir, exist_ok=True)
    name = lambda k: os.path.join(save_dir, f"{k}{'_' + suffix if suffix else ''}.mp4")
    out = {"rgb": _to8b(np.stack(rgbs, 0))}
    disp = np.stack(disps, 0)
    out["disp"] = _to8b(disp / np.max(disp))
    if "semantics" in ret and sems:
        out["sem"] = _to8b(np.stack(sems, 0))
    if ret_cluster and clusters:
        out["clus"] = (np.stack(clusters, 0) * 255).astype(np.uint8)
    for k, frames in out.items():
        writer(name(k), frames, fps=fps, quality=quality)
    return out


def export_density(model, extents=(2.0, 2.0, 2.0), voxel_size=2. / 256., save_dir="", device=None, chunk=1 << 22):
    """eval.py:279-304: queries `model.nerf_fine` on a regular grid (x14, zero view directions) and returns
    relu(raw[..., -1]) as a numpy volume [W, H, D] -- the LAST raw channel, exactly as the reference does (with a segmentation
    head that is the last semantic logit, not sigma: raw = [rgb(3), sigma, sem...], nerf_mlp.py:94; use `channel=` semantics by
    slicing `model.nerf_fine(pts, viewdirs=...)` yourself if sigma is wanted).  The query runs through nsos_mlp_query_dir (tensor cores;
    nsos_mlp_query for mode='simt') in slices
    of `chunk` points; `density.npy` is written when `save_dir` is given (the reference's .mrc / .ply writers need mrc / open3d)."""
    model.eval()
    device = device or next(model.parameters()).device
    with torch.no_grad():
        h, w, d = extents
        pts = torch.stack(torch.meshgrid(torch.linspace(-w / 2, w / 2, int(w / voxel_size)), torch.linspace(-h / 2, h / 2, int(h / voxel_size)),
                                         torch.linspace(-d / 2, d / 2, int(d / voxel_size)), indexing="ij"), dim=-1).to(device).float()
        pts = pts * 14
        flat = pts.reshape(-1, 3)
        sig = torch.empty(flat.shape[0], device=device)
        mode = model.resolve_mode() if hasattr(model, "resolve_mode") else 0
        for i in range(0, flat.shape[0], chunk):
            x = flat[i:i + chunk]
            if mode != 0 and hasattr(model.nerf_fine, "query_dir"):      # one view direction for all points: the tcgen05 query
                raw = model.nerf_fine.query_dir(x, (0.0, 0.0, 0.0), mode)
            else:
                raw = model.nerf_fine(x, viewdirs=torch.zeros_like(x))
            sig[i:i + chunk] = raw[..., -1].clamp_min(0)
        sigma = sig.reshape(pts.shape[:-1]).cpu().numpy()
    if save_dir:
        os.makedirs(save_dir, exist_ok=True)
        np.save(os.path.join(save_dir, "density.npy"), sigma)
    return sigma
