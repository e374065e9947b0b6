"""Synthetic-scene writer in the reference's on-disk format (no datasets are reachable offline).

An analytic forward-facing scene -- a coloured sphere (the 'object', mask = 1) in front of a textured back plane -- is
ray-traced at LLFF-like poses; rays follow utils/ray.py:12-22 (`d = ((i-W/2)/f, -(j-H/2)/f, -1)` rotated by c2w,
un-normalised, `o = c2w[:3,3]`).  Output: meta.json, rays_/rgbs_/masks_/poses_<split>.npy for train / val / test /
exhibit, exactly what data/datasets.py reads (gen_dataset.py:212-250 of the reference)."""
from __future__ import annotations

import json
import os

import numpy as np


def persp_rays(H, W, focal, c2w):
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    dirs = np.stack([(i - W / 2) / focal, -(j - H / 2) / focal, -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape)
    return np.stack([rays_o, rays_d], 2).astype(np.float32)                      # [H, W, 2, 3]


def _shade(rays, centre, radius, plane_z):
    o, d = rays[..., 0, :], rays[..., 1, :]
    oc = o - centre
    a, b, c = (d * d).sum(-1), 2 * (oc * d).sum(-1), (oc * oc).sum(-1) - radius ** 2
    disc = b * b - 4 * a * c
    t_s = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), 1e6)
    hit = (disc > 0) & (t_s > 0) & (t_s < 1e5)
    t_p = (plane_z - o[..., 2]) / d[..., 2]
    p = o + d * t_p[..., None]
    checker = ((np.floor(p[..., 0] * 2) + np.floor(p[..., 1] * 2)) % 2)[..., None]
    plane_rgb = 0.25 + 0.5 * checker * np.array([0.9, 0.8, 0.6], np.float32)
    ps = o + d * t_s[..., None]
    n = (ps - centre) / radius
    lam = np.clip((n * np.array([0.3, 0.5, 0.8], np.float32)).sum(-1, keepdims=True), 0.1, 1.0)
    sphere_rgb = lam * np.array([0.9, 0.2, 0.2], np.float32)
    rgb = np.where(hit[..., None], sphere_rgb, plane_rgb).astype(np.float32)
    return rgb, hit[..., None].astype(np.float32)


def write_synthetic_scene(root, n_train=6, n_val=1, n_test=2, n_exhibit=2, H=96, W=128, focal=None, near=1.2, far=12.0, seed=0):
    os.makedirs(root, exist_ok=True)
    focal = float(focal if focal is not None else 815.0 * W / 1008.0)            # LLFF flower: f=815 at 1008x756
    rng = np.random.default_rng(seed)
    centre, radius, plane_z = np.array([0.0, 0.0, -4.0], np.float32), 0.8, -8.0
    with open(os.path.join(root, "meta.json"), "w") as f:
        json.dump({"near": near, "far": far, "focal": focal, "H": H, "W": W, "generator": "nerfsos_b200.data.synthetic"}, f)
    for split, n in (("train", n_train), ("val", n_val), ("test", n_test), ("exhibit", n_exhibit)):
        rays, rgbs, masks, poses = [], [], [], []
        for _ in range(n):
            c2w = np.concatenate([np.eye(3, dtype=np.float32), rng.uniform(-0.3, 0.3, (3, 1)).astype(np.float32)], 1)
            r = persp_rays(H, W, focal, c2w)
            rgb, m = _shade(r, centre, radius, plane_z)
            rays.append(r); rgbs.append(rgb); masks.append(m); poses.append(c2w)
        np.save(os.path.join(root, f"rays_{split}.npy"), np.stack(rays))
        np.save(os.path.join(root, f"masks_{split}.npy"), np.stack(masks))
        np.save(os.path.join(root, f"poses_{split}.npy"), np.stack(poses))
        if split != "exhibit":
            np.save(os.path.join(root, f"rgbs_{split}.npy"), np.stack(rgbs))
    return root
