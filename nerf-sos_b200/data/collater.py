"""Drop-ins for data/collater.py of the reference (same names, same return tuples):
RayBatchCollater :7-28, PatchBatchCollater :30-60, ViewBatchCollater :62-83, ExhibitCollater :85-99."""
from __future__ import annotations

import torch


def _stack(xs, key, cat=False):
    if key not in xs[0]:
        return None
    ts = [torch.as_tensor(x[key]) for x in xs]
    return torch.cat(ts, 0) if cat else torch.stack(ts, 0)


class RayBatchCollater:
    def __call__(self, xs):
        rays = _stack(xs, "rays").transpose(0, 1)                      # [2, B, 3]
        out = (rays, _stack(xs, "target_s"), _stack(xs, "masks"))
        return out + (_stack(xs, "cam_id"),) if "cam_id" in xs[0] else out


class PatchBatchCollater:
    """[B, P*P, 2, 3] rays, [B, P*P, 3] rgbs, masks, poses, start_idx -- the batch train_one_step consumes."""
    def __call__(self, xs):
        rays, rgbs, masks = _stack(xs, "rays"), _stack(xs, "target_s"), _stack(xs, "masks")
        if "cam_id" in xs[0]:
            return rays, rgbs, masks, _stack(xs, "cam_id")
        return rays, rgbs, masks, _stack(xs, "poses"), _stack(xs, "start_idx")


class ViewBatchCollater:
    def __call__(self, xs):
        rays = _stack(xs, "rays", cat=True).transpose(0, 1)
        rgbs = _stack(xs, "target_s", cat=True)
        if "cam_id" in xs[0]:
            ids = torch.cat([torch.full((rays.shape[0],), int(x["cam_id"]), dtype=torch.int64) for x in xs], 0)
            return rays, rgbs, ids
        return rays, rgbs


class ExhibitCollater:
    def __init__(self, H, W):
        self.H, self.W = H, W

    def __call__(self, xs):
        rays = _stack(xs, "rays").transpose(0, 1)
        rays = rays.reshape(rays.shape[0], self.H, self.W, rays.shape[-1])
        rgbs = _stack(xs, "target_s")
        if rgbs is not None:
            rgbs = rgbs.reshape(self.H, self.W, rgbs.shape[-1])
        return rays, rgbs
