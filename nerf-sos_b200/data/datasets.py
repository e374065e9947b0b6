"""Drop-ins for data/datasets.py of the reference: the same on-disk scene format and class surface.

Scene directory (written by the reference's gen_dataset.py:212-250, or by data/synthetic.py here):
    meta.json                 {"near", "far", ...optional "focal", "H", "W"}
    rays_<split>[_x<s>].npy   [N, H, W, 2, 3]   (origin, un-normalised direction)
    rgbs_<split>[_x<s>].npy   [N, H, W, 3]
    masks_<split>.npy         [N, H, W, 1]      (optional; binarised at bin_thres like datasets.py:66-69)
    poses_<split>.npy         [N, 3, 4]         (optional, ret_k)

B200-first difference: PatchNeRFDataset can hold the whole split on the GPU (`device=`) and `sample_batch` gathers B strided
PxP crops with one device-side index op -- no DataLoader worker processes, no per-step host->device copy
(the reference uses 8 workers + collate + .to(device) per step, run_nerf.py:262-268).  The map-style
__len__/__getitem__ interface of the reference is kept for the stock DataLoader + PatchBatchCollater route."""
from __future__ import annotations

import json
import math
import os
import random

import numpy as np
import torch


def _load(root, name):
    p = os.path.join(root, name)
    return np.load(p) if os.path.exists(p) else None


class BaseNeRFDataset(torch.utils.data.Dataset):
    """datasets.py:12-119 -- metadata + arrays; subclasses reshape."""

    def __init__(self, root_dir, args=None, split="train", subsample=0, cam_id=False, rgb=True, use_masks=True, bin_thres=0.3,
                 ret_k=False):
        super().__init__()
        self.split = split
        meta = os.path.join(root_dir, "meta.json")
        if not os.path.exists(meta):
            raise IOError(f"{root_dir}: no meta.json (generate the scene first, e.g. data.synthetic.write_synthetic_scene)")
        with open(meta) as f:
            self.meta_dict = json.load(f)
        if not all(k in self.meta_dict for k in ("near", "far")):
            raise IOError("Missing required meta data")
        sfx = f"_x{subsample}" if subsample != 0 else ""
        self.rays = _load(root_dir, f"rays_{split}{sfx}.npy")                   # [N, H, W, 2, 3]
        if self.rays is None:
            raise IOError(f"{root_dir}: rays_{split}{sfx}.npy not found")
        n, h, w = self.rays.shape[:3]
        self.rgbs = _load(root_dir, f"rgbs_{split}{sfx}.npy") if rgb else None
        self.masks = None
        if use_masks:
            m = _load(root_dir, f"masks_{split}.npy")
            if m is None:
                m = np.ones([n, h, w, 1])                                       # datasets.py:61-64
            self.masks = (m > bin_thres).astype(np.int64) if bin_thres != -1 else m.astype(np.float32)
        poses = _load(root_dir, f"poses_{split}.npy") if ret_k else None
        self.poses = poses if poses is not None else np.zeros([n, 3, 4])
        if ret_k:
            K = np.eye(3, dtype=np.float32)
            K[0, 0] = K[1, 1] = self.meta_dict["focal"]
            K[0, -1], K[1, -1] = self.meta_dict["W"] / 2.0, self.meta_dict["H"] / 2.0
            self.K = torch.from_numpy(K)
        self.has_cam_id = cam_id
        if cam_id:
            self.cam_ids = np.arange(n, dtype=np.int64)
        self.height, self.width, self.image_count = h, w, n
        self.image_step = h * w

    def num_images(self):
        return self.image_count

    def height_width(self):
        return self.height, self.width

    def near_far(self):
        return self.meta_dict["near"], self.meta_dict["far"]

    def radii(self):
        return 2.0 / max(self.height, self.width) * 2 / math.sqrt(12)


class RayNeRFDataset(BaseNeRFDataset):
    """datasets.py:121-171 -- one ray per item for training, one view per item otherwise."""

    def __init__(self, root_dir, args=None, split="train", subsample=0, cam_id=False, use_masks=True, bin_thres=0.3):
        super().__init__(root_dir, args, split, subsample, cam_id, True, use_masks, bin_thres)
        self.rays = torch.from_numpy(self.rays).float()
        self.rgbs = torch.from_numpy(self.rgbs).float()
        self.masks = torch.from_numpy(self.masks).long() if use_masks else torch.zeros_like(self.rgbs)[..., :1].long()
        if split == "train":
            self.rays = self.rays.reshape(-1, 2, 3)
            self.rgbs = self.rgbs.reshape(-1, 3)
            self.masks = self.masks.reshape(-1, self.masks.shape[-1])
        else:
            self.rays = self.rays.permute(0, 3, 1, 2, 4)                        # [N, 2, H, W, 3]

    def __len__(self):
        return self.rays.shape[0]

    def __getitem__(self, i):
        d = dict(rays=self.rays[i], target_s=self.rgbs[i])
        if self.split == "train" and self.has_cam_id:
            d["cam_id"] = torch.as_tensor(self.cam_ids[i // self.image_step])
        else:
            d["masks"] = self.masks[i]
        return d


class PatchNeRFDataset(BaseNeRFDataset):
    """datasets.py:173-254 -- a random crop_size x crop_size window sub-sampled with patch_stride per item
    (crop 384 / stride 6 -> the 64x64 patches of the shipped recipes)."""

    def __init__(self, root_dir, args=None, split="train", subsample=0, cam_id=False, use_masks=True, crop_size=32, patch_stride=1,
                 bin_thres=0.3, ret_k=False, device=None):
        super().__init__(root_dir, args, split, subsample, cam_id, True, use_masks, bin_thres, ret_k)
        self.use_masks, self.crop_size, self.patch_stride, self.ret_k = use_masks, crop_size, patch_stride, ret_k
        self.rays = torch.from_numpy(self.rays).float()
        self.rgbs = torch.from_numpy(self.rgbs).float()
        if use_masks:
            self.masks = torch.from_numpy(self.masks)
            self.masks = self.masks.long() if bin_thres != -1 else self.masks.float()
        else:
            self.masks = torch.zeros_like(self.rgbs)[..., :1].long()
        self.poses = torch.from_numpy(np.asarray(self.poses)).float()
        if split != "train":
            self.rays = self.rays.permute(0, 3, 1, 2, 4)                        # [N, 2, H, W, 3] (datasets.py:215)
        self.device = torch.device(device) if device is not None else None
        if self.device is not None:                                             # whole split resident in HBM
            self.rays, self.rgbs, self.masks, self.poses = (t.to(self.device) for t in (self.rays, self.rgbs, self.masks, self.poses))

    def __len__(self):
        return self.rays.shape[0]

    def patch_side(self):
        return len(range(0, self.crop_size, self.patch_stride))

    def __getitem__(self, i):
        if self.split != "train":
            return dict(rays=self.rays[i], target_s=self.rgbs[i], masks=self.masks[i])
        h0 = random.randint(0, self.height - self.crop_size)                    # datasets.py:236-237 (inclusive bounds)
        w0 = random.randint(0, self.width - self.crop_size)
        sl = (slice(h0, h0 + self.crop_size, self.patch_stride), slice(w0, w0 + self.crop_size, self.patch_stride))
        return dict(rays=self.rays[i][sl].reshape(-1, 2, 3), target_s=self.rgbs[i][sl].reshape(-1, 3),
                    masks=self.masks[i][sl].reshape(-1, self.masks.shape[-1]), poses=self.poses[i],
                    start_idx=torch.tensor([h0, w0], dtype=torch.float32))

    def sample_batch(self, batch_size, generator=None):
        """B random (view, window) crops gathered in one index op on whatever device holds the split.
        Returns the PatchBatchCollater tuple (rays [B,P*P,2,3], rgbs [B,P*P,3], masks [B,P*P,1], poses [B,3,4], start_idx [B,2])."""
        assert self.split == "train"
        dev = self.rays.device
        g = generator
        view = torch.randint(0, self.image_count, (batch_size,), generator=g)
        h0 = torch.randint(0, self.height - self.crop_size + 1, (batch_size,), generator=g)
        w0 = torch.randint(0, self.width - self.crop_size + 1, (batch_size,), generator=g)
        off = torch.arange(0, self.crop_size, self.patch_stride)
        hh = (h0[:, None] + off[None, :]).to(dev)                               # [B, P]
        ww = (w0[:, None] + off[None, :]).to(dev)
        vi = view.to(dev)[:, None, None]
        idx = (vi, hh[:, :, None], ww[:, None, :])
        B = batch_size
        return (self.rays[idx].reshape(B, -1, 2, 3), self.rgbs[idx].reshape(B, -1, 3),
                self.masks[idx].reshape(B, -1, self.masks.shape[-1]), self.poses[view.to(dev)],
                torch.stack([h0, w0], 1).float().to(dev))


class ViewNeRFDataset(BaseNeRFDataset):
    """datasets.py:256-314 -- N_rand random pixels of one view per item (optional centre pre-crop)."""

    def __init__(self, root_dir, batch_size, args=None, split="train", subsample=0, cam_id=False, precrop_iters=0, precrop_frac=0.5,
                 start_iters=0):
        super().__init__(root_dir, args, split, subsample, cam_id, True, False)
        self.batch_size, self.precrop_iters, self.precrop_frac = batch_size, precrop_iters, precrop_frac
        self.counter = self.start_iters = start_iters
        self.rays = torch.from_numpy(self.rays).float().permute(0, 3, 1, 2, 4)   # [N, 2, H, W, 3]
        self.rgbs = torch.from_numpy(self.rgbs).float()

    def __len__(self):
        return self.rays.shape[0]

    def __getitem__(self, i):
        self.counter += 1
        H, W = self.height, self.width
        if self.counter < self.precrop_iters:
            dH, dW = int(H // 2 * self.precrop_frac), int(W // 2 * self.precrop_frac)
            hs, ws = torch.arange(H // 2 - dH, H // 2 + dH), torch.arange(W // 2 - dW, W // 2 + dW)
        else:
            hs, ws = torch.arange(H), torch.arange(W)
        coords = torch.stack(torch.meshgrid(hs, ws, indexing="ij"), -1).reshape(-1, 2)
        sel = coords[np.random.choice(coords.shape[0], size=[self.batch_size], replace=False)]
        rays_o, rays_d = self.rays[i]
        d = dict(rays=torch.stack([rays_o[sel[:, 0], sel[:, 1]], rays_d[sel[:, 0], sel[:, 1]]], 1),
                 target_s=self.rgbs[i][sel[:, 0], sel[:, 1]])
        if self.split == "train" and self.has_cam_id:
            d["cam_id"] = torch.as_tensor(self.cam_ids[i])
        return d


class ExhibitNeRFDataset(BaseNeRFDataset):
    """datasets.py:317-332 -- rays only, for free-viewpoint rendering."""

    def __init__(self, root_dir, args=None, subsample=0, use_semantics=False):
        super().__init__(root_dir, args, "exhibit", subsample, False, False, use_semantics)
        self.rays = torch.from_numpy(self.rays).float().permute(0, 3, 1, 2, 4)

    def __len__(self):
        return self.rays.shape[0]

    def __getitem__(self, i):
        return dict(rays=self.rays[i])
