"""Host logic either side of the render path: on-disk scene format, dataset classes, collaters,
device-side patch sampler, and the evaluation metrics against sklearn / a direct numpy restatement.  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

import nerfsos_b200  # noqa: F401
from nerfsos_b200.data import (ExhibitNeRFDataset, PatchBatchCollater, PatchNeRFDataset, RayBatchCollater, RayNeRFDataset,
                               ViewNeRFDataset, write_synthetic_scene)
from nerfsos_b200.utils import metrics as M


@pytest.fixture(scope="module")
def scene(tmp_path_factory):
    return write_synthetic_scene(str(tmp_path_factory.mktemp("scene")), n_train=3, n_val=1, n_test=1, n_exhibit=1, H=48, W=64)


def test_scene_format_matches_reference_layout(scene):
    meta = json.load(open(os.path.join(scene, "meta.json")))
    assert {"near", "far", "focal", "H", "W"} <= set(meta)
    rays = np.load(os.path.join(scene, "rays_train.npy"))
    assert rays.shape == (3, 48, 64, 2, 3) and rays.dtype == np.float32
    assert np.load(os.path.join(scene, "rgbs_train.npy")).shape == (3, 48, 64, 3)
    assert np.load(os.path.join(scene, "masks_train.npy")).shape == (3, 48, 64, 1)
    assert not os.path.exists(os.path.join(scene, "rgbs_exhibit.npy"))
    # rays follow utils/ray.py:12-22: d_z = -1 in camera frame (identity rotation), un-normalised; o constant per view
    assert np.allclose(rays[..., 1, 2], -1.0) and np.ptp(rays[0, ..., 0, :].reshape(-1, 3), 0).max() == 0
    assert np.abs(np.linalg.norm(rays[..., 1, :], axis=-1) - 1).max() > 1e-3
    m = np.load(os.path.join(scene, "masks_train.npy"))
    assert 0.01 < m.mean() < 0.5                                               # the object covers part of every view


def test_patch_dataset_items_and_collater(scene):
    ds = PatchNeRFDataset(scene, split="train", crop_size=24, patch_stride=6)
    assert len(ds) == 3 and ds.near_far() == (1.2, 12.0) and ds.patch_side() == 4
    it = ds[1]
    assert it["rays"].shape == (16, 2, 3) and it["target_s"].shape == (16, 3) and it["masks"].shape == (16, 1)
    h0, w0 = (int(v) for v in it["start_idx"])
    full = torch.from_numpy(np.load(os.path.join(scene, "rays_train.npy")))[1]
    assert torch.equal(it["rays"].reshape(4, 4, 2, 3), full[h0:h0 + 24:6, w0:w0 + 24:6])
    rays, rgbs, masks, poses, idx = PatchBatchCollater()([ds[0], ds[2]])
    assert rays.shape == (2, 16, 2, 3) and rgbs.shape == (2, 16, 3) and masks.shape == (2, 16, 1) and poses.shape == (2, 3, 4)
    assert idx.shape == (2, 2) and masks.dtype == torch.int64
    val = PatchNeRFDataset(scene, split="val", crop_size=24, patch_stride=6)
    assert val[0]["rays"].shape == (2, 48, 64, 3)


def test_device_side_patch_sampler_equals_slicing(scene):
    ds = PatchNeRFDataset(scene, split="train", crop_size=24, patch_stride=6)
    g = torch.Generator().manual_seed(3)
    rays, rgbs, masks, poses, idx = ds.sample_batch(5, generator=g)
    assert rays.shape == (5, 16, 2, 3) and rgbs.shape == (5, 16, 3) and masks.shape == (5, 16, 1) and poses.shape == (5, 3, 4)
    g = torch.Generator().manual_seed(3)
    view = torch.randint(0, 3, (5,), generator=g)
    for b in range(5):
        h0, w0 = (int(v) for v in idx[b])
        assert 0 <= h0 <= 48 - 24 and 0 <= w0 <= 64 - 24
        assert torch.equal(rays[b].reshape(4, 4, 2, 3), ds.rays[view[b]][h0:h0 + 24:6, w0:w0 + 24:6])
        assert torch.equal(rgbs[b].reshape(4, 4, 3), ds.rgbs[view[b]][h0:h0 + 24:6, w0:w0 + 24:6])


def test_ray_view_exhibit_datasets(scene):
    tr = RayNeRFDataset(scene, split="train")
    assert len(tr) == 3 * 48 * 64 and tr[5]["rays"].shape == (2, 3)
    rays, rgbs, masks = RayBatchCollater()([tr[i] for i in range(7)])
    assert rays.shape == (2, 7, 3) and rgbs.shape == (7, 3) and masks.shape == (7, 1)
    te = RayNeRFDataset(scene, split="test")
    assert te[0]["rays"].shape == (2, 48, 64, 3)
    vw = ViewNeRFDataset(scene, 32, split="train")
    assert vw[0]["rays"].shape == (32, 2, 3) and vw[0]["target_s"].shape == (32, 3)
    ex = ExhibitNeRFDataset(scene)
    assert ex[0]["rays"].shape == (2, 48, 64, 3) and "target_s" not in ex[0]
    with pytest.raises(IOError):
        PatchNeRFDataset(os.path.join(scene, "nope"))


def test_adjusted_rand_index_matches_sklearn():
    from sklearn.metrics import adjusted_rand_score as sk
    rng = np.random.default_rng(0)
    for n, ka, kb in ((1000, 2, 2), (5000, 2, 3), (64, 4, 2)):
        a, b = rng.integers(0, ka, n), rng.integers(0, kb, n)
        b[: n // 2] = a[: n // 2] % kb
        assert abs(float(M.adjusted_rand_score(torch.from_numpy(a), torch.from_numpy(b))) - sk(a, b)) < 1e-6
    z = torch.zeros(10)
    assert float(M.adjusted_rand_score(z, z)) == 1.0 == sk(z.numpy(), z.numpy())
    a = torch.tensor([0, 0, 1, 1])
    assert abs(float(M.adjusted_rand_score(a, 1 - a)) - 1.0) < 1e-12           # label permutation invariance


def test_ssim_matches_direct_restatement():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(20, 24, 3, generator=g)
    y = (x + 0.1 * torch.randn(20, 24, 3, generator=g)).clamp(0, 1)
    assert abs(float(M.ssim(x, x, format="HWC")) - 1.0) < 1e-6
    got = float(M.ssim(x, y, format="HWC"))
    # direct numpy: zero-padded 11x11 gaussian (sigma 1.5) windows, per channel (utils/ssim.py:17-40)
    k = np.exp(-(np.arange(11) - 5) ** 2 / (2 * 1.5 ** 2)); k /= k.sum()
    w = np.outer(k, k).astype(np.float32)
    def filt(a):
        p = np.pad(a, ((5, 5), (5, 5), (0, 0)))
        out = np.zeros_like(a)
        for i in range(a.shape[0]):
            for j in range(a.shape[1]):
                out[i, j] = (p[i:i + 11, j:j + 11] * w[..., None]).sum((0, 1))
        return out
    a, b = x.numpy(), y.numpy()
    mu1, mu2 = filt(a), filt(b)
    s1, s2, s12 = filt(a * a) - mu1 ** 2, filt(b * b) - mu2 ** 2, filt(a * b) - mu1 * mu2
    ref = (((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 ** 2 + mu2 ** 2 + 1e-4) * (s1 + s2 + 9e-4))).mean()
    assert abs(got - ref) < 1e-5, (got, ref)


def test_two_means_matches_sklearn_partition():
    from sklearn.cluster import KMeans
    from sklearn.metrics import adjusted_rand_score as sk
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.normal([0.8, 0.2], 0.05, (300, 2)), rng.normal([0.3, 0.7], 0.08, (500, 2))]).astype(np.float32)
    x = x[rng.permutation(800)].reshape(20, 40, 2)
    lab = M.kmeans_labels(torch.from_numpy(x), 2)
    assert lab.shape == (20, 40, 1) and lab.dtype == torch.int64
    ref = KMeans(n_clusters=2, random_state=0, n_init=10).fit(x.reshape(-1, 2)).labels_
    assert sk(ref, lab.reshape(-1).numpy()) == 1.0
    gt = torch.from_numpy(ref).reshape(20, 40, 1)
    assert float(M.binary_iou(lab, gt)) == 1.0 and float(M.binary_iou(1 - lab, gt)) == 1.0


def test_metrics_match_the_reference_metric_code():
    """utils/metrics.py against outputs of the reference's OWN metric code (utils/ssim.py:7-40, sklearn ARI as called by
    engines/eval.py:70-74, utils/misc.py:40-50 KMeans, utils/get_metrics.py:15-26 compute_iou), generated through the shim by
    oracle/make_golden.py:metrics."""
    import numpy as np
    import torch
    from conftest import load_golden
    from nerfsos_b200.utils import metrics as M
    g = load_golden("metrics_ref")
    a, b = torch.from_numpy(g["img1"]), torch.from_numpy(g["img2"])
    assert abs(float(M.ssim(a, b)) - float(g["ssim"])) <= 1e-6
    np.testing.assert_allclose(M.ssim(a, b, size_average=False).numpy(), g["ssim_each"], rtol=0, atol=1e-6)
    assert abs(float(M.ssim(a[0].permute(1, 2, 0), b[0].permute(1, 2, 0), format="HWC")) - float(g["ssim_each"][0])) <= 1e-6
    logits, gt = torch.from_numpy(g["logits"]), torch.from_numpy(g["gt"]).long()
    prob = logits.softmax(-1)
    clus = M.kmeans_labels(prob, n_clusters=2)
    ref_clus = torch.from_numpy(g["clus"]).long()
    assert clus.shape == ref_clus.shape
    same = (clus == ref_clus).float().mean().item()
    assert max(same, 1 - same) >= 0.999                                   # the same partition (labels may be swapped)
    sem_pred = prob.argmax(-1, keepdim=True)
    assert torch.equal(sem_pred, torch.from_numpy(g["sem_pred"]).long())
    fg = gt == 1
    for k, (x, y) in dict(clus_ari=(gt, ref_clus), clus_ari_fg=(gt[fg], ref_clus[fg]), sem_ari=(gt, sem_pred), sem_ari_fg=(gt[fg], sem_pred[fg])).items():
        assert abs(float(M.adjusted_rand_score(x, y)) - float(g[k])) <= 1e-6, k
    assert abs(float(M.binary_iou(ref_clus, gt)) - float(g["iou_fg"])) <= 1e-6
