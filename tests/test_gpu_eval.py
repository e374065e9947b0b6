"""Full-view evaluation and a dataset-fed training loop on a synthetic scene in the reference's on-disk format
(BASELINE configs[3]/[4] at toy size)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def scene(tmp_path_factory):
    from nerfsos_b200.data import write_synthetic_scene
    return write_synthetic_scene(str(tmp_path_factory.mktemp("scene")), n_train=3, n_val=1, n_test=2, n_exhibit=1, H=48, W=64)


def _net(**kw):
    import dist_gpu_worker as W
    return W.make_net(DEV)


def test_evaluate_full_views_device_metrics(scene, tmp_path):
    from nerfsos_b200.data import RayNeRFDataset
    from nerfsos_b200.engines.eval import eval_one_view, evaluate
    from nerfsos_b200.utils.image import img2mse, mse2psnr
    net = _net().eval()
    ds = RayNeRFDataset(scene, split="test")
    res = evaluate(net, ds, DEV, save_dir=str(tmp_path))
    assert len(res["all"]["mse"]) == 2 and all(np.isfinite(res[k]) for k in ("mse", "psnr", "ssim", "clus_ari", "sem_ari", "seg_iou"))
    assert np.isnan(res["lpips"])                                              # no LPIPS weights offline
    assert os.path.exists(tmp_path / "log.json") and os.path.exists(tmp_path / "view_001.npz")
    # one view by hand: same render, same numbers
    b = ds[0]
    ret, m = eval_one_view(net, b, ds.near_far(), ds.radii(), DEV)
    with torch.no_grad():
        direct = net(b["rays"].to(DEV), ds.near_far(), retraw=False)
    assert ret["rgb"].shape == (48, 64, 3) and torch.equal(ret["rgb"], direct["rgb"])
    mse = img2mse(direct["rgb"], b["target_s"].to(DEV))
    assert abs(float(m["psnr"]) - float(mse2psnr(mse))) < 1e-5 and abs(res["all"]["mse"][0] - float(mse)) < 1e-7
    assert ret["clustering"].shape == (48, 64, 1) and set(ret["clustering"].unique().tolist()) <= {0, 1}
    assert ret["sem"].shape == (48, 64, 1)


def test_training_loop_fed_by_device_resident_patch_sampler(scene):
    import dist_gpu_worker as W
    from nerfsos_b200.data import PatchNeRFDataset
    from nerfsos_b200.engines.lr import LRScheduler
    a = W.Args()
    a.use_correlation = True
    a.patch_size = 8
    ds = PatchNeRFDataset(scene, split="train", crop_size=8 * 5, patch_stride=5, device=DEV)
    assert ds.rays.is_cuda and ds.patch_side() == 8

    class Loader:
        dataset = ds
    net = _net()
    opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=2e-3)
    sched = LRScheduler(opt, 2e-3, 0.1, 250000)
    losses = [None, None, W.CorrelationLoss(a), W.GeoCorrelationLoss(a)]
    g = torch.Generator().manual_seed(0)
    for step in range(3):
        rays, rgbs, masks, poses, idx = ds.sample_batch(4, generator=g)
        assert rays.is_cuda and rays.shape == (4, 64, 2, 3)
        out = W.train_one_step((rays, rgbs, masks), [net, W.FakeDino()], opt, sched, Loader(), step + 1, losses, DEV, a)
        assert torch.isfinite(out["loss"])


def test_device_metrics_match_reference_metric_code():
    """The same fixture as tests/test_data_eval_cpu.py (reference ssim / sklearn ARI / KMeans / compute_iou outputs), with the
    tensors on the GPU -- the path eval_one_view takes."""
    import numpy as np
    from conftest import load_golden
    from nerfsos_b200.utils import metrics as M
    g = load_golden("metrics_ref")
    dev = "cuda:0"
    a, b = torch.from_numpy(g["img1"]).to(dev), torch.from_numpy(g["img2"]).to(dev)
    assert abs(float(M.ssim(a, b)) - float(g["ssim"])) <= 2e-6
    logits, gt = torch.from_numpy(g["logits"]).to(dev), torch.from_numpy(g["gt"]).long().to(dev)
    prob = logits.softmax(-1)
    clus = M.kmeans_labels(prob, n_clusters=2)
    ref_clus = torch.from_numpy(g["clus"]).long().to(dev)
    same = (clus == ref_clus).float().mean().item()
    assert max(same, 1 - same) >= 0.999
    sem_pred = prob.argmax(-1, keepdim=True)
    fg = gt == 1
    for k, (x, y) in dict(clus_ari=(gt, clus), clus_ari_fg=(gt[fg], clus[fg]), sem_ari=(gt, sem_pred), sem_ari_fg=(gt[fg], sem_pred[fg])).items():
        assert abs(float(M.adjusted_rand_score(x, y)) - float(g[k])) <= 1e-3, k
    assert abs(float(M.binary_iou(clus, gt)) - float(g["iou_fg"])) <= 1e-3


def test_render_video_export_density_save_checkpoint(scene, tmp_path):
    """The remaining engine entry points of the boundary: render_video (eval.py:214-270), export_density (:279-304)
    and save_checkpoint (trainer.py:216-222)."""
    from nerfsos_b200.data import ExhibitNeRFDataset
    from nerfsos_b200.engines.eval import export_density, render_video
    from nerfsos_b200.engines.optim import FusedAdam
    from nerfsos_b200.engines.trainer import save_checkpoint
    from oracle import nerf_oracle as O
    net = _net().eval()
    ds = ExhibitNeRFDataset(scene)
    written = {}
    out = render_video(net, ds, DEV, str(tmp_path / "vid"), suffix="t", find_fg=False,
                       writer=lambda path, frames, **kw: written.__setitem__(os.path.basename(path), (frames.shape, frames.dtype, kw["fps"])))
    assert set(written) == {"rgb_t.mp4", "disp_t.mp4", "sem_t.mp4", "clus_t.mp4"}
    assert written["rgb_t.mp4"] == ((1, 48, 64, 3), np.uint8, 30) and out["clus"].shape == (1, 48, 64, 1)
    with torch.no_grad():
        direct = net(ds[0]["rays"].to(DEV), ds.near_far(), retraw=False)
    assert np.array_equal(out["rgb"][0], (255 * np.clip(direct["rgb"].cpu().numpy(), 0, 1)).astype(np.uint8))
    # default writer offline: frame stacks as .npy
    render_video(net, ds, DEV, str(tmp_path / "vid2"), ret_cluster=False, find_fg=False)
    assert os.path.exists(tmp_path / "vid2" / "rgb.npy") or os.path.exists(tmp_path / "vid2" / "rgb.mp4")

    # export_density: the reference's grid (x14), zero view directions, relu of the LAST raw channel
    sig = export_density(net, extents=(0.25, 0.25, 0.25), voxel_size=2. / 64., save_dir=str(tmp_path / "dens"), device=DEV, chunk=300)
    assert sig.shape == (8, 8, 8) and os.path.exists(tmp_path / "dens" / "density.npy")
    lin = torch.linspace(-0.125, 0.125, 8)
    pts = (torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1) * 14).reshape(-1, 3).numpy()
    sd = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
    _, fine = O.split_state_dict(sd)
    mk = dict(D=net.nerf_fine.mlp.D, skips=tuple(net.nerf_fine.mlp.skips))
    ref = O.mlp_forward(fine, O.encode(pts, 10), O.encode(np.zeros_like(pts), 4), **mk)
    np.testing.assert_allclose(sig.reshape(-1), np.maximum(ref[:, -1], 0), rtol=1e-4, atol=5e-4)

    opt = FusedAdam([p for p in net.parameters()], lr=1e-3)
    path = str(tmp_path / "ck.ckpt")
    save_checkpoint(path, 7, net, opt)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert ck["global_step"] == 7 and set(ck) == {"global_step", "model", "optimizer"}
    assert set(ck["model"]) == set(net.state_dict()) and "param_groups" in ck["optimizer"]
