"""Single-GPU training step through the trainer drop-in (engines/trainer.py:32 of the reference):
kernel A forward + backward, kernel B losses, Adam on the semantic head (--fix_backbone recipe)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def test_train_steps_update_only_semantic_head_and_reduce_loss():
    import dist_gpu_worker as W
    dev = torch.device("cuda:0")
    a = W.Args()
    a.use_correlation = True
    net = W.make_net(dev)
    before = {n: p.detach().clone() for n, p in net.named_parameters()}
    from nerfsos_b200.engines.optim import FusedAdam
    opt = FusedAdam([p for p in net.parameters() if p.requires_grad], lr=2e-3)
    from nerfsos_b200.engines.lr import LRScheduler
    sched = LRScheduler(opt, 2e-3, 0.1, 250000)
    losses = [None, None, W.CorrelationLoss(a), W.GeoCorrelationLoss(a)]
    g = W.load_golden("flower_eval_256")
    B, Ps = 2, a.patch_size
    rays = torch.from_numpy(g["rays"])[:, :B * Ps * Ps].permute(1, 0, 2).reshape(B, Ps * Ps, 2, 3)
    gt = torch.rand(B, Ps * Ps, 3, generator=torch.Generator().manual_seed(0))
    torch.manual_seed(0)
    vals = []
    for step in range(6):
        torch.manual_seed(0)                     # same Philox seed and same sample coordinates every step
        out = W.train_one_step((rays, gt), [net, W.FakeDino()], opt, sched, W.Loader(), step + 1, losses, dev, a)
        vals.append(out["loss"].item())
        assert torch.isfinite(out["loss"])
    sem_part = lambda o: (o["corr0"] + o["corr1"] + o["geo_corr0"] + o["geo_corr1"]).item()
    assert vals[-1] < vals[0], vals                                    # the trainable head reduces the correlation terms
    for n, p in net.named_parameters():
        changed = not torch.equal(p.detach(), before[n])
        assert changed == ("semantic_linear" in n), n
    assert set(out) >= {"loss", "psnr", "img0", "img1", "corr0", "corr1", "geo_corr0", "geo_corr1"}
    assert abs(opt.param_groups[0]["lr"] - 2e-3 * 0.1 ** (6 / 250000)) < 1e-12
