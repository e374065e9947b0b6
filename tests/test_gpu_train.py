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


@pytest.mark.parametrize("mode,tol", [("simt", 1e-4), ("exact", 2e-4)])
def test_five_steps_match_the_reference_trainer(mode, tol):
    """SURVEY Appendix B 'end-to-end': five consecutive steps of the drop-in trainer (kernel A fwd/bwd, kernel B, fused Adam,
    LR schedule) against the UNMODIFIED reference trainer on CPU (tests/golden/flower_train_steps.npz, produced through the shim
    by oracle/make_golden.py:train_steps): same weights, batch, stand-in feature provider and -- injected -- every random draw
    the reference made.  Each loss term of each step within 1e-4 (fp32 path) / 2e-4 (fp16 hi/lo path: its importance samples
    differ in ill-conditioned low-mass bins), the trained semantic heads within Adam's step noise."""
    import numpy as np
    import dist_gpu_worker as W
    from nerfsos_b200.engines.lr import LRScheduler
    from nerfsos_b200.engines.optim import FusedAdam
    from nerfsos_b200.models.nerf_net import NeRFNet
    dev = torch.device("cuda:0")
    g = W.load_golden("flower_train_steps")
    a = W.Args()
    a.use_correlation = True; a.patch_size = 8; a.batch_size = 2
    net = NeRFNet(N_samples=64, N_importance=128, use_semantics=True, sem_with_coord=True, sem_dim=2, perturb=1.0, raw_noise_std=1.0,
                  mode=mode)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in W.load_golden("flower_weights")["sd"].items()})
    net = net.to(dev)
    for n, p in net.named_parameters():
        p.requires_grad_("semantic_linear" in n)
    opt = FusedAdam([p for p in net.parameters() if p.requires_grad], lr=5e-4, betas=(0.9, 0.999))
    sched = LRScheduler(opt, 5e-4, 0.1, 250000)
    losses = [None, None, W.CorrelationLoss(a), W.GeoCorrelationLoss(a)]
    rays, gt = torch.from_numpy(g["rays"]), torch.from_numpy(g["gt"])
    names = ("loss", "img0", "img1", "corr0", "corr1", "geo_corr0", "geo_corr1")
    for k in range(5):
        r = g["rnd"]
        rnd = {n: torch.from_numpy(r[f"s{k}_{n}"]).to(dev) for n in ("t_rand", "noise0", "u", "noise1")}
        coords = torch.stack([torch.stack([torch.from_numpy(r[f"s{k}_c{i}_{j}"]) for j in (1, 2)]) for i in (0, 1)]).to(dev) * 2 - 1
        out = W.train_one_step((rays, gt), [net, W.FakeDino()], opt, sched, W.Loader(), 2 + k, losses, dev, a, randoms=rnd, coords=coords)
        for n, ref in zip(names, g["steps"][k]):
            got = float(out[n])
            assert abs(got - ref) <= tol * max(1.0, abs(ref)), (mode, k, n, got, ref)
    assert abs(opt.param_groups[0]["lr"] - float(g["lr_last"])) <= 1e-12
    for n, p in net.named_parameters():
        if p.requires_grad:
            d = np.abs(p.detach().cpu().numpy() - g["final"][n])
            assert (d <= 2e-5).mean() >= 0.98 and d.max() <= 5e-3, (n, (d <= 2e-5).mean(), d.max())


def test_random_negatives_are_step_seeded():
    """--rand_neg (image.py:349-356): the permutation of negatives comes from a step-seeded host generator so that every rank of
    a data-parallel run picks the same ones: same step -> same loss terms, another step -> another permutation."""
    import dist_gpu_worker as W
    dev = torch.device("cuda:0")
    a = W.Args()
    a.use_correlation = True
    a.rand_neg = True
    g = W.load_golden("flower_eval_256")
    B, Ps = 4, a.patch_size
    n = B * Ps * Ps
    rays = torch.from_numpy(g["rays"]).repeat(1, (n + 255) // 256, 1)[:, :n].permute(1, 0, 2).reshape(B, Ps * Ps, 2, 3).contiguous()
    rays[..., 0, :] += torch.linspace(0, 0.3, B).view(B, 1, 1)               # four different patches
    gt = torch.rand(B, Ps * Ps, 3, generator=torch.Generator().manual_seed(0))
    from nerfsos_b200.engines.lr import LRScheduler
    from nerfsos_b200.engines.optim import FusedAdam

    def geo_terms(step):
        net = W.make_net(dev)
        opt = FusedAdam([p for p in net.parameters() if p.requires_grad], lr=0.0)
        losses = [None, None, W.CorrelationLoss(a), W.GeoCorrelationLoss(a)]
        torch.manual_seed(0)
        out = W.train_one_step((rays, gt), [net, W.FakeDino()], opt, LRScheduler(opt, 0.0, 0.1, 250000), W.Loader(), step, losses, dev, a)
        return [out[k].item() for k in ("corr0", "corr1", "geo_corr0", "geo_corr1")]

    t1, t1b, t2 = geo_terms(1), geo_terms(1), geo_terms(2)
    assert t1 == t1b
    assert any(abs(x - y) > 1e-7 for x, y in zip(t1, t2))
