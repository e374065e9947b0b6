"""Multi-GPU equivalence worker (launched by tests/test_gpu_dist.py through torchrun, one rank per GPU):
  1. ray-sharded render == single-GPU render, bitwise per ray;
  2. one data-parallel training step (patches sharded; ONE packed all-gather of the per-patch tensors, kernel B evaluated
     per rank on its own patches with the batch-wide old_mean all-reduced, code gradients of remote negatives all-reduced,
     flat NCCL gradient all-reduce) reproduces the single-GPU loss (1e-6 rel) and gradients (1e-5 rel of max) on the
     global batch -- appearance AND geometry correlation losses on, same sample coordinates on every rank.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nerfsos_b200  # noqa: E402,F401
from nerfsos_b200 import parallel as P  # noqa: E402
from nerfsos_b200.engines.trainer import train_one_step  # noqa: E402
from nerfsos_b200.models.nerf_net import NeRFNet  # noqa: E402
from nerfsos_b200.utils.image import CorrelationLoss, GeoCorrelationLoss  # noqa: E402
from conftest import load_golden  # noqa: E402


class Args:
    patch_tune = True; patch_size = 8; patch_stride = 6; batch_size = 4
    use_dino = True; use_correlation = True; use_geoCorr = True; use_contrast = False
    rgb_w = 1.0; correlation_w = 1.0; Gcorrelation_w = 0.01; contrast_w = 0.0
    rand_neg = False; self_corr_w = 1; use_sim_matrix = True
    app_corr_params = [0.18, 1, 0.46, 1]; geo_corr_params = [0.5, 1, 3, 1]


class FakeDino:
    """Deterministic stand-in feature provider with the extractor's interface (no DINO weights offline)."""
    def get_vit_attn_feat(self, x):
        B = x.shape[0]
        p = torch.nn.functional.adaptive_avg_pool2d(x, 14).reshape(B, 3, 196).permute(0, 2, 1)      # [B,196,3]
        g = torch.Generator().manual_seed(0)
        proj = torch.randn(3, 384, generator=g).to(x.device)
        feat = p @ proj
        return {"attn": feat[..., :1].permute(0, 2, 1), "cls_": feat.mean(1), "feat": feat}


class DS:
    def near_far(self): return 1.2, 12.0
    def radii(self): return None


class Loader:
    dataset = DS()


def make_net(dev, all_params=False):
    sd = load_golden("flower_weights")["sd"]
    net = NeRFNet(N_samples=64, N_importance=128, use_semantics=True, sem_with_coord=True, sem_dim=2, perturb=1.0,
                  raw_noise_std=1.0, mode="exact")
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    net = net.to(dev)
    for n, p in net.named_parameters():
        p.requires_grad_(all_params or "semantic_linear" in n)        # --fix_backbone unless all_params (stage-1 training)
    return net


def main():
    rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    a = Args()
    B, Ps = a.batch_size, a.patch_size
    g = load_golden("flower_eval_256")
    rays = torch.from_numpy(g["rays"]).to(dev)                        # [2,256,3]
    # ---- 1. sharded render
    net = make_net(dev).eval()
    full = net(rays, (1.2, 12.0))
    sh = P.render_sharded(net, rays, (1.2, 12.0))
    for k in sh:
        assert torch.equal(sh[k], full[k]), k
    # ---- 2. DP train step vs single GPU on the global batch (injected randoms so both runs draw the same noise)
    gen = torch.Generator().manual_seed(1)
    N = B * Ps * Ps
    rnd = {"t_rand": torch.rand(N, 64, generator=gen), "noise0": torch.randn(N, 64, generator=gen),
           "u": torch.rand(N, 128, generator=gen), "noise1": torch.randn(N, 192, generator=gen)}
    rnd = {k: v.to(dev) for k, v in rnd.items()}
    rays_p = rays[:, :N].permute(1, 0, 2).reshape(B, Ps * Ps, 2, 3)    # [B, P*P, 2, 3] as PatchBatchCollater yields
    gt = torch.rand(B, Ps * Ps, 3, generator=gen).to(dev)

    def run(lo, hi, group, all_params=False):
        net = make_net(dev, all_params)
        from nerfsos_b200.engines.optim import FusedAdam
        opt = FusedAdam([p for p in net.parameters() if p.requires_grad], lr=5e-4)
        losses = [None, None, CorrelationLoss(a), GeoCorrelationLoss(a)]
        sl = slice(lo * Ps * Ps, hi * Ps * Ps)
        r = {k: v[sl] for k, v in rnd.items()}
        out = train_one_step((rays_p[lo:hi], gt[lo:hi]), [net, FakeDino()], opt, None, Loader(), 1, losses, dev, a,
                             group=group, randoms=r)
        grads = torch.cat([p.grad.reshape(-1) for p in net.parameters() if p.requires_grad])
        return out["loss"].item(), grads

    class _Solo:                                                     # run the reference single-GPU step without collectives
        pass
    lo, hi = P.shard_bounds(B, rank, ws)
    l_dp, g_dp = run(lo, hi, None)                                   # default group: all ranks
    # single-GPU run of the whole batch: temporarily behave as world_size 1
    orig = P.world
    P.world = lambda group=None: (0, 1)
    l_1, g_1 = run(0, B, None)
    P.world = orig
    rel_l = abs(l_dp - l_1) / max(1e-12, abs(l_1))
    rel_g = (g_dp - g_1).abs().max().item() / g_1.abs().max().item()
    print(f"rank {rank}: loss dp={l_dp:.8f} single={l_1:.8f} rel={rel_l:.2e}; grad rel err {rel_g:.2e}", flush=True)
    assert rel_l <= 1e-6 and rel_g <= 1e-5, (rel_l, rel_g)
    # ---- 3. the same with EVERY parameter trainable: the 5.1 MB gradient all-reduce of north_star, tensor-core backward
    l_dp, g_dp = run(lo, hi, None, all_params=True)
    P.world = lambda group=None: (0, 1)
    l_1, g_1 = run(0, B, None, all_params=True)
    P.world = orig
    rel_l = abs(l_dp - l_1) / max(1e-12, abs(l_1))
    rel_g = (g_dp - g_1).abs().max().item() / g_1.abs().max().item()
    print(f"rank {rank}: all parameters ({g_1.numel()} gradients): loss rel={rel_l:.2e}; grad rel err {rel_g:.2e}", flush=True)
    assert g_1.numel() == 1274124 and rel_l <= 1e-6 and rel_g <= 1e-4, (rel_l, rel_g)
    dist.barrier()
    if rank == 0:
        print("DIST_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
