"""Small all-parameter training step for compute-sanitizer: replay-all kernel (MODE 3), tcgen05 row GEMM and weight-gradient
kernels, narrow-head kernels, kernel B sharded phases, fused Adam.
    compute-sanitizer --tool memcheck python tools/san_allparams.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nerfsos_b200  # noqa
import dist_gpu_worker as W
from nerfsos_b200.engines.optim import FusedAdam
dev = torch.device("cuda:0")
a = W.Args(); a.patch_size = 4; a.batch_size = 2
net = W.make_net(dev, all_params=True)
opt = FusedAdam([p for p in net.parameters() if p.requires_grad], lr=5e-4)
losses = [None, None, W.CorrelationLoss(a), W.GeoCorrelationLoss(a)]
g = W.load_golden("flower_eval_256")
B, Ps = 2, a.patch_size
rays = torch.from_numpy(g["rays"])[:, :B * Ps * Ps].permute(1, 0, 2).reshape(B, Ps * Ps, 2, 3)
gt = torch.rand(B, Ps * Ps, 3, generator=torch.Generator().manual_seed(0))
for step in range(2):
    out = W.train_one_step((rays, gt), [net, W.FakeDino()], opt, None, W.Loader(), step + 1, losses, dev, a)
    torch.cuda.synchronize()
    print("step", step, float(out["loss"]))
