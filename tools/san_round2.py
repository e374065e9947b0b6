"""compute-sanitizer target: the kernels rewritten late in round 2 -- ragged blocked saves + kernel C (odd point count), the split geometry
pair kernels, the tcgen05 point query, the train-mode bitonic resampling.
    compute-sanitizer --tool memcheck python tools/san_round2.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import nerfsos_b200  # noqa
from nerfsos_b200 import _lib
from nerfsos_b200.models.nerf_net import NeRFNet
from nerfsos_b200.utils.image import GeoCorrelationLoss
dev = "cuda:0"
torch.manual_seed(0)
net = NeRFNet(N_samples=64, N_importance=128, use_semantics=True, sem_with_coord=True, sem_dim=2, mode="exact", perturb=1.0, raw_noise_std=1.0).to(dev)
n = 33
g = torch.Generator().manual_seed(1)
o = torch.rand(n, 3, generator=g) * 0.6 - 0.3
d = torch.cat([torch.rand(n, 2, generator=g) - 0.5, -torch.ones(n, 1)], -1)
rays = torch.stack([o, d], 0).to(dev)
net.train()
for nm, p in net.named_parameters():
    p.requires_grad_("semantic_linear" in nm)
out = net(rays, (1.2, 12.0), N_samples=40, N_importance=25)          # 33 x 40 and 33 x 65 (odd) points: ragged groups, odd tail
(out["semantics"].sum() + out["semantics0"].sum()).backward()
torch.cuda.synchronize()
print("ragged train ok", float(out["rgb"].detach().mean()))
pts = torch.rand(64 * 3 + 7, 3, generator=g).to(dev)
raw = net.nerf_fine.query_dir(pts, (0.0, 0.0, 0.0), _lib.MODE_TC_EXACT)
torch.cuda.synchronize()
print("query ok", float(raw.mean()))


class A:
    rand_neg = False; self_corr_w = 1; use_sim_matrix = True; patch_stride = 6
    app_corr_params = [0.18, 0.67, 0.46, 0.63]; geo_corr_params = [0.18, 0.67, 0.46, 0.63]


B, Ps = 3, 24
geo = GeoCorrelationLoss(A())
depth = torch.rand(B, 1, Ps, Ps, generator=g).to(dev) * 10
code = torch.randn(B, 2, Ps, Ps, generator=g).to(dev).requires_grad_(True)
ro, rd = torch.randn(B, 3, Ps, Ps, generator=g).to(dev), torch.randn(B, 3, Ps, Ps, generator=g).to(dev)
sim = torch.rand(B, B, generator=g).to(dev)
loss = geo(depth, code, [ro, rd, None], sim)
loss.backward()
torch.cuda.synchronize()
print("geo ok", float(loss))
