"""Small render for compute-sanitizer (SURVEY.md appendix B "hygiene"): 70 rays (odd pair count per CTA, one padding ray),
eval and train mode, exact tcgen05 path, flower-shaped net with random weights.
    compute-sanitizer --tool memcheck  python tools/san_render.py
    compute-sanitizer --tool racecheck python tools/san_render.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import nerfsos_b200  # noqa
from nerfsos_b200.models.nerf_net import NeRFNet
dev = "cuda:0"
torch.manual_seed(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
net = NeRFNet(N_samples=64, N_importance=128, use_semantics=True, sem_with_coord=True, sem_dim=2, mode=mode).to(dev)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 71
g = torch.Generator().manual_seed(1)
o = torch.rand(n, 3, generator=g) * 0.6 - 0.3
d = torch.cat([torch.rand(n, 2, generator=g) - 0.5, -torch.ones(n, 1)], -1)
rays = torch.stack([o, d], 0).to(dev)
net.eval()
with torch.no_grad():
    out = net(rays, (1.2, 12.0), retz=True)
torch.cuda.synchronize()
print("eval ok", float(out["rgb"].mean()), float(out["acc"].mean()))
net.train()
for nm, p in net.named_parameters():
    p.requires_grad_("semantic_linear" in nm)
out = net(rays, (1.2, 12.0))
(out["semantics"].sum() + out["semantics0"].sum()).backward()
torch.cuda.synchronize()
print("train ok", float(out["rgb"].mean()))
