"""Diagnostics behind the parity tolerances (run on the GPU box): which entries of which rays exceed 1e-4, and the
per-parameter error of the all-parameter backward.  python tools/diag_parity.py [exact|simt]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
import test_gpu_render as T

mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
DEV = "cuda:0"
for tag in ("flower", "fortress", "co3d_apple"):
    net, g = T.fixture_net(tag, mode)
    net.eval()
    with torch.no_grad():
        out = net(torch.from_numpy(g["rays"]).to(DEV), (float(g["near"]), float(g["far"])), retz=True)
    ref, st = g["out"], g["stage"]
    inds = out["inds"].cpu().numpy()
    flip = inds != st["inds"]
    cdf = st["cdf"]
    below = np.maximum(st["inds"] - 1, 0); above = np.minimum(st["inds"], cdf.shape[1] - 1)
    den = np.take_along_axis(cdf, above, 1) - np.take_along_axis(cdf, below, 1)
    thr = (np.abs(den - 1e-5) < 2e-6).any(-1)
    ok = ~flip.any(-1)
    dzs = np.abs(out["z_samples"].cpu().numpy() - st["z_samples"])
    print(f"[{mode} {tag}] flips interior {flip[:, :-1].mean():.2e} last {flip[:, -1].mean():.3f}; den-threshold rays {thr.sum()}; "
          f"max |dz_samples| on ok rays {dzs[ok].max():.2e}")
    for k in ("rgb", "acc", "semantics", "weights", "depth"):
        d = np.abs(out[k].cpu().numpy() - ref[k])
        bad = (d > 1e-4 + 1e-4 * np.abs(ref[k]))
        rays_bad = np.where(bad.reshape(bad.shape[0], -1).any(-1))[0]
        print(f"   {k}: max diff ok-rays {d[ok].max():.2e}, all {d.max():.2e}; rays over tol {list(rays_bad[:12])} "
              f"(flipped: {[bool(not ok[r]) for r in rays_bad[:12]]}, den-thr: {[bool(thr[r]) for r in rays_bad[:12]]})")
    r0 = np.abs(np.maximum(out["raw0"][..., 3].cpu().numpy(), 0) - np.maximum(ref["raw0"][..., 3], 0))
    print(f"   coarse relu(sigma): max abs {r0.max():.2e}, max rel {(r0 / np.maximum(np.abs(ref['raw0'][..., 3]), 1)).max():.2e}; sigma max {ref['raw0'][..., 3].max():.1f}")

# all-parameter gradients
g, gs = load_golden("flower_train_64_semgrads"), load_golden("flower_train_64_allgrads")
net = T.flower_net(mode, perturb=1.0, raw_noise_std=1.0).train()
rnd = {k: torch.from_numpy(v).to(DEV) for k, v in g["rnd"].items()}
out = net(torch.from_numpy(g["rays"]).to(DEV), (1.2, 12.0), randoms=rnd)
loss = sum((out[k] * torch.from_numpy(v).to(DEV)).sum() for k, v in gs["gout"].items())
print("loss", loss.item(), float(gs["loss"]))
loss.backward()
for n, p in net.named_parameters():
    ref = gs["grads"][n]
    err = np.abs(p.grad.cpu().numpy() - ref)
    print(f"   {n:45s} max err / max ref = {err.max() / max(np.abs(ref).max(), 1e-12):.2e}   median rel {np.median(err) / max(np.abs(ref).max(), 1e-12):.1e}")
