#!/usr/bin/env python
"""GPU bring-up diagnostics (run on the B200 box): exercises each layer of the stack separately and
prints numbers instead of asserting, so one gpurun call yields maximal information."""
import os, sys, time, traceback
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nerfsos_b200
from nerfsos_b200 import _lib
from nerfsos_b200.models.nerf_net import NeRFNet
from conftest import load_golden

dev = torch.device("cuda:0")
L = _lib.lib()
print("device:", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)


SEL = sys.argv[1] if len(sys.argv) > 1 else None
STEPS = []


def step(name):
    def deco(fn):
        STEPS.append(name)
        if SEL is None or SEL == "--all" or SEL not in name:
            return
        print(f"\n=== {name} ===", flush=True)
        t0 = time.time()
        try:
            fn()
        except Exception:
            traceback.print_exc()
        torch.cuda.synchronize()
        print(f"--- {name}: {time.time() - t0:.2f}s", flush=True)
    return deco


def cmp(name, a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.abs(a - b)
    print(f"  {name:12s} max|d|={d.max():.3e} med={np.median(d):.3e} ref_max={np.abs(b).max():.3e} nan={np.isnan(a).sum()}", flush=True)


for _tm in (1, 0):
  @step(f"selftest umma a_in_tmem={_tm}")
  def _(_tm=_tm):
    g = torch.Generator().manual_seed(0)
    for mode, mname in ((1, "exact"), (2, "fast")):
        for a_in_tmem, N, K in ((1, 256, 64), (1, 256, 256), (1, 128, 128), (1, 32, 64), (1, 192, 192), (0, 256, 64), (0, 256, 63), (0, 128, 64), (0, 32, 20)):
            if a_in_tmem != _tm:
                continue
            a = (torch.rand(128, K, generator=g) * 4).to(dev)
            w = (torch.randn(N, K, generator=g) * 0.3).to(dev)
            d = torch.full((128, N), float("nan"), device=dev)
            scratch = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
            rc = L.nsos_selftest_umma(_lib.ptr(a), _lib.ptr(w), _lib.ptr(d), N, K, a_in_tmem, mode, _lib.ptr(scratch), scratch.numel(), None)
            torch.cuda.synchronize()
            ref = a.double() @ w.double().T
            err = (d.double() - ref).abs()
            print(f"  mode={mname} a_in_tmem={a_in_tmem} N={N} K={K} rc={rc} max_err={err.max().item():.3e} "
                  f"rel={err.max().item() / ref.abs().max().item():.3e} nan={torch.isnan(d).sum().item()}", flush=True)
            if err.max().item() > 1e-2 * ref.abs().max().item() and not torch.isnan(d).any():
                # locate the error pattern
                bad = (err > 1e-2 * ref.abs().max()).nonzero()
                print("    first bad (row,col):", bad[:8].tolist(), " rows bad:", len(set(bad[:, 0].tolist())), " cols bad:", len(set(bad[:, 1].tolist())))


def make_net(kind, mode):
    if kind == "cfg1":
        g = load_golden("cfg1_d4w64_eval")
        net = NeRFNet(netdepth=4, netwidth=64, netdepth_fine=4, netwidth_fine=64, N_samples=64, N_importance=0,
                      use_semantics=True, sem_with_coord=True, mode=mode)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in g["sd"].items()})
        return net.to(dev).eval(), g
    g = load_golden("flower_eval_256")
    sd = load_golden("flower_weights")["sd"]
    net = NeRFNet(N_samples=64, N_importance=128, use_semantics=True, sem_with_coord=True, sem_dim=2, mode=mode)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return net.to(dev).eval(), g


for kind in ("cfg1", "flower"):
    for mode in ("simt", "exact", "fast"):
        @step(f"forward {kind} mode={mode}")
        def _():
            net, g = make_net(kind, mode)
            rays = torch.from_numpy(g["rays"]).to(dev)
            with torch.no_grad():
                ret = net(rays, (1.2, 12.0), retz=True)
            torch.cuda.synchronize()
            for k in ("rgb", "acc", "depth", "semantics", "weights", "raw", "rgb0", "semantics0", "weights0", "raw0", "z_std"):
                if k in ret and k in g["out"]:
                    cmp(k, ret[k].cpu().numpy(), g["out"][k])
            if "inds" in ret:
                fl = ret["inds"].cpu().numpy() != g["stage"]["inds"]
                print(f"  index flip rate {fl.mean():.2e} (rays {fl.any(-1).mean():.2e})")


@step("throughput 4096 rays flower (exact/fast/simt)")
def _():
    print("  NSOS_CLUSTER =", os.environ.get("NSOS_CLUSTER", "default(2)"))
    for mode in ("exact", "fast") + (() if os.environ.get("NSOS_CLUSTER") else ("simt",)):
        net, g = make_net("flower", mode)
        rays = torch.from_numpy(np.tile(g["rays"], (1, 16, 1))).to(dev)
        with torch.no_grad():
            for _ in range(2):
                net(rays, (1.2, 12.0))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 5
            for _ in range(n):
                net(rays, (1.2, 12.0))
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"  mode={mode}: {ms:.3f} ms / 4096 rays -> {4096 / ms * 1e3:.0f} rays/s, {4096 * 324.86e6 / ms / 1e9:.1f} TFLOP/s algorithmic", flush=True)


if SEL == "--all":
    import subprocess
    for name in STEPS:
        r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=600)
        print(r.stdout[-6000:]); print(r.stderr[-1500:] if r.returncode else "", flush=True)


@step("trainstep timing 32768 rays (--fix_backbone recipe, B=8 patches 64x64)")
def _():
    """fwd / bwd / full train_one_step time at the shipped batch (8 patches x 64x64 rays)."""
    import dist_gpu_worker as W
    a = W.Args()
    a.use_correlation = True
    net = W.make_net(dev)
    opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=2e-3)
    from nerfsos_b200.engines.lr import LRScheduler
    sched = LRScheduler(opt, 2e-3, 0.1, 250000)
    losses = [None, None, W.CorrelationLoss(a), W.GeoCorrelationLoss(a)]
    B, Ps = 8, 64
    g = load_golden("flower_eval_256")
    base = torch.from_numpy(g["rays"])
    rays = base.repeat(1, B * Ps * Ps // base.shape[1], 1).permute(1, 0, 2).reshape(B, Ps * Ps, 2, 3).contiguous()
    rays[..., 1, :] += 0.02 * torch.randn(B, Ps * Ps, 3, generator=torch.Generator().manual_seed(1))
    gt = torch.rand(B, Ps * Ps, 3, generator=torch.Generator().manual_seed(0))
    a.patch_size = Ps
    for variant in ("tc", "simt"):
        if variant == "simt":
            os.environ["NSOS_BWD_SIMT"] = "1"
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ts = []
        for it in range(3):
            torch.cuda.synchronize()
            ev[0].record()
            out = W.train_one_step((rays, gt), [net, W.FakeDino()], opt, sched, W.Loader(), it + 1, losses, dev, a)
            ev[1].record()
            torch.cuda.synchronize()
            ts.append(ev[0].elapsed_time(ev[1]))
        print(f"  backward={variant}: train_one_step ms = {[round(t, 1) for t in ts]}  loss={out['loss'].item():.4f}  "
              f"=> {B * Ps * Ps / (min(ts) * 1e-3):.0f} rays/s fwd+bwd+losses+Adam", flush=True)
        # forward / backward split of the render alone
        r = rays.reshape(-1, 2, 3).permute(1, 0, 2).contiguous().to(dev)
        net.train()
        for it in range(2):
            torch.cuda.synchronize(); ev[0].record()
            o = net(r, (1.2, 12.0))
            ev[1].record(); torch.cuda.synchronize()
            tf = ev[0].elapsed_time(ev[1])
            l = o["semantics"].sum() + o["semantics0"].sum()
            torch.cuda.synchronize(); ev[0].record()
            l.backward()
            ev[1].record(); torch.cuda.synchronize()
            tb = ev[0].elapsed_time(ev[1])
        print(f"  backward={variant}: render fwd {tf:.1f} ms, render bwd {tb:.1f} ms", flush=True)
    os.environ.pop("NSOS_BWD_SIMT", None)


@step("corrloss timing (kernel B: geometry + appearance correlation losses, B=8 patches 64x64)")
def _():
    """ms per forward+backward of the two loss modules at the shipped batch; pairs/s for the all-pairs geometry loss."""
    from nerfsos_b200.utils import image as I

    class A:
        rand_neg = False; self_corr_w = 1; use_sim_matrix = True; patch_stride = 6
        app_corr_params = [0.18, 1, 0.46, 1]; geo_corr_params = [0.5, 1, 3, 1]
    g = torch.Generator().manual_seed(0)
    B, Pp = 8, 64
    code = torch.randn(B, 2, Pp, Pp, generator=g).to(dev).requires_grad_(True)
    ray_o = (torch.rand(B, 3, 1, 1, generator=g) * 0.6 - 0.3).expand(B, 3, Pp, Pp).contiguous().to(dev)
    ray_d = torch.randn(B, 3, Pp, Pp, generator=g).to(dev)
    depth = (torch.rand(B, 1, Pp, Pp, generator=g) * 10 + 1.2).to(dev)
    sim = I.get_similarity_matrix(torch.randn(B, 384, generator=g).to(dev))
    feat = torch.randn(B, 384, 14, 14, generator=g).to(dev)
    geo, app = I.GeoCorrelationLoss(A()), I.CorrelationLoss(A())
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for name, fn in (("geometry", lambda: geo(depth.clone(), code, [ray_o, ray_d, None], sim)),
                     ("appearance", lambda: app(feat.permute(0, 2, 3, 1).reshape(B, 196, 384) if False else feat, code, sim))):
        ts = []
        for it in range(5):
            code.grad = None
            torch.cuda.synchronize(); ev[0].record()
            l = fn()
            l.backward()
            ev[1].record(); torch.cuda.synchronize()
            ts.append(ev[0].elapsed_time(ev[1]))
        extra = ""
        if name == "geometry":
            pairs = 2 * B * (Pp * Pp) ** 2          # negative + self term, all pixel pairs
            extra = f"  {pairs / (min(ts) * 1e-3) / 1e9:.1f} G pairs/s (fwd+bwd)"
        print(f"  {name}: fwd+bwd ms = {[round(t, 3) for t in ts]}{extra}", flush=True)
