#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Run in the build container only (the reference does not travel to the GPU box):
    python oracle/make_golden.py
The reference modules are imported read-only through `oracle/ref_shim.py` (stub modules for the
missing non-arithmetic imports).  Nothing from the reference is copied into this repo; only its
numerical outputs on seeded synthetic inputs are stored.
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402  (puts /root/reference on sys.path with stubs)

torch.autograd.set_detect_anomaly(False)
from models.nerf_net import NeRFNet  # noqa: E402
from utils.ray import get_persp_rays, get_persp_intrinsic  # noqa: E402
from utils import image as ref_image  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)
CKPT = ("/root/reference/logs/flower_B8_P64_PS6_1corr0.18,1,0.46,1_0.01geoCorr0.5,1,3,1_noTrainSFM/"
        "checkpoints/latest.ckpt")


def llff_rays(n, seed=0, patch=None):
    """get_persp_rays(756,1008,K(focal 815), [I|t]), t~U(-0.3,0.3)^3, seeded."""
    g = torch.Generator().manual_seed(seed)
    K = get_persp_intrinsic(756, 1008, 815.0)
    t = torch.rand(3, generator=g) * 0.6 - 0.3
    c2w = torch.cat([torch.eye(3), t[:, None]], 1)
    rays = get_persp_rays(756, 1008, K, c2w)            # [2,H,W,3]
    if patch is None:
        idx = torch.randperm(756 * 1008, generator=g)[:n]
        return rays.reshape(2, -1, 3)[:, idx].contiguous()
    P, stride = patch
    y0 = int(torch.randint(0, 756 - P * stride, (1,), generator=g))
    x0 = int(torch.randint(0, 1008 - P * stride, (1,), generator=g))
    return rays[:, y0:y0 + P * stride:stride, x0:x0 + P * stride:stride].contiguous()  # [2,P,P,3]


class RecordRandom:
    """Record every torch.rand / torch.randn draw made inside the reference (call order preserved)."""

    def __enter__(self):
        self.draws = []
        self._rand, self._randn = torch.rand, torch.randn

        def rand(*a, **k):
            r = self._rand(*a, **k); self.draws.append(("rand", r.clone())); return r

        def randn(*a, **k):
            r = self._randn(*a, **k); self.draws.append(("randn", r.clone())); return r

        torch.rand, torch.randn = rand, randn
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randn = self._rand, self._randn


def npsd(model):
    return {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}


def save(name, **arrs):
    flat = {}
    for k, v in arrs.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                flat[f"{k}/{kk}"] = np.asarray(vv)
        else:
            flat[k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **flat)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB, {len(flat)} arrays")


def outs(ret):
    return {k: v.detach().numpy() for k, v in ret.items()}


# ---------------------------------------------------------------------------------------------
def cfg1():
    """BASELINE config[0]: 512 rays, 64 coarse only, D=4 W=64, seeded default init, eval mode."""
    torch.manual_seed(0)
    net = NeRFNet(netdepth=4, netwidth=64, netdepth_fine=4, netwidth_fine=64, N_samples=64, N_importance=0,
                  use_semantics=True, sem_with_coord=True)
    net.eval()
    rays = llff_rays(512, seed=1)
    with torch.no_grad():
        ret = net(rays, (1.2, 12.0))
    save("cfg1_d4w64_eval", sd=npsd(net), rays=rays.numpy(), near=1.2, far=12.0, out=outs(ret))

    # same tiny net, hierarchical (64+32), train mode with recorded randoms + full-parameter gradients
    torch.manual_seed(1)
    net = NeRFNet(netdepth=4, netwidth=64, netdepth_fine=4, netwidth_fine=64, N_samples=64, N_importance=32,
                  use_semantics=True, sem_with_coord=True, perturb=1.0, raw_noise_std=1.0)
    net.train()
    rays = llff_rays(96, seed=2)
    with RecordRandom() as rec:
        ret = net(rays, (1.2, 12.0))
    kinds = [k for k, _ in rec.draws]
    assert kinds == ["rand", "randn", "rand", "randn"], kinds
    rnd = dict(t_rand=rec.draws[0][1], noise0=rec.draws[1][1], u=rec.draws[2][1], noise1=rec.draws[3][1])
    g = torch.Generator().manual_seed(3)
    tgt = {k: torch.randn(ret[k].shape, generator=g) for k in ("rgb", "rgb0", "semantics", "semantics0", "acc", "depth0")}
    loss = sum((ret[k] * tgt[k]).sum() for k in tgt)
    loss.backward()
    grads = {k: p.grad.numpy().copy() for k, p in net.named_parameters()}
    save("cfg1_d4w64_train_grads", sd=npsd(net), rays=rays.numpy(), near=1.2, far=12.0, out=outs(ret),
         rnd={k: v.numpy() for k, v in rnd.items()}, gout={k: v.numpy() for k, v in tgt.items()}, grads=grads,
         loss=float(loss))


def flower():
    """BASELINE config[1] weights: shipped stage-2 flower checkpoint, D=8 W=256 + seg head, 64+128."""
    ck = torch.load(CKPT, map_location="cpu")
    kw = dict(N_samples=64, N_importance=128, use_semantics=True, sem_with_coord=True, sem_dim=2, sem_layer=2)
    net = NeRFNet(perturb=1.0, raw_noise_std=1.0, **kw)
    net.load_state_dict(ck["model"], strict=True)

    # weights fixture (fp32, exact) -- shared by every 'flower' test, GPU parity tests and bench.py
    save("flower_weights", sd=npsd(net), global_step=ck["global_step"])

    # eval mode, 256 random LLFF rays + one 16x16 stride-6 patch; keep all intermediates
    net.eval()
    rays = llff_rays(256, seed=0)
    with torch.no_grad():
        ret = net(rays, (1.2, 12.0))
        # stage-wise intermediates for the 'exact indices given identical cdf/u' contract
        z = net.point_sampler(rays[0], rays[1], torch.tensor([[1.2, 12.0]]).expand(256, 2), zvals_only=True, perturb=0.0)
        smp = net.importance_sampler
        mid = .5 * (z[..., 1:] + z[..., :-1])
        w = ret["weights0"][..., 1:-1] + 1e-5
        pdf = w / torch.sum(w, -1, keepdim=True)
        cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
        u = torch.linspace(0., 1., steps=128).expand(256, 128).contiguous()
        inds = torch.searchsorted(cdf, u, right=True)
        z_samples = smp.sample_pdf(mid, ret["weights0"][..., 1:-1], det=True)
        z_fine, _ = torch.sort(torch.cat([z, z_samples], -1), -1)
    save("flower_eval_256", rays=rays.numpy(), near=1.2, far=12.0, out=outs(ret),
         stage=dict(z=z.numpy(), mid=mid.numpy(), cdf=cdf.numpy(), u=u.numpy(), inds=inds.numpy(),
                    z_samples=z_samples.numpy(), z_fine=z_fine.numpy()))

    # train mode (perturb=1, raw_noise_std=1, configs/flower_full.txt:14), 64 rays, recorded randoms,
    # seg-head-only gradients as under --fix_backbone (run_nerf.py:307-318)
    net.train()
    for n, p in net.named_parameters():
        p.requires_grad_("semantic_linear" in n)
    rays = llff_rays(64, seed=5)
    with RecordRandom() as rec:
        ret = net(rays, (1.2, 12.0))
    rnd = dict(t_rand=rec.draws[0][1], noise0=rec.draws[1][1], u=rec.draws[2][1], noise1=rec.draws[3][1])
    g = torch.Generator().manual_seed(7)
    tgt = {k: torch.randn(ret[k].shape, generator=g) for k in ("rgb", "rgb0", "semantics", "semantics0")}
    loss = sum((ret[k] * tgt[k]).sum() for k in tgt)
    loss.backward()
    grads = {k: p.grad.numpy().copy() for k, p in net.named_parameters() if p.grad is not None}
    keep = {k: v for k, v in outs(ret).items() if not k.startswith("raw")}
    save("flower_train_64_semgrads", rays=rays.numpy(), near=1.2, far=12.0, out=keep,
         rnd={k: v.numpy() for k, v in rnd.items()}, gout={k: v.numpy() for k, v in tgt.items()}, grads=grads,
         loss=float(loss))



class ReplayRandom:
    """Replay recorded torch.rand / torch.randn draws (in call order) inside the reference."""

    def __init__(self, draws):
        self.draws, self.i = list(draws), 0

    def __enter__(self):
        self._rand, self._randn = torch.rand, torch.randn

        def nxt(*a, **k):
            r = self.draws[self.i]; self.i += 1
            return r.clone()

        torch.rand, torch.randn = nxt, nxt
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randn = self._rand, self._randn


def cam_rays(n, seed, t):
    """n pixels of a 1008x756 / focal 815 view with c2w = [I | t] (same generator as llff_rays, caller-chosen origin)."""
    g = torch.Generator().manual_seed(seed)
    K = get_persp_intrinsic(756, 1008, 815.0)
    c2w = torch.cat([torch.eye(3), torch.tensor(t, dtype=torch.float32)[:, None]], 1)
    rays = get_persp_rays(756, 1008, K, c2w)
    idx = torch.randperm(756 * 1008, generator=g)[:n]
    return rays.reshape(2, -1, 3)[:, idx].contiguous()


def stage1_checkpoint(tag, fname, rays, near, far):
    """The other shipped stage-1 checkpoints (pretrained_ckpt/*.ckpt): strict=False load (the 8 semantic_linear tensors keep
    their seeded default init, exactly what run_nerf.py does at the start of stage 2), eval maps + stage intermediates on
    256 rays, and the largest hidden activation per trunk layer (range evidence for the fp16 split of the tcgen05 path)."""
    ck = torch.load("/root/reference/pretrained_ckpt/" + fname, map_location="cpu")
    torch.manual_seed(0)
    net = NeRFNet(N_samples=64, N_importance=128, use_semantics=True, sem_with_coord=True, sem_dim=2, sem_layer=2)
    missing = net.load_state_dict(ck["model"], strict=False)
    assert len(missing.missing_keys) == 8 and all("semantic_linear" in k for k in missing.missing_keys), missing
    net.eval()
    amax = {}

    def hook(name):
        def f(mod, inp, out):
            amax[name] = max(amax.get(name, 0.0), float(torch.relu(out).abs().max()))
        return f

    hs = []
    for pre, m in (("nerf", net.nerf.mlp), ("nerf_fine", net.nerf_fine.mlp)):
        for i, lin in enumerate(m.pts_linears):
            hs.append(lin.register_forward_hook(hook(f"{pre}.pts_linears.{i}")))
        hs.append(m.views_linears[0].register_forward_hook(hook(f"{pre}.views_linears.0")))
        hs.append(m.semantic_linear[0].register_forward_hook(hook(f"{pre}.semantic_linear.0")))
    n = rays.shape[1]
    with torch.no_grad():
        ret = net(rays, (near, far))
        z = net.point_sampler(rays[0], rays[1], torch.tensor([[near, far]]).expand(n, 2), zvals_only=True, perturb=0.0)
        mid = .5 * (z[..., 1:] + z[..., :-1])
        w = ret["weights0"][..., 1:-1] + 1e-5
        pdf = w / torch.sum(w, -1, keepdim=True)
        cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
        u = torch.linspace(0., 1., steps=128).expand(n, 128).contiguous()
        inds = torch.searchsorted(cdf, u, right=True)
        z_samples = net.importance_sampler.sample_pdf(mid, ret["weights0"][..., 1:-1], det=True)
    for h in hs:
        h.remove()
    print(tag, "max hidden activation per layer:", {k: round(v, 2) for k, v in amax.items()})
    keep = {k: v for k, v in outs(ret).items() if k != "raw"}          # raw0 kept (coarse per-sample check), fine raw dropped (size)
    save(tag + "_eval_256", sd=npsd(net), rays=rays.numpy(), near=near, far=far, out=keep,
         stage=dict(z=z.numpy(), mid=mid.numpy(), cdf=cdf.numpy(), u=u.numpy(), inds=inds.numpy(), z_samples=z_samples.numpy()),
         amax={k: np.float32(v) for k, v in amax.items()}, global_step=ck["global_step"])


def other_checkpoints():
    # fortress: LLFF forward-facing, far = 14.72 (the value quoted in models/sampler.py:45)
    stage1_checkpoint("fortress", "fortress_00150000.ckpt", llff_rays(256, seed=31), 1.2, 14.72)
    # co3d apple: object-centric; camera 4 units in front of the object, some rays miss it (acc < 1, depth -> 1e10 path)
    stage1_checkpoint("co3d_apple", "co3d_apple_110_00140000.ckpt", cam_rays(256, 32, [0.1, -0.05, 4.0]), 1.0, 8.0)


def safe_ray_mask(ret, u, margin=2e-5):
    """Rays whose importance samples sit safely inside their cdf bins: every u at least `margin` away from both neighbouring
    cdf knots, and no selected bin with |denom - 1e-5| < margin/10 (sampler.py:117-132).  On these rays a 1-ulp difference in
    the coarse weights cannot flip a searchsorted index or the denom<1e-5 switch, so fine-pass outputs and gradients are
    comparable at fp32 tolerance; the cotangents of the other rays are zeroed in the gradient fixtures."""
    w = ret["weights0"].detach()[..., 1:-1] + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    below = torch.max(torch.zeros_like(inds - 1), inds - 1)
    above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(inds), inds)
    cb, ca = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    d_lo = (u - cb).abs()
    d_hi = torch.where(inds < cdf.shape[-1], (ca - u).abs(), torch.ones_like(u))
    den = ca - cb
    ok = (d_lo > margin) & (d_hi > margin) & ((den - 1e-5).abs() > margin / 10)
    return ok.all(-1)


def _grads_safe(name, net, rays, draws, keys, seed):
    net.train()
    cap = {}
    orig = net.importance_sampler.sample_pdf

    def sample_pdf(*a, **k):                     # capture the reference's own importance samples (sampler.py:132)
        cap["z_samples"] = orig(*a, **k)
        return cap["z_samples"]

    net.importance_sampler.sample_pdf = sample_pdf
    with ReplayRandom(draws) as rp:
        ret = net(rays, (1.2, 12.0))
    net.importance_sampler.sample_pdf = orig
    assert rp.i == 4
    safe = safe_ray_mask(ret, draws[2])
    g = torch.Generator().manual_seed(seed)
    tgt = {k: torch.randn(ret[k].shape, generator=g) * safe.reshape(-1, *([1] * (ret[k].dim() - 1))) for k in keys}
    loss = sum((ret[k] * tgt[k]).sum() for k in tgt)
    loss.backward()
    grads = {k: p.grad.numpy().copy() for k, p in net.named_parameters()}
    print(name, f"safe rays: {int(safe.sum())} of {safe.numel()}")
    save(name, safe=safe.numpy(), gout={k: v.numpy() for k, v in tgt.items()}, grads=grads, loss=float(loss.detach()),
         z_samples=cap["z_samples"].detach().numpy())


def flower_all_param_grads():
    """All-parameter autograd gradients (stage-1 training, engines/trainer.py:201 without --fix_backbone) of the shipped
    flower net on the rays / random draws of flower_train_64_semgrads.npz (replayed), cotangents restricted to the rays
    whose importance samples cannot flip (safe_ray_mask).  The semantic_linear entries double as the --fix_backbone
    gradients (they do not depend on which other parameters require grad)."""
    g = np.load(os.path.join(OUT, "flower_train_64_semgrads.npz"))
    ck = torch.load(CKPT, map_location="cpu")
    net = NeRFNet(perturb=1.0, raw_noise_std=1.0, N_samples=64, N_importance=128, use_semantics=True, sem_with_coord=True,
                  sem_dim=2, sem_layer=2)
    net.load_state_dict(ck["model"], strict=True)
    draws = [torch.from_numpy(g["rnd/" + k]) for k in ("t_rand", "noise0", "u", "noise1")]
    _grads_safe("flower_train_64_allgrads", net, torch.from_numpy(g["rays"]), draws, ("rgb", "rgb0", "semantics", "semantics0", "acc", "depth0"), 17)


def cfg1_grads_safe():
    """Same for the tiny D=4 W=64 net of cfg1_d4w64_train_grads.npz."""
    g = np.load(os.path.join(OUT, "cfg1_d4w64_train_grads.npz"))
    net = NeRFNet(netdepth=4, netwidth=64, netdepth_fine=4, netwidth_fine=64, N_samples=64, N_importance=32,
                  use_semantics=True, sem_with_coord=True, perturb=1.0, raw_noise_std=1.0)
    net.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}, strict=True)
    draws = [torch.from_numpy(g["rnd/" + k]) for k in ("t_rand", "noise0", "u", "noise1")]
    _grads_safe("cfg1_d4w64_train_grads_safe", net, torch.from_numpy(g["rays"]), draws, ("rgb", "rgb0", "semantics", "semantics0", "acc", "depth0"), 19)


class _Args:
    rand_neg = False
    self_corr_w = 1
    use_sim_matrix = True
    patch_stride = 6
    app_corr_params = ["0.18", "1", "0.46", "1"]
    geo_corr_params = ["0.5", "1", "3", "1"]


def losses():
    """CorrelationLoss / GeoCorrelationLoss (utils/image.py:263-482), small shapes, with code gradients."""
    g = torch.Generator().manual_seed(11)
    B, P = 4, 16
    feat = torch.randn(B, 384, 14, 14, generator=g)
    cls_ = torch.randn(B, 384, generator=g)
    sim = ref_image.get_similarity_matrix(cls_)
    code = torch.randn(B, 2, P, P, generator=g, requires_grad=True)
    app = ref_image.CorrelationLoss(_Args())
    with RecordRandom() as rec:
        la = app(feat, code, sim)
    la.backward()
    coords1, coords2 = rec.draws[0][1], rec.draws[1][1]          # raw U[0,1) draws; the loss maps them to *2-1
    ga = code.grad.clone(); code.grad = None

    rays = torch.stack([llff_rays(0, seed=20 + b, patch=(P, 6)) for b in range(B)], 1)   # [2,B,P,P,3]
    ray_o = rays[0].permute(0, 3, 1, 2).contiguous(); ray_d = rays[1].permute(0, 3, 1, 2).contiguous()
    depth = (torch.rand(B, 1, P, P, generator=g) * 16.0 + 1.2)  # some values > max_depth=15 to hit the clip
    geo = ref_image.GeoCorrelationLoss(_Args())
    depth_in = depth.clone()
    lg = geo(depth_in, code, [ray_o, ray_d, None], sim)
    lg.backward()
    gg = code.grad.clone()
    save("losses_b4_p16", feat=feat.numpy(), cls=cls_.numpy(), sim=sim.numpy(), code=code.detach().numpy(),
         rand1=coords1.numpy(), rand2=coords2.numpy(), app_loss=float(la), app_gcode=ga.numpy(),
         ray_o=ray_o.numpy(), ray_d=ray_d.numpy(), depth=depth.numpy(), depth_clipped=depth_in.numpy(),
         geo_loss=float(lg), geo_gcode=gg.numpy(),
         app_params=np.array([0.18, 1, 0.46, 1.0]), geo_params=np.array([0.5, 1, 3, 1.0]))


class StandInDino:
    """Deterministic feature provider with the extractor's interface (no DINO weights offline); duplicated in tests/."""

    def __init__(self):
        self.proj = torch.randn(3, 384, generator=torch.Generator().manual_seed(0))

    def get_vit_attn_feat(self, x):
        B = x.shape[0]
        p = torch.nn.functional.adaptive_avg_pool2d(x, 14).reshape(B, 3, 196).permute(0, 2, 1)      # [B,196,3]
        feat = p @ self.proj.to(x.device)
        return {"attn": feat[..., :1].permute(0, 2, 1), "cls_": feat.mean(1), "feat": feat}


def train_steps():
    """SURVEY Appendix B 'end-to-end': five consecutive steps of the UNMODIFIED reference trainer (engines/trainer.py:32-213,
    engines/lr.py, torch.optim.Adam) in the shipped stage-2 recipe (--fix_backbone, both correlation losses) on the flower
    checkpoint, CPU.  Every random draw is recorded per step (4 render draws + 2x2 sample-coordinate draws of the appearance
    loss); the per-step loss terms and the trained semantic-head tensors are the targets."""
    import engines.trainer as RT
    from engines.lr import LRScheduler as RefLR
    ck = torch.load(CKPT, map_location="cpu")
    net = NeRFNet(perturb=1.0, raw_noise_std=1.0, N_samples=64, N_importance=128, use_semantics=True, sem_with_coord=True,
                  sem_dim=2, sem_layer=2)
    net.load_state_dict(ck["model"], strict=True)
    for n, p in net.named_parameters():
        p.requires_grad_("semantic_linear" in n)
    opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=5e-4, betas=(0.9, 0.999))
    sched = RefLR(opt, 5e-4, 0.1, 250000)

    class A(_Args):
        patch_tune = True; patch_size = 8; batch_size = 2; use_dino = True; use_correlation = True; use_geoCorr = True
        use_contrast = False; rgb_w = 1.0; correlation_w = 1.0; Gcorrelation_w = 0.01; contrast_w = 0.0; i_print = 100000
        clus_no_sfm = False; N_cluster = 2

    class DS:
        height, width, K = 756, 1008, None
        def near_far(self): return 1.2, 12.0
        def radii(self): return None

    class Loader:
        dataset = DS()

    losses = [None, None, ref_image.CorrelationLoss(A()), ref_image.GeoCorrelationLoss(A())]
    B, P = A.batch_size, A.patch_size
    rays = torch.stack([llff_rays(0, seed=70 + b, patch=(P, 6)) for b in range(B)], 0)          # [B,2,P,P,3]
    batch_rays = rays.permute(0, 2, 3, 1, 4).reshape(B, P * P, 2, 3).contiguous()
    g = torch.Generator().manual_seed(71)
    gt = torch.rand(B, P * P, 3, generator=g)
    masks = torch.zeros(B, P * P, 1)
    steps, rec = [], {}
    dino = StandInDino()
    for k in range(5):
        with RecordRandom() as rr:
            out = RT.train_one_step((batch_rays, gt, masks), [net, dino], opt, sched, Loader(), 2 + k, losses, "cpu", A())
        kinds = [a for a, _ in rr.draws]
        assert kinds == ["rand", "randn", "rand", "randn", "rand", "rand", "rand", "rand"], kinds
        for name, (_, t) in zip(("t_rand", "noise0", "u", "noise1", "c0_1", "c0_2", "c1_1", "c1_2"), rr.draws):
            rec[f"s{k}_{name}"] = t.numpy()
        steps.append([float(out[n].detach()) for n in ("loss", "img0", "img1", "corr0", "corr1", "geo_corr0", "geo_corr1")])
        print("step", 2 + k, steps[-1])
    final = {n: p.detach().numpy().copy() for n, p in net.named_parameters() if p.requires_grad}
    save("flower_train_steps", rays=batch_rays.numpy(), gt=gt.numpy(), steps=np.array(steps), rnd=rec, final=final,
         lr_last=opt.param_groups[0]["lr"])


def metrics():
    """Evaluation metrics exactly as the reference computes them (engines/eval.py:60-95): utils/ssim.py:7-40, sklearn
    adjusted_rand_score, utils/misc.py:40-50 (sklearn KMeans, random_state=0) and utils/get_metrics.py:15-26 (compute_iou)."""
    from sklearn.metrics import adjusted_rand_score
    from utils import ssim as ref_ssim
    # utils/get_metrics.py runs a hard-coded evaluation script at import time: take its compute_iou function alone
    import ast
    from sklearn.metrics import confusion_matrix
    src = open(os.path.join(ref_shim.REF, "utils", "get_metrics.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "compute_iou")
    ns = {"np": np, "confusion_matrix": confusion_matrix}
    exec(compile(ast.Module([fn], []), "utils/get_metrics.py", "exec"), ns)
    compute_iou = ns["compute_iou"]
    from utils.misc import segmap_cluster
    g = torch.Generator().manual_seed(41)
    H, W = 48, 64
    img1 = torch.rand(2, 3, H, W, generator=g)
    img2 = (img1 + 0.1 * torch.randn(2, 3, H, W, generator=g)).clamp(0, 1)
    s_all = float(ref_ssim.ssim(img1, img2))
    s_each = ref_ssim.ssim(img1, img2, size_average=False).numpy()
    # a blobby ground-truth mask, logits that mostly agree with it
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    gt = (((yy - 20) ** 2 + (xx - 30) ** 2) < 15 ** 2).long()
    logits = torch.stack([1.5 - 3.0 * gt.float(), 3.0 * gt.float() - 1.5], -1) + 0.8 * torch.randn(H, W, 2, generator=g)
    prob = logits.softmax(-1)
    clus = segmap_cluster(prob.numpy(), n_clusters=2)                       # [H,W,1] int
    sem_pred = prob.argmax(-1, keepdim=True).numpy()
    gt_np = gt.numpy()[..., None]
    fg = gt_np == 1
    out = dict(clus_ari=adjusted_rand_score(gt_np.reshape(-1), clus.reshape(-1)), clus_ari_fg=adjusted_rand_score(gt_np[fg].reshape(-1), clus[fg].reshape(-1)),
               sem_ari=adjusted_rand_score(gt_np.reshape(-1), sem_pred.reshape(-1)), sem_ari_fg=adjusted_rand_score(gt_np[fg].reshape(-1), sem_pred[fg].reshape(-1)))
    iou_a, iou_b = compute_iou(clus, gt_np), compute_iou(1 - clus, gt_np)   # clustering has no polarity: both assignments
    save("metrics_ref", img1=img1.numpy(), img2=img2.numpy(), ssim=s_all, ssim_each=s_each, logits=logits.numpy(), gt=gt_np, clus=clus,
         sem_pred=sem_pred, iou_fg=max(float(iou_a[1]), float(iou_b[1])), **{k: float(v) for k, v in out.items()})


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg1", "flower", "losses", "other_checkpoints", "flower_all_param_grads", "cfg1_grads_safe", "metrics", "train_steps"]
    for w in which:
        globals()[w]()
