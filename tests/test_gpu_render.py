"""GPU parity tests for kernel A (run on the B200 box: pytest -m gpu).

Every call goes through the C ABI (include/nerfsos.h) via the ctypes drop-in classes.  The checker is
the numpy oracle / the reference-generated golden fixtures; tolerances follow BASELINE.json's north_star:
1e-4 on rgb / density-derived maps, exact inverse-CDF indices given identical cdf/u (stage-wise), and an
end-to-end index flip rate <= 1e-3 on the 127 interior samples.  The 128th deterministic sample is
u = 1.0 exactly, i.e. ON the end point of the cdf (cumsum(pdf)[-1] = 1 +- 1 ulp by construction): searchsorted(right=True)
returns 63 or 62 depending on the last ulp of the sum, in 10-26 % of the rays for the numpy oracle itself against the
reference (every other oracle index is identical) -- that column is reported separately and the affected rays are the
"flipped" ones.  Every NON-flipped ray is held to the 1e-4 contract, gradients to 1e-3 on rays that cannot flip.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _imports():
    import nerfsos_b200  # noqa: F401
    from nerfsos_b200 import _lib
    from nerfsos_b200.models.nerf_net import NeRFNet
    return _lib, NeRFNet


def flower_net(mode, **kw):
    _lib, NeRFNet = _imports()
    sd = load_golden("flower_weights")["sd"]
    args = dict(N_samples=64, N_importance=128, use_semantics=True, sem_with_coord=True, sem_dim=2, sem_layer=2, mode=mode)
    args.update(kw)
    net = NeRFNet(**args)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return net.to(DEV)


def fixture_net(tag, mode, **kw):
    """flower: shipped stage-2 checkpoint; fortress / co3d_apple: shipped stage-1 checkpoints + seeded semantic heads."""
    if tag == "flower":
        return flower_net(mode, **kw), load_golden("flower_eval_256")
    _lib, NeRFNet = _imports()
    g = load_golden(tag + "_eval_256")
    args = dict(N_samples=64, N_importance=128, use_semantics=True, sem_with_coord=True, sem_dim=2, sem_layer=2, mode=mode)
    args.update(kw)
    net = NeRFNet(**args)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in g["sd"].items()}, strict=True)
    return net.to(DEV), g


def cfg1_net(mode, golden, n_importance=0, **kw):
    _lib, NeRFNet = _imports()
    net = NeRFNet(netdepth=4, netwidth=64, netdepth_fine=4, netwidth_fine=64, N_samples=64, N_importance=n_importance,
                  use_semantics=True, sem_with_coord=True, mode=mode, **kw)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in golden["sd"].items()}, strict=True)
    return net.to(DEV)


def close(a, b, rtol=1e-4, atol=1e-4, frac=1.0):
    a = a.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    ok = np.abs(a - b) <= atol + rtol * np.abs(b)
    assert ok.mean() >= frac, f"only {ok.mean():.4f} within tol; max abs diff {np.abs(a - b).max():.3e}"


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode,tol", [(1, 2e-6), (2, 2e-3)])
@pytest.mark.parametrize("a_in_tmem,N,K", [(1, 256, 256), (1, 256, 64), (1, 128, 192), (1, 32, 64), (0, 256, 63), (0, 32, 20)])
def test_umma_building_blocks(mode, tol, a_in_tmem, N, K):
    """tcgen05 self test: pack -> bulk copy -> UMMA (TMEM-A / SMEM-A) -> TMEM read-out vs fp64 matmul."""
    _lib, _ = _imports()
    L = _lib.lib()
    g = torch.Generator().manual_seed(N * 1000 + K)
    a = (torch.rand(128, K, generator=g) * 4 - 1).to(DEV)
    w = (torch.randn(N, K, generator=g) * 0.3).to(DEV)
    d = torch.full((128, N), float("nan"), device=DEV)
    scratch = torch.zeros(1 << 20, dtype=torch.uint8, device=DEV)
    _lib.check(L.nsos_selftest_umma(_lib.ptr(a), _lib.ptr(w), _lib.ptr(d), N, K, a_in_tmem, mode, _lib.ptr(scratch),
                                    scratch.numel(), None), "selftest")
    torch.cuda.synchronize()
    ref = a.double() @ w.double().T
    err = (d.double() - ref).abs().max().item()
    assert err <= tol * ref.abs().max().item(), err


def test_invert_cdf_exact_indices():
    """Stage-wise contract: given the reference's own cdf/u the indices are bit-exact (sampler.py:117-132)."""
    _lib, _ = _imports()
    L = _lib.lib()
    st = load_golden("flower_eval_256")["stage"]
    bins, cdf, u = (torch.from_numpy(st[k]).to(DEV).contiguous() for k in ("mid", "cdf", "u"))
    n, M = cdf.shape
    K = u.shape[1]
    samples = torch.empty(n, K, device=DEV)
    inds = torch.empty(n, K, dtype=torch.int64, device=DEV)
    _lib.check(L.nsos_invert_cdf(_lib.ptr(bins), _lib.ptr(cdf), _lib.ptr(u), _lib.ptr(samples), _lib.ptr(inds), n, M, K, None), "invert_cdf")
    assert np.array_equal(inds.cpu().numpy(), st["inds"])
    assert np.array_equal(samples.cpu().numpy(), st["z_samples"])      # same fp32 op sequence -> bit-exact samples too
    # random u (train mode) against the numpy oracle
    from oracle import nerf_oracle as O
    ur = torch.rand(n, K, generator=torch.Generator().manual_seed(3)).to(DEV)
    _lib.check(L.nsos_invert_cdf(_lib.ptr(bins), _lib.ptr(cdf), _lib.ptr(ur), _lib.ptr(samples), _lib.ptr(inds), n, M, K, None), "invert_cdf")
    s_ref, i_ref = O.invert_cdf(st["mid"], st["cdf"], ur.cpu().numpy())
    assert np.array_equal(inds.cpu().numpy(), i_ref)
    np.testing.assert_allclose(samples.cpu().numpy(), s_ref, rtol=0, atol=1e-6)


@pytest.mark.parametrize("mode", ["simt", "exact"])
def test_cfg1_eval(mode):
    """BASELINE config[0]: 512 rays, 64 coarse samples, D=4 W=64."""
    g = load_golden("cfg1_d4w64_eval")
    net = cfg1_net(mode, g).eval()
    with torch.no_grad():
        out = net(torch.from_numpy(g["rays"]).to(DEV), (1.2, 12.0))
    assert set(out) == set(g["out"])
    for k in ("rgb", "acc", "semantics", "weights", "disp"):
        close(out[k], g["out"][k], rtol=1e-4, atol=1e-4)
    close(out["depth"], g["out"]["depth"], rtol=1e-4, atol=1e-3)
    close(out["raw"], g["out"]["raw"], rtol=1e-4, atol=2e-4)


def _same_samples(z_ours, z_ref, tol=1e-5):
    """Rays whose importance samples equal the reference's to `tol`.  The inverse cdf is ill-conditioned in low-mass bins
    (d sample / d cdf = bin width / denom, up to 1e4: a 1e-7 difference in a coarse weight moves such a sample by 1e-3 with
    no index flip), so end-to-end fine-pass comparisons are made on these rays; the fine stage itself is pinned on ALL rays
    by injecting the reference's samples (NsosRandoms.z_samples)."""
    z_ours = z_ours.detach().cpu().numpy() if torch.is_tensor(z_ours) else z_ours
    return np.abs(z_ours - z_ref).max(-1) <= tol


@pytest.mark.parametrize("tag", ["flower", "fortress", "co3d_apple"])
@pytest.mark.parametrize("mode", ["simt", "exact"])
def test_checkpoint_eval(mode, tag):
    """BASELINE config[1] geometry (64+128 samples, D=8 W=256 + seg head) on every kind of shipped checkpoint: the stage-2
    flower net, and the stage-1 fortress (configs[2]) / CO3D apple (configs[4]) nets with seeded semantic heads.
    Stage-wise contract: (a) coarse pass vs the reference, (b) sampler indices, (c) fine pass on the reference's own sample
    positions -- each at 1e-4 on every ray -- then (d) the end-to-end maps."""
    net, g = fixture_net(tag, mode)
    net.eval()
    bounds = (float(g["near"]), float(g["far"]))
    rays = torch.from_numpy(g["rays"]).to(DEV)
    ref, st = g["out"], g["stage"]
    with torch.no_grad():
        out = net(rays, bounds, retz=True)
        fin = net(rays, bounds, retz=True, randoms={"z_samples": torch.from_numpy(st["z_samples"])})
    # (a) coarse pass
    assert np.array_equal(out["z_vals0"].cpu().numpy(), st["z"])                 # coarse sample positions: bit-exact
    for k in ("rgb0", "acc0", "semantics0", "weights0"):
        close(out[k], ref[k], rtol=1e-4, atol=1e-4)
    # per-sample density: compare relu(sigma) with abs+rel tolerance (raw sigma crosses 0)
    close(torch.relu(out["raw0"][..., 3]), np.maximum(ref["raw0"][..., 3], 0), rtol=1e-4, atol=1e-3)
    close(torch.sigmoid(out["raw0"][..., :3]), 1 / (1 + np.exp(-ref["raw0"][..., :3])), rtol=1e-4, atol=1e-4)
    # (b) sampler: interior indices (the u = 1.0 column sits on the cdf's end point, see the module docstring)
    flip = out["inds"].cpu().numpy() != st["inds"]
    assert flip[:, :-1].mean() <= 1e-3 and flip[:, -1].mean() <= 0.5, (flip[:, :-1].mean(), flip[:, -1].mean())
    # (c) fine pass on the reference's samples: every ray, every map, per-sample weights
    zf = np.sort(np.concatenate([st["z"], st["z_samples"]], -1), -1)
    assert np.array_equal(fin["z_vals"].cpu().numpy(), zf)
    for k in ("rgb", "acc", "semantics", "weights"):
        close(fin[k], ref[k], rtol=1e-4, atol=1e-4)
    hit = ref["acc"][:, 0] > 1e-3                                                # depth is 1e10 where acc <= 1e-10 (renderer.py:72)
    close(fin["depth"][torch.from_numpy(hit).to(DEV)], ref["depth"][hit], rtol=1e-4, atol=1e-3)
    close(fin["z_std"], ref["z_std"], rtol=1e-5, atol=1e-5)
    # (d) end to end with the kernel's own sampler
    same = _same_samples(out["z_samples"], st["z_samples"]) & ~flip.any(-1)
    assert same.mean() >= 0.2, same.mean()
    sd_ = torch.from_numpy(same).to(DEV)
    for k in ("rgb", "acc", "semantics"):                                        # same sample set: the full contract on the maps
        close(out[k][sd_], ref[k][same], rtol=1e-4, atol=1e-4)                   # (per-sample weights: pinned in (c); at a sharp
                                                                                 # surface d w / d z ~ 30, so 1e-5 in z is 3e-4 in w)
    # every other ray drew at least one sample elsewhere (moved low-mass sample, or the u = 1.0 knife edge): whatever it drew,
    # its fine maps must equal the oracle's fine pass on the SAME sample depths
    from oracle import nerf_oracle as O
    oth = np.where(~same)[0]
    if len(oth):
        own = O.fine_pass_on(g["sd"] if tag != "flower" else load_golden("flower_weights")["sd"], g["rays"][:, oth],
                             out["z_vals"][torch.from_numpy(oth).to(DEV)].cpu().numpy())
        for k in ("rgb", "acc", "semantics", "weights"):
            close(out[k][torch.from_numpy(oth).to(DEV)], own[k], rtol=1e-4, atol=1e-4)
    for k in ref:
        assert tuple(out[k].shape) == ref[k].shape, k
    mse = float(((out["rgb"].cpu().numpy() - ref["rgb"]) ** 2).mean())
    assert -10 * np.log10(max(mse, 1e-20)) > 80.0                                # PSNR(ours, reference)
    assert not net.range_overflow()                                             # fp16 hi/lo activation planes stayed in range


def test_activation_range_guard():
    """Exact mode keeps activations as fp16(16*a): |a| > 4094 cannot be represented.  The shipped nets peak at 59 (fortress
    fine layer 7, recorded in the fixtures); a net that does overflow must say so (sticky device flag + non-finite maps),
    never return plausible numbers (inf/NaN planes alone would not do: fmaxf in the next ReLU maps NaN to 0)."""
    for tag in ("fortress", "co3d_apple"):
        am = load_golden(tag + "_eval_256")["amax"]
        assert max(float(v) for v in am.values()) < 4094 / 16                    # >= 16x headroom on every shipped layer
    net, g = fixture_net("flower", "exact")
    net.eval()
    rays = torch.from_numpy(g["rays"][:, :32]).to(DEV)
    with torch.no_grad():
        net(rays, (1.2, 12.0))
        assert not net.range_overflow()
        net.nerf_fine.mlp.pts_linears[3].weight.mul_(3000.0)                     # hidden activations ~1e4
        out = net(rays, (1.2, 12.0))
    assert net.range_overflow()
    assert torch.isnan(out["rgb"]).any() and torch.isnan(out["semantics"]).any()   # poisoned, not silently zeroed by the ReLUs
    assert not net.range_overflow()                                              # the flag is cleared by reading it


def test_full_size_vs_numpy_oracle():
    """BASELINE configs[1] at its real size -- the bench's own 4096 rays -- against the numpy oracle (outside the repo's
    kernels): persistent-CTA multi-iteration path, 14 ray pairs per CTA, tile tails.  Same stage-wise protocol as above."""
    from oracle import nerf_oracle as O
    import bench
    rays_np = bench.llff_rays(4096, 100)
    rays = torch.from_numpy(rays_np).to(DEV)
    net = flower_net("exact").eval()
    ref = O.nerfnet_forward(load_golden("flower_weights")["sd"], rays_np, (bench.NEAR, bench.FAR), extras=True)
    with torch.no_grad():
        out = net(rays, (bench.NEAR, bench.FAR), retz=True)
        fin = net(rays, (bench.NEAR, bench.FAR), retz=True, randoms={"z_samples": torch.from_numpy(ref["z_samples"])})
    assert np.array_equal(out["z_vals0"].cpu().numpy(), ref["z_vals0"])
    for k in ("rgb0", "acc0", "semantics0", "weights0"):
        close(out[k], ref[k], rtol=1e-4, atol=1e-4)
    flip = out["inds"].cpu().numpy() != ref["inds"]
    assert flip[:, :-1].mean() <= 1e-3 and flip[:, -1].mean() <= 0.5, (flip[:, :-1].mean(), flip[:, -1].mean())
    assert np.array_equal(fin["z_vals"].cpu().numpy(), ref["z_vals"])
    for k in ("rgb", "acc", "semantics", "weights"):                             # fine pass on the oracle's samples: all 4096 rays
        close(fin[k], ref[k], rtol=1e-4, atol=1e-4)
    close(fin["depth"], ref["depth"], rtol=1e-4, atol=1e-3)
    same = _same_samples(out["z_samples"], ref["z_samples"]) & ~flip.any(-1)
    assert same.mean() >= 0.2, same.mean()
    sd_ = torch.from_numpy(same).to(DEV)
    for k in ("rgb", "acc", "semantics"):
        close(out[k][sd_], ref[k][same], rtol=1e-4, atol=1e-4)
    oth = np.where(~same)[0]                                                     # rays that drew other samples: oracle on THEIR samples
    own = O.fine_pass_on(load_golden("flower_weights")["sd"], rays_np[:, oth], out["z_vals"][torch.from_numpy(oth).to(DEV)].cpu().numpy())
    for k in ("rgb", "acc", "semantics", "weights"):
        close(out[k][torch.from_numpy(oth).to(DEV)], own[k], rtol=1e-4, atol=1e-4)
    mse = float(((out["rgb"].cpu().numpy() - ref["rgb"]) ** 2).mean())
    assert -10 * np.log10(max(mse, 1e-20)) > 60.0     # one u = 1.0 knife-edge ray in 4096 can differ by 5e-2 (67 dB); checked above


def test_flower_fast_mode_psnr():
    g = load_golden("flower_eval_256")
    net = flower_net("fast").eval()
    with torch.no_grad():
        out = net(torch.from_numpy(g["rays"]).to(DEV), (1.2, 12.0))
    mse = float(((out["rgb"].cpu().numpy() - g["out"]["rgb"]) ** 2).mean())
    assert -10 * np.log10(mse) > 35.0, -10 * np.log10(mse)


@pytest.mark.parametrize("mode", ["simt", "exact"])
def test_flower_train_injected_randoms(mode):
    g = load_golden("flower_train_64_semgrads")
    net = flower_net(mode, perturb=1.0, raw_noise_std=1.0).train()
    rnd = {k: torch.from_numpy(v).to(DEV) for k, v in g["rnd"].items()}
    out = net(torch.from_numpy(g["rays"]).to(DEV), (1.2, 12.0), randoms=rnd)
    for k in ("rgb0", "acc0", "semantics0", "weights0"):
        close(out[k], g["out"][k], rtol=1e-4, atol=1e-4)
    d = np.abs(out["rgb"].detach().cpu().numpy() - g["out"]["rgb"]).max(-1)
    assert np.median(d) < 5e-5 and d.max() < 2e-2, (np.median(d), d.max())
    # fine pass on the reference's own importance samples: every ray at the full contract
    rnd["z_samples"] = torch.from_numpy(load_golden("flower_train_64_allgrads")["z_samples"]).to(DEV)
    fin = net(torch.from_numpy(g["rays"]).to(DEV), (1.2, 12.0), randoms=rnd)
    for k in ("rgb", "acc", "semantics", "weights"):
        close(fin[k], g["out"][k], rtol=1e-4, atol=1e-4)


def _safe_grad_run(net, g, gs):
    """Forward with the recorded draws and the reference's importance samples, loss = <outputs, cotangents> (the cotangents
    are additionally zero on rays whose u sit within 2e-5 of a cdf knot, oracle/make_golden.py:safe_ray_mask)."""
    rnd = {k: torch.from_numpy(v).to(DEV) for k, v in g["rnd"].items()}
    rnd["z_samples"] = torch.from_numpy(gs["z_samples"]).to(DEV)       # stage-wise: the fine pass runs on the reference's samples
    out = net(torch.from_numpy(g["rays"]).to(DEV), (1.2, 12.0), randoms=rnd)
    loss = sum((out[k] * torch.from_numpy(v).to(DEV)).sum() for k, v in gs["gout"].items())
    assert abs(loss.item() - float(gs["loss"])) <= 1e-4 * max(1.0, abs(float(gs["loss"])))
    loss.backward()
    return {n: p.grad for n, p in net.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("mode", ["simt", "exact"])
def test_flower_semantic_head_gradients(mode):
    """--fix_backbone recipe: only semantic_linear.{0,2} of both nets receive gradients (run_nerf.py:307-318); 1e-3 of max
    against reference autograd on both passes."""
    g, gs = load_golden("flower_train_64_semgrads"), load_golden("flower_train_64_allgrads")
    net = flower_net(mode, perturb=1.0, raw_noise_std=1.0).train()
    for n, p in net.named_parameters():
        p.requires_grad_("semantic_linear" in n)
    got = _safe_grad_run(net, g, gs)
    assert set(got) == {n for n in gs["grads"] if "semantic_linear" in n} and len(got) == 8
    for n, gr in got.items():
        ref = gs["grads"][n]
        err = np.abs(gr.cpu().numpy() - ref).max()
        assert err <= 1e-3 * np.abs(ref).max(), (n, err, np.abs(ref).max())


@pytest.mark.parametrize("mode", ["simt", "exact"])
def test_flower_all_parameter_gradients(mode):
    """Stage-1 training (engines/trainer.py:201 with every parameter trainable) at D=8 W=256: all 1,274,124 gradients vs
    reference autograd, 1e-3 of each tensor's max."""
    g, gs = load_golden("flower_train_64_semgrads"), load_golden("flower_train_64_allgrads")
    net = flower_net(mode, perturb=1.0, raw_noise_std=1.0).train()
    got = _safe_grad_run(net, g, gs)
    assert set(got) == set(gs["grads"])
    assert sum(v.numel() for v in got.values()) == 1274124
    for n, gr in got.items():
        ref = gs["grads"][n]
        err = np.abs(gr.cpu().numpy() - ref).max()
        assert err <= 1e-3 * max(np.abs(ref).max(), 1e-8), (n, err, np.abs(ref).max())


@pytest.mark.parametrize("n,ns,ni", [(4096 + 37, 64, 128), (333, 40, 25)])
def test_semantic_head_backward_tensor_core_paths_match_fp32_recompute(monkeypatch, n, ns, ni):
    """--fix_backbone backward at a size that crosses the 4096-ray chunk with an odd tail, and at sample counts whose point totals
    (333 x 40, 333 x 65 -- odd) end inside a 32-point group of the saved activations' blocked layout.  Three implementations of the
    same gradients: (tc) activations saved by the forward kernel + weight gradients on tcgen05, (replay) trunk replayed on
    tcgen05 in backward + the same weight-gradient kernel, (wsimt) trunk replay on tcgen05 + fp32 CUDA-core GEMMs, (simt)
    everything recomputed in fp32 on CUDA cores."""
    g = load_golden("flower_eval_256")
    rays = torch.from_numpy(g["rays"]).to(DEV)
    rays = rays.repeat(1, n // rays.shape[1] + 1, 1)[:, :n].contiguous()
    rays[1] += 0.01 * torch.randn(n, 3, device=DEV, generator=torch.Generator(DEV).manual_seed(3))
    gen = torch.Generator(DEV).manual_seed(5)
    gsem = torch.randn(n, 2, device=DEV, generator=gen)
    gsem0 = torch.randn(n, 2, device=DEV, generator=gen)
    grads = {}
    for which, env, val in (("tc", None, ""), ("replay", "NSOS_SAVE_ACT_GB", "0"), ("wsimt", "NSOS_WGRAD_SIMT", "1"),
                            ("simt", "NSOS_BWD_SIMT", "1")):
        if env:
            monkeypatch.setenv(env, val)
        net = flower_net("exact", perturb=1.0, raw_noise_std=0.5).train()
        for nme, p in net.named_parameters():
            p.requires_grad_("semantic_linear" in nme)
        torch.manual_seed(11)                       # same Philox seed for all runs
        out = net(rays, (1.2, 12.0), N_samples=ns, N_importance=ni)
        ((out["semantics"] * gsem).sum() + (out["semantics0"] * gsem0).sum()).backward()
        grads[which] = {nme: p.grad.clone() for nme, p in net.named_parameters() if p.grad is not None}
        if env:
            monkeypatch.delenv(env)
    assert len(grads["tc"]) == 8 and set(grads["tc"]) == set(grads["simt"]) == set(grads["wsimt"]) == set(grads["replay"])
    for nme, ref in grads["wsimt"].items():
        # identical inputs (same trunk activations), bf16 hi+lo tensor-core contraction vs fp32 FMA: 2e-5 of max
        scale = ref.abs().max().item()
        for which in ("tc", "replay"):
            err = (grads[which][nme] - ref).abs().max().item()
            assert err <= 2e-5 * scale + 1e-6, (which, nme, err, scale)
    for nme, ref in grads["simt"].items():
        # different trunk arithmetic: a ReLU of semantic_linear.0 sitting at 0 +- 1 ulp flips its mask and moves one row of
        # the weight gradient by one point's contribution -> robust criterion plus a loose bound on the outliers
        scale = ref.abs().max().item()
        d = (grads["tc"][nme] - ref).abs()
        assert d.median().item() <= 2e-5 * scale + 1e-6, (nme, d.median().item(), scale)
        assert (d > 1e-3 * scale).float().mean().item() <= 0.02, (nme, scale)
        assert d.max().item() <= 2e-2 * scale, (nme, d.max().item(), scale)


def test_cfg1_full_gradients():
    """All-parameter backward (trunk dgrad/wgrad) vs reference autograd, tiny net, train mode, injected randoms."""
    g, gs = load_golden("cfg1_d4w64_train_grads"), load_golden("cfg1_d4w64_train_grads_safe")
    net = cfg1_net("simt", g, n_importance=32, perturb=1.0, raw_noise_std=1.0).train()
    rnd = {k: torch.from_numpy(v).to(DEV) for k, v in g["rnd"].items()}
    rnd["z_samples"] = torch.from_numpy(gs["z_samples"]).to(DEV)
    out = net(torch.from_numpy(g["rays"]).to(DEV), (1.2, 12.0), randoms=rnd)
    for k in ("rgb0", "semantics0", "acc0", "rgb", "semantics", "acc"):
        close(out[k], g["out"][k], rtol=1e-4, atol=1e-4)
    loss = sum((out[k] * torch.from_numpy(v).to(DEV)).sum() for k, v in gs["gout"].items())
    loss.backward()
    for n, p in net.named_parameters():
        ref = gs["grads"][n]
        err = np.abs(p.grad.cpu().numpy() - ref).max()
        assert err <= 1e-3 * max(np.abs(ref).max(), 1e-6), (n, err, np.abs(ref).max())


# ---- edge cases the reference exercises ---------------------------------------------------------------
@pytest.mark.parametrize("mode", ["simt", "exact"])
def test_edge_shapes_and_bounds(mode):
    from oracle import nerf_oracle as O
    g = load_golden("flower_eval_256")
    net = flower_net(mode).eval()
    sd = load_golden("flower_weights")["sd"]
    rays = g["rays"][:, :15]                                        # odd count: last pair half empty
    with torch.no_grad():
        a = net(torch.from_numpy(rays).to(DEV), (1.2, 12.0))
        # arbitrary leading shape [2,3,5,3] and tensor near/far
        r2 = torch.from_numpy(rays.reshape(2, 3, 5, 3)).to(DEV)
        near = torch.full((15, 1), 1.2, device=DEV); far = torch.full((15, 1), 12.0, device=DEV)
        b = net(r2, (near, far))
        one = net(torch.from_numpy(rays[:, :1]).to(DEV), (1.2, 12.0))
    for k in ("rgb", "semantics", "acc", "rgb0"):
        close(a[k], g["out"][k][:15], rtol=1e-4, atol=1e-4)
        assert torch.equal(a[k].reshape(b[k].shape), b[k])
        assert b[k].shape[:2] == (3, 5)
        close(one[k], g["out"][k][:1], rtol=1e-4, atol=1e-4)
    assert b["raw"].shape == (3, 5, 192, 6) and b["z_std"].shape == (3, 5) and b["weights0"].shape == (3, 5, 64)
    # N_importance=0 kwarg gates the fine pass off: only coarse keys come back (nerf_net.py:104)
    with torch.no_grad():
        c = net(torch.from_numpy(rays).to(DEV), (1.2, 12.0), N_importance=0)
    assert "rgb0" not in c and "z_std" not in c and c["weights"].shape == (15, 64)
    close(c["rgb"], g["out"]["rgb0"][:15], rtol=1e-4, atol=1e-4)
    # far rays that see nothing: acc ~ 0 -> depth 1e10 path (renderer.py:72) must not produce NaN
    with torch.no_grad():
        e = net(torch.from_numpy(rays).to(DEV), (50.0, 60.0))
    ref = O.nerfnet_forward(sd, rays, (50.0, 60.0))
    assert torch.isfinite(e["rgb"]).all()
    close(e["acc"], ref["acc"], rtol=1e-3, atol=1e-4)
    close(e["rgb"], ref["rgb"], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("mode", ["simt", "exact"])
@pytest.mark.parametrize("variant", ["white_bkgd", "no_coord", "no_sem", "imp64", "sem4", "sc128"])
def test_config_variants_vs_oracle(mode, variant):
    """Seeded random-init nets in configurations the shipped checkpoints do not cover."""
    from oracle import nerf_oracle as O
    _lib, NeRFNet = _imports()
    kw = dict(netdepth=8, netwidth=128, netdepth_fine=6, netwidth_fine=128, N_samples=64, N_importance=128,
              use_semantics=True, sem_with_coord=True, mode=mode)
    okw = dict(D=8, D_fine=6, n_samples=64, n_importance=128, use_semantics=True, sem_with_coord=True)
    if variant == "white_bkgd":
        kw["white_bkgd"] = True; okw["white_bkgd"] = True
    elif variant == "no_coord":
        kw["sem_with_coord"] = False; okw["sem_with_coord"] = False
    elif variant == "no_sem":
        kw["use_semantics"] = False; okw["use_semantics"] = False
    elif variant == "imp64":
        kw["N_importance"] = 64; okw["n_importance"] = 64
    elif variant == "sc128":                                         # largest supported sampling: 128 coarse (two coarse tiles per ray
        kw["N_samples"] = 128; kw["N_importance"] = 128                 # pair) + 128 importance = 256 fine samples (four fine tiles)
        okw["n_samples"] = 128; okw["n_importance"] = 128
    elif variant == "sem4":                                          # wide semantic head (own epilogue kind on the tcgen05 path)
        kw["sem_dim"] = 4
    torch.manual_seed(5)
    net = NeRFNet(**kw)
    # make the field non-trivial: scale the density head so that alpha is not ~0 everywhere
    with torch.no_grad():
        for m in (net.nerf.mlp, net.nerf_fine.mlp):
            m.alpha_linear.weight.mul_(30.0)
    net = net.to(DEV).eval()
    g = load_golden("flower_eval_256")
    rays = g["rays"][:, :64]
    with torch.no_grad():
        out = net(torch.from_numpy(rays).to(DEV), (1.2, 12.0), retz=True)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    ref = O.nerfnet_forward(sd, rays, (1.2, 12.0), extras=True, **okw)
    assert set(out) <= set(ref), set(out) - set(ref)
    sem = ("semantics",) if variant != "no_sem" else ()
    for k in ("rgb0", "acc0", "weights0") + tuple(k + "0" for k in sem):
        close(out[k], ref[k], rtol=1e-4, atol=1e-4)
    okn = ~(out["inds"].cpu().numpy() != ref["inds"]).any(-1) & _same_samples(out["z_samples"], ref["z_samples"])
    ok = torch.from_numpy(okn).to(DEV)
    assert ok.float().mean().item() >= 0.2
    for k in ("rgb", "acc") + sem:                                   # every ray with the reference's sample set: the full contract
        close(out[k][ok], ref[k][ok.cpu().numpy()], rtol=1e-4, atol=1e-4)
    if variant == "no_sem":
        assert "semantics" not in out and out["raw"].shape[-1] == 4


# ---- size-independent properties at BASELINE.json's full size ---------------------------------------
def _big_rays(n):
    g = load_golden("flower_eval_256")
    r = np.tile(g["rays"], (1, (n + 255) // 256, 1))[:, :n].copy()
    rng = np.random.default_rng(0)
    r[0] += rng.uniform(-0.05, 0.05, r[0].shape).astype(np.float32)
    return torch.from_numpy(r).to(DEV)


def test_full_size_exact_vs_same_device_fp32():
    """4096 rays x (64+128): tcgen05 exact mode vs the fp32 CUDA-core path on the same device."""
    rays = _big_rays(4096)
    with torch.no_grad():
        a = flower_net("exact").eval()(rays, (1.2, 12.0), retz=True)
        b = flower_net("simt").eval()(rays, (1.2, 12.0), retz=True)
    flip = (a["inds"] != b["inds"])
    assert flip[:, :-1].float().mean().item() <= 1e-3 and flip[:, -1].float().mean().item() <= 0.5
    for k in ("rgb0", "acc0", "semantics0"):
        close(a[k], b[k].cpu().numpy(), rtol=1e-4, atol=1e-4)
    ok = ~flip.any(-1) & ((a["z_samples"] - b["z_samples"]).abs().amax(-1) <= 1e-5)
    assert ok.float().mean().item() >= 0.2
    for k in ("rgb", "acc", "semantics"):
        close(a[k][ok], b[k][ok].cpu().numpy(), rtol=1e-4, atol=1e-4)
        close(a[k], b[k].cpu().numpy(), rtol=1e-4, atol=5e-3)
    # determinism and ray-permutation equivariance (rays are independent units)
    net = flower_net("exact").eval()
    with torch.no_grad():
        c = net(rays, (1.2, 12.0))
        perm = torch.randperm(4096, generator=torch.Generator().manual_seed(1)).to(DEV)
        d = net(rays[:, perm], (1.2, 12.0))
    assert torch.equal(a["rgb"], c["rgb"]) and torch.equal(a["semantics"], c["semantics"])
    assert torch.equal(c["rgb"][perm], d["rgb"]) and torch.equal(c["weights"][perm], d["weights"])
    # weights are a sub-probability distribution; z sorted
    assert (a["weights"] >= 0).all() and (a["weights"].sum(-1) <= 1 + 1e-4).all()
    assert (a["z_vals"][:, 1:] >= a["z_vals"][:, :-1]).all()


@pytest.mark.parametrize("mode", ["simt", "exact"])
def test_train_mode_philox(mode):
    """Production train mode (no injected randoms): in-kernel Philox; bounded, sorted, reproducible per seed."""
    net = flower_net(mode, perturb=1.0, raw_noise_std=1.0).train()
    rays = _big_rays(512)
    torch.manual_seed(123)
    with torch.no_grad():
        a = net(rays, (1.2, 12.0), retz=True)
    torch.manual_seed(123)
    with torch.no_grad():
        b = net(rays, (1.2, 12.0), retz=True)
    torch.manual_seed(124)
    with torch.no_grad():
        c = net(rays, (1.2, 12.0), retz=True)
    assert torch.equal(a["rgb"], b["rgb"]) and not torch.equal(a["rgb"], c["rgb"])
    z = a["z_vals"]
    assert torch.isfinite(a["rgb"]).all() and (z >= 1.2 - 1e-4).all() and (z <= 12.0 + 1e-4).all()
    assert (z[:, 1:] >= z[:, :-1]).all()
    # stratified jitter: one sample per coarse bin
    z0 = a["z_vals0"]
    t = torch.linspace(0, 1, 64, device=DEV); zl = 1.2 * (1 - t) + 12.0 * t
    mids = 0.5 * (zl[1:] + zl[:-1])
    lo = torch.cat([zl[:1], mids]); hi = torch.cat([mids, zl[-1:]])
    assert (z0 >= lo - 1e-5).all() and (z0 <= hi + 1e-5).all()
    assert 0.2 < ((z0 - lo) / (hi - lo)).mean().item() < 0.8


def test_mlp_query_matches_oracle():
    """NeRFMLP.forward as called by export_density (engines/eval.py:297)."""
    from oracle import nerf_oracle as O
    net = flower_net("simt").eval()
    g = torch.Generator().manual_seed(2)
    pts = (torch.rand(1000, 3, generator=g) * 2 - 1)
    vd = torch.nn.functional.normalize(torch.randn(1000, 3, generator=g), dim=-1)
    raw = net.nerf_fine(pts.to(DEV), viewdirs=vd.to(DEV))
    _, fine = O.split_state_dict(load_golden("flower_weights")["sd"])
    ref = O.mlp_forward(fine, O.encode(pts.numpy(), 10), O.encode(vd.numpy(), 4))
    close(raw, ref, rtol=1e-4, atol=5e-4)


@pytest.mark.parametrize("mode,tol", [("exact", 1e-4), ("fast", 2e-2)])
def test_mlp_query_dir_on_tensor_cores(mode, tol):
    """nsos_mlp_query_dir (export_density's query: many points, ONE view direction) through the tcgen05 replay mode, against the
    numpy oracle's MLP and against the fp32 CUDA-core query; ragged point count (tail of the last 64-point group)."""
    from oracle import nerf_oracle as O
    _lib, _ = _imports()
    net = flower_net(mode).eval()
    g = torch.Generator().manual_seed(5)
    n = 64 * 37 + 19
    pts = (torch.rand(n, 3, generator=g) * 2 - 1) * 3
    for vd in ((0.0, 0.0, 0.0), (0.6, -0.48, 0.64)):
        raw = net.nerf_fine.query_dir(pts.to(DEV), vd, _lib.MODES[mode])
        vdt = torch.tensor(vd).expand(n, 3).contiguous()
        _, fine = O.split_state_dict(load_golden("flower_weights")["sd"])
        ref = O.mlp_forward(fine, O.encode(pts.numpy(), 10), O.encode(vdt.numpy(), 4))
        simt = net.nerf_fine(pts.to(DEV), viewdirs=vdt.to(DEV))
        assert raw.shape == (n, 6)
        scale = np.abs(ref).max(0)
        err = (raw.cpu().numpy() - ref) / np.maximum(scale, 1.0)
        assert np.abs(err).max() <= tol * 5, np.abs(err).max()              # per channel, relative to the channel's range
        close(simt, ref, rtol=1e-4, atol=5e-4)
    # shape handling like forward: leading dims are kept
    grid = pts[:60].reshape(3, 4, 5, 3).to(DEV)
    assert net.nerf_fine.query_dir(grid, (0.0, 0.0, 0.0), _lib.MODES[mode]).shape == (3, 4, 5, 6)


def test_capture_eval_graph_replays_the_direct_call():
    """NeRFNet.capture_eval: H2D + render + D2H as one CUDA graph; replays give the bits of the direct call, for new rays too."""
    net = flower_net("exact").eval()
    g = load_golden("flower_eval_256")
    rays = torch.from_numpy(g["rays"])                                    # [2, 256, 3]
    cap = net.capture_eval(256, 1.2, 12.0)
    for shift in (0.0, 0.05):
        r = rays.clone()
        r[0] += shift
        cap.rays_host.copy_(r)
        cap.replay()
        torch.cuda.synchronize()
        with torch.no_grad():
            direct = net(r.to(DEV), (1.2, 12.0), retraw=False, retmaps=True)["maps"]
        assert torch.equal(cap.maps_host, direct.cpu())
    net.train()
    with pytest.raises(ValueError):
        net.capture_eval(256, 1.2, 12.0)


def test_cpu_tensors_fail_loudly():
    _lib, NeRFNet = _imports()
    net = NeRFNet(netdepth=2, netwidth=64, netdepth_fine=2, netwidth_fine=64, N_samples=8, N_importance=0)
    with pytest.raises(_lib.NsosError):
        net(torch.zeros(2, 4, 3), (1.0, 2.0))


def test_wide_encodings_fail_loudly():
    """multires > 10 / multires_views > 4 are CLI-exposed in the reference but outside both kernels' row pitch: refuse, never
    truncate the encoding silently."""
    _lib, NeRFNet = _imports()
    rays = torch.rand(2, 8, 3, device=DEV)
    for kw in (dict(multires=11), dict(multires_views=5)):
        net = NeRFNet(netdepth=2, netwidth=64, netdepth_fine=2, netwidth_fine=64, N_samples=8, N_importance=8, **kw).to(DEV).eval()
        with pytest.raises(_lib.NsosError):
            with torch.no_grad():
                net(rays, (1.0, 2.0))
        with pytest.raises(_lib.NsosError):
            net.nerf(torch.rand(5, 3, device=DEV), viewdirs=torch.rand(5, 3, device=DEV))


@pytest.mark.parametrize("K,N,P,trans,opts", [(256, 256, 1000, True, "mask"), (128, 256, 300, True, "acc"), (256, 256, 64, False, "bias_relu"),
                                              (64, 128, 129, False, ""),
                                              # several tiles per persistent CTA (148 x 128 rows = one wave): rotating A images, tail tile
                                              (256, 256, 148 * 128 * 2 + 77, True, "mask"), (128, 256, 148 * 128 + 5, True, "acc"),
                                              (192, 64, 148 * 128 * 3 + 130, False, "bias_relu")])
def test_rowgemm_building_block(K, N, P, trans, opts):
    """tcgen05 row GEMM of the all-parameter backward (bf16 hi/lo, 3 MMAs per product) vs fp64."""
    _lib, _ = _imports()
    L = _lib.lib()
    g = torch.Generator().manual_seed(K + N + P)
    a = (torch.randn(P, K, generator=g) * torch.logspace(-6, 0, K)[None]).to(DEV)          # gradient-like dynamic range
    w = (torch.randn(K, N, generator=g) * 0.2).to(DEV)                                      # B(k,n)
    bt = w.t().contiguous() if trans else w                                                  # stored [N,K] (k contiguous) or [K,N]
    b_rs, b_cs = (1, K) if trans else (N, 1)
    c = torch.randn(P, N, generator=g).to(DEV)
    c0 = c.clone()
    mask = torch.randn(P, N, generator=g).to(DEV) if "mask" in opts else None
    bias = torch.randn(N, generator=g).to(DEV) if "bias" in opts else None
    scratch = torch.zeros(2 * K * N * 2 + 4096, dtype=torch.uint8, device=DEV)
    _lib.check(L.nsos_selftest_rowgemm(_lib.ptr(a), K, K, _lib.ptr(bt), b_rs, b_cs, _lib.ptr(c), N, N, _lib.ptr(mask), N, _lib.ptr(bias),
                                       int("relu" in opts), int("acc" in opts), P, _lib.ptr(scratch), scratch.numel(), None), "rowgemm")
    torch.cuda.synchronize()
    ref = a.double() @ w.double()
    if bias is not None:
        ref = ref + bias.double()
    if "relu" in opts:
        ref = ref.clamp_min(0)
    if mask is not None:
        ref = torch.where(mask > 0, ref, torch.zeros_like(ref))
    if "acc" in opts:
        ref = ref + c0.double()
    err = (c.double() - ref).abs().max().item()
    assert err <= 3e-5 * ref.abs().max().item(), (err, ref.abs().max().item())


@pytest.mark.parametrize("Mo,main,aux_w,P", [(256, True, 0, 5000), (256, True, 63, 777), (256, False, 63, 130), (128, True, 0, 64)])
def test_wgrad_building_block(Mo, main, aux_w, P):
    """tcgen05 weight-gradient kernel (contraction over points, both operands transposed on the fly) vs fp64."""
    _lib, _ = _imports()
    L = _lib.lib()
    g = torch.Generator().manual_seed(Mo + aux_w + P)
    dy = (torch.randn(P, Mo, generator=g) * torch.logspace(-5, 0, Mo)[None]).to(DEV)
    x = torch.relu(torch.randn(P, 256, generator=g)).to(DEV) if main else None
    e = torch.randn(P, 64, generator=g).to(DEV) if aux_w else None
    ldw = (aux_w if aux_w else 0) + (256 if main else 0) + 5
    dw = torch.randn(Mo, ldw, generator=g).to(DEV)
    dw0 = dw.clone()
    aux_col, main_col = 2, 2 + aux_w
    db = torch.randn(Mo, generator=g).to(DEV)
    db0 = db.clone()
    scr = torch.empty(L.nsos_selftest_wgrad_scratch_bytes(), dtype=torch.uint8, device=DEV)     # per-CTA partial sums
    _lib.check(L.nsos_selftest_wgrad(_lib.ptr(dy), Mo, Mo, _lib.ptr(x), 256, main_col, _lib.ptr(e), 64, aux_w, aux_col, _lib.ptr(dw), ldw,
                                     _lib.ptr(db), P, _lib.ptr(scr), scr.numel(), None), "wgrad")
    torch.cuda.synchronize()
    refb = db0.double() + dy.double().sum(0)                               # bias gradient from the constant-one feature
    assert (db.double() - refb).abs().max().item() <= 3e-5 * refb.abs().max().item()
    ref = dw0.double()
    if main:
        ref[:, main_col:main_col + 256] += dy.double().t() @ x.double()
    if aux_w:
        ref[:, aux_col:aux_col + aux_w] += dy.double().t() @ e.double()[:, :aux_w]
    err = (dw.double() - ref).abs().max().item()
    assert err <= 3e-5 * ref.abs().max().item(), (err, ref.abs().max().item())


def test_disp_depth_acc_gradients_vs_torch_autograd():
    """Losses on the disparity / depth / acc maps train in the reference (renderer.py:69-74 are differentiable); the compositing
    backward carries all three.  Coarse pass, all parameters, fp32 path vs torch autograd through the op-for-op port."""
    from oracle import torch_port as TP
    g = load_golden("flower_eval_256")
    rays = torch.from_numpy(g["rays"][:, :32])
    net = flower_net("simt").eval()
    for p in net.parameters():
        p.requires_grad_(True)
    gen = torch.Generator().manual_seed(9)
    c = {k: torch.randn(32, 1, generator=gen) for k in ("disp", "depth", "acc")}
    out = net(rays.to(DEV), (1.2, 12.0), N_importance=0)
    loss = sum((out[k] * c[k].to(DEV)).sum() for k in c)
    loss.backward()
    sd = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in load_golden("flower_weights")["sd"].items()}
    o, d = rays[0], rays[1]
    t = torch.linspace(0., 1., 64)
    z = (1.2 * (1. - t) + 12.0 * t).expand(32, 64)
    raw = TP._query(sd, "nerf.mlp", o[:, None] + d[:, None] * z[..., None], d / torch.norm(d, dim=-1, keepdim=True), 8, True, True, 1 << 20)
    ref = TP._composite(raw, z, d, True)
    lref = sum((ref[k] * c[k]).sum() for k in c)
    assert abs(loss.item() - lref.item()) <= 1e-4 * max(1.0, abs(lref.item()))
    lref.backward()
    for n, p in net.named_parameters():
        if n.startswith("nerf.") and "semantic" not in n:
            r = sd[n].grad.numpy()
            err = np.abs(p.grad.cpu().numpy() - r).max()
            assert err <= 1e-3 * max(np.abs(r).max(), 1e-8), (n, err, np.abs(r).max())
