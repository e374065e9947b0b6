"""GPU parity tests for kernel B (patch-wise correlation losses), through the C ABI via the drop-in
CorrelationLoss / GeoCorrelationLoss modules.  Tolerances: 1e-4 rel on the scalar loss,
1e-3 rel (of the largest gradient entry) on gradients."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class Args:
    rand_neg = False
    self_corr_w = 1
    use_sim_matrix = True
    patch_stride = 6
    app_corr_params = ["0.18", "1", "0.46", "1"]
    geo_corr_params = ["0.5", "1", "3", "1"]


def _mods():
    import nerfsos_b200  # noqa: F401
    from nerfsos_b200.utils import image as I
    return I


def test_appearance_loss_matches_reference():
    I = _mods()
    g = load_golden("losses_b4_p16")
    feat = torch.from_numpy(g["feat"]).to(DEV)
    code = torch.from_numpy(g["code"]).to(DEV).requires_grad_(True)
    sim = torch.from_numpy(g["sim"]).to(DEV)
    c1 = torch.from_numpy(g["rand1"]).to(DEV) * 2 - 1
    c2 = torch.from_numpy(g["rand2"]).to(DEV) * 2 - 1
    loss = I.CorrelationLoss(Args())(feat, code, sim, coords=(c1, c2))
    ref = float(g["app_loss"])
    assert abs(loss.item() - ref) <= 1e-4 * max(1.0, abs(ref)), (loss.item(), ref)
    loss.backward()
    gr = g["app_gcode"]
    err = np.abs(code.grad.cpu().numpy() - gr).max()
    assert err <= 1e-3 * np.abs(gr).max(), (err, np.abs(gr).max())
    np.testing.assert_allclose(I.get_similarity_matrix(torch.from_numpy(g["cls"]).to(DEV)).cpu().numpy(), g["sim"], rtol=1e-5, atol=1e-6)


def test_geometry_loss_matches_reference():
    I = _mods()
    g = load_golden("losses_b4_p16")
    code = torch.from_numpy(g["code"]).to(DEV).requires_grad_(True)
    sim = torch.from_numpy(g["sim"]).to(DEV)
    depth = torch.from_numpy(g["depth"]).to(DEV)
    ray_o, ray_d = torch.from_numpy(g["ray_o"]).to(DEV), torch.from_numpy(g["ray_d"]).to(DEV)
    loss = I.GeoCorrelationLoss(Args())(depth, code, [ray_o, ray_d, None], sim)
    ref = float(g["geo_loss"])
    assert abs(loss.item() - ref) <= 1e-4 * max(1.0, abs(ref)), (loss.item(), ref)
    np.testing.assert_allclose(depth.cpu().numpy(), g["depth_clipped"], rtol=1e-6, atol=1e-6)   # in-place clip (image.py:455)
    loss.backward()
    gr = g["geo_gcode"]
    err = np.abs(code.grad.cpu().numpy() - gr).max()
    assert err <= 1e-3 * np.abs(gr).max(), (err, np.abs(gr).max())


def _geo_reference_fp64(xyz, code, neg, params, max_corr=15.0):
    """Blockwise fp64 restatement of GeoCorrelationLoss (image.py:404-482) for the full 64x64 patch size,
    one patch pair at a time (the reference's own [B,M,M] tensors would need ~9 GB)."""
    self_shift, self_w, neg_shift, neg_w = params
    B, _, M = xyz.shape
    chat = code / code.norm(dim=1, keepdim=True).clamp_min(1e-10)

    def corr(a, b):                                        # [C,M],[C,M] -> [M,M]
        return (1.0 / ((a[:, :, None] - b[:, None, :]).abs().sum(0) + 5e-2)).clamp(max=max_corr)

    total = 0.0
    for second_of, shift, w in ((lambda n: int(neg[n]), neg_shift, neg_w), (lambda n: n, self_shift, self_w)):
        fds = [corr(xyz[n], xyz[second_of(n)]) for n in range(B)]
        old = torch.stack([f.mean() for f in fds]).mean()
        acc = 0.0
        for n in range(B):
            fd = fds[n] - fds[n].mean(1, keepdim=True)
            fds[n] = None
            cd = corr(chat[n], chat[second_of(n)])
            acc = acc + (-(cd.clamp(min=0)) * (fd + old - shift)).sum()
        total = total + w * acc / (B * M * M)
    return total


def test_geometry_loss_full_size_vs_fp64_torch():
    """B=8 patches of 64x64 (M=4096, the shipped recipe): 134 M pairs per helper, never materialised."""
    I = _mods()
    g = torch.Generator().manual_seed(0)
    B, P = 8, 64
    code = torch.randn(B, 2, P, P, generator=g).to(DEV).requires_grad_(True)
    ray_o = (torch.rand(B, 3, 1, 1, generator=g) * 0.6 - 0.3).expand(B, 3, P, P).to(DEV)
    ii, jj = torch.meshgrid(torch.arange(P), torch.arange(P), indexing="xy")
    d = torch.stack([(ii * 6 - 192) / 815.0, -(jj * 6 - 192) / 815.0, -torch.ones(P, P)], 0)
    ray_d = d[None].expand(B, 3, P, P).contiguous().to(DEV)
    depth = (torch.rand(B, 1, P, P, generator=g) * 10 + 1.2).to(DEV)
    sim = I.get_similarity_matrix(torch.randn(B, 384, generator=g).to(DEV))
    mod = I.GeoCorrelationLoss(Args())
    loss = mod(depth.clone(), code, [ray_o, ray_d, None], sim)
    loss.backward()
    code64 = code.detach().double().reshape(B, 2, -1).requires_grad_(True)
    xyz = (ray_o + ray_d * depth).double().reshape(B, 3, -1)
    neg = torch.min(sim, dim=0)[1]
    ref = _geo_reference_fp64(xyz, code64, neg, (0.5, 1.0, 3.0, 1.0))
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-4 * max(1.0, abs(ref.item())), (loss.item(), ref.item())
    gr = code64.grad.reshape(B, 2, P, P).float()
    err = (code.grad - gr).abs().max().item()
    assert err <= 1e-3 * gr.abs().max().item(), (err, gr.abs().max().item())
    # linearity in the loss weights (size-independent property): doubling both weights doubles the loss
    class A2(Args):
        geo_corr_params = ["0.5", "2", "3", "2"]
    l2 = I.GeoCorrelationLoss(A2())(depth.clone(), code.detach(), [ray_o, ray_d, None], sim)
    assert abs(l2.item() - 2 * loss.item()) <= 1e-5 * max(1.0, abs(loss.item()))


def test_losses_refuse_cpu():
    I = _mods()
    from nerfsos_b200 import _lib
    with pytest.raises(_lib.NsosError):
        I.GeoCorrelationLoss(Args())(torch.ones(2, 1, 4, 4), torch.randn(2, 2, 4, 4), [torch.zeros(2, 3, 4, 4), torch.ones(2, 3, 4, 4), None],
                                     torch.eye(2))


@pytest.mark.parametrize("split", [(0, 2, 4), (0, 1, 4), (0, 3, 4)])
def test_sharded_phases_sum_to_the_single_call(split):
    """Data-parallel evaluation of both losses (NsosLossShard): two 'ranks' evaluate disjoint query ranges of the same gathered
    batch; the all-reduce of the old_mean sums is emulated by adding the two partial sums.  Sum of the shares == one call on
    the whole batch, for the loss (1e-6) and for the code gradient of every patch (1e-5 of max) -- including gradients that
    land on a negative patch owned by the other 'rank'."""
    I = _mods()
    g = load_golden("losses_b4_p16")
    feat = torch.from_numpy(g["feat"]).to(DEV)
    sim = torch.from_numpy(g["sim"]).to(DEV)
    c1 = torch.from_numpy(g["rand1"]).to(DEV) * 2 - 1
    c2 = torch.from_numpy(g["rand2"]).to(DEV) * 2 - 1
    ray_o, ray_d = torch.from_numpy(g["ray_o"]).to(DEV), torch.from_numpy(g["ray_d"]).to(DEV)
    app, geo = I.CorrelationLoss(Args()), I.GeoCorrelationLoss(Args())
    code = torch.from_numpy(g["code"]).to(DEV).requires_grad_(True)
    whole = app(feat, code, sim, coords=(c1, c2)) + 0.5 * geo(torch.from_numpy(g["depth"]).to(DEV), code, [ray_o, ray_d, None], sim)
    whole.backward()
    g_whole = code.grad.clone()
    bounds = list(zip(split[:-1], split[1:]))
    code2 = torch.from_numpy(g["code"]).to(DEV).requires_grad_(True)
    pa = [app.begin(feat, code2, sim, lo, hi - lo, coords=(c1, c2)) for lo, hi in bounds]
    pg = [geo.begin(torch.from_numpy(g["depth"]).to(DEV), code2, [ray_o, ray_d, None], sim, lo, hi - lo) for lo, hi in bounds]
    for group in (pa, pg):
        tot = sum(p.sums for p in group)                     # the all-reduce
        for p in group:
            p.sums = tot.clone()
    shares = sum(app.finish(p) for p in pa) + 0.5 * sum(geo.finish(p) for p in pg)
    assert abs(shares.item() - whole.item()) <= 1e-6 * max(1.0, abs(whole.item())), (shares.item(), whole.item())
    shares.backward()
    err = (code2.grad - g_whole).abs().max().item()
    assert err <= 1e-5 * g_whole.abs().max().item(), (err, g_whole.abs().max().item())


def test_fused_adam_matches_torch_adam():
    """One launch per step for all tensors (run_nerf.py:320 + engines/lr.py) == torch.optim.Adam, 20 steps with a decaying lr;
    state_dict keys are torch's."""
    import nerfsos_b200  # noqa: F401
    from nerfsos_b200.engines.lr import LRScheduler
    from nerfsos_b200.engines.optim import FusedAdam
    gen = torch.Generator().manual_seed(0)
    shapes = [(128, 319), (128,), (2, 128), (2,), (7,), (300, 257)]
    pa = [torch.randn(s, generator=gen).to(DEV).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa, ob = FusedAdam(pa, lr=5e-4, betas=(0.9, 0.999)), torch.optim.Adam(pb, lr=5e-4, betas=(0.9, 0.999))
    sa, sb = LRScheduler(oa, 5e-4, 0.1, 50), LRScheduler(ob, 5e-4, 0.1, 50)
    for step in range(1, 21):
        for x, y in zip(pa, pb):
            gr = torch.randn(x.shape, generator=gen).to(DEV) * (1.0 + step)
            x.grad, y.grad = gr.clone(), gr.clone()
        oa.step(); ob.step(); sa.step(step); sb.step(step)
    for x, y in zip(pa, pb):
        assert (x - y).abs().max().item() <= 2e-6 * y.abs().max().item()
    ka = oa.state_dict()["state"][0]
    assert set(ka) == {"step", "exp_avg", "exp_avg_sq"} and int(ka["step"]) == 20
    ob2 = torch.optim.Adam(pb, lr=5e-4)
    ob2.load_state_dict(oa.state_dict())                      # checkpoints are interchangeable
    v = pa[0]._version
    pa[0].grad = torch.zeros_like(pa[0]); oa.step()
    assert pa[0]._version > v                                 # the weight-pack cache sees the update
