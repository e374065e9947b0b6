"""Live cross-check against the UNMODIFIED reference when it is present (the build container: /root/reference; skipped on
the GPU box, where only the committed fixtures travel): the reference, run here through oracle/ref_shim.py on the fixture
inputs, still reproduces the stored golden outputs, and the oracle agrees with it on inputs the fixtures do not contain."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import nerf_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference is not present")


def _ref_net(sd, **kw):
    from models.nerf_net import NeRFNet
    net = NeRFNet(**kw)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return net.eval()


def test_reference_reproduces_the_cfg1_fixture():
    g = load_golden("cfg1_d4w64_eval")
    net = _ref_net(g["sd"], netdepth=4, netwidth=64, netdepth_fine=4, netwidth_fine=64, N_samples=64, N_importance=0,
                   use_semantics=True, sem_with_coord=True)
    with torch.no_grad():
        out = net(torch.from_numpy(g["rays"]), (1.2, 12.0))
    for k in ("rgb", "acc", "depth", "semantics", "weights"):
        assert np.array_equal(out[k].numpy(), g["out"][k]), k                     # eval mode is bit-deterministic on CPU


def test_oracle_matches_reference_on_fresh_rays(flower_sd):
    """Shipped flower weights, 32 rays that are in no fixture (other seed, other bounds)."""
    net = _ref_net(flower_sd, N_samples=64, N_importance=128, use_semantics=True, sem_with_coord=True, sem_dim=2, sem_layer=2)
    gen = torch.Generator().manual_seed(123)
    o = torch.rand(32, 3, generator=gen) * 0.4 - 0.2
    d = torch.cat([torch.rand(32, 2, generator=gen) * 0.8 - 0.4, -torch.ones(32, 1)], -1)
    rays = torch.stack([o, d], 0)
    with torch.no_grad():
        ref = net(rays, (1.0, 9.0))
    mine = O.nerfnet_forward({k: np.asarray(v) for k, v in flower_sd.items()}, rays.numpy(), (1.0, 9.0))
    for k in ("rgb0", "acc0", "semantics0"):
        np.testing.assert_allclose(mine[k], ref[k].numpy(), rtol=1e-4, atol=1e-5, err_msg=k)
    same = np.abs(mine["rgb"] - ref["rgb"].numpy()).max(-1) < 1e-4               # fine pass: identical unless an index flipped
    assert same.mean() >= 0.9


class _LossArgs:      # what CorrelationLoss / GeoCorrelationLoss read from args (image.py:276-283, 385-393)
    rand_neg = False; self_corr_w = 1.0; use_sim_matrix = True; patch_stride = 6
    app_corr_params = [0.25, 0.8, 0.4, 1.3]; geo_corr_params = [0.25, 1, 1, 1]      # CO3D values / off-fixture weights


def test_oracle_losses_match_reference_on_fresh_inputs():
    """Other batch size, patch size and loss parameters than the losses_b4_p16 fixture."""
    from utils import image as ref_image
    g = torch.Generator().manual_seed(77)
    B, Pp = 3, 8
    feat = torch.randn(B, 384, 14, 14, generator=g)
    cls_ = torch.randn(B, 384, generator=g)
    sim = ref_image.get_similarity_matrix(cls_)
    code = torch.randn(B, 2, Pp, Pp, generator=g)
    rand = []
    _rand = torch.rand

    def rec(*a, **k):
        r = _rand(*a, **k); rand.append(r.clone()); return r
    torch.rand = rec
    try:
        la = float(ref_image.CorrelationLoss(_LossArgs())(feat, code, sim))
    finally:
        torch.rand = _rand
    c1, c2 = rand[0].numpy() * 2 - 1, rand[1].numpy() * 2 - 1
    mine = O.correlation_loss(feat.numpy(), code.numpy(), sim.numpy(), c1, c2, tuple(_LossArgs.app_corr_params))
    assert abs(mine - la) <= 1e-5 * max(1.0, abs(la)), (mine, la)

    ray_o = torch.rand(B, 3, Pp, Pp, generator=g) * 0.2
    ray_d = torch.cat([torch.rand(B, 2, Pp, Pp, generator=g) - 0.5, -torch.ones(B, 1, Pp, Pp)], 1)
    depth = torch.rand(B, 1, Pp, Pp, generator=g) * 17.0 + 1.0                      # some beyond max_depth = 15
    lg = float(ref_image.GeoCorrelationLoss(_LossArgs())(depth.clone(), code, [ray_o, ray_d, None], sim))
    mine_g, _ = O.geo_correlation_loss(depth.numpy(), code.numpy(), ray_o.numpy(), ray_d.numpy(), sim.numpy(),
                                       tuple(_LossArgs.geo_corr_params))
    assert abs(mine_g - lg) <= 1e-4 * max(1.0, abs(lg)), (mine_g, lg)


def test_contrast_loss_matches_reference_module():
    """utils/image.py:192-218 (NeRFContrastive, the optional contrast term of the trainer) against the reference class."""
    import torch
    from utils.image import NeRFContrastive as Ref
    import nerfsos_b200  # noqa: F401
    from nerfsos_b200.utils.image import NeRFContrastive
    x = torch.randn(8, 384, generator=torch.Generator().manual_seed(0)).requires_grad_(True)
    a = NeRFContrastive(device="cpu")(x)
    g1, = torch.autograd.grad(a, x)
    b = Ref(device="cpu")(x)
    g2, = torch.autograd.grad(b, x)
    assert torch.allclose(a, b, rtol=1e-6, atol=1e-7) and torch.allclose(g1, g2, rtol=1e-5, atol=1e-7)
