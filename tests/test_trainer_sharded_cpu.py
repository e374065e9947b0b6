"""The data-parallel training step (engines/trainer.py) on CPU: world_size-2 gloo groups, no GPU.

Kernel A and kernel B are replaced by small torch stand-ins that keep the exact interfaces the trainer uses -- a model
returning the reference's dict, and loss objects with the two-phase begin()/finish() protocol whose loss couples every
patch with a NEGATIVE patch that may live on the other rank and whose centring needs a batch-wide mean (`old_mean`,
utils/image.py:316-319).  Checked: loss value and parameter gradients of the 2-rank step equal the single-process step
on the global batch; and the N=1 path runs the same code."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nerfsos_b200  # noqa: F401
from nerfsos_b200.engines.trainer import train_one_step

PS, SD = 4, 2


class ToyNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.w = torch.nn.Parameter(torch.randn(6, 8, generator=g) * 0.5)

    def forward(self, rays, bounds, **kw):
        x = torch.cat([rays[0], rays[1]], -1)                       # [N, 6]
        h = torch.tanh(x @ self.w)
        return {"rgb": torch.sigmoid(h[:, :3]), "rgb0": torch.sigmoid(h[:, 3:6]), "depth": 2.0 + h[:, 6:7].detach() ** 2,
                "semantics": h[:, 4:6] * 1.5, "semantics0": h[:, 2:4]}


class ToyDino:
    def get_vit_attn_feat(self, x):
        B = x.shape[0]
        p = torch.nn.functional.adaptive_avg_pool2d(x, 2).reshape(B, 3, 4).permute(0, 2, 1)          # [B,4,3]
        proj = torch.linspace(-1, 1, 3 * 5).reshape(3, 5)
        f = p @ proj
        return {"cls_": f.mean(1), "feat": f}


class _Pend:
    pass


class ToyCorr:
    """loss = mean_{n,p,q} -(cd)(fd - rowmean_q fd + mean fd - shift); cd = <code[n,:,p], code[neg n,:,q]>, fd from `feats`."""
    feature_samples = 3

    def __init__(self, shift):
        self.shift = shift

    @staticmethod
    def _neg(sim):
        return torch.min(sim, dim=0)[1]

    def _fd(self, feats, neg, rows):
        a = feats[rows].flatten(2)                                      # [nq, C, M]
        b = feats[neg[rows]].flatten(2)
        return torch.einsum("ncp,ncq->npq", a, b).detach()

    def _loss(self, fd, old, code, neg, rows, B_total):
        side = int(round(fd.shape[1] ** 0.5))
        code = torch.nn.functional.adaptive_avg_pool2d(code, side)      # stand-in for the grid_sample of the appearance loss
        a, b = code[rows].flatten(2), code[neg[rows]].flatten(2)
        cd = torch.einsum("ncp,ncq->npq", a, b)
        t = fd - fd.mean(-1, keepdim=True) + old - self.shift
        return -(cd * t).sum() / (B_total * fd.shape[1] * fd.shape[2])

    def __call__(self, feats, code, sim, coords=None):                 # single-call form on the global batch
        neg, rows = self._neg(sim), slice(None)
        fd = self._fd(feats, neg, rows)
        return self._loss(fd, fd.mean(), code, neg, rows, feats.shape[0])

    def begin(self, feats, code, sim, q0, nq, coords=None, neg=None):
        p = _Pend()
        p.neg, p.rows, p.code, p.B = self._neg(sim), slice(q0, q0 + nq), code, feats.shape[0]
        p.fd = self._fd(feats, p.neg, p.rows)
        p.sums = torch.stack([p.fd.mean(-1).sum().double(), torch.zeros((), dtype=torch.float64)])   # sum of my row means
        return p

    def finish(self, p):
        old = (p.sums[0] / (p.B * p.fd.shape[1])).float()
        return self._loss(p.fd, old, p.code, p.neg, p.rows, p.B)


class ToyGeo(ToyCorr):
    def __call__(self, depth, code, rays, sim):
        return super().__call__(rays[0] + rays[1] * depth, code, sim)

    def begin(self, depth, code, rays, sim, q0, nq, neg=None):
        return super().begin(rays[0] + rays[1] * depth, code, sim, q0, nq)


class Args:
    patch_tune = True; patch_size = PS; patch_stride = 2; use_dino = True; use_correlation = True; use_geoCorr = True
    use_contrast = False; rgb_w = 1.0; correlation_w = 1.0; Gcorrelation_w = 0.3; contrast_w = 0.0; i_print = 0


class Loader:
    class dataset:
        @staticmethod
        def near_far(): return 1.0, 2.0
        @staticmethod
        def radii(): return None


def _batch(B):
    g = torch.Generator().manual_seed(3)
    return torch.randn(B, PS * PS, 2, 3, generator=g), torch.rand(B, PS * PS, 3, generator=g)


def _step(rays, gt, step=1):
    net = ToyNet()
    opt = torch.optim.SGD(net.parameters(), lr=0.0)
    out = train_one_step((rays, gt), [net, ToyDino()], opt, None, Loader(), step, [None, None, ToyCorr(0.2), ToyGeo(0.4)], torch.device("cpu"), Args())
    return {k: float(v) for k, v in out.items() if torch.is_tensor(v)}, net.w.grad.clone()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        B = 4
        rays, gt = _batch(B)
        per = B // ws
        out, grad = _step(rays[rank * per:(rank + 1) * per], gt[rank * per:(rank + 1) * per])
        q.put((rank, out, grad.tolist()))
    finally:
        dist.destroy_process_group()


def test_single_process_matches_plain_global_loss():
    """N=1: the sharded code path (q0=0, nq=B, identities for the collectives) equals the plainly written global loss."""
    rays, gt = _batch(4)
    out, grad = _step(rays, gt)
    net = ToyNet()
    ret = net(rays.reshape(-1, 2, 3).permute(1, 0, 2), None)
    p = lambda t: t.reshape(4, PS, PS, -1)
    mse = ((p(ret["rgb"]) - gt.reshape(4, PS, PS, 3)) ** 2).mean() + ((p(ret["rgb0"]) - gt.reshape(4, PS, PS, 3)) ** 2).mean()
    with torch.no_grad():
        x = torch.nn.functional.interpolate(p(ret["rgb"]).permute(0, 3, 1, 2), (PS * 2, PS * 2))
        from nerfsos_b200.engines.trainer import normalize_batch
        d = ToyDino().get_vit_attn_feat(normalize_batch(x))
    from nerfsos_b200.utils.image import get_similarity_matrix
    sim = get_similarity_matrix(d["cls_"])
    feat = d["feat"].reshape(4, 2, 2, 5).permute(0, 3, 1, 2)
    s0, s1 = p(ret["semantics0"]).permute(0, 3, 1, 2), p(ret["semantics"]).permute(0, 3, 1, 2)
    rb = rays.reshape(-1, 2, 3).permute(1, 0, 2)
    ro, rd = p(rb[0]).permute(0, 3, 1, 2), p(rb[1]).permute(0, 3, 1, 2)
    dep = p(ret["depth"]).permute(0, 3, 1, 2)
    total = mse + ToyCorr(0.2)(feat, s0, sim) + ToyCorr(0.2)(feat, s1, sim) + 0.3 * (ToyGeo(0.4)(dep, s0, [ro, rd], sim) + ToyGeo(0.4)(dep, s1, [ro, rd], sim))
    total.backward()
    assert abs(out["loss"] - float(total.detach())) <= 1e-6 * max(1.0, abs(float(total.detach())))
    assert torch.allclose(grad, net.w.grad, rtol=1e-5, atol=1e-7)


def test_world2_gloo_step_equals_single_process_global_batch():
    rays, gt = _batch(4)
    ref_out, ref_grad = _step(rays, gt)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, out, grad in res:
        for k in ("loss", "img0", "img1", "corr0", "corr1", "geo_corr0", "geo_corr1"):
            assert abs(out[k] - ref_out[k]) <= 1e-6 * max(1.0, abs(ref_out[k])), (rank, k, out[k], ref_out[k])
        grad = torch.tensor(grad)
        assert torch.allclose(grad, ref_grad, rtol=1e-5, atol=1e-7), (rank, (grad - ref_grad).abs().max())
