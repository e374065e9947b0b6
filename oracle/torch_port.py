"""CPU baseline port of the render hot path on PyTorch CPU ops.  TEST / BENCH INFRASTRUCTURE ONLY.

`bench.py`'s `cpu_baseline` leg and `bench.py --impl reference` time THIS restatement on the GPU box's
host cores: the reference itself is pure Python and lives in /root/reference, which does not exist on
the GPU box, so it cannot be run there.  The port issues the same ATen operators the reference issues
(linspace, sin/cos, addmm via F.linear, relu, cumprod, searchsorted, sort, gather), in the same order and
with the same chunking (ray_chunk 32768, pts_chunk as configured), so its throughput on N cores is what
the reference's own CPU path would deliver; `tests/test_oracle_golden.py::test_torch_port_*` pins it
against the reference-generated fixtures.  Never imported by the product path.

File:line citations refer to /root/reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _encode(x, L):                                            # models/embedder.py:34-48
    freqs = 2. ** torch.linspace(0., L - 1, L)
    xf = x[..., None].expand(x.shape + (L,)) * freqs
    emb = torch.stack([torch.sin(xf).transpose(-1, -2), torch.cos(xf).transpose(-1, -2)], -2)
    return torch.cat([x, emb.reshape(x.shape[:-1] + (-1,))], -1)


def _mlp(p, pre, enc, encd, D, use_sem, sem_coord):          # models/nerf_mlp.py:67-100
    lin = lambda h, n: F.linear(h, p[f"{pre}.{n}.weight"], p[f"{pre}.{n}.bias"])
    h = enc
    for i in range(D):
        h = F.relu(lin(h, f"pts_linears.{i}"))
        if i == 4 and D > 5:
            h = torch.cat([enc, h], -1)
    alpha = lin(h, "alpha_linear")
    outs = []
    if use_sem:
        s_in = torch.cat([h, enc], -1) if sem_coord else h
        outs = [lin(F.relu(lin(s_in, "semantic_linear.0")), "semantic_linear.2")]
    hv = F.relu(lin(torch.cat([lin(h, "feature_linear"), encd], -1), "views_linears.0"))
    return torch.cat([lin(hv, "rgb_linear"), alpha] + outs, -1)


def _query(p, pre, pts, vdirs, D, use_sem, sem_coord, pts_chunk):   # models/nerf_mlp.py:179-215
    flat = pts.reshape(-1, 3)
    dflat = vdirs[:, None, :].expand(pts.shape).reshape(-1, 3)
    out = []
    for i in range(0, flat.shape[0], pts_chunk):
        out.append(_mlp(p, pre, _encode(flat[i:i + pts_chunk], 10), _encode(dflat[i:i + pts_chunk], 4), D, use_sem, sem_coord))
    return torch.cat(out, 0).reshape(pts.shape[:-1] + (-1,))


def _composite(raw, z, d, use_sem):                           # models/renderer.py:35-85 (eval: no noise)
    dists = torch.cat([z[..., 1:] - z[..., :-1], 1e10 * torch.ones_like(z[..., :1])], -1)
    dists = dists * torch.linalg.norm(d[..., None, :], ord=2, dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    alpha = 1. - torch.exp(-F.relu(raw[..., 3]) * dists)
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1. - alpha + 1e-10], -1), -1)[..., :-1]
    w = alpha * T
    ret = dict(rgb=torch.sum(w[..., None] * rgb, -2), weights=w)
    if use_sem:
        ret["semantics"] = torch.sum(w[..., None] * raw[..., 4:], -2)
    depth = torch.sum(w * z, -1, keepdim=True)
    acc = torch.sum(w, -1, keepdim=True)
    depth[acc <= 1e-10] = 1e10
    ret.update(depth=depth, acc=acc, disp=1. / torch.max(torch.full_like(depth, 1e-10), depth / acc))
    return ret


def _resample(z, w, K):                                       # models/sampler.py:91-170 (det=True)
    mid = .5 * (z[..., 1:] + z[..., :-1])
    w = w[..., 1:-1] + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    u = torch.linspace(0., 1., K).expand(list(cdf.shape[:-1]) + [K]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below, above = torch.clamp(inds - 1, min=0), torch.clamp(inds, max=cdf.shape[-1] - 1)
    g = torch.stack([below, above], -1)
    shp = [g.shape[0], g.shape[1], cdf.shape[-1]]
    cg = torch.gather(cdf.unsqueeze(1).expand(shp), 2, g)
    bg = torch.gather(mid.unsqueeze(1).expand(shp), 2, g)
    den = cg[..., 1] - cg[..., 0]
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    zs = bg[..., 0] + (u - cg[..., 0]) / den * (bg[..., 1] - bg[..., 0])
    return torch.sort(torch.cat([z, zs], -1), -1)[0], zs


@torch.no_grad()
def render_eval(sd: dict, rays_o, rays_d, near, far, *, n_samples=64, n_importance=128, D=8, D_fine=8, use_sem=True,
                sem_coord=True, ray_chunk=1024 * 32, pts_chunk=1024 * 256) -> dict:
    """NeRFNet.forward in eval mode (models/nerf_net.py:132-195 -> :71-130).  sd: torch state_dict (CPU)."""
    parts = []
    for i in range(0, rays_o.shape[0], ray_chunk):
        o, d = rays_o[i:i + ray_chunk], rays_d[i:i + ray_chunk]
        vd = d / torch.norm(d, dim=-1, keepdim=True)
        t = torch.linspace(0., 1., n_samples)
        z = (near * (1. - t) + far * t).expand(o.shape[0], n_samples)
        raw = _query(sd, "nerf.mlp", o[:, None] + d[:, None] * z[..., None], vd, D, use_sem, sem_coord, pts_chunk)
        ret = _composite(raw, z, d, use_sem)
        if n_importance > 0:
            ret0 = ret
            z, zs = _resample(z, ret0["weights"], n_importance)
            raw = _query(sd, "nerf_fine.mlp", o[:, None] + d[:, None] * z[..., None], vd, D_fine, use_sem, sem_coord, pts_chunk)
            ret = _composite(raw, z, d, use_sem)
            ret["z_std"] = torch.std(zs, dim=-1, unbiased=False)
            ret.update({k + "0": v for k, v in ret0.items()})
        parts.append(ret)
    return {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}
