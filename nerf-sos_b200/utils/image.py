"""Drop-ins for the loss side of utils/image.py of the reference (same class names, ctor `args`
attributes and forward signatures); the all-pairs arithmetic runs in libnerfsos.so (kernel B).

    img2mse / mse2psnr / get_similarity_matrix   <-> utils/image.py:125-137, 187-190   (tiny torch ops)
    CorrelationLoss(args)(feats, code, sim)                           <-> utils/image.py:263-370
    GeoCorrelationLoss(args)(depth, code, [ray_o, ray_d, gt], sim)    <-> utils/image.py:373-482

What stays in torch (host glue, autograd-visible): the random sample coordinates and F.grid_sample
(image.py:343-362), argmin over the similarity matrix (:354), the in-place depth clip (:455) and
XYZ = o + d*depth (:443).  Extra kwargs for tests: coords=(coords1, coords2) injects the two random draws.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib


def img2mse(x, y, reduction='mean'):
    diff = torch.mean((x - y) ** 2, -1)
    if reduction == 'mean':
        return torch.mean(diff)
    if reduction == 'sum':
        return torch.sum(diff)
    return diff


_LOG10 = {}


def mse2psnr(x):
    if isinstance(x, float):
        x = torch.tensor([x])
    if x.device not in _LOG10:                         # log(10) as the reference forms it (fp32 tensor), created once per device
        _LOG10[x.device] = torch.log(torch.tensor([10.], device=x.device))
    return -10. * torch.log(x) / _LOG10[x.device]


def get_similarity_matrix(x):
    return F.cosine_similarity(x.unsqueeze(0), x.unsqueeze(1), dim=2)


_WS = {}


class NeRFContrastive(nn.Module):
    """image.py:192-218 -- the optional contrast term of trainer.py:168-170 on the DINO class tokens [B, D]: with the largest and
    the smallest off-diagonal cosine similarity of the batch, loss = -log(max / (max + min)).  Plain PyTorch (B x B, off the hot
    path); only the min/max variant exists in the reference."""

    def __init__(self, temperature=1, device=None, verbose=False, min_max_contrast=True):
        super().__init__()
        if not min_max_contrast:
            raise NotImplementedError("only min_max_contrast=True is implemented (as in the reference)")
        self.device, self.verbose = device, verbose
        self.register_buffer("temperature", torch.tensor(temperature))

    def forward(self, embeddings):
        sim = F.cosine_similarity(embeddings.unsqueeze(1), embeddings.unsqueeze(0), dim=2)
        off = sim[~torch.eye(sim.shape[0], dtype=torch.bool, device=sim.device)]
        hi, lo = off.max(), off.min()
        return -torch.log(hi / (hi + lo))


def _workspace(nbytes, device):
    ws = _WS.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = _WS[device] = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
    return ws


def _params4(p):
    return (C.c_float * 4)(*[float(v) for v in p])


def _shard_struct(phase, q0, nq, B_total, sums):
    return _lib.LossShard(int(q0), int(nq), int(B_total), int(phase), _lib.ptr(sums))


class _Pending:
    """Phase-1 state of a sharded loss call (row means in `ws`, partial old_mean sums in `sums`)."""
    __slots__ = ("kind", "ws", "sums", "args", "params", "q0", "nq", "B_total")


class _GeoCorrFn(torch.autograd.Function):
    """GeoCorrelationLoss on [B,3,M] points / [B,C,M] code.  pend=None: the whole batch in one call (phase 0).
    pend=_Pending: phase 2 of a sharded call -- this rank's share of the loss, gradient for every patch it touched."""

    @staticmethod
    def forward(ctx, xyz, code, neg_idx, params, pend=None):
        if xyz.device.type != "cuda":
            raise _lib.NsosError("GeoCorrelationLoss runs on CUDA only (no CPU fallback)")
        L = _lib.lib()
        B, Cc = code.shape[0], code.shape[1]
        M = code.shape[2] * code.shape[3]
        xyz_f = xyz.detach().reshape(B, 3, M).float().contiguous()
        code_f = code.detach().reshape(B, Cc, M).float().contiguous()
        neg = neg_idx.to(torch.int64).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=code.device)
        need_g = code.requires_grad
        g = torch.empty_like(code_f) if need_g else None
        nb = L.nsos_geo_corr_workspace_bytes(B, Cc, M)
        if nb == 0:
            raise _lib.NsosError("geo correlation loss: unsupported sizes")
        with torch.cuda.device(code.device):        # the library launches on the calling thread's current device
            if pend is None:
                ws = _workspace(nb, code.device)
                _lib.check(L.nsos_geo_corr_loss(_lib.ptr(xyz_f), _lib.ptr(code_f), _lib.ptr(neg), _params4(params), _lib.ptr(loss), _lib.ptr(g),
                                                B, Cc, M, _lib.ptr(ws), ws.numel(), _lib.cur_stream(code.device)), "nsos_geo_corr_loss")
            else:
                sh = _shard_struct(2, pend.q0, pend.nq, pend.B_total, pend.sums)
                _lib.check(L.nsos_geo_corr_loss_sharded(_lib.ptr(xyz_f), _lib.ptr(code_f), _lib.ptr(neg), _params4(params), _lib.ptr(loss),
                                                        _lib.ptr(g), B, Cc, M, C.byref(sh), _lib.ptr(pend.ws), pend.ws.numel(),
                                                        _lib.cur_stream(code.device)), "nsos_geo_corr_loss_sharded")
        ctx.g, ctx.shape = g, code.shape
        return loss[0]

    @staticmethod
    def backward(ctx, gl):
        g = None if ctx.g is None else (ctx.g * gl).reshape(ctx.shape)
        return None, g, None, None, None


class _AppCorrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, nfeats, code, ncode, params, pend=None):
        if code.device.type != "cuda":
            raise _lib.NsosError("CorrelationLoss runs on CUDA only (no CPU fallback)")
        L = _lib.lib()
        B, Cf, Cc = feats.shape[0], feats.shape[1], code.shape[1]
        S = feats.shape[2] * feats.shape[3]
        f = lambda t, c: t.detach().reshape(B, c, S).float().contiguous()
        f1, f2, c1, c2 = f(feats, Cf), f(nfeats, Cf), f(code, Cc), f(ncode, Cc)
        loss = torch.empty(1, dtype=torch.float32, device=code.device)
        need_g = code.requires_grad or ncode.requires_grad
        g1 = torch.empty_like(c1) if need_g else None
        g2 = torch.empty_like(c2) if need_g else None
        with torch.cuda.device(code.device):
            if pend is None:
                ws = _workspace(L.nsos_app_corr_workspace_bytes(B, Cf, Cc, S), code.device)
                _lib.check(L.nsos_app_corr_loss(_lib.ptr(f1), _lib.ptr(f2), _lib.ptr(c1), _lib.ptr(c2), _params4(params), _lib.ptr(loss),
                                                _lib.ptr(g1), _lib.ptr(g2), B, Cf, Cc, S, _lib.ptr(ws), ws.numel(),
                                                _lib.cur_stream(code.device)), "nsos_app_corr_loss")
            else:
                sh = _shard_struct(2, 0, B, pend.B_total, pend.sums)
                _lib.check(L.nsos_app_corr_loss_sharded(_lib.ptr(f1), _lib.ptr(f2), _lib.ptr(c1), _lib.ptr(c2), _params4(params),
                                                        _lib.ptr(loss), _lib.ptr(g1), _lib.ptr(g2), B, Cf, Cc, S, C.byref(sh),
                                                        _lib.ptr(pend.ws), pend.ws.numel(), _lib.cur_stream(code.device)),
                           "nsos_app_corr_loss_sharded")
        ctx.g1, ctx.g2, ctx.shape = g1, g2, code.shape
        return loss[0]

    @staticmethod
    def backward(ctx, gl):
        if ctx.g1 is None:
            return None, None, None, None, None, None
        return None, None, (ctx.g1 * gl).reshape(ctx.shape), (ctx.g2 * gl).reshape(ctx.shape), None, None


class CorrelationLoss(nn.Module):
    """STEGO-style feature-correspondence loss between DINO features and the rendered semantic code."""

    def __init__(self, args=None):
        super().__init__()
        self.feature_samples = 11
        self.self_shift, self.self_weight, self.neg_shift, self.neg_weight = 0.18, 0.67, 0.46, 0.63
        self.verbose = False
        self.rand_neg = args.rand_neg
        self.self_corr_w = args.self_corr_w
        self.use_sim_matrix = args.use_sim_matrix
        self.self_shift, self.self_weight, self.neg_shift, self.neg_weight = [float(x) for x in args.app_corr_params]

    def sample(self, t, coords):
        return F.grid_sample(t, coords.permute(0, 2, 1, 3), padding_mode='border', align_corners=True)

    def super_perm(self, size, device):
        perm = torch.randperm(size, device=device, dtype=torch.long)
        perm[perm == torch.arange(size, device=device)] += 1
        return perm % size

    def _neg_index(self, sim_matrix, n, device):
        if sim_matrix is None:
            neg = self.super_perm(n, device)
        else:
            assert len(sim_matrix.shape) == 2
            neg = torch.min(sim_matrix, dim=0)[1]
        if self.rand_neg:
            neg = torch.randperm(sim_matrix.shape[0], device=device, dtype=torch.long)
        return neg

    def _sampled(self, orig_feats, orig_code, sim_matrix, coords, rows, neg=None):
        """The four sampled tensors of image.py:343-362 for the query patches `rows` (slice) of the batch.
        `neg` [B]: negatives chosen by the caller (data-parallel training: random choices must agree on all ranks)."""
        B = orig_feats.shape[0]
        shape = [B, self.feature_samples, self.feature_samples, 2]
        if coords is None:
            coords1 = torch.rand(shape, device=orig_feats.device) * 2 - 1
            coords2 = torch.rand(shape, device=orig_feats.device) * 2 - 1
        else:
            coords1, coords2 = coords
        neg = (self._neg_index(sim_matrix, B, orig_feats.device) if neg is None else neg.to(orig_feats.device))[rows]
        c1, c2 = coords1[rows], coords2[rows]
        with torch.no_grad():
            feats = self.sample(orig_feats[rows], c1)
            neg_feats = self.sample(orig_feats[neg], c2)
        code = self.sample(orig_code[rows], c1)
        neg_code = self.sample(orig_code[neg], c2)
        return feats, neg_feats, code, neg_code

    def forward(self, orig_feats, orig_code, sim_matrix, coords=None):
        feats, neg_feats, code, neg_code = self._sampled(orig_feats, orig_code, sim_matrix, coords, slice(None))
        params = (self.self_shift, self.self_weight, self.neg_shift, self.neg_weight)
        return _AppCorrFn.apply(feats, neg_feats, code, neg_code, params)

    # ---- sharded evaluation (data-parallel training): begin() on every rank, all-reduce the `sums` of all pending calls in
    # one collective, finish() -> this rank's share of the loss (sum over ranks == forward() on the global batch)
    def begin(self, orig_feats, orig_code, sim_matrix, q0, nq, coords=None, neg=None):
        """orig_feats / orig_code / sim_matrix / coords (/ neg) describe the GLOBAL batch (gathered); this rank's queries are
        the patches [q0, q0+nq)."""
        L = _lib.lib()
        t = self._sampled(orig_feats, orig_code, sim_matrix, coords, slice(q0, q0 + nq), neg)
        B, Cf, Cc, S = nq, t[0].shape[1], t[2].shape[1], t[0].shape[2] * t[0].shape[3]
        p = _Pending()
        p.kind, p.args, p.q0, p.nq, p.B_total = "app", t, q0, nq, orig_feats.shape[0]
        p.params = (self.self_shift, self.self_weight, self.neg_shift, self.neg_weight)
        dev = orig_code.device
        p.ws = torch.empty(L.nsos_app_corr_workspace_bytes(B, Cf, Cc, S), dtype=torch.uint8, device=dev)
        p.sums = torch.zeros(2, dtype=torch.float64, device=dev)
        f = lambda x, c: x.detach().reshape(B, c, S).float().contiguous()
        f1, f2, c1, c2 = f(t[0], Cf), f(t[1], Cf), f(t[2], Cc), f(t[3], Cc)
        sh = _shard_struct(1, 0, B, p.B_total, p.sums)
        with torch.cuda.device(dev):
            _lib.check(L.nsos_app_corr_loss_sharded(_lib.ptr(f1), _lib.ptr(f2), _lib.ptr(c1), _lib.ptr(c2), _params4(p.params), None, None, None,
                                                    B, Cf, Cc, S, C.byref(sh), _lib.ptr(p.ws), p.ws.numel(), _lib.cur_stream(dev)),
                       "nsos_app_corr_loss_sharded")
        return p

    def finish(self, p):
        if p.kind == "app":
            return _AppCorrFn.apply(*p.args, p.params, p)
        xyz, code, neg = p.args
        return _GeoCorrFn.apply(xyz, code, neg, p.params, p)


class GeoCorrelationLoss(CorrelationLoss):
    """Same loss with inverse-L1 proximity of back-projected 3-D points as the (gradient-free) target."""

    def __init__(self, args=None):
        super().__init__(args)
        self.max_depth = 15
        self.ps = args.patch_stride
        self.self_shift, self.self_weight, self.neg_shift, self.neg_weight = [float(x) for x in args.geo_corr_params]

    def depth2pts(self, depth, batch_rays):
        ray_o, ray_d = batch_rays[0], batch_rays[1]
        return ray_o + ray_d * depth

    def _points(self, depth, batch_rays):
        with torch.no_grad():
            # depth[depth > 15] = depth[depth < 15].max(), in place like image.py:455 -- without the host read of `far.any()`
            near_max = torch.where(depth < self.max_depth, depth, depth.new_full((), -float("inf"))).max()
            depth.copy_(torch.where(depth > self.max_depth, near_max, depth))
            return self.depth2pts(depth, batch_rays)

    def forward(self, orig_feats, orig_code, batch_rays, sim_matrix):
        xyz = self._points(orig_feats, batch_rays)
        neg = self._neg_index(sim_matrix, orig_code.shape[0], orig_code.device)
        params = (self.self_shift, self.self_weight, self.neg_shift, self.neg_weight)
        return _GeoCorrFn.apply(xyz, orig_code, neg, params)

    def begin(self, orig_feats, orig_code, batch_rays, sim_matrix, q0, nq, neg=None):
        """Sharded evaluation, see CorrelationLoss.begin: depth / code / rays / sim (/ neg) describe the GLOBAL batch."""
        L = _lib.lib()
        xyz = self._points(orig_feats, batch_rays)
        B, Cc = orig_code.shape[0], orig_code.shape[1]
        M = orig_code.shape[2] * orig_code.shape[3]
        neg = (self._neg_index(sim_matrix, B, orig_code.device) if neg is None else neg.to(orig_code.device)).to(torch.int64).contiguous()
        p = _Pending()
        p.kind, p.args, p.q0, p.nq, p.B_total = "geo", (xyz, orig_code, neg), q0, nq, B
        p.params = (self.self_shift, self.self_weight, self.neg_shift, self.neg_weight)
        dev = orig_code.device
        p.ws = torch.empty(L.nsos_geo_corr_workspace_bytes(B, Cc, M), dtype=torch.uint8, device=dev)
        p.sums = torch.zeros(2, dtype=torch.float64, device=dev)
        xyz_f = xyz.detach().reshape(B, 3, M).float().contiguous()
        code_f = orig_code.detach().reshape(B, Cc, M).float().contiguous()
        sh = _shard_struct(1, q0, nq, B, p.sums)
        with torch.cuda.device(dev):
            _lib.check(L.nsos_geo_corr_loss_sharded(_lib.ptr(xyz_f), _lib.ptr(code_f), _lib.ptr(neg), _params4(p.params), None, None,
                                                    B, Cc, M, C.byref(sh), _lib.ptr(p.ws), p.ws.numel(), _lib.cur_stream(dev)),
                       "nsos_geo_corr_loss_sharded")
        return p
