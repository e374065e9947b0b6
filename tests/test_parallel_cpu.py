"""Host-side logic of the data-parallel path on CPU: world_size-2 gloo groups (no GPU needed).
The CUDA kernels are replaced by a small differentiable torch stand-in: what is tested here is the
sharding, the autograd-aware gather of per-patch tensors and the single flat gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nerfsos_b200  # noqa: F401
from nerfsos_b200 import parallel as P


def test_shard_bounds_cover_in_order():
    for n in (0, 1, 7, 8, 4096, 762048):
        for ws in (1, 2, 3, 8):
            prev = 0
            for r in range(ws):
                lo, hi = P.shard_bounds(n, r, ws)
                assert lo == prev and hi >= lo and hi - lo in (n // ws, n // ws + 1)
                prev = hi
            assert prev == n
    assert P.gather_cat(torch.ones(3)) is not None          # identity when torch.distributed is not initialised


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _toy_render(w, rays):            # stand-in for kernel A: per-ray "code" [n,2] and "rgb" [n,3], differentiable in w
    h = torch.tanh(rays @ w[:3])
    return h[:, :2], torch.sigmoid(h[:, 2:5])


def _global_loss(code, rgb, gt, neg):
    """Batch-mean image loss + a toy pairwise 'correlation' term that couples patch n with patch neg[n]."""
    B = code.shape[0]
    pair = (code[:, None, :] * code[neg][:, None, :]).sum(-1).mean()
    return ((rgb - gt) ** 2).mean() + 0.1 * pair + 0.01 * (code.reshape(B, -1) ** 2).mean()


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        g = torch.Generator().manual_seed(0)
        B, R = 5, 16                                       # 5 patches (ragged over 2 ranks: 3 + 2) of 16 rays
        rays = torch.randn(B, R, 3, generator=g)
        gt = torch.rand(B, R, 3, generator=g)
        w0 = torch.randn(3, 5, generator=g)
        neg = torch.tensor([3, 4, 0, 1, 2])                 # negatives live on the other rank
        # ---- single-process reference on the global batch
        w_ref = w0.clone().requires_grad_(True)
        c, rgb = _toy_render(w_ref, rays.reshape(-1, 3))
        l_ref = _global_loss(c.reshape(B, R, 2), rgb.reshape(B, R, 3), gt, neg)
        l_ref.backward()
        # ---- data parallel: patches sharded, per-patch tensors gathered, identical global loss on every rank
        lo, hi = P.shard_bounds(B, rank, ws)
        w = w0.clone().requires_grad_(True)
        c, rgb = _toy_render(w, rays[lo:hi].reshape(-1, 3))
        c_all = P.gather_cat(c.reshape(hi - lo, R, 2))
        rgb_all = P.gather_cat(rgb.reshape(hi - lo, R, 3))
        assert c_all.shape == (B, R, 2)
        loss = _global_loss(c_all, rgb_all, gt, neg)
        loss.backward()
        n = P.allreduce_gradients([w])
        assert n == w.numel()
        ok = (abs(loss.item() - l_ref.item()) <= 1e-6 * max(1, abs(l_ref.item()))
              and torch.allclose(w.grad, w_ref.grad, rtol=1e-5, atol=1e-7))
        # ---- sharded render: every rank gets the full, ordered result
        class Net:
            def __call__(self, rb, bounds, **kw):
                c, rgb = _toy_render(w0, rb[0] + rb[1])
                return {"rgb": rgb, "semantics": c}
        rb = torch.stack([rays.reshape(-1, 3), torch.zeros(B * R, 3)], 0)
        out = P.render_sharded(Net(), rb, (1.0, 2.0))
        c_full, rgb_full = _toy_render(w0, rays.reshape(-1, 3))
        ok = ok and torch.equal(out["rgb"], rgb_full) and torch.equal(out["semantics"], c_full)
        # fewer rays than ranks: the rank with an empty shard renders one ray for the shapes, contributes no rows and does not hang
        near1, far1 = torch.full((1, 1), 1.0), torch.full((1, 1), 2.0)
        out1 = P.render_sharded(Net(), rb[:, :1], (near1, far1))
        ok = ok and out1["rgb"].shape == (1, 3) and torch.equal(out1["rgb"], rgb_full[:1]) and torch.equal(out1["semantics"], c_full[:1])
        q.put((rank, bool(ok), loss.item(), l_ref.item()))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_equivalence():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
