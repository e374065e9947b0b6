"""Single-node data parallelism for the render path: one process per GPU, torch.distributed (NCCL over
NVLink 5 / NVSwitch on the B200 box, gloo in the CPU tests).

The reference has no distributed code on its live path: rays are independent units, so
  * evaluation / rendering shards the flattened ray list into contiguous 1/G slices with NO data-path
    collective; the per-ray outputs are all-gathered once at the end (`render_sharded`);
  * training shards whole 64x64 patches across ranks.  The correlation losses pick their negative from
    ANOTHER patch of the global batch (utils/image.py:354,359,473), so the small per-patch tensors
    (semantic code, depth, rays, DINO features; 0.5 MB per patch) travel in ONE packed all-gather
    (`all_gather_rows`, static shapes, no host synchronisation); every rank then evaluates only the loss
    rows of ITS OWN patches (kernel B's sharded phases, one all-reduce of the batch-wide `old_mean` sums) and
    the code gradients that land on remote negatives come home in one all-reduce (`engines/trainer.py`);
  * parameter gradients are then summed with ONE all-reduce over a single flat fp32 buffer
    (`allreduce_gradients`; 0.33 MB under --fix_backbone, 5.1 MB for all parameters).
  Four collectives per training step in total, none of them followed by a host read.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world(group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_bounds(n: int, rank: int, world_size: int):
    """Contiguous slice [lo, hi) of n units for `rank`: sizes differ by at most one, order preserved."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """[n, L] on every rank (same n everywhere) -> [G*n, L], rank-major.  One collective, static shapes, no gradient."""
    rank, ws = world(group)
    if ws == 1:
        return x
    x = x.contiguous()
    out = x.new_empty((ws * x.shape[0],) + tuple(x.shape[1:]))
    dist.all_gather_into_tensor(out, x, group=group)
    return out


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if world(group)[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class _GatherCat(torch.autograd.Function):
    """all-gather along dim 0 (ragged sizes allowed).  Backward: every rank holds the gradient of the SAME
    global loss w.r.t. the gathered tensor, so it keeps its own slice -- no collective in backward.
    `sizes` (rows per rank) makes the gather static; without it the sizes are exchanged first (one host sync)."""

    @staticmethod
    def forward(ctx, x, group, sizes=None):
        rank, ws = world(group)
        if sizes is None:
            n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
            sz = [torch.zeros_like(n) for _ in range(ws)]
            dist.all_gather(sz, n, group=group)
            sizes = [int(s.item()) for s in sz]
        mx = max(sizes)
        pad = x.new_zeros((mx,) + tuple(x.shape[1:]))
        pad[:x.shape[0]] = x
        parts = [torch.empty_like(pad) for _ in range(ws)]
        dist.all_gather(parts, pad.contiguous(), group=group)
        ctx.lo = sum(sizes[:rank])
        ctx.n = x.shape[0]
        return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)

    @staticmethod
    def backward(ctx, g):
        return g[ctx.lo:ctx.lo + ctx.n].contiguous(), None, None


def gather_cat(x: torch.Tensor, group=None, sizes=None) -> torch.Tensor:
    """Differentiable all-gather + concat along dim 0; identity when not distributed."""
    if world(group)[1] == 1:
        return x
    return _GatherCat.apply(x, group, sizes)


def allreduce_gradients(params, group=None, average: bool = False):
    """Sum (or average) the gradients of `params` across ranks with ONE collective on a flat fp32 buffer.
    Parameters without a gradient contribute zeros so that every rank reduces the same layout."""
    rank, ws = world(group)
    params = [p for p in params if p.requires_grad]
    if ws == 1 or not params:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= ws
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat.numel()


@torch.no_grad()
def render_sharded(net, ray_batch, bound_batch, group=None, keys=("rgb", "depth", "acc", "semantics", "rgb0"), **kwargs):
    """Render a (large) ray batch with contiguous ray sharding; returns the gathered per-ray outputs on every
    rank, in the original ray order.  ray_batch [2, ..., 3] must be identical on all ranks."""
    rank, ws = world(group)
    rays_o, rays_d = ray_batch
    lead = rays_d.shape[:-1]
    ro, rd = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    n = ro.shape[0]
    lo, hi = shard_bounds(n, rank, ws)
    near, far = bound_batch
    empty = hi == lo                    # more ranks than rays: render one ray for the output shapes and contribute none of it
    sl = slice(0, 1) if empty else slice(lo, hi)
    if torch.is_tensor(near):
        near = near.reshape(n, -1)[sl]
    if torch.is_tensor(far):
        far = far.reshape(n, -1)[sl]
    out = net(torch.stack([ro[sl], rd[sl]], 0), (near, far), **kwargs)
    sizes = [b - a for a, b in (shard_bounds(n, r, ws) for r in range(ws))]        # static: no size exchange, no host sync
    res = {}
    for k in (keys if keys is not None else out.keys()):
        if k in out:
            part = out[k][:0] if empty else out[k]
            full = gather_cat(part.contiguous(), group, sizes)
            res[k] = full.reshape(*lead, *full.shape[1:])
    return res
