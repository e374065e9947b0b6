"""CPU oracle for the NeRF-SOS volumetric-rendering hot path.  TEST INFRASTRUCTURE ONLY.

This is a plain numpy (fp32) restatement of the reference algorithm.  It is imported only by
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs, and
only as the checker or the timed CPU baseline -- never by the product path under
`nerf-sos_b200/`, which must fail loudly when the CUDA library is missing.

Pinning: the reference ships no tests, so this oracle is pinned against outputs of
the UNMODIFIED reference modules imported from /root/reference in the build container; the
generating script is `oracle/make_golden.py` and the fixtures live in `tests/golden/*.npz`.
`tests/test_oracle_golden.py` re-checks every function here against those fixtures.

Every function cites the reference file:line it restates (paths relative to /root/reference).
All arithmetic is float32 unless stated; arrays are C-contiguous numpy.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def linspace01(steps: int) -> np.ndarray:
    """torch.linspace(0., 1., steps) bit-for-bit (fp32).

    ATen evaluates `start + step*i` for the lower half and `end - step*(steps-1-i)` for the upper
    half with `step = (end-start)/(steps-1)` in fp32; the upper half is a FUSED multiply-add (one
    rounding) -- verified against torch 2.11 CPU in tests/test_oracle_golden.py; CUDA code must use
    fmaf(-step, k, 1.0f).  Used at models/sampler.py:46 (t_vals) and :99 (deterministic u).
    """
    if steps == 1:
        return np.zeros(1, F32)
    step = F32(1.0) / F32(steps - 1)
    i = np.arange(steps)
    lo = (step * i.astype(F32)).astype(F32)
    hi = (1.0 - np.float64(step) * (steps - 1 - i)).astype(F32)     # fma: exact product, one rounding
    return np.where(i < steps // 2, lo, hi).astype(F32)


def _f32(x):
    return np.ascontiguousarray(x, dtype=F32)


# --------------------------------------------------------------------------------------
# a3  StratifiedSampler.forward            models/sampler.py:25-74
# --------------------------------------------------------------------------------------
def stratified_z(near, far, n_samples: int, perturb: float = 0.0, t_rand=None) -> np.ndarray:
    """z_vals [N, n_samples].  near/far: [N,1] fp32.  t_rand: [N, n_samples] U[0,1) if perturb>0.

    models/sampler.py:46-49  z = near*(1-t) + far*t   (lindisp=False is hard-wired, nerf_net.py:31)
    models/sampler.py:54-69  mids / upper / lower / lower + (upper-lower)*t_rand
    """
    near = _f32(near).reshape(-1, 1)
    far = _f32(far).reshape(-1, 1)
    t = linspace01(n_samples)[None, :]
    z = (near * (F32(1.0) - t) + far * t).astype(F32)
    if perturb > 0.0:
        assert t_rand is not None, "perturb>0 needs injected t_rand (the reference draws torch.rand)"
        mids = (F32(0.5) * (z[:, 1:] + z[:, :-1])).astype(F32)
        upper = np.concatenate([mids, z[:, -1:]], -1)
        lower = np.concatenate([z[:, :1], mids], -1)
        z = (lower + (upper - lower) * _f32(t_rand)).astype(F32)
    return z


def points(rays_o, rays_d, z) -> np.ndarray:
    """models/sampler.py:71 and :167 -- pts = o + d*z with the UN-normalised direction."""
    return (_f32(rays_o)[:, None, :] + _f32(rays_d)[:, None, :] * _f32(z)[:, :, None]).astype(F32)


# --------------------------------------------------------------------------------------
# a4  PositionEncoder.forward              models/embedder.py:34-48
# --------------------------------------------------------------------------------------
def encode(x, n_freqs: int) -> np.ndarray:
    """[..., 3] -> [..., 3 + 6*n_freqs]: [x, sin(2^0 x), cos(2^0 x), ..., sin(2^{L-1} x), cos(2^{L-1} x)].

    Frequency-major, then (sin, cos), then xyz (embedder.py:39-43: stack(-2) of the transposed
    [.., N_freq, 3] blocks, then reshape).  freq_bands = 2**linspace(0, L-1, L) are exact powers of
    two (embedder.py:26), so 2^k*x is an exact fp32 product.
    """
    x = _f32(x)
    out = [x]
    for k in range(n_freqs):
        xf = (x * F32(2.0 ** k)).astype(F32)
        out.append(np.sin(xf).astype(F32))
        out.append(np.cos(xf).astype(F32))
    return np.concatenate(out, -1).astype(F32)


# --------------------------------------------------------------------------------------
# a5/a6  NeRFMLP.forward + MLP.forward      models/nerf_mlp.py:179-215, 67-100
# --------------------------------------------------------------------------------------
def _lin(h, params, name):
    return (h @ params[name + ".weight"].T + params[name + ".bias"]).astype(F32)


def mlp_forward(params: dict, enc_pts, enc_dirs, *, D=8, skips=(4,), use_viewdirs=True,
                use_semantics=True, sem_with_coord=True, return_acts=False):
    """MLP.forward (nerf_mlp.py:67-100).  params: state_dict of one `MLP` (keys without the
    'nerf.mlp.' prefix, numpy fp32, weights [out,in]).  enc_pts [P,63], enc_dirs [P,27].
    Returns raw [P, 4(+sem_dim)] ordered [rgb(3), sigma(1), sem(sem_dim)] (nerf_mlp.py:94).
    """
    h = _f32(enc_pts)
    inp = h
    acts = {}
    for i in range(D):
        h = np.maximum(_lin(h, params, f"pts_linears.{i}"), F32(0))        # :71-72
        if i in skips:
            h = np.concatenate([inp, h], -1)                               # :74  [enc, h]
    if not use_viewdirs:
        return _lin(h, params, "output_linear")                           # :98
    alpha = _lin(h, params, "alpha_linear")                               # :77
    sem = None
    if use_semantics:
        sem_in = np.concatenate([h, inp], -1) if sem_with_coord else h    # :79  [h, enc]
        s0 = np.maximum(_lin(sem_in, params, "semantic_linear.0"), F32(0))
        sem = _lin(s0, params, "semantic_linear.2")                       # :80
        acts["sem_in"], acts["s0"] = sem_in, s0
    feat = _lin(h, params, "feature_linear")                              # :86
    hv = np.concatenate([feat, _f32(enc_dirs)], -1)                       # :87  [feature, dir_enc]
    hv = np.maximum(_lin(hv, params, "views_linears.0"), F32(0))          # :89-90
    rgb = _lin(hv, params, "rgb_linear")                                  # :92
    out = np.concatenate([rgb, alpha] + ([sem] if sem is not None else []), -1).astype(F32)
    return (out, acts) if return_acts else out


def nerf_mlp(params: dict, pts, viewdirs, *, multires=10, multires_views=4, chunk=1 << 18, **mlp_kw):
    """NeRFMLP.forward (nerf_mlp.py:179-215): flatten, chunk loop, encode, cat, MLP.
    pts [N,S,3], viewdirs [N,3] (expanded per point exactly as nerf_net.py:94 does)."""
    N, S, _ = pts.shape
    flat = _f32(pts).reshape(-1, 3)
    dirs = np.broadcast_to(_f32(viewdirs)[:, None, :], (N, S, 3)).reshape(-1, 3)
    outs = []
    for i in range(0, flat.shape[0], chunk):                               # :190
        e = encode(flat[i:i + chunk], multires)                            # :193
        ed = encode(dirs[i:i + chunk], multires_views)                     # :202
        outs.append(mlp_forward(params, e, ed, **mlp_kw))                  # :209
    out = np.concatenate(outs, 0)
    return out.reshape(N, S, -1)


# --------------------------------------------------------------------------------------
# a7  VolumetricRenderer.forward           models/renderer.py:21-85
# --------------------------------------------------------------------------------------
def composite(raw, z, rays_d, noise=None, white_bkgd=False, use_semantics=True) -> dict:
    """raw [N,S,C], z [N,S], rays_d [N,3]; noise [N,S] already multiplied by raw_noise_std or None."""
    raw, z, rays_d = _f32(raw), _f32(z), _f32(rays_d)
    dists = z[:, 1:] - z[:, :-1]                                           # :35
    dists = np.concatenate([dists, np.full_like(dists[:, :1], 1e10)], -1)  # :37
    dnorm = np.sqrt((rays_d * rays_d).sum(-1, dtype=F32)).astype(F32)[:, None]
    dists = (dists * dnorm).astype(F32)                                    # :38
    rgb = (F32(1) / (F32(1) + np.exp(-raw[..., :3]))).astype(F32)          # :41 sigmoid
    sig = raw[..., 3] + (_f32(noise) if noise is not None else F32(0))     # :50
    alpha = (F32(1) - np.exp(-np.maximum(sig, F32(0)) * dists)).astype(F32)  # :52
    Ts = np.concatenate([np.ones_like(alpha[:, :1]), F32(1) - alpha + F32(1e-10)], -1)  # :57
    Ts = np.cumprod(Ts, -1, dtype=F32)[:, :-1]                             # :58
    w = (alpha * Ts).astype(F32)                                           # :61
    out = {}
    out["rgb"] = (w[..., None] * rgb).sum(-2, dtype=F32)                   # :62
    if use_semantics:
        out["semantics"] = (w[..., None] * raw[..., 4:]).sum(-2, dtype=F32)  # :65-66 (logits)
    depth = (w * z).sum(-1, dtype=F32)[:, None]                            # :69
    acc = w.sum(-1, dtype=F32)[:, None]                                    # :71
    depth = np.where(acc <= F32(1e-10), F32(1e10), depth).astype(F32)      # :72
    with np.errstate(divide="ignore", invalid="ignore"):
        disp = (F32(1) / np.maximum(F32(1e-10), depth / acc)).astype(F32)  # :74
    if white_bkgd:                                                         # :77-81
        out["rgb"] = out["rgb"] + (F32(1) - acc)
        if use_semantics:
            out["semantics"] = out["semantics"] + (F32(1) - acc)
    out.update(disp=disp, acc=acc, weights=w, depth=depth)
    return out


# --------------------------------------------------------------------------------------
# a8  ImportanceSampler.sample_pdf         models/sampler.py:91-134
# --------------------------------------------------------------------------------------
def pdf_cdf(weights) -> np.ndarray:
    """cdf [N, M+1] from weights [N, M] (sampler.py:93-96).

    ATen's CPU cumsum accumulates fp32 inputs in DOUBLE and rounds each prefix once (measured: an
    fp64-accumulated cumsum of the reference's own pdf reproduces torch.cumsum bit-for-bit, an fp32
    sequential one only on 19 % of entries).  The normalising torch.sum is a vectorised fp32 cascade
    whose order is not reproducible here; we take it in fp64 and round once, which differs from ATen
    in the last ulp on about half the rays -- hence 'exact sample indices' is a stage-wise contract
    (SURVEY.md section 7): exact given identical cdf/u, see invert_cdf()."""
    w = (_f32(weights) + F32(1e-5)).astype(F32)
    tot = w.astype(np.float64).sum(-1, keepdims=True).astype(F32)
    pdf = (w / tot).astype(F32)
    cdf = np.cumsum(pdf.astype(np.float64), -1).astype(F32)
    return np.concatenate([np.zeros_like(cdf[:, :1]), cdf], -1).astype(F32)


def invert_cdf(bins, cdf, u):
    """Given bins [N,M+1], cdf [N,M+1], u [N,K] -> (samples [N,K] fp32, inds [N,K] int64).
    sampler.py:117-132: searchsorted(right=True), below/above clamp, denom<1e-5 -> 1, lerp.
    This stage is EXACT: integer indices must match bit-for-bit given identical cdf/u."""
    bins, cdf, u = _f32(bins), _f32(cdf), _f32(u)
    M1 = cdf.shape[-1]
    inds = (cdf[:, None, :] <= u[:, :, None]).sum(-1).astype(np.int64)      # == searchsorted(right=True)
    below = np.maximum(0, inds - 1)
    above = np.minimum(M1 - 1, inds)
    cb = np.take_along_axis(cdf, below, -1)
    ca = np.take_along_axis(cdf, above, -1)
    bb = np.take_along_axis(bins, below, -1)
    ba = np.take_along_axis(bins, above, -1)
    denom = (ca - cb).astype(F32)
    denom = np.where(denom < F32(1e-5), F32(1), denom).astype(F32)
    t = ((u - cb) / denom).astype(F32)
    samples = (bb + t * (ba - bb)).astype(F32)
    return samples, inds


def importance_z(z_vals, weights, n_importance: int, perturb: float = 0.0, u=None):
    """ImportanceSampler.forward (sampler.py:136-170) -> (z_fine [N,S+K] sorted, z_samples, inds, cdf, u)."""
    z_vals = _f32(z_vals)
    mid = (F32(0.5) * (z_vals[:, 1:] + z_vals[:, :-1])).astype(F32)         # :157
    cdf = pdf_cdf(_f32(weights)[:, 1:-1])                                    # :158
    if perturb == 0.0:
        u = np.broadcast_to(linspace01(n_importance)[None, :], (z_vals.shape[0], n_importance))
    else:
        assert u is not None, "perturb>0 needs injected u (the reference draws torch.rand)"
    u = _f32(u)
    z_samples, inds = invert_cdf(mid, cdf, u)
    z_fine = np.sort(np.concatenate([z_vals, z_samples], -1), -1).astype(F32)  # :161
    return z_fine, z_samples, inds, cdf, u


# --------------------------------------------------------------------------------------
# a2  NeRFNet.render_rays / a1 NeRFNet.forward     models/nerf_net.py:71-130, 132-195
# --------------------------------------------------------------------------------------
def split_state_dict(sd: dict):
    """{'nerf.mlp.X': ..} -> (coarse, fine) dicts keyed 'X' (numpy fp32)."""
    c, f = {}, {}
    for k, v in sd.items():
        a = np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=F32)
        if k.startswith("nerf.mlp."):
            c[k[len("nerf.mlp."):]] = a
        elif k.startswith("nerf_fine.mlp."):
            f[k[len("nerf_fine.mlp."):]] = a
    return c, (f if f else c)


def render_rays(coarse: dict, fine: dict, rays_o, rays_d, near, far, *, n_samples=64, n_importance=128,
                perturb=0.0, raw_noise_std=0.0, white_bkgd=False, randoms: dict | None = None,
                D=8, D_fine=8, use_semantics=True, sem_with_coord=True, multires=10, multires_views=4,
                extras=False) -> dict:
    """nerf_net.py:71-130 with viewdirs = d/||d|| from nerf_net.py:163-166.
    randoms (train mode): {'t_rand':[N,Sc], 'noise0':[N,Sc], 'u':[N,K], 'noise1':[N,Sc+K]} -- the four
    draws the reference makes, in call order (sampler.py:61, renderer.py:47, sampler.py:103, renderer.py:47);
    noise arrays are standard normal and are scaled by raw_noise_std here."""
    rays_o, rays_d = _f32(rays_o), _f32(rays_d)
    N = rays_o.shape[0]
    near = np.broadcast_to(_f32(near).reshape(-1, 1), (N, 1))
    far = np.broadcast_to(_f32(far).reshape(-1, 1), (N, 1))
    rnd = randoms or {}
    nrm = np.sqrt((rays_d * rays_d).sum(-1, dtype=F32, keepdims=True)).astype(F32)
    viewdirs = (rays_d / nrm).astype(F32)                                    # nerf_net.py:165
    kw = dict(use_semantics=use_semantics, sem_with_coord=sem_with_coord,
              multires=multires, multires_views=multires_views)
    z = stratified_z(near, far, n_samples, perturb, rnd.get("t_rand"))       # :93
    raw = nerf_mlp(coarse, points(rays_o, rays_d, z), viewdirs, D=D, **kw)   # :95
    n0 = _f32(rnd["noise0"]) * F32(raw_noise_std) if raw_noise_std > 0 else None
    ret = composite(raw, z, rays_d, n0, white_bkgd, use_semantics)           # :96
    ret["raw"] = raw
    if extras:
        ret["z_vals"] = z
    if n_importance > 0:
        ret0 = ret
        z_fine, z_samples, inds, cdf, u = importance_z(z, ret0["weights"], n_importance, perturb, rnd.get("u"))
        raw = nerf_mlp(fine, points(rays_o, rays_d, z_fine), viewdirs, D=D_fine, **kw)   # :113
        n1 = _f32(rnd["noise1"]) * F32(raw_noise_std) if raw_noise_std > 0 else None
        ret = composite(raw, z_fine, rays_d, n1, white_bkgd, use_semantics)  # :115
        ret["raw"] = raw
        zs64 = z_samples.astype(np.float64)
        ret["z_std"] = np.sqrt(((zs64 - zs64.mean(-1, keepdims=True)) ** 2).mean(-1)).astype(F32)  # :124
        if extras:
            ret.update(z_vals=z_fine, z_samples=z_samples, inds=inds, cdf=cdf, u=u)
        for k in list(ret0):
            ret[k + "0"] = ret0[k]                                           # :127-128
    return ret


def fine_pass_on(sd: dict, ray_batch, z_fine, *, D_fine=8, white_bkgd=False, use_semantics=True, sem_with_coord=True,
                 multires=10, multires_views=4) -> dict:
    """The fine pass of render_rays (nerf_net.py:113-115) on GIVEN sorted sample depths z_fine [N, S]: fine MLP + compositing.
    Stage-wise checker for rays whose importance samples legitimately differ from the reference's (ill-conditioned inverse
    cdf in low-mass bins): whatever samples a kernel drew, its fine maps must equal this on the same samples."""
    _, fine = split_state_dict(sd)
    rays_o, rays_d = _f32(ray_batch[0]).reshape(-1, 3), _f32(ray_batch[1]).reshape(-1, 3)
    nrm = np.sqrt((rays_d * rays_d).sum(-1, dtype=F32, keepdims=True)).astype(F32)
    viewdirs = (rays_d / nrm).astype(F32)
    z_fine = _f32(z_fine)
    raw = nerf_mlp(fine, points(rays_o, rays_d, z_fine), viewdirs, D=D_fine, use_semantics=use_semantics,
                   sem_with_coord=sem_with_coord, multires=multires, multires_views=multires_views)
    ret = composite(raw, z_fine, rays_d, None, white_bkgd, use_semantics)
    ret["raw"] = raw
    return ret


def nerfnet_forward(sd: dict, ray_batch, bounds, *, ray_chunk=1 << 15, **kw) -> dict:
    """NeRFNet.forward (nerf_net.py:132-195): flatten any leading ray shape, chunk loop, cat, unflatten.
    sd: full state_dict (torch tensors or numpy); ray_batch [2, ..., 3]."""
    coarse, fine = split_state_dict(sd)
    rays_o, rays_d = np.asarray(ray_batch[0], F32), np.asarray(ray_batch[1], F32)
    assert rays_o.shape == rays_d.shape                                      # :155
    lead = rays_d.shape[:-1]
    ro, rd = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    near, far = bounds
    N = ro.shape[0]
    near = np.full((N, 1), near, F32) if np.isscalar(near) else _f32(near).reshape(N, 1)
    far = np.full((N, 1), far, F32) if np.isscalar(far) else _f32(far).reshape(N, 1)
    rnd = kw.pop("randoms", None)
    parts = []
    for i in range(0, N, ray_chunk):                                         # :177
        sl = slice(i, min(i + ray_chunk, N))
        r = {k: v[sl] for k, v in rnd.items()} if rnd else None
        parts.append(render_rays(coarse, fine, ro[sl], rd[sl], near[sl], far[sl], randoms=r, **kw))
    out = {k: np.concatenate([p[k] for p in parts], 0) for k in parts[0]}    # :188
    return {k: v.reshape(*lead, *v.shape[1:]) for k, v in out.items()}       # :191-193


# --------------------------------------------------------------------------------------
# backward of a7 (compositing) -- SURVEY.md Appendix A.1, checked against autograd in the fixtures
# --------------------------------------------------------------------------------------
def composite_backward(raw, z, rays_d, g_rgb, g_sem, g_depth=None, g_acc=None, noise=None):
    """d(loss)/d(raw) [N,S,C] for upstream grads on rgb [N,3], semantics [N,2], depth [N,1], acc [N,1].
    fp64 internally (this is a checker, not a throughput path).  No white_bkgd term; depth's masked
    overwrite (renderer.py:72) blocks g_depth on rays with acc<=1e-10."""
    raw = np.asarray(raw, np.float64)
    z = np.asarray(z, np.float64)
    N, S, C = raw.shape
    dn = np.sqrt((np.asarray(rays_d, np.float64) ** 2).sum(-1))[:, None]
    # the reference forms dists in fp32 (1e10 * ||d||); keep that rounding
    d32 = np.concatenate([(_f32(z)[:, 1:] - _f32(z)[:, :-1]), np.full((N, 1), 1e10, F32)], -1) * _f32(dn)
    dist = d32.astype(np.float64)
    sig = raw[..., 3] + (np.asarray(noise, np.float64) if noise is not None else 0.0)
    a = 1.0 - np.exp(-np.maximum(sig, 0) * dist)
    om = 1.0 - a + 1e-10
    T = np.cumprod(np.concatenate([np.ones((N, 1)), om], -1), -1)[:, :-1]
    w = a * T
    c = 1.0 / (1.0 + np.exp(-raw[..., :3]))
    g_rgb = np.asarray(g_rgb, np.float64)
    G = (c * g_rgb[:, None, :]).sum(-1)
    if g_sem is not None and C > 4:
        G = G + (raw[..., 4:] * np.asarray(g_sem, np.float64)[:, None, :]).sum(-1)
    if g_depth is not None:
        acc = w.sum(-1, keepdims=True)
        G = G + np.where(acc <= 1e-10, 0.0, np.asarray(g_depth, np.float64).reshape(N, 1)) * z
    if g_acc is not None:
        G = G + np.asarray(g_acc, np.float64).reshape(N, 1)
    wG = w * G
    suffix = np.concatenate([np.cumsum(wG[:, ::-1], -1)[:, ::-1][:, 1:], np.zeros((N, 1))], -1)  # sum_{k>i}
    dalpha = T * G - suffix / om
    dsig = dalpha * dist * (1.0 - a) * (sig > 0)
    g_raw = np.zeros_like(raw)
    g_raw[..., :3] = w[..., None] * g_rgb[:, None, :] * c * (1 - c)
    g_raw[..., 3] = dsig
    if g_sem is not None and C > 4:
        g_raw[..., 4:] = w[..., None] * np.asarray(g_sem, np.float64)[:, None, :]
    return g_raw


def sem_head_backward(params: dict, sem_in, s0, g_sem_pts):
    """Gradients of the 4 `semantic_linear` tensors (the only trainable ones under --fix_backbone,
    run_nerf.py:307-318) given d(loss)/d(sem) per point [P,sem_dim].  fp64."""
    g = np.asarray(g_sem_pts, np.float64)
    s0 = np.asarray(s0, np.float64)
    x = np.asarray(sem_in, np.float64)
    W2 = params["semantic_linear.2.weight"].astype(np.float64)
    gW2 = g.T @ s0
    gb2 = g.sum(0)
    gs0 = (g @ W2) * (s0 > 0)
    gW0 = gs0.T @ x
    gb0 = gs0.sum(0)
    return {"semantic_linear.0.weight": gW0, "semantic_linear.0.bias": gb0,
            "semantic_linear.2.weight": gW2, "semantic_linear.2.bias": gb2}


# --------------------------------------------------------------------------------------
# a12  img2mse / mse2psnr / get_similarity_matrix     utils/image.py:125-137, 187-190
# --------------------------------------------------------------------------------------
def img2mse(x, y):
    return F32(np.mean(np.mean((_f32(x) - _f32(y)) ** 2, -1, dtype=F32), dtype=F32))


def mse2psnr(m):
    return F32(-10.0) * np.log(F32(m)) / np.log(F32(10.0))


def similarity_matrix(cls):
    """F.cosine_similarity(x[None], x[:,None], dim=2) (image.py:187-190), eps=1e-8 on each norm."""
    x = np.asarray(cls, np.float64)
    n = np.maximum(np.sqrt((x * x).sum(-1)), 1e-8)
    return ((x @ x.T) / (n[:, None] * n[None, :])).astype(F32)


# --------------------------------------------------------------------------------------
# a10/a11  CorrelationLoss / GeoCorrelationLoss       utils/image.py:263-482
# --------------------------------------------------------------------------------------
def _normalize_c(t, eps=1e-10):
    """F.normalize(t, dim=1, eps) (image.py:300-301)."""
    n = np.sqrt((t * t).sum(1, keepdims=True))
    return t / np.maximum(n, eps)


def grid_sample_bilinear_border(t, coords):
    """F.grid_sample(t, coords.permute(0,2,1,3), padding_mode='border', align_corners=True)
    (image.py:303-304).  t [B,C,H,W]; coords [B,h,w,2] in [-1,1], last dim (x,y).
    Output [B,C,w,h] -- note the permute: out[b,c,i,j] samples at coords[b,j,i]."""
    t = np.asarray(t, np.float64)
    B, C, H, W = t.shape
    g = np.asarray(coords, np.float64).transpose(0, 2, 1, 3)
    x = (g[..., 0] + 1) * 0.5 * (W - 1)
    y = (g[..., 1] + 1) * 0.5 * (H - 1)
    x = np.clip(x, 0, W - 1)
    y = np.clip(y, 0, H - 1)
    x0 = np.floor(x).astype(np.int64); y0 = np.floor(y).astype(np.int64)
    x1 = np.minimum(x0 + 1, W - 1); y1 = np.minimum(y0 + 1, H - 1)
    wx = x - x0; wy = y - y0
    out = np.empty((B, C) + x.shape[1:], np.float64)
    for b in range(B):
        v00 = t[b][:, y0[b], x0[b]]; v01 = t[b][:, y0[b], x1[b]]
        v10 = t[b][:, y1[b], x0[b]]; v11 = t[b][:, y1[b], x1[b]]
        out[b] = (v00 * (1 - wx[b]) * (1 - wy[b]) + v01 * wx[b] * (1 - wy[b])
                  + v10 * (1 - wx[b]) * wy[b] + v11 * wx[b] * wy[b])
    return out


def _centre(fd):
    """image.py:316-319 / :420-424 (pointwise=True)."""
    old = fd.mean()
    fd = fd - fd.mean(axis=(3, 4), keepdims=True)
    return fd - fd.mean() + old


def _corr_dot(a, b):
    return np.einsum("nchw,ncij->nhwij", a, b)                               # image.py:298


def _corr_invl1(a, b, max_depth=15.0):
    """GeoCorrelationLoss.tensor_correlation (image.py:404-413)."""
    x = a[:, :, :, :, None, None]
    y = b[:, :, None, None, :, :]
    r = np.abs(x - y).sum(1)
    r = 1.0 / (r + 5e-2)
    return np.minimum(r, max_depth)


def neg_index(sim):
    """torch.min(sim, dim=0)[1] (image.py:354): argmin over rows for each column."""
    return np.argmin(np.asarray(sim), axis=0)


def correlation_loss(feats, code, sim, coords1, coords2, params=(0.18, 1.0, 0.46, 1.0), neg_idx=None):
    """CorrelationLoss.forward (image.py:335-370) with injected coords (the reference draws torch.rand).
    feats [B,Cf,hf,wf], code [B,Cc,H,W], params=(self_shift,self_weight,neg_shift,neg_weight). fp64."""
    self_shift, self_w, neg_shift, neg_w = params
    if neg_idx is None:
        neg_idx = neg_index(sim)
    f1 = grid_sample_bilinear_border(feats, coords1)
    c1 = grid_sample_bilinear_border(code, coords1)
    f2 = grid_sample_bilinear_border(np.asarray(feats)[neg_idx], coords2)
    c2 = grid_sample_bilinear_border(np.asarray(code)[neg_idx], coords2)

    def helper(fa, fb, ca, cb, shift):
        fd = _centre(_corr_dot(_normalize_c(fa), _normalize_c(fb)))
        cd = _corr_dot(_normalize_c(ca), _normalize_c(cb))
        return (-np.maximum(cd, 0.0) * (fd - shift)).mean()                  # :323-331 zero_clamp

    return neg_w * helper(f1, f2, c1, c2, neg_shift) + self_w * helper(f1, f1, c1, c1, self_shift)


def geo_correlation_loss(depth, code, ray_o, ray_d, sim, params=(0.5, 1.0, 3.0, 1.0), neg_idx=None,
                         max_depth=15.0):
    """GeoCorrelationLoss.forward (image.py:448-482).  depth [B,1,P,P], code [B,Cc,P,P],
    ray_o/ray_d [B,3,P,P].  Returns (loss, clipped_depth) -- the reference clips `depth` IN PLACE
    (image.py:455).  O(B*P^4) memory: small P only."""
    self_shift, self_w, neg_shift, neg_w = params
    depth = np.array(depth, np.float64)
    if (depth > max_depth).any():
        depth[depth > max_depth] = depth[depth < max_depth].max()           # :455
    if neg_idx is None:
        neg_idx = neg_index(sim)
    X = np.asarray(ray_o, np.float64) + np.asarray(ray_d, np.float64) * depth  # :443
    code = np.asarray(code, np.float64)

    def helper(fa, fb, ca, cb, shift):
        fd = _centre(_corr_invl1(fa, fb, max_depth))                         # :418-424
        cd = _corr_invl1(_normalize_c(ca), _normalize_c(cb), max_depth)     # :426 -> subclass override
        return (-np.maximum(cd, 0.0) * (fd - shift)).mean()

    loss = neg_w * helper(X, X[neg_idx], code, code[neg_idx], neg_shift) + self_w * helper(X, X, code, code, self_shift)
    return loss, depth
