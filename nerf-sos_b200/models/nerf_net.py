"""Drop-in for models/nerf_net.py of the reference: same ctor (nerf_net.py:22-24), same
forward(ray_batch, bound_batch, **kwargs) -> dict (:132-195), same render_rays (:71-130), same
state_dict keys; the body is one call into libnerfsos.so per forward (nsos_render_fwd) and one per
backward (nsos_render_bwd).

Extra, non-reference keyword arguments (all optional):
    mode     'auto' | 'exact' | 'fast' | 'simt'   (ctor or forward kwarg; env NSOS_MODE)
             exact = tcgen05 fp16 hi/lo split (fp32-equivalent), fast = single fp16 pass,
             simt = fp32 CUDA-core path, auto = exact when the geometry is covered else simt
    randoms  dict(t_rand, noise0, u, noise1) of CUDA tensors: inject the four random draws the reference
             makes (parity tests); default is the in-kernel Philox stream seeded from torch's CPU generator.
             Optional fifth entry z_samples [N, N_importance]: stage-wise hook, use these importance samples
             instead of inverting the kernel's own cdf (fine-pass parity on the reference's sample positions)
    retz     also return 'z_vals'/'z_vals0', 'z_samples', 'inds'
    retmaps  also return 'maps', the packed [N, 2*(6+sem_dim)+1] per-ray output row the kernel writes
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn

from .. import _lib
from .nerf_mlp import NeRFMLP

_MAP_KEYS = ("rgb", "disp", "acc", "depth")


def _save_acts(net, params, cfg, n_points, grad_mode) -> bool:
    """Save h_last / s_hid in the forward pass?  Only when exactly the semantic heads train (run_nerf.py:307-318), on a
    tcgen05 mode, for the W=256 geometry the weight-gradient kernel covers, and within NSOS_SAVE_ACT_GB (default 24)."""
    if cfg.mode == _lib.MODE_SIMT or not grad_mode or os.environ.get("NSOS_BWD_SIMT") or os.environ.get("NSOS_WGRAD_SIMT"):
        return False
    m = net.nerf.mlp
    if not (m.use_semantics and m.W == 256 and m.sem_dim <= 4 and net.nerf_fine.mlp.W == 256):
        return False
    sem_ids = {id(p) for mm in (net.nerf.mlp, net.nerf_fine.mlp) if mm.use_semantics for p in mm.semantic_linear.parameters()}
    req = [p for p in params if p.requires_grad]
    if not req or any(id(p) not in sem_ids for p in req):
        return False
    cap = float(os.environ.get("NSOS_SAVE_ACT_GB", "24")) * 2 ** 30
    return n_points * (m.W + m.W // 2) * 4 <= cap


def _cfg(net: "NeRFNet", n_samples, n_importance, perturb, raw_noise_std, mode) -> _lib.RenderCfg:
    return _lib.RenderCfg(net.nerf.desc(), net.nerf_fine.desc(), int(n_samples), int(n_importance), float(perturb),
                          float(raw_noise_std), int(bool(net.white_bkgd)), int(mode))


class _RenderFn(torch.autograd.Function):
    """Kernel A forward/backward as one autograd node.  Inputs after `ctx_args` are the parameters of the
    coarse net then the fine net (flat-buffer order) so autograd routes gradients to the nn.Parameters."""

    @staticmethod
    def launch(net, cfg, rays_o, rays_d, near, far, rnd, seed, want, params):
        """Allocate the outputs and launch kernel A; returns (out dict, saved activations or {})."""
        L = _lib.lib()
        dev = rays_o.device
        N = rays_o.shape[0]
        fine = cfg.n_importance > 0
        Sc, K = cfg.n_samples, cfg.n_importance
        Sf = Sc + K
        sem = net.nerf.mlp.sem_dim if net.nerf.mlp.use_semantics else 0
        Cr, C6 = 4 + sem, 6 + sem
        f32 = dict(dtype=torch.float32, device=dev)
        maps = torch.empty(N, 2 * C6 + 1, **f32)
        S_last = Sf if fine else Sc
        out = dict(maps=maps, weights=torch.empty(N, S_last, **f32))
        if fine:
            out["weights0"] = torch.empty(N, Sc, **f32)
        # --fix_backbone training on the tcgen05 path: keep the two activations the semantic-head gradients need, so that
        # the backward pass does not recompute the trunk (393 kB per ray at 64+192 samples; beyond the cap it replays)
        acts = {}
        if _save_acts(net, params, cfg, N * (Sc + (Sf if fine else 0)), want.get("grad", False)):
            # opaque to Python: the library writes them in its blocked layout (groups of 32 points, csrc/internal.h
            # sem_saves_blocked), hence the point count rounded up to whole groups
            Wd = net.nerf.mlp.W
            pts = lambda S: (N * S + 31) // 32 * 32
            acts["h_last"] = torch.empty(pts(S_last) * Wd, **f32)
            acts["s_hid"] = torch.empty(pts(S_last) * (Wd // 2), **f32)
            if net.nerf.mlp.sem_with_coord:                 # gamma(x) per point: saves the backward its own encoding pass
                acts["enc"] = torch.empty(pts(S_last) * 64, **f32)
            if fine:
                acts["h_last0"] = torch.empty(pts(Sc) * Wd, **f32)
                acts["s_hid0"] = torch.empty(pts(Sc) * (Wd // 2), **f32)
                if net.nerf.mlp.sem_with_coord:
                    acts["enc0"] = torch.empty(pts(Sc) * 64, **f32)
        if want["raw"] or acts:
            out["raw"] = torch.empty(N, S_last, Cr, **f32)
            if fine:
                out["raw0"] = torch.empty(N, Sc, Cr, **f32)
        need_z = want["z"] or (want.get("grad", False) and any(p.requires_grad for p in params))   # the backward needs the depths
        if need_z:
            out["z_vals"] = torch.empty(N, S_last, **f32)
            if fine:
                out["z_vals0"] = torch.empty(N, Sc, **f32)
        if want["z"] and fine:
            out["z_samples"] = torch.empty(N, K, **f32)
            out["inds"] = torch.empty(N, K, dtype=torch.int64, device=dev)
        ro = _lib.RenderOut(*[_lib.ptr(out.get(k)) for k in ("maps", "weights0", "weights", "raw0", "raw", "z_vals0", "z_vals",
                                                             "z_samples", "inds")],
                            *[_lib.ptr(acts.get(k)) for k in ("h_last0", "s_hid0", "h_last", "s_hid", "enc0", "enc")], _lib.ptr(net._status(dev)))
        rs = _lib.Randoms(*[_lib.ptr(rnd.get(k)) for k in ("t_rand", "noise0", "u", "noise1", "z_samples")])
        flat_c, flat_f = want.get("flat") or (net.nerf.flat_params(), net.nerf_fine.flat_params())
        pk_c = net.nerf.packed(cfg.mode, force=net.training, flat=flat_c)
        pk_f = net.nerf_fine.packed(cfg.mode, force=net.training, flat=flat_f) if fine else pk_c
        wsz = L.nsos_render_workspace_bytes(cfg, N)
        ws = net._workspace(wsz, dev)
        with torch.cuda.device(dev):        # the library launches on the calling thread's current device
            _lib.check(L.nsos_render_fwd(cfg, _lib.ptr(flat_c), _lib.ptr(flat_f), _lib.ptr(pk_c), _lib.ptr(pk_f), _lib.ptr(rays_o),
                                         _lib.ptr(rays_d), _lib.ptr(near), _lib.ptr(far), C.byref(rs), seed, C.byref(ro), _lib.ptr(ws),
                                         ws.numel(), N, _lib.cur_stream(dev)), "nsos_render_fwd")
        return out, acts

    @staticmethod
    def forward(ctx, net, cfg, rays_o, rays_d, near, far, rnd, seed, want, n_coarse_params, *params):
        out, acts = _RenderFn.launch(net, cfg, rays_o, rays_d, near, far, rnd, seed, want, params)
        ctx.acts = dict(acts, raw=out.get("raw"), raw0=out.get("raw0")) if acts else None
        ctx.net, ctx.cfg, ctx.rnd, ctx.seed, ctx.n_coarse = net, cfg, rnd, seed, n_coarse_params
        ctx.shapes = [p.shape for p in params]
        ctx.req = [p.requires_grad for p in params]
        ctx.save_for_backward(rays_o, rays_d, out.get("z_vals0"), out.get("z_vals"))
        ctx.names = list(out.keys())
        outs = tuple(out[k] for k in ctx.names)
        ctx.mark_non_differentiable(*[o for k, o in zip(ctx.names, outs) if k != "maps"])
        ctx.set_materialize_grads(False)
        want["names"] = ctx.names          # read back by render_rays (a Function may only return tensors cleanly)
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        g_maps = gouts[ctx.names.index("maps")]
        n_fixed = 10
        if g_maps is None or not any(ctx.req):
            return (None,) * (n_fixed + len(ctx.req))
        L = _lib.lib()
        net, cfg = ctx.net, ctx.cfg
        rays_o, rays_d, z0, z1 = ctx.saved_tensors
        dev = rays_o.device
        N = rays_o.shape[0]
        fine = cfg.n_importance > 0
        flat_c, flat_f = net.nerf.flat_params(), net.nerf_fine.flat_params()
        g_c = torch.zeros_like(flat_c)
        g_f = torch.zeros_like(flat_f) if fine else g_c
        names_c = [n for n, _ in net.nerf.mlp.named_parameters()]
        # trunk gradients are needed iff any non-semantic parameter requires grad (run_nerf.py:307-318)
        plist = list(net.nerf._flat.params) + (list(net.nerf_fine._flat.params) if fine else [])
        sem_ids = set()
        for m in ([net.nerf.mlp] + ([net.nerf_fine.mlp] if fine else [])):
            if m.use_semantics:
                sem_ids |= {id(p) for p in m.semantic_linear.parameters()}
        trunk = int(any(r and id(p) not in sem_ids for p, r in zip(plist, ctx.req)))
        rs = _lib.Randoms(*[_lib.ptr(ctx.rnd.get(k)) for k in ("t_rand", "noise0", "u", "noise1", "z_samples")])
        wsz = L.nsos_render_bwd_workspace_bytes(cfg, N, trunk)
        ws = net._workspace(wsz, dev)
        # the packed images of the forward call (parameters are unchanged between forward and backward)
        pk_c = net.nerf.packed(cfg.mode)
        pk_f = net.nerf_fine.packed(cfg.mode) if fine else pk_c
        saved = None
        if ctx.acts and not trunk:
            saved = C.byref(_lib.RenderOut(None, None, None, _lib.ptr(ctx.acts.get("raw0")), _lib.ptr(ctx.acts.get("raw")), None, None,
                                           None, None, *[_lib.ptr(ctx.acts.get(k)) for k in ("h_last0", "s_hid0", "h_last", "s_hid", "enc0", "enc")], None))
        with torch.cuda.device(dev):
            _lib.check(L.nsos_render_bwd(cfg, _lib.ptr(flat_c), _lib.ptr(flat_f), _lib.ptr(pk_c), _lib.ptr(pk_f), _lib.ptr(rays_o),
                                         _lib.ptr(rays_d), _lib.ptr(z0),
                                         _lib.ptr(z1), C.byref(rs), ctx.seed, _lib.ptr(g_maps.contiguous()), _lib.ptr(g_c), _lib.ptr(g_f),
                                         trunk, saved, _lib.ptr(ws), ws.numel(), N, _lib.cur_stream(dev)), "nsos_render_bwd")
        ctx.acts = None
        grads = []
        offs = list(net.nerf._flat.offsets) + (list(net.nerf_fine._flat.offsets) if fine else [])
        for i, (shape, req) in enumerate(zip(ctx.shapes, ctx.req)):
            if not req:
                grads.append(None)
                continue
            g = g_c if i < ctx.n_coarse else g_f
            n = 1
            for s in shape:
                n *= s
            grads.append(g[offs[i]:offs[i] + n].view(shape))
        return (None,) * n_fixed + tuple(grads)


class NeRFNet(nn.Module):

    def __init__(self, netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, N_samples=64, N_importance=64,
                 viewdirs=True, use_embed=True, multires=10, multires_views=4, conv_embed=False, ray_chunk=1024 * 32,
                 pts_chuck=1024 * 64, perturb=1., raw_noise_std=0., white_bkgd=False, use_semantics=False, sem_layer=2,
                 sem_dim=2, sem_with_coord=False, sem_with_geo=False, mode=None):
        super().__init__()
        self.use_semantics = use_semantics
        self.N_samples, self.N_importance = N_samples, N_importance
        self.white_bkgd = white_bkgd
        self.chunk = ray_chunk            # accepted for compatibility; results never depended on it
        self.use_viewdirs = viewdirs
        self.mode = mode or os.environ.get("NSOS_MODE", "auto")
        kw = dict(input_dim=3, output_dim=4, skips=[4], viewdirs=viewdirs, use_embed=use_embed, multires=multires,
                  multires_views=multires_views, conv_embed=conv_embed, netchunk=pts_chuck, use_semantics=use_semantics,
                  sem_with_coord=sem_with_coord, sem_layer=sem_layer, sem_dim=sem_dim, sem_with_geo=sem_with_geo)
        self.nerf = NeRFMLP(net_depth=netdepth, net_width=netwidth, **kw)
        self.nerf_fine = self.nerf                                                    # nerf_net.py:49
        if N_importance > 0:
            self.nerf_fine = NeRFMLP(net_depth=netdepth_fine, net_width=netwidth_fine, **kw)
        self.render_kwargs_train = {'N_importance': N_importance, 'N_samples': N_samples, 'perturb': perturb,
                                    'raw_noise_std': raw_noise_std, 'retraw': True, 'retpts': False}
        self.render_kwargs_test = self.render_kwargs_train.copy()
        self.render_kwargs_test['perturb'] = 0.
        self.render_kwargs_test['raw_noise_std'] = 0.
        self._ws = None

    # ---- helpers -----------------------------------------------------------------------------------
    def _const_bound(self, value, like):
        key = (value, like.shape[0], like.device)
        cache = self.__dict__.setdefault("_bound_cache", {})
        t = cache.get(key)
        if t is None:
            if len(cache) >= 8:
                cache.clear()
            t = cache[key] = torch.full((like.shape[0], 1), value, dtype=torch.float32, device=like.device)
        return t

    def _workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        return self._ws

    def _status(self, device):
        st = self.__dict__.get("_status_word")
        if st is None or st.device != device:
            st = self.__dict__["_status_word"] = torch.zeros(1, dtype=torch.int32, device=device)
        return st

    def range_overflow(self) -> bool:
        """True if, since the last call, a tcgen05-mode render met a hidden activation beyond the fp16 range of its
        activation planes (|a| > 4094; the affected maps are non-finite).  Reads and clears the kernel's sticky status word
        (one device synchronisation) -- call it after an evaluation pass, not per step.  Such a net needs mode='simt'."""
        st = self.__dict__.get("_status_word")
        if st is None:
            return False
        v = int(st.item())
        st.zero_()
        return bool(v & 1)

    def capture_eval(self, n_rays: int, near, far, **kwargs) -> "CapturedEval":
        """One evaluation call -- pinned host rays in, the [N, 2*(6+sem_dim)+1] map rows out to pinned host memory -- captured as a
        CUDA graph: host-to-device copy, the fused render launch and the device-to-host copy replay as one submission (the per-call
        Python / launch overhead of a 4096-ray batch is ~10 % of the kernel).  Eval mode only (perturb = 0, no noise: nothing random
        is baked into the graph); the weights are read at replay time, re-capture after changing the parameters' storage.
        Not part of the reference's surface: forward() is the drop-in; this is the batch-serving variant."""
        return CapturedEval(self, n_rays, near, far, **kwargs)

    def resolve_mode(self, mode=None, n_samples=None, n_importance=None) -> int:
        mode = mode or self.mode
        if mode != "auto":
            return _lib.MODES[mode]
        L = _lib.lib()
        ns = self.N_samples if n_samples is None else n_samples
        ni = self.N_importance if n_importance is None else n_importance
        ok = (L.nsos_packed_bytes(self.nerf.desc(), _lib.MODE_TC_EXACT) > 0
              and L.nsos_packed_bytes(self.nerf_fine.desc(), _lib.MODE_TC_EXACT) > 0
              and 2 <= ns <= 128 and ns + ni <= 256)
        return _lib.MODE_TC_EXACT if ok else _lib.MODE_SIMT

    # ---- render_rays (nerf_net.py:71-130) ---------------------------------------------------------------
    def render_rays(self, rays_o, rays_d, near, far, viewdirs=None, raw_noise_std=0., verbose=False, retraw=False,
                    retpts=False, pytest=False, **kwargs):
        """viewdirs is accepted and ignored: the kernel forms d/|d| itself exactly as NeRFNet.forward does."""
        if retpts:
            raise NotImplementedError("retpts=True is never used by the reference callers")
        dev = rays_o.device
        if dev.type != "cuda":
            raise _lib.NsosError("nerfsos_b200.NeRFNet renders on CUDA only (no CPU fallback)")
        N = rays_o.shape[0]
        n_samples = kwargs.get('N_samples', self.N_samples)                     # sampler.py:41
        n_imp_gate = kwargs.get('N_importance', self.N_importance)            # nerf_net.py:104 (gate only)
        n_importance = self.N_importance if (self.N_importance > 0 and n_imp_gate > 0) else 0   # sampler.py:100
        perturb = kwargs.get('perturb', self.render_kwargs_train['perturb'])
        mode = self.resolve_mode(kwargs.get('mode'), n_samples, n_importance)
        cfg = _cfg(self, n_samples, n_importance, perturb, raw_noise_std, mode)
        f = lambda t: t.reshape(N, -1).to(dev, torch.float32).contiguous()
        rays_o, rays_d = f(rays_o), f(rays_d)
        near, far = f(near).reshape(N), f(far).reshape(N)
        rnd = {k: v.to(dev, torch.float32).contiguous() for k, v in (kwargs.get('randoms') or {}).items()}
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if (perturb > 0 or raw_noise_std > 0) else 0
        want = dict(raw=bool(retraw), z=bool(kwargs.get('retz', False)), grad=torch.is_grad_enabled())
        fine = n_importance > 0
        pc = list(self.nerf._flat.params)
        pf = list(self.nerf_fine._flat.params) if fine else []
        want["flat"] = (self.nerf.flat_params(), self.nerf_fine.flat_params())    # validated once per call, reused below
        if want["grad"] and any(p.requires_grad for p in pc + pf):
            tensors = _RenderFn.apply(self, cfg, rays_o, rays_d, near, far, rnd, seed, want, len(pc), *(pc + pf))
            out = dict(zip(want["names"], tensors))
        else:                                            # nothing to differentiate: skip the autograd node
            out, _ = _RenderFn.launch(self, cfg, rays_o, rays_d, near, far, rnd, seed, want, pc + pf)
        maps = out.pop("maps")
        sem = self.nerf.mlp.sem_dim if self.nerf.mlp.use_semantics else 0
        C6 = 6 + sem

        def block(b, sfx, ret):
            o = b * C6
            ret['rgb' + sfx] = maps[:, o:o + 3]
            ret['disp' + sfx] = maps[:, o + 3:o + 4]
            ret['acc' + sfx] = maps[:, o + 4:o + 5]
            ret['depth' + sfx] = maps[:, o + 5:o + 6]
            if sem:
                ret['semantics' + sfx] = maps[:, o + 6:o + 6 + sem]

        ret = {}
        block(0, '', ret)
        ret['weights'] = out['weights']
        if retraw:
            ret['raw'] = out['raw']
        if fine:
            ret['z_std'] = maps[:, 2 * C6]
            block(1, '0', ret)
            ret['weights0'] = out['weights0']
            if retraw:
                ret['raw0'] = out['raw0']
        if kwargs.get('retmaps', False):
            ret['maps'] = maps        # the kernel's single per-ray output row: [fine 6+sem | coarse 6+sem | z_std]
        if want["z"]:
            for k in ("z_vals", "z_vals0", "z_samples", "inds"):
                if k in out:
                    ret[k] = out[k]
        return ret

    # ---- forward (nerf_net.py:132-195) ----------------------------------------------------------------------
    def forward(self, ray_batch, bound_batch, **kwargs):
        render_kwargs = (self.render_kwargs_train if self.training else self.render_kwargs_test).copy()
        render_kwargs.update(kwargs)
        rays_o, rays_d = ray_batch
        assert rays_o.shape == rays_d.shape                                  # :155
        old_shape = rays_d.shape
        rays_o = torch.reshape(rays_o, [-1, rays_o.shape[-1]]).float()
        rays_d = torch.reshape(rays_d, [-1, rays_d.shape[-1]]).float()
        near, far = bound_batch
        if isinstance(near, (int, float)):                                   # :167-170; the [N,1] constants are cached: two fills
            near = self._const_bound(float(near), rays_d)                    # and two multiplies per call otherwise
        if isinstance(far, (int, float)):
            far = self._const_bound(float(far), rays_d)
        if rays_o.shape[0] == 0:
            raise ValueError("empty ray batch")                              # the reference fails in torch.cat here too
        # one fused launch instead of the serial ray_chunk loop (:177-188)
        ret = self.render_rays(rays_o, rays_d, near, far, **render_kwargs)
        for k in ret:                                                        # :191-193 unflatten
            ret[k] = torch.reshape(ret[k], list(old_shape[:-1]) + list(ret[k].shape[1:]))
        return ret


class CapturedEval:
    """See NeRFNet.capture_eval.  `rays_host` [2, N, 3] and `maps_host` [N, ML] are pinned buffers owned by this object:
    fill rays_host, call replay() (asynchronous on the current stream), synchronise, read maps_host."""

    def __init__(self, net: NeRFNet, n_rays: int, near, far, **kwargs):
        if net.training:
            raise ValueError("capture_eval: put the net in eval() mode (a captured graph would replay the same random draws)")
        dev = next(net.parameters()).device
        self.net, self.near, self.far = net, near, far
        kwargs = dict(kwargs, retraw=False, retmaps=True)
        self.rays_host = torch.zeros(2, n_rays, 3).pin_memory()
        self.rays_host[1, :, 2] = -1.0                                       # a harmless direction for the warm-up calls
        self._rays = torch.empty(2, n_rays, 3, device=dev)
        with torch.no_grad():
            stream = torch.cuda.Stream(device=dev)
            stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(stream):                                  # warm-up on a side stream: workspaces, packed weights, attributes
                for _ in range(2):
                    self._rays.copy_(self.rays_host, non_blocking=True)
                    out = net(self._rays, (near, far), **kwargs)
            torch.cuda.current_stream(dev).wait_stream(stream)
            self.maps_host = torch.empty(out["maps"].shape).pin_memory()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):     # other threads (NCCL watchdog, data loaders) may call CUDA
                self._rays.copy_(self.rays_host, non_blocking=True)
                out = net(self._rays, (near, far), **kwargs)
                self.maps_host.copy_(out["maps"], non_blocking=True)
            self._maps = out["maps"]                                         # keeps the graph's output alive

    def replay(self):
        self.graph.replay()
        return self.maps_host
