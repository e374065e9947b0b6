"""Parameter containers mirroring models/nerf_mlp.py of the reference (same attribute names, shapes,
default initialisation and state_dict keys), with the arithmetic delegated to libnerfsos.so.

    MLP      <-> models/nerf_mlp.py:24-100    (ctor :24-65)
    NeRFMLP  <-> models/nerf_mlp.py:132-215   (forward :179 == nsos_mlp_query, used by export_density,
                                               engines/eval.py:297)

Only the configurations that NeRFNet can reach from the reference CLI are accepted: sem_layer <= 2,
sem_with_geo=False, conv_embed=False, use_embed=True (anything else raises NotImplementedError).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from .. import _lib


class MLP(nn.Module):
    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False,
                 use_semantics=True, sem_layer=2, sem_dim=2, sem_with_coord=False, sem_with_geo=False):
        super().__init__()
        if sem_with_geo:
            raise NotImplementedError("sem_with_geo is not wired through NeRFNet by the reference CLI (run_nerf.py:301-305)")
        if use_semantics and sem_layer > 2:
            raise NotImplementedError("sem_layer > 2 is not implemented")
        if list(skips) not in ([4], []):
            raise NotImplementedError("only skips=[4] (NeRFNet's hard-wired value) is implemented")
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.skips = list(skips)
        self.use_viewdirs = use_viewdirs
        self.use_semantics = use_semantics and use_viewdirs   # the reference emits no semantics without viewdirs (:96-98)
        self.sem_with_coord = sem_with_coord
        self.sem_dim = sem_dim
        # same construction order as the reference so that torch.manual_seed(s) gives identical weights
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] + [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W)
                                        for i in range(D - 1)])
        if use_viewdirs:
            self.alpha_linear = nn.Linear(W, 1)
            self.feature_linear = nn.Linear(W, W)
            self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
            self.rgb_linear = nn.Linear(W // 2, output_ch - 1)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        if use_semantics:
            sem_in_dim = W + input_ch if sem_with_coord else W
            self.semantic_linear = nn.Sequential(nn.Linear(sem_in_dim, W // 2), nn.ReLU(), nn.Linear(W // 2, sem_dim))
            self.geo_map_sem = None

    def ordered_params(self):
        """Parameters in the flat-buffer order of nsos_param_layout (== state_dict order)."""
        ps = []
        for l in self.pts_linears:
            ps += [l.weight, l.bias]
        if self.use_viewdirs:
            for l in (self.alpha_linear, self.feature_linear, self.views_linears[0], self.rgb_linear):
                ps += [l.weight, l.bias]
            if self.use_semantics:
                ps += [self.semantic_linear[0].weight, self.semantic_linear[0].bias,
                       self.semantic_linear[2].weight, self.semantic_linear[2].bias]
        else:
            ps += [self.output_linear.weight, self.output_linear.bias]
        return ps


class FlatParams:
    """One contiguous fp32 device buffer per net; every nn.Parameter is a view into it, so the kernels
    (and the NCCL gradient all-reduce) see a single flat array and optimiser updates need no gather."""

    def __init__(self, params):
        self.params = list(params)
        self.flat = None
        self.offsets = []

    def ensure(self):
        dev = self.params[0].device
        if dev.type != "cuda":
            raise _lib.NsosError("nerfsos_b200 runs on CUDA only (no CPU fallback): move the model with .cuda()")
        ok = self.flat is not None and self.flat.device == dev
        if ok:
            base = self.flat.data_ptr()
            for p, off in zip(self.params, self.offsets):
                if p.data_ptr() != base + 4 * off or p.dtype != torch.float32:
                    ok = False
                    break
        if ok:
            return self.flat
        total = sum(p.numel() for p in self.params)
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        offs, off = [], 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                flat[off:off + n].copy_(p.detach().reshape(-1).float())
                p.data = flat[off:off + n].view(p.shape)
                offs.append(off)
                off += n
        self.flat, self.offsets = flat, offs
        return flat

    def version(self):
        return tuple(p._version for p in self.params) + (self.flat.data_ptr() if self.flat is not None else 0,)


class NeRFMLP(nn.Module):
    def __init__(self, input_dim=3, output_dim=4, net_depth=8, net_width=256, skips=[4], viewdirs=True, use_embed=True,
                 multires=10, multires_views=4, conv_embed=False, netchunk=1024 * 64, use_semantics=False, sem_layer=2,
                 sem_dim=2, sem_with_coord=False, sem_with_geo=False):
        super().__init__()
        if not use_embed or conv_embed or input_dim != 3 or output_dim != 4:
            raise NotImplementedError("only use_embed=True, conv_embed=False, input_dim=3, output_dim=4 are implemented")
        self.chunk = netchunk
        self.multires, self.multires_views = multires, multires_views
        input_ch = 3 + 6 * multires
        input_ch_views = (3 + 6 * multires_views) if viewdirs else 0
        self.mlp = MLP(net_depth, net_width, skips=skips, input_ch=input_ch, output_ch=output_dim,
                       input_ch_views=input_ch_views, use_viewdirs=viewdirs, use_semantics=use_semantics,
                       sem_layer=sem_layer, sem_dim=sem_dim, sem_with_coord=sem_with_coord, sem_with_geo=sem_with_geo)
        self._flat = FlatParams(self.mlp.ordered_params())
        self._packed = {}   # mode -> (version, tensor)

    # ---- descriptors / buffers -------------------------------------------------------------------
    def desc(self) -> _lib.NetDesc:
        d = self.__dict__.get("_desc")
        if d is None:                                    # the geometry is fixed at construction: build the struct once
            m = self.mlp
            d = self.__dict__["_desc"] = _lib.NetDesc(m.D, m.W, 4 if m.skips else -1, self.multires, self.multires_views,
                                                      int(m.use_viewdirs), int(m.use_semantics),
                                                      m.sem_dim if m.use_semantics else 0, int(m.sem_with_coord))
        return d

    def flat_params(self) -> torch.Tensor:
        flat = self._flat.ensure()
        if self.__dict__.get("_count_ok") != flat.numel():
            n = _lib.lib().nsos_param_count(self.desc())
            if n != flat.numel():
                raise _lib.NsosError(f"parameter layout mismatch: library expects {n} floats, module holds {flat.numel()}")
            self.__dict__["_count_ok"] = n
        return flat

    def packed(self, mode: int, force: bool = False, flat: torch.Tensor = None):
        """fp16 hi/lo swizzled weight image for the tcgen05 kernel; re-packed when the parameters changed.
        `flat`: the result of a flat_params() call made earlier in the same forward (skips re-validating the views)."""
        if mode == _lib.MODE_SIMT:
            return None
        if flat is None:
            flat = self.flat_params()
        ver = self._flat.version()
        hit = self._packed.get(mode)
        if hit is not None and hit[0] == ver and not force:
            return hit[1]
        L = _lib.lib()
        d = self.desc()
        nbytes = L.nsos_packed_bytes(d, mode)
        if nbytes == 0:
            raise _lib.NsosError("this net geometry is not covered by the tcgen05 path (use mode='simt')")
        buf = hit[1] if hit is not None and hit[1].numel() == nbytes and hit[1].device == flat.device else \
            torch.empty(nbytes, dtype=torch.uint8, device=flat.device)
        with torch.cuda.device(flat.device):        # the library launches on the calling thread's current device
            _lib.check(L.nsos_pack_weights(d, _lib.ptr(flat), _lib.ptr(buf), mode, _lib.cur_stream(flat.device)), "nsos_pack_weights")
        self._packed[mode] = (ver, buf)
        return buf

    # ---- the same query for points that share ONE view direction, on the tensor cores --------------
    @torch.no_grad()
    def query_dir(self, inputs, viewdir=(0.0, 0.0, 0.0), mode: int = _lib.MODE_TC_EXACT):
        """forward(inputs, viewdirs=viewdir.expand_as(inputs)) through the tcgen05 render kernel's replay mode
        (nsos_mlp_query_dir): what export_density (engines/eval.py:290-297) asks for -- a grid of points with viewdirs = 0.
        `viewdir` is used as given (the caller normalises, as for forward).  Raises NsosError for net geometries the
        tcgen05 path does not cover (use forward then)."""
        if not self.mlp.use_viewdirs:
            raise _lib.NsosError("query_dir: the net takes no view direction")
        flat = self.flat_params()
        pk = self.packed(mode, flat=flat)
        sh = inputs.shape
        pts = inputs.reshape(-1, 3).to(flat.device, torch.float32).contiguous()
        n = pts.shape[0]
        Cc = 4 + (self.mlp.sem_dim if self.mlp.use_semantics else 0)
        out = torch.empty(n, Cc, dtype=torch.float32, device=flat.device)
        vd = (C.c_float * 3)(*[float(v) for v in viewdir])
        with torch.cuda.device(flat.device):
            _lib.check(_lib.lib().nsos_mlp_query_dir(self.desc(), _lib.ptr(pk), _lib.ptr(pts), vd, _lib.ptr(out), mode, n,
                                                     _lib.cur_stream(flat.device)), "nsos_mlp_query_dir")
        return out.reshape(*sh[:-1], Cc)

    # ---- NeRFMLP.forward (nerf_mlp.py:179-215): raw network query ---------------------------------
    @torch.no_grad()
    def forward(self, inputs, viewdirs=None):
        flat = self.flat_params()
        sh = inputs.shape
        pts = inputs.reshape(-1, 3).to(flat.device, torch.float32).contiguous()
        vd = None
        if self.mlp.use_viewdirs:
            if viewdirs is None:
                raise ValueError("viewdirs required")
            vd = viewdirs.reshape(-1, 3).to(flat.device, torch.float32).contiguous()
        n = pts.shape[0]
        C = 4 + (self.mlp.sem_dim if self.mlp.use_semantics else 0)
        out = torch.empty(n, C, dtype=torch.float32, device=flat.device)
        L = _lib.lib()
        d = self.desc()
        step = 1 << 18
        ws = torch.empty(L.nsos_mlp_workspace_bytes(d, min(n, step)), dtype=torch.uint8, device=flat.device)
        for i in range(0, n, step):
            m = min(step, n - i)
            with torch.cuda.device(flat.device):
                _lib.check(L.nsos_mlp_query(d, _lib.ptr(flat), _lib.ptr(pts[i:i + m]), _lib.ptr(vd[i:i + m]) if vd is not None else None,
                                            _lib.ptr(out[i:i + m]), _lib.ptr(ws), ws.numel(), m, _lib.cur_stream(flat.device)),
                           "nsos_mlp_query")
        return out.reshape(*sh[:-1], C)
