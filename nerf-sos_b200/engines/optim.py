"""Drop-in for the optimiser of the reference's training loop (run_nerf.py:320: torch.optim.Adam(grad_vars, lr, betas=(0.9, 0.999)))
that updates every trainable tensor with ONE launch of libnerfsos.so's multi-tensor Adam kernel.  Same constructor, `param_groups`
(engines/lr.py writes group['lr'] every step), `zero_grad`, `state_dict` / `load_state_dict` with torch.optim.Adam's state names
(`step`, `exp_avg`, `exp_avg_sq`), so checkpoints written by either optimiser load into the other (trainer.py:216-226)."""
from __future__ import annotations

import torch

from .. import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("FusedAdam covers the reference's configuration only: weight_decay=0, amsgrad=False")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("closures are not used by the reference trainer")
        L = _lib.lib()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            if dev.type != "cuda":
                raise _lib.NsosError("FusedAdam updates CUDA tensors only (no CPU fallback)")
            arr = (_lib.AdamTensor * len(ps))()
            keep = []
            step = None
            for i, p in enumerate(ps):
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32)          # host scalar, like torch's capturable=False
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                s = int(st["step"])
                step = s if step is None else step
                if s != step:
                    raise _lib.NsosError("FusedAdam: parameters of one group must share the step count")
                g = p.grad if p.grad.is_contiguous() and p.grad.dtype == torch.float32 else p.grad.float().contiguous()
                if not p.is_contiguous() or p.dtype != torch.float32:
                    raise _lib.NsosError("FusedAdam: parameters must be contiguous fp32")
                keep.append(g)
                arr[i] = _lib.AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel())
            b1, b2 = group["betas"]
            with torch.cuda.device(dev):
                _lib.check(L.nsos_adam_multi(arr, len(ps), float(group["lr"]), float(b1), float(b2), float(group["eps"]), step,
                                             _lib.cur_stream(dev)), "nsos_adam_multi")
            for p in ps:                                          # the kernel wrote through raw pointers: tell autograd / the
                torch.autograd.graph.increment_version(p)        # weight-pack cache (FlatParams.version) that p changed
        return None
