"""Fused multi-tensor SGD (momentum + weight decay) — one kernel launch for all 247 parameter tensors.

Same update rule and constructor as `torch.optim.SGD(params, lr, momentum, weight_decay)` used at
main.py:132-137 (dampening 0, no nesterov): d = g + wd*p; buf = mu*buf + d (buf starts at zero, i.e. buf = d on a parameter's first step); p -= lr*buf.
`state_dict()` uses torch's layout (`momentum_buffer` per parameter), so checkpoints are interchangeable.
"""
import ctypes

import torch

from . import _lib


class SGD(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0, weight_decay=0.0, nesterov=False):
        if dampening != 0 or nesterov:
            raise NotImplementedError("selavi_b200.optim.SGD implements the reference's configuration (no dampening/nesterov)")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, dampening=0, nesterov=False))
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.lib()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            # a missing momentum buffer is created as zeros: mu*0 + d == d is torch's first-step rule, per parameter
            # (a per-group "first step" flag would reset the momentum of every other tensor when one parameter
            # receives its first gradient late: find_unused_parameters, add_param_group, partially loaded state)
            for p in ps:
                st = self.state[p]
                if "momentum_buffer" not in st or st["momentum_buffer"] is None:
                    st["momentum_buffer"] = torch.zeros_like(p)
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                    raise ValueError("selavi_b200.optim.SGD needs contiguous fp32 CUDA parameters and gradients")
            n = len(ps)
            arr = ctypes.c_void_p * n
            params = arr(*[p.data_ptr() for p in ps])
            grads = arr(*[p.grad.data_ptr() for p in ps])
            bufs = arr(*[self.state[p]["momentum_buffer"].data_ptr() for p in ps])
            sizes = (ctypes.c_longlong * n)(*[p.numel() for p in ps])
            with torch.cuda.device(ps[0].device):
                _lib.check(lib.selavi_sgd_step_host(params, grads, bufs, sizes, n, float(group["lr"]), float(group["momentum"]),
                                                    float(group["weight_decay"]), 0, _lib.stream_ptr()),
                           "selavi_sgd_step_host")
            # the kernel wrote the parameters behind autograd's back: bump their version counters so that
            # version-keyed caches (the packed tensor-core weights in engine.py) see the update
            torch.autograd.graph.increment_version(ps)
        return loss
