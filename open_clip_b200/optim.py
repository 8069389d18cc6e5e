"""NativeAdamW — drop-in for the `torch.optim.AdamW` the reference's optimizer factory builds
(open_clip_train/optim.py:336-454: decoupled weight decay, parameter groups with their own lr / weight_decay): same
constructor, same `param_groups` / `state` / `state_dict()` layout (so LR schedulers — open_clip_train/scheduler.py
assigns `param_group["lr"]` — and checkpoints keep working), but `step()` is ONE launch of libclipn's multi-tensor
kernel over every parameter of the model instead of torch's per-dtype multi-tensor launches.

Arithmetic: torch's fused AdamW (fp32 math, one rounding per stored element, bf16 moments for bf16 parameters).
CUDA tensors only; amsgrad / maximize / capturable / sparse gradients are not supported (the reference's CLIP training
path uses none of them) and raise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, Tuple

import torch

from . import _lib as L
from . import ops


class NativeAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, amsgrad: bool = False, maximize: bool = False, **unused):
        if amsgrad or maximize:
            raise L.ClipnError("NativeAdamW: amsgrad / maximize are not supported")
        if not 0.0 <= lr or not 0.0 <= eps or not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError("NativeAdamW: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        groups = {}
        for gi, group in enumerate(self.param_groups):
            beta1, beta2 = group["betas"]
            entries = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise L.ClipnError("NativeAdamW does not support sparse gradients")
                if not p.is_cuda:
                    raise L.ClipnError("NativeAdamW runs on CUDA tensors only; there is no CPU fallback")
                if p.dtype not in (torch.bfloat16, torch.float32):
                    raise L.ClipnError(f"NativeAdamW: unsupported parameter dtype {p.dtype}")
                st = self.state[p]
                if len(st) == 0:  # same state layout as torch.optim.AdamW
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad
                if g.dtype != p.dtype or not g.is_contiguous():
                    g = g.to(p.dtype).contiguous()
                if not p.is_contiguous():
                    raise L.ClipnError("NativeAdamW: parameters must be contiguous")
                entries.append((p, g, st))
            if not entries:
                continue
            # parameters of one group that were created together share the step count; group by (betas, eps, step)
            for p, g, st in entries:
                key = (beta1, beta2, group["eps"], int(st["step"]))
                groups.setdefault(key, []).append((p, g, st, float(group["lr"]), float(group["weight_decay"])))
        for (beta1, beta2, eps, step), items in groups.items():
            arr = (L.AdamTensor * len(items))()
            for i, (p, g, st, lr, wd) in enumerate(items):
                a = arr[i]
                a.param, a.grad = p.data_ptr(), g.data_ptr()
                a.exp_avg, a.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                a.numel, a.lr, a.weight_decay = p.numel(), lr, wd
                a.is_bf16 = int(p.dtype == torch.bfloat16)
            bc1 = 1.0 - beta1 ** step
            bc2_sqrt = math.sqrt(1.0 - beta2 ** step)
            ops._call(L.lib().clipn_adamw_multi(arr, len(items), beta1, beta2, eps, bc1, bc2_sqrt, ops._stream()),
                      (len(items) + 511) // 512)
        return loss
