"""AdamW for the head parameters on the device: torch.optim.AdamW semantics, one elementwise
kernel per parameter tensor, step counter on the device so that `step()` can be captured in a CUDA graph together with the
head step (the reference builds torch.optim.AdamW(lr=1e-4, weight_decay=0.01) over the same parameters)."""
from __future__ import annotations

import torch

from ._lib import check, load, ptr, require_cuda, stream_ptr


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._step = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = load()
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                require_cuda(p, p.grad)
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("b200clip.FusedAdamW: parameters must be contiguous float32 CUDA tensors")
                if self._step is None:
                    self._step = torch.zeros((), dtype=torch.float32, device=p.device)
                st = self.state[p]
                if not st:
                    st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(p), torch.zeros_like(p)
        if self._step is None:
            return loss
        check(lib.b200clip_adamw_tick(ptr(self._step), stream_ptr()), "adamw_tick")
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                g = p.grad if (p.grad.dtype == torch.float32 and p.grad.is_contiguous()) else p.grad.float().contiguous()
                check(lib.b200clip_adamw_step(ptr(p), ptr(g), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(), float(group["lr"]),
                                              float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                              ptr(self._step), stream_ptr()), "adamw_step")
        return loss
