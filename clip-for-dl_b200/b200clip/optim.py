"""AdamW for the head parameters on the device: torch.optim.AdamW semantics, one elementwise
kernel per parameter tensor, step counter on the device so that `step()` can be captured in a CUDA graph together with the
head step (the reference builds torch.optim.AdamW(lr=1e-4, weight_decay=0.01) over the same parameters and checkpoints
`optimizer.state_dict()`, 0426/train.py:846-860, :670).

State layout = torch.optim.AdamW's: state[p] = {"step": 0-dim float32 tensor, "exp_avg", "exp_avg_sq"}, so a checkpoint written
by either optimizer resumes in the other.  All parameters share ONE device counter (every state[p]["step"] is a view of it);
load_state_dict() re-creates the sharing from the loaded values."""
from __future__ import annotations

import torch

from ._lib import check, load, ptr, require_cuda, stream_ptr


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._step = None

    def _counter(self, device) -> torch.Tensor:
        """The shared device step counter; adopts the value of already-present per-parameter `step` entries (resume)."""
        if self._step is None:
            steps = [float(st["step"]) for st in self.state.values() if "step" in st]
            if steps and max(steps) != min(steps):
                raise RuntimeError("b200clip.FusedAdamW: parameters with different step counts are not supported "
                                   f"(found {min(steps)} .. {max(steps)})")
            self._step = torch.full((), steps[0] if steps else 0.0, dtype=torch.float32, device=device)
            for st in self.state.values():
                if "step" in st:
                    st["step"] = self._step
        return self._step

    def state_dict(self):
        """torch.optim.AdamW's layout: every parameter gets its OWN `step` tensor (a clone of the shared device counter) -- torch's
        foreach implementation increments each state's tensor, so a shared one would be advanced once per parameter."""
        sd = super().state_dict()
        sd["state"] = {k: ({**v, "step": v["step"].detach().clone()} if "step" in v else dict(v)) for k, v in sd["state"].items()}
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._step = None                       # rebuilt from the loaded per-parameter steps on the next step()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = load()
        todo = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                require_cuda(p, p.grad)
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("b200clip.FusedAdamW: parameters must be contiguous float32 CUDA tensors")
                step = self._counter(p.device)
                st = self.state[p]
                if "exp_avg" not in st:
                    st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(p), torch.zeros_like(p)
                st["step"] = step
                todo.append((group, p, st))
        if not todo:
            return loss
        check(lib.b200clip_adamw_tick(ptr(self._step), stream_ptr()), "adamw_tick")
        for group, p, st in todo:
            b1, b2 = group["betas"]
            g = p.grad if (p.grad.dtype == torch.float32 and p.grad.is_contiguous()) else p.grad.float().contiguous()
            check(lib.b200clip_adamw_step(ptr(p), ptr(g), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(), float(group["lr"]),
                                          float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                          ptr(self._step), stream_ptr()), "adamw_step")
        return loss
