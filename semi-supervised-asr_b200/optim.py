"""Fused global-norm clip + Adam / AMSGrad over flat f32 buffers (las_grad_norm + las_adam_step).

Replaces `torch.nn.utils.clip_grad_norm_(params, max_norm)` followed by `torch.optim.Adam.step()`
(solver.py:152-153 amsgrad + wd 1e-6, 171-173 plain, 296-297, 384-385, 488-489). The parameters of
the module are re-pointed at views of ONE flat buffer (so are their .grad), which makes the whole
update three launches and the data-parallel gradient exchange one NCCL all-reduce.
`state_dict()` / `load_state_dict()` speak torch.optim.Adam's format so reference `.opt`
checkpoints load unchanged.
"""
import torch

from ._lib import call, ptr
from .functional import advance_dropout_seed


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        self.params = [p for p in params]
        assert self.params, "no parameters"
        dev = self.params[0].device
        assert dev.type == "cuda", "FusedAdam needs CUDA parameters (there is no CPU path)"
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)]
        sizes = [p.numel() for p in self.params]
        self.offsets = [0]
        for n in sizes:
            self.offsets.append(self.offsets[-1] + (n + 3) // 4 * 4)      # 16-byte aligned views
        total = self.offsets[-1]
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        for p, o in zip(self.params, self.offsets):
            v = self.flat[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.max_exp_avg_sq = torch.zeros_like(self.flat) if amsgrad else None
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.norm_dev = torch.zeros(1, device=dev, dtype=torch.float32)
        self._partials = torch.zeros(256, device=dev, dtype=torch.float64)

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()
        for p, o in zip(self.params, self.offsets):      # re-attach in case somebody set .grad = None
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)

    def clip_and_step(self, max_norm, grad_scale=1.0):
        """clip_grad_norm_(max_norm) + step in three launches. Returns the (device) pre-clip norm."""
        g = self.param_groups[0]
        n = self.flat.numel()
        call("las_grad_norm", ptr(self.flat_grad), n, ptr(self._partials), ptr(self.norm_dev))
        self.step_dev += 1
        call("las_adam_step", ptr(self.flat), ptr(self.flat_grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
             ptr(self.max_exp_avg_sq), n, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
             float(g["weight_decay"]), ptr(self.step_dev), float(max_norm if max_norm else 0.0), ptr(self.norm_dev),
             float(grad_scale))
        advance_dropout_seed()                         # the next forward draws new dropout masks
        return self.norm_dev

    def step(self):
        return self.clip_and_step(0.0)

    # ---- torch.optim.Adam-compatible checkpoint format (solver.py:38-41, 55-61)
    def state_dict(self):
        state = {}
        step = int(self.step_dev.item())
        if step > 0:
            for i, (p, o) in enumerate(zip(self.params, self.offsets)):
                sl = slice(o, o + p.numel())
                st = {"step": torch.tensor(float(step)), "exp_avg": self.exp_avg[sl].view(p.shape).clone(),
                      "exp_avg_sq": self.exp_avg_sq[sl].view(p.shape).clone()}
                if self.max_exp_avg_sq is not None:
                    st["max_exp_avg_sq"] = self.max_exp_avg_sq[sl].view(p.shape).clone()
                state[i] = st
        g = dict(self.param_groups[0])
        g["params"] = list(range(len(self.params)))
        return {"state": state, "param_groups": [g]}

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in g:
                self.param_groups[0][k] = g[k]
        step = 0
        for i, st in sd["state"].items():
            i = int(i)
            p, o = self.params[i], self.offsets[i]
            sl = slice(o, o + p.numel())
            self.exp_avg[sl].copy_(st["exp_avg"].flatten())
            self.exp_avg_sq[sl].copy_(st["exp_avg_sq"].flatten())
            if self.max_exp_avg_sq is not None and "max_exp_avg_sq" in st:
                self.max_exp_avg_sq[sl].copy_(st["max_exp_avg_sq"].flatten())
            step = int(float(st["step"]))
        self.step_dev.fill_(step)
