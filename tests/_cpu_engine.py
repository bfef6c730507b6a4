"""CPU stand-in for Eagle3Engine used by the host-logic tests (no CUDA library calls): the same flat-buffer attributes and
the arithmetic of sf_optimizer_step restated with torch ops (bf16 gradient buffer -> grad norm -> clip -> AdamW on fp32
masters -> bf16 write-back; optimizer.py:95-168).  Test infrastructure only."""
import math

import torch


class CpuFlatEngine:
    def __init__(self, shapes: dict, seed: int = 0):
        self.device = torch.device("cpu")
        self.shapes = dict(shapes)
        self.offsets, self.sizes, o = {}, {}, 0
        for n, shp in shapes.items():
            self.offsets[n], self.sizes[n] = o, math.prod(shp)
            o += self.sizes[n]
        self.n_params = o
        g = torch.Generator().manual_seed(seed)
        self.params = (torch.randn(o, generator=g) * 0.05).to(torch.bfloat16)
        self.grads_f32 = torch.zeros(o)
        self.grads_bf16 = torch.zeros(o, dtype=torch.bfloat16)
        self.master = self.exp_avg = self.exp_avg_sq = None
        self.opt_step = 0

    def param_view(self, name, buf=None):
        buf = self.params if buf is None else buf
        o, n = self.offsets[name], self.sizes[name]
        return buf[o:o + n].view(self.shapes[name])

    def ensure_optimizer_state(self):
        if self.master is None:
            self.master = self.params.float()
            self.exp_avg = torch.zeros_like(self.master)
            self.exp_avg_sq = torch.zeros_like(self.master)

    def grads_to_bf16(self, scale=None, first=0, count=None):
        s = 1.0 if scale is None else float(scale)
        count = self.n_params - first if count is None else count
        self.grads_bf16[first:first + count].copy_((self.grads_f32[first:first + count] * s).to(torch.bfloat16))
        return self.grads_bf16

    def optimizer_step(self, lr, *, grad_scale=1.0, max_grad_norm=0.5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.ensure_optimizer_state()
        self.opt_step += 1
        g = (self.grads_bf16.float() * grad_scale).to(torch.bfloat16).float()
        norm = g.norm()
        g = g * torch.clamp(max_grad_norm / (norm + 1e-6), max=1.0)
        b1, b2 = betas
        self.master.mul_(1 - lr * weight_decay)
        self.exp_avg.lerp_(g, 1 - b1)
        self.exp_avg_sq.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1, bc2 = 1 - b1 ** self.opt_step, 1 - b2 ** self.opt_step
        self.master.addcdiv_(self.exp_avg, self.exp_avg_sq.sqrt() / math.sqrt(bc2) + eps, value=-lr / bc1)
        self.params.copy_(self.master.to(torch.bfloat16))
        return norm.reshape(1)
