"""B200BF16Optimizer — the reference's BF16Optimizer (specforge/optimizer.py:12-229) with the arithmetic moved into
`sf_optimizer_step` (one fused clip + AdamW + bf16 write-back kernel over the flat buffers of the engine).

Same constructor arguments, `step()`, `get_learning_rate()`, `state_dict()` / `load_state_dict()`.  The state dict has the
REFERENCE layout (optimizer.py:221-229), so a checkpoint written here resumes under the reference optimizer and vice versa:

    {"optimizer_state_dict": torch.optim.AdamW layout — state[i] = {step, exp_avg, exp_avg_sq} for the i-th trainable parameter
                             in the reference module's parameter order, one param group,
     "scheduler_state_dict": the reference warm-up scheduler's state (lr_scheduler.py:56-147),
     "lr_scheduler_type", "max_grad_norm", "fp32_params": [fp32 master per parameter, on the CPU]}

`load_state_dict` validates `lr_scheduler_type` and `max_grad_norm` like the reference (optimizer.py:175-190).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

# trainable parameters in the order LlamaForCausalLMEagle3.parameters() yields them (modeling/draft/llama3_eagle.py:1653-1703)
REFERENCE_PARAM_ORDER = [
    "midlayer.self_attn.q_proj.weight", "midlayer.self_attn.k_proj.weight", "midlayer.self_attn.v_proj.weight",
    "midlayer.self_attn.o_proj.weight", "midlayer.mlp.gate_proj.weight", "midlayer.mlp.up_proj.weight",
    "midlayer.mlp.down_proj.weight", "midlayer.hidden_norm.weight", "midlayer.input_layernorm.weight",
    "midlayer.post_attention_layernorm.weight", "fc.weight", "fc_norm.0.weight", "fc_norm.1.weight", "fc_norm.2.weight",
    "norm.weight", "lm_head.weight",
]


class WarmupSchedule:
    """The reference's CosineAnnealingWarmupLR / ConstantWarmupLR (lr_scheduler.py:56-147) as a closed form of the number of
    optimizer steps taken, plus the counters its `state_dict()` carries, so that the two are interchangeable in a checkpoint.

    lr while n < warmup: base * (n + 1) / warmup; afterwards constant, or cosine over T_max = total - warmup with t = n - warmup.
    The cosine phase reproduces what the reference actually yields: its nested torch CosineAnnealingLR is first evaluated
    through the chainable (recursive) form at last_epoch 0, which scales the whole phase by 2 / (1 + cos(pi / T_max))
    (1 + 4e-12 at the default 800 000 steps, 0.3 % at 30): lr_t = eta_min + (base - eta_min) (1 + cos(pi t / T)) / (1 + cos(pi / T))."""

    def __init__(self, base_lr: float, total_steps: int, warmup_steps: int, kind: str = "cosine", eta_min: float = 0.0):
        if kind not in ("cosine", "constant"):
            raise ValueError(f"unsupported lr_scheduler={kind!r}; expected one of ['constant', 'cosine']")
        if kind == "constant":
            if total_steps <= 0:
                raise ValueError(f"total_steps must be positive, got {total_steps}")
            if not 0 <= warmup_steps < total_steps:
                raise ValueError(f"warmup_steps must be in [0, total_steps), got {warmup_steps} for total_steps={total_steps}")
        self.base_lr, self.total_steps, self.warmup, self.kind, self.eta_min = float(base_lr), int(total_steps), int(warmup_steps), kind, eta_min
        self.n = 0          # scheduler.step() calls so far == optimizer steps taken

    def lr_at(self, n: int) -> float:
        if n < self.warmup:
            return (n + 1) / self.warmup * self.base_lr
        if self.kind == "constant":
            return self.base_lr
        t, tmax = n - self.warmup, self.total_steps - self.warmup
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * t / tmax)) / (1 + math.cos(math.pi / tmax))

    @property
    def lr(self) -> float:
        return self.lr_at(self.n)

    def step(self) -> None:
        self.n += 1

    # ---- the reference scheduler's state dict -------------------------------------------------------------------
    def state_dict(self) -> dict:
        n, w = self.n, self.warmup
        finished = n >= w                       # get_lr() flips it the first time last_epoch reaches warmup_epochs
        outer_epoch = min(n, w)                 # once finished only the after-scheduler advances (_WarmupScheduler.step)
        after_epoch = max(0, n - w)
        after = {"base_lrs": [self.base_lr], "last_epoch": after_epoch, "_step_count": after_epoch + 1, "_is_initial": False,
                 "_get_lr_called_within_step": False,
                 "_last_lr": [self.lr_at(w + after_epoch) if finished or after_epoch else self.base_lr]}
        if self.kind == "cosine":
            after = {"T_max": self.total_steps - w, "eta_min": self.eta_min, **after}
        return {"warmup_epochs": w, "finished": finished, "base_lrs": [self.base_lr], "last_epoch": outer_epoch,
                "_step_count": outer_epoch + 1, "_is_initial": False, "_get_lr_called_within_step": False, "_last_lr": [self.lr],
                "after_scheduler_type": "CosineAnnealingLR" if self.kind == "cosine" else "_FlatLR", "after_scheduler_dict": after}

    def load_state_dict(self, state: dict) -> None:
        w = int(state.get("warmup_epochs", self.warmup))
        if w != self.warmup:
            raise ValueError(f"checkpoint scheduler has warmup_epochs={w} but this run has {self.warmup}")
        after = state.get("after_scheduler_dict") or {}
        if self.kind == "cosine" and "T_max" in after and int(after["T_max"]) != self.total_steps - self.warmup:
            raise ValueError(f"checkpoint scheduler has T_max={after['T_max']} but this run has {self.total_steps - self.warmup}")
        if state.get("finished"):
            self.n = w + int(after.get("last_epoch", 0))
        else:
            self.n = int(state.get("last_epoch", 0))


def _adamw_param_group_template() -> dict:
    """The keys torch.optim.AdamW puts in a param group on THIS torch version (so a reference AdamW accepts the dict)."""
    opt = torch.optim.AdamW([torch.zeros(1, requires_grad=True)], lr=1.0)
    return dict(opt.state_dict()["param_groups"][0])


class B200BF16Optimizer:
    def __init__(self, model, lr, weight_decay=0.0, max_grad_norm=0.5, total_steps=800_000, warmup_ratio=0.015,
                 lr_scheduler="cosine", offload_master=False):
        engine = getattr(model, "engine", None)
        if engine is None and hasattr(model, "draft_model"):
            engine = getattr(model.draft_model, "engine", None)
        if engine is None:
            raise TypeError("B200BF16Optimizer needs the B200 draft module (bind_engine() first): it updates the engine's flat buffers")
        if offload_master:
            raise NotImplementedError("optimizer_cpu_offload: the fused step keeps the fp32 masters in HBM (4.8 GB for an 8B draft)")
        self.model, self.engine = model, engine
        self.max_grad_norm = max_grad_norm
        self.weight_decay = weight_decay
        self.lr_scheduler_type = lr_scheduler
        self.scheduler = WarmupSchedule(lr, total_steps, int(warmup_ratio * total_steps), lr_scheduler)
        self.last_grad_norm: Optional[torch.Tensor] = None
        # per-parameter state is listed in the order the reference module yields its trainable parameters; engines of other
        # draft families (DFlash) fall back to their own flat-buffer order
        self.names: List[str] = [n for n in REFERENCE_PARAM_ORDER if n in engine.offsets] or list(getattr(engine, "names", engine.offsets))
        self.grad_scale = 1.0          # 1 / world for data-parallel averaging (set by the backend)

    # ---- reference surface ----------------------------------------------------------------------------------------
    def configure_grad_norm_reduction(self, *, process_group=None, enabled: bool = True) -> None:
        """Replicated (DDP-style) gradients: every rank holds the full, already all-reduced gradient, so the norm needs no
        collective (the reference disables it for NO_SHARD too, training/backend.py:299-308)."""

    def get_learning_rate(self) -> float:
        return self.scheduler.lr

    def step(self) -> torch.Tensor:
        """Clip + AdamW + bf16 write-back on the engine's bf16 gradient buffer (which the backend has filled and all-reduced)."""
        gn = self.engine.optimizer_step(self.scheduler.lr, grad_scale=self.grad_scale, max_grad_norm=self.max_grad_norm,
                                        weight_decay=self.weight_decay)
        self.scheduler.step()
        self.last_grad_norm = gn
        return gn

    # ---- reference-layout state -------------------------------------------------------------------------------------
    def state_dict(self) -> dict:
        eng = self.engine
        eng.ensure_optimizer_state()
        step_t = torch.tensor(float(eng.opt_step), dtype=torch.float32)
        state = {}
        for i, n in enumerate(self.names):
            state[i] = {"step": step_t.clone(), "exp_avg": eng.param_view(n, eng.exp_avg).detach().cpu().clone(),
                        "exp_avg_sq": eng.param_view(n, eng.exp_avg_sq).detach().cpu().clone()}
        group = _adamw_param_group_template()
        group.update(lr=self.scheduler.lr, weight_decay=self.weight_decay, initial_lr=self.scheduler.base_lr,
                     params=list(range(len(self.names))))
        if eng.opt_step == 0:
            state = {}                           # torch.optim.AdamW has no per-parameter state before its first step
        return {
            "optimizer_state_dict": {"state": state, "param_groups": [group]},
            "scheduler_state_dict": self.scheduler.state_dict(),
            "lr_scheduler_type": self.lr_scheduler_type,
            "max_grad_norm": self.max_grad_norm,
            "fp32_params": [eng.param_view(n, eng.master).detach().cpu().clone() for n in self.names],
        }

    def load_state_dict(self, state_dict: dict) -> None:
        saved_type = state_dict.get("lr_scheduler_type", "cosine")
        if saved_type != self.lr_scheduler_type:
            raise ValueError(f"checkpoint optimizer used lr_scheduler={saved_type!r} but this run has lr_scheduler={self.lr_scheduler_type!r}")
        saved_norm = state_dict.get("max_grad_norm")
        if saved_norm is not None and float(saved_norm) != float(self.max_grad_norm):
            raise ValueError(f"checkpoint optimizer used max_grad_norm={saved_norm} but this run has max_grad_norm={self.max_grad_norm}")
        eng = self.engine
        eng.ensure_optimizer_state()
        osd = state_dict["optimizer_state_dict"]
        per_param: Dict[int, dict] = osd.get("state", {})
        if per_param and len(per_param) != len(self.names):
            raise ValueError(f"checkpoint carries optimizer state for {len(per_param)} parameters but this model has {len(self.names)}")
        steps = set()
        with torch.no_grad():
            for i, n in enumerate(self.names):
                st = per_param.get(i, per_param.get(str(i)))
                if st is None:
                    eng.param_view(n, eng.exp_avg).zero_()
                    eng.param_view(n, eng.exp_avg_sq).zero_()
                    continue
                for key, buf in (("exp_avg", eng.exp_avg), ("exp_avg_sq", eng.exp_avg_sq)):
                    dst = eng.param_view(n, buf)
                    if tuple(st[key].shape) != tuple(dst.shape):
                        raise ValueError(f"optimizer state {key} of parameter {i} ({n}): checkpoint {tuple(st[key].shape)} vs {tuple(dst.shape)}")
                    dst.copy_(st[key].to(dst.device, torch.float32))
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"per-parameter AdamW step counts differ in the checkpoint: {sorted(steps)}")
        eng.opt_step = steps.pop() if steps else 0
        self.scheduler.load_state_dict(state_dict["scheduler_state_dict"])
        saved_fp32 = state_dict.get("fp32_params")
        with torch.no_grad():
            if saved_fp32 is not None:
                if len(saved_fp32) != len(self.names):
                    raise ValueError(f"checkpoint carries {len(saved_fp32)} fp32 master params but this rank has {len(self.names)}")
                for i, (n, saved) in enumerate(zip(self.names, saved_fp32)):
                    dst = eng.param_view(n, eng.master)
                    if tuple(saved.shape) != tuple(dst.shape):
                        raise ValueError(f"fp32 master param {i} shape mismatch: checkpoint {tuple(saved.shape)} vs current {tuple(dst.shape)}")
                    dst.copy_(saved.to(dst.device, torch.float32))
            else:   # same fallback as the reference: re-clone the masters from the bf16 weights (not numerically faithful)
                eng.master.copy_(eng.params.float())
