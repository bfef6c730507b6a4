"""B200TrainingBackend — the TrainingBackend seam (specforge/training/backend.py:126-148) for the CUDA path.

DDP semantics of the reference (backend.py:233-253, NO_SHARD): replicated weights, ONE gradient all-reduce per
optimizer step over the flat bf16 gradient buffer (NCCL over NVLink/NVSwitch), gradients averaged over ranks, then
BF16Optimizer.step (optimizer.py:140-168) — here fused into sf_optimizer_step — identically on every rank.
`no_sync` accumulation semantics: non-boundary micro-steps only accumulate into the fp32 buffer."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn


@dataclass
class LRSchedule:
    """Warm-up + cosine / constant, mirroring specforge/lr_scheduler.py:93-147 as used by BF16Optimizer
    (optimizer.py:56-61: warmup_steps = int(warmup_ratio * total_steps))."""
    base_lr: float
    total_steps: int = 800_000
    warmup_ratio: float = 0.015
    kind: str = "cosine"
    eta_min: float = 0.0

    def lr_at(self, step: int) -> float:
        warm = int(self.warmup_ratio * self.total_steps)
        if step < warm:
            return self.base_lr * (step + 1) / warm
        if self.kind == "constant":
            return self.base_lr
        t = step - warm
        tmax = max(1, self.total_steps - warm)
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * t / tmax)) / 2


class B200TrainingBackend:
    name = "b200"
    optimizer_state_is_replicated = True

    def __init__(self, *, lr: float, max_grad_norm: float = 0.5, weight_decay: float = 0.0, total_steps: int = 800_000,
                 warmup_ratio: float = 0.015, lr_scheduler: str = "cosine", process_group=None):
        self.schedule = LRSchedule(lr, total_steps, warmup_ratio, lr_scheduler)
        self.max_grad_norm = max_grad_norm
        self.weight_decay = weight_decay
        self.process_group = process_group
        self.strategy = None
        self.engine = None
        self.module: Optional[nn.Module] = None
        self._step = 0
        self._comm_stream: Optional[torch.cuda.Stream] = None
        self._reduced_elems = 0
        self.overlap_allreduce = True

    @property
    def parallel_config(self):
        """What the reference controller reads off a backend (training/controller.py:379-380,413-414,692; the reference
        type is training/backend.py:31-54): pure data parallelism — every group is the data-parallel group."""
        from types import SimpleNamespace
        pg = self.process_group
        return SimpleNamespace(world_size=self.world_size, tp_size=1, sp_ulysses_size=1, sp_ring_size=1, sharding_strategy="NO_SHARD",
                               param_dtype=torch.bfloat16, fsdp_process_group=pg, dp_group=pg, draft_dp_group=pg, tp_group=None,
                               sp_ulysses_group=None, sp_ring_group=None, draft_sp_group=None, device_mesh=None, tp_device_mesh=None, extra={})

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.process_group) if dist.is_available() and dist.is_initialized() else 1

    def attach(self, strategy) -> None:
        self.strategy = strategy
        self.engine = strategy.engine
        strategy.grad_ready_hook = self._on_grad_slice_ready

    # ---- overlapped gradient all-reduce ---------------------------------------------------------------
    # On the boundary micro-step the C library reports each contiguous slice of the flat gradient as soon as the last
    # kernel writing it has been enqueued (norms, lm_head, down, gate+up, o, qkv, fc).  Each slice is converted to
    # bf16 and all-reduced on a communication stream while the remaining weight-gradient GEMMs still run.
    def _on_grad_slice_ready(self, first: int, count: int) -> None:
        if self.world_size == 1 or not self.overlap_allreduce:
            return
        eng = self.engine
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=eng.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(eng.device))
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(ev)
            eng.grads_to_bf16(scale=self.strategy._last_grad_out, first=first, count=count)
            dist.all_reduce(eng.grads_bf16[first:first + count], op=dist.ReduceOp.SUM, group=self.process_group)
        self._reduced_elems += count

    def prepare_model(self, model: nn.Module, *, wrap: bool = True, optimizer_target=None) -> nn.Module:
        self.module = model
        if self.world_size > 1 and self.engine is not None:
            dist.broadcast(self.engine.params, src=0, group=self.process_group)  # identical replicas
        return model

    def backward(self, loss: torch.Tensor, *, is_boundary: bool = True) -> None:
        self.strategy._is_boundary = is_boundary
        self._reduced_elems = 0
        loss.backward()  # -> _Eagle3StepFn.backward -> sf_eagle3_backward(_ex) (accumulates into the fp32 flat buffer)

    def scale_gradients(self, factor: torch.Tensor) -> None:
        self.engine.grads_f32.mul_(factor)

    def step(self) -> torch.Tensor:
        eng, st = self.engine, self.strategy
        world = self.world_size
        if world > 1 and self._reduced_elems == eng.n_params:
            # every slice was converted + all-reduced on the comm stream during backward: just join it
            torch.cuda.current_stream(eng.device).wait_stream(self._comm_stream)
        else:
            g = eng.grads_to_bf16(scale=st._last_grad_out)   # x (1/accumulation_steps) from autograd, on device
            if world > 1:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.process_group)
        self._reduced_elems = 0
        lr = self.schedule.lr_at(self._step)
        gn = eng.optimizer_step(lr, grad_scale=1.0 / world, max_grad_norm=self.max_grad_norm,
                                weight_decay=self.weight_decay)
        self._step += 1
        st._micro_in_window = 0
        return gn

    def get_learning_rate(self) -> float:
        return self.schedule.lr_at(self._step)

    @property
    def optimizer(self):
        return self  # exposes get_learning_rate() like BF16Optimizer (controller.py:684-687)

    def state_dict(self) -> dict:
        eng = self.engine
        return {
            "model": self.strategy.draft_model.state_dict(),
            "optimizer": {
                "step": self._step, "opt_step": eng.opt_step,
                "fp32_params": None if eng.master is None else eng.master.cpu(),
                "exp_avg": None if eng.exp_avg is None else eng.exp_avg.cpu(),
                "exp_avg_sq": None if eng.exp_avg_sq is None else eng.exp_avg_sq.cpu(),
                "lr_scheduler_type": self.schedule.kind, "max_grad_norm": self.max_grad_norm,
            },
            "rng": {"torch": torch.get_rng_state(), "cuda": torch.cuda.get_rng_state(eng.device)},
        }

    def load_state_dict(self, state: dict) -> None:
        eng = self.engine
        if state.get("model") is not None:
            self.strategy.draft_model.load_state_dict(state["model"], strict=False)
        opt = state.get("optimizer")
        if opt is not None:
            if opt.get("lr_scheduler_type", self.schedule.kind) != self.schedule.kind:
                raise ValueError("checkpoint lr_scheduler differs from this run")
            self._step = int(opt["step"])
            eng.opt_step = int(opt["opt_step"])
            if opt.get("fp32_params") is not None:
                eng.master = opt["fp32_params"].to(eng.device)
                eng.exp_avg = opt["exp_avg"].to(eng.device)
                eng.exp_avg_sq = opt["exp_avg_sq"].to(eng.device)
        rng = state.get("rng")
        if rng is not None:
            torch.set_rng_state(rng["torch"])
            if rng.get("cuda") is not None:
                torch.cuda.set_rng_state(rng["cuda"], eng.device)
