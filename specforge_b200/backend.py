"""B200TrainingBackend — the TrainingBackend seam (specforge/training/backend.py:126-148) for the CUDA path.

DDP semantics of the reference (backend.py:233-253, NO_SHARD): replicated weights, ONE gradient all-reduce per
optimizer step over the flat bf16 gradient buffer (NCCL over NVLink/NVSwitch), gradients averaged over ranks, then
BF16Optimizer.step (optimizer.py:140-168) — here `B200BF16Optimizer`, one fused kernel — identically on every rank.
`no_sync` accumulation semantics: non-boundary micro-steps only accumulate into the fp32 buffer.

Two ways to build it:
  * directly:            B200TrainingBackend(lr=..., max_grad_norm=..., total_steps=...) ; attach(strategy) ; prepare_model(m)
  * the reference's way:  B200TrainingBackend(parallel_config, optimizer_factory=f) ; prepare_model(model, optimizer_target=draft)
    — the call `Trainer` makes at specforge/training/trainer.py:421-425 (see specforge_b200/train.py); the strategy built
    afterwards by the step provider finds this backend through the draft module and attaches itself.
`state_dict()` is the reference's {"model", "optimizer", "rng"} with the optimizer entry in BF16Optimizer's own layout."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from .optimizer import B200BF16Optimizer, WarmupSchedule


@dataclass
class LRSchedule:
    """Warm-up + cosine / constant as BF16Optimizer configures it (optimizer.py:56-61: warmup_steps =
    int(warmup_ratio * total_steps)); thin front of optimizer.WarmupSchedule kept for callers that only want lr_at()."""
    base_lr: float
    total_steps: int = 800_000
    warmup_ratio: float = 0.015
    kind: str = "cosine"
    eta_min: float = 0.0

    def lr_at(self, step: int) -> float:
        return WarmupSchedule(self.base_lr, self.total_steps, int(self.warmup_ratio * self.total_steps), self.kind, self.eta_min).lr_at(step)


class B200TrainingBackend:
    name = "b200"
    optimizer_state_is_replicated = True

    def __init__(self, parallel_config=None, *, optimizer_factory: Optional[Callable] = None, lr: Optional[float] = None,
                 max_grad_norm: float = 0.5, weight_decay: float = 0.0, total_steps: int = 800_000, warmup_ratio: float = 0.015,
                 lr_scheduler: str = "cosine", process_group=None):
        if optimizer_factory is None and lr is None:
            raise TypeError("B200TrainingBackend needs either lr=... or optimizer_factory=...")
        self._parallel_config = parallel_config
        self._optimizer_factory = optimizer_factory
        self._opt_kwargs = dict(lr=lr, max_grad_norm=max_grad_norm, weight_decay=weight_decay, total_steps=total_steps,
                                warmup_ratio=warmup_ratio, lr_scheduler=lr_scheduler)
        self.process_group = process_group if process_group is not None else getattr(parallel_config, "fsdp_process_group", None)
        self.strategy = None
        self.engine = None
        self.module: Optional[nn.Module] = None
        self.optimizer: Optional[B200BF16Optimizer] = None
        if optimizer_factory is None:                      # the schedule is known up front: expose the learning rate at once
            self._lr_probe = WarmupSchedule(lr, total_steps, int(warmup_ratio * total_steps), lr_scheduler)
        self._comm_stream = None
        self._reduced_elems = 0
        self.overlap_allreduce = True

    # ---- what the reference controller reads off a backend ----------------------------------------------------------
    @property
    def parallel_config(self):
        """training/controller.py:379-380,413-414,692 (reference type: training/backend.py:31-54): pure data parallelism —
        every group is the data-parallel group."""
        if self._parallel_config is not None:
            return self._parallel_config
        from types import SimpleNamespace
        pg = self.process_group
        return SimpleNamespace(world_size=self.world_size, tp_size=1, sp_ulysses_size=1, sp_ring_size=1, sharding_strategy="NO_SHARD",
                               param_dtype=torch.bfloat16, fsdp_process_group=pg, dp_group=pg, draft_dp_group=pg, tp_group=None,
                               sp_ulysses_group=None, sp_ring_group=None, draft_sp_group=None, device_mesh=None, tp_device_mesh=None, extra={})

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.process_group) if dist.is_available() and dist.is_initialized() else 1

    def get_learning_rate(self) -> float:
        return self.optimizer.get_learning_rate() if self.optimizer is not None else self._lr_probe.lr

    # ---- wiring -------------------------------------------------------------------------------------------------------
    def attach(self, strategy) -> None:
        self.strategy = strategy
        self.engine = strategy.engine
        strategy.grad_ready_hook = self._on_grad_slice_ready
        if self.optimizer is None and self._optimizer_factory is None:
            self._build_optimizer(strategy.draft_model)

    def _build_optimizer(self, target) -> None:
        if self._optimizer_factory is not None:
            opt = self._optimizer_factory(target)
            if not isinstance(opt, B200BF16Optimizer):
                raise TypeError(f"optimizer_factory returned {type(opt).__name__}; the B200 backend steps the engine's flat buffers and "
                                "needs a specforge_b200.optimizer.B200BF16Optimizer (train.py installs a compatible factory)")
        else:
            opt = B200BF16Optimizer(target, **self._opt_kwargs)
        self.optimizer = opt
        self.engine = opt.engine

    def prepare_model(self, model: nn.Module, *, wrap: bool = True, optimizer_target=None) -> nn.Module:
        """Replicas instead of a wrapper: rank 0's flat parameter buffer is broadcast, the module is returned as is."""
        self.module = model
        target = optimizer_target if optimizer_target is not None else getattr(model, "draft_model", None)
        if target is not None and getattr(target, "engine", None) is not None:
            if self.optimizer is None:
                self._build_optimizer(target)
            target._b200_backend = self                    # the strategy built next attaches itself through this
        if self.world_size > 1 and self.engine is not None:
            dist.broadcast(self.engine.params, src=0, group=self.process_group)  # identical replicas
        return model

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

    def backward(self, loss: torch.Tensor, *, is_boundary: bool = True) -> None:
        self.strategy._is_boundary = is_boundary
        self._reduced_elems = 0
        loss.backward()  # -> the strategy's autograd.Function -> sf_*_backward (accumulates into the fp32 flat buffer)

    def scale_gradients(self, factor: torch.Tensor) -> None:
        """The reference controller calls this between backward() and step() when the strategy reports loss_terms
        (controller.py:375-398).  If slices were already converted and all-reduced during backward, the factor has to reach
        the bf16 buffer the optimizer reads; otherwise it is applied to the fp32 accumulators."""
        eng = self.engine
        if self._reduced_elems > 0:
            torch.cuda.current_stream(eng.device).wait_stream(self._comm_stream)
            eng.grads_bf16.mul_(factor.to(eng.grads_bf16.device))
        eng.grads_f32.mul_(factor.to(eng.grads_f32.device))

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
        self.optimizer.grad_scale = 1.0 / world
        gn = self.optimizer.step()
        st._micro_in_window = 0
        return gn

    # ---- checkpoint state (training/backend.py:338-400): {"model", "optimizer", "rng"} ---------------------------------
    def state_dict(self) -> dict:
        eng = self.engine
        rng = {"torch": torch.get_rng_state()}
        if eng.device.type == "cuda":
            rng["cuda"] = torch.cuda.get_rng_state(eng.device)
        # the composite's state dict ("draft_model."-prefixed keys, what the reference gathers from its wrapped module and
        # hands to strategy.checkpoint_state_filter, controller.py:843-853)
        model = self.module.state_dict() if self.module is not None else (self.strategy.draft_model.state_dict() if self.strategy else {})
        return {"model": model, "optimizer": self.optimizer.state_dict(), "rng": rng}

    def load_state_dict(self, state: dict) -> None:
        eng = self.engine
        if state.get("model"):
            model = {k.replace("draft_model.", "", 1): v for k, v in state["model"].items()}
            self.strategy.draft_model.load_state_dict(model, strict=False)
        if state.get("optimizer") is not None:
            self.optimizer.load_state_dict(state["optimizer"])
        rng = state.get("rng")
        if rng is not None:
            torch.set_rng_state(rng["torch"])
            if rng.get("cuda") is not None and eng.device.type == "cuda":
                torch.cuda.set_rng_state(rng["cuda"], eng.device)
