"""B200Eagle3TrainStrategy — drop-in for Eagle3TrainStrategy (specforge/training/strategies/base.py:123-319).

Same `name`, `required_features`, `forward_loss(batch, ctx) -> StepOutput` (same metric keys, each a length-T list of
0-dim tensors, controller.py:216-252), `ploss_decay`, `checkpoint_state_filter`; works under torch.no_grad() (eval,
controller.py:794-815).  The whole step (teacher, TTT forward, loss, and later backward) runs in libspecforge_b200."""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from .contracts import StepContext, StepOutput, TrainBatch
from .draft import B200Eagle3DraftModel


class _Eagle3StepFn(torch.autograd.Function):
    """loss = f(params) with the forward/backward done by the C-ABI library.
    Accumulation contract: the engine accumulates UNSCALED micro-batch gradients in fp32 and the upstream factor (`grad_out`, i.e. the
    controller's 1 / accumulation_steps, controller.py:345) is applied ONCE, when the window's sum is converted to bf16 — so every
    micro-step of a window must arrive with the same `grad_out` (true for the reference controller; a caller that weights
    micro-batches differently must scale its losses' gradients itself).  The flat bf16 parameter buffer is the
    single differentiable input; backward leaves dLoss/dparams in the engine's fp32 accumulator (the backend reads
    it) and, for stock optimizers/DDP, also returns it to autograd as a bf16 tensor."""

    @staticmethod
    def forward(ctx, flat_params, strategy, batch_tensors, need_grad):
        eng = strategy.engine
        loss, _ = eng.forward(batch_tensors, need_grad=need_grad)
        ctx.strategy = strategy
        return loss.clone().reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        strategy = ctx.strategy
        eng = strategy.engine
        strategy._last_grad_out = grad_out.detach()
        hook = strategy.grad_ready_hook if strategy._is_boundary else None
        eng.backward(loss_scale=1.0, accumulate=strategy._micro_in_window > 0, on_ready=hook)
        strategy._micro_in_window += 1
        if strategy.return_autograd_grads:
            g = eng.grads_to_bf16(scale=grad_out)
            return g.clone(), None, None, None
        return None, None, None, None


class _Trainable(nn.Module):
    """What `trainable_module()` returns: exposes `.draft_model` like OnlineEagle3Model (trainer.py:205,425)."""

    def __init__(self, draft_model: B200Eagle3DraftModel):
        super().__init__()
        self.draft_model = draft_model


class B200Eagle3TrainStrategy:
    name = "eagle3"
    required_features = {"input_ids", "attention_mask", "loss_mask", "hidden_state", "target"}

    def __init__(self, draft_model, *, target_head_weight: Optional[torch.Tensor] = None, target_head: Optional[nn.Module] = None,
                 ploss_decay: float = 0.8, compact_teacher: bool = False, compact_teacher_chunk_size: Optional[int] = None,
                 return_autograd_grads: bool = False):
        """`draft_model`: the B200 draft module, or the composite the reference hands a step provider (anything exposing
        `.draft_model`, algorithms/eagle3/providers.py:45-52).  The frozen head is `target_head_weight` [V, H_t] or the
        reference's `TargetHead` module (its `fc.weight`, modeling/target/target_head.py:100-101).
        `compact_teacher` / `compact_teacher_chunk_size` (strategies/base.py:148-215) are accepted and need nothing: the teacher
        statistics are always streamed inside the target-head GEMM epilogue, the [B, S, V] logits never exist on this path."""
        composite = draft_model if hasattr(draft_model, "draft_model") else None
        if composite is not None:
            draft_model = composite.draft_model
        if getattr(draft_model, "engine", None) is None:
            raise RuntimeError("bind_engine() must be called on the draft model first")
        if target_head_weight is None:
            if target_head is None:
                raise ValueError("target_repr='hidden_state' requires a target_head to re-run the lm_head projection")
            target_head_weight = target_head.fc.weight.detach()
        self.compact_teacher = bool(compact_teacher)
        self.compact_teacher_chunk_size = compact_teacher_chunk_size
        self.target_head = target_head
        self.draft_model = draft_model
        self.engine = draft_model.engine
        if abs(self.engine.ploss_decay - ploss_decay) > 1e-12:
            raise ValueError("ploss_decay differs from the value the engine was bound with")
        self.ploss_decay = ploss_decay
        self.return_autograd_grads = return_autograd_grads
        self._module = composite if composite is not None else _Trainable(draft_model)
        self.eagle3_model = self._module
        self._micro_in_window = 0
        self._last_grad_out: Optional[torch.Tensor] = None
        self._is_boundary = True            # set by the backend before loss.backward()
        self.grad_ready_hook = None         # backend: called with (first_elem, n_elems) as gradient slices complete
        draft_model.sync_frozen(target_head_weight)
        backend = getattr(draft_model, "_b200_backend", None)   # a backend prepared before us (Trainer order, trainer.py:421-431)
        if backend is not None:
            backend.attach(self)

    def trainable_module(self) -> nn.Module:
        return self._module

    def validate_batch(self, batch: TrainBatch) -> None:
        missing = {f for f in self.required_features if f not in batch.tensors}
        if missing:
            raise ValueError(f"{self.name} batch missing required features {sorted(missing)}; present={sorted(batch.tensors)}")

    def forward_loss(self, batch: TrainBatch, ctx: Optional[StepContext] = None) -> StepOutput:
        self.validate_batch(batch)
        target_repr = batch.metadata.get("target_repr", "hidden_state") or "hidden_state"
        if target_repr != "hidden_state":
            if self.compact_teacher:
                raise ValueError("compact teacher is offline-only and requires target_repr='hidden_state'")
            raise ValueError(f"target_repr={target_repr!r}: the CUDA path implements the offline hidden-state teacher only "
                             "(online 'logits' / 'pruned_logits' capture is outside the replaced hot path)")
        self._check_position_ids(batch.tensors.get("position_ids"), batch.tensors["input_ids"])
        need_grad = torch.is_grad_enabled()
        eng = self.engine
        flat = eng.params if not need_grad else eng.params.detach().requires_grad_(True)
        loss = _Eagle3StepFn.apply(flat, self, batch.tensors, need_grad)
        m = eng.metrics.clone()  # [T, 8] on device; slicing below creates views, no host sync
        T = eng.T
        acc = m[:, 1] / m[:, 2]    # one launch for all T positions
        metrics = {
            "plosses": [m[j, 0] for j in range(T)],
            "acces": [acc[j] for j in range(T)],
            "acceptance_rates": [m[j, 3] for j in range(T)],
            "acc_corrects": [m[j, 1] for j in range(T)],
            "acc_denoms": [m[j, 2] for j in range(T)],
            "metric_losses": [m[j, 0] for j in range(T)],
            "metric_loss_denoms": [m[j, 6] for j in range(T)],
        }
        return StepOutput(loss=loss, metrics=metrics)

    @staticmethod
    def _check_position_ids(position_ids, input_ids) -> None:
        """The kernels rotate row s of every sequence at position s + j (the reference's default when a batch carries no
        `position_ids`, eagle3/model.py:335-345).  Explicit positions are accepted only if they say the same thing: [B, S] or, for
        mrope, [3, B, S] with all three axes equal to arange(S) — text tokens; anything else (packed sequences, vision tokens) is
        refused rather than silently mis-rotated."""
        if position_ids is None:
            return
        S = input_ids.shape[-1]
        p = position_ids.reshape(-1, S)
        if not bool((p == torch.arange(S, device=p.device, dtype=p.dtype)).all()):
            raise NotImplementedError("position_ids other than arange(seq_len) on every axis (packed / multimodal positions) are not "
                                      "implemented on the CUDA path")

    def checkpoint_state_filter(self, state_dict: Dict[str, Any]) -> Dict[str, Any]:
        """Same rule as the reference (strategies/base.py:306-319): draft weights without the `draft_model.` prefix,
        frozen embedding dropped."""
        prefixed = any("draft_model." in k for k in state_dict)     # composite state dict vs the bare draft module's
        return {k.replace("draft_model.", ""): v for k, v in state_dict.items()
                if (not prefixed or "draft_model." in k) and "embed" not in k.lower()}
