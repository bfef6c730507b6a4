"""Data contracts at the strategy boundary, mirroring the reference's own (same field names and meaning):
TrainBatch  <- specforge/runtime/contracts.py:118-129
StepOutput  <- specforge/training/strategies/base.py:29-42
StepContext <- specforge/training/strategies/base.py:45-53
When the reference package is importable its classes are used as-is, so objects flow through unchanged."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import torch

try:  # pragma: no cover - only where the reference is installed next to us
    from specforge.runtime.contracts import TrainBatch  # type: ignore
    from specforge.training.strategies.base import StepContext, StepOutput  # type: ignore
except Exception:  # the GPU box / a standalone install

    @dataclass
    class TrainBatch:
        sample_ids: List[str]
        strategy: str
        tensors: Dict[str, "torch.Tensor"]
        metadata: Dict[str, Any] = field(default_factory=dict)

    @dataclass(frozen=True)
    class StepOutput:
        loss: torch.Tensor
        metrics: Dict[str, Any]
        ratio_metrics: Dict[str, Tuple[Any, Any]] = field(default_factory=dict)
        loss_terms: Optional[Tuple[torch.Tensor, torch.Tensor]] = None

    @dataclass(frozen=True)
    class StepContext:
        global_step: int = 0
        total_steps: Optional[int] = None
