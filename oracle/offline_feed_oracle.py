"""TEST INFRASTRUCTURE — CPU restatement of the reference's offline EAGLE3 input pipe (per-sample normalisation + batch
collation).  Only tests/ may import this; the product path (specforge_b200/shards.py) never does.

Pinned against the reference's own functions by tests/test_integration_reference.py (build container, where
/root/reference is importable); restated here so the GPU box and the CPU suite can check shards without it.
"""
from __future__ import annotations

from typing import Dict, List

import torch


def normalize_offline_sample(raw: Dict[str, torch.Tensor], max_len: int) -> Dict[str, torch.Tensor]:
    """algorithms/eagle3/data.py:10-27 (same as data/preprocessing.py:642-665): truncate to max_len, swap names
    (hidden_state <- aux_hidden_state, target <- hidden_state), zero the last kept loss-mask position, all-ones mask."""
    hidden_state = raw["aux_hidden_state"].squeeze(0)[:max_len].unsqueeze(0)
    target = raw["hidden_state"].squeeze(0)[:max_len].unsqueeze(0)
    input_ids = raw["input_ids"][:max_len].unsqueeze(0)
    loss_mask = raw["loss_mask"][:max_len].clone().unsqueeze(0)
    if loss_mask.numel() > 0:
        loss_mask[0, -1] = 0
    return {"attention_mask": torch.ones_like(loss_mask, dtype=torch.long), "loss_mask": loss_mask, "target": target,
            "hidden_state": hidden_state, "input_ids": input_ids}


def collate_with_padding(features: List[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """data/utils.py:106-200 for sp_degree == 1: zero-pad every feature on the token axis to the longest sample, cat on
    the batch axis."""
    n = max(f["input_ids"].shape[1] for f in features)

    def pad(t: torch.Tensor) -> torch.Tensor:
        shape = list(t.shape)
        shape[1] = n - t.shape[1]
        return torch.cat((t, torch.zeros(shape, dtype=t.dtype)), dim=1)

    return {k: torch.cat([pad(f[k]) for f in features]) for k in ("input_ids", "attention_mask", "loss_mask", "hidden_state", "target")}
