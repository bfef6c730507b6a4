"""TEST INFRASTRUCTURE — CPU restatement of the reference's offline EAGLE3 input pipe (per-sample normalisation + batch
collation).  Only tests/ may import this; the product path (specforge_b200/shards.py) never does.

Pinned against the reference's own functions by tests/test_integration_reference.py (build container, where
/root/reference is importable); restated here so the GPU box and the CPU suite can check shards without it.
"""
from __future__ import annotations

from typing import Dict, List

import torch


def normalize_offline_sample(raw: Dict[str, torch.Tensor], max_len: int) -> Dict[str, torch.Tensor]:
    """algorithms/eagle3/data.py:10-27 (same as data/preprocessing.py:642-665): keep the first max_len tokens of every
    feature, rename (the stored final-layer state becomes `target`, the stored auxiliary layers become `hidden_state`),
    zero the LAST kept loss-mask position, attention mask of ones; every tensor gets a leading batch axis of 1."""
    def head(t: torch.Tensor) -> torch.Tensor:           # [1, L, W] or [L] -> first max_len tokens, batch axis restored
        t = t.squeeze(0) if t.dim() == 3 else t
        return t[:max_len].unsqueeze(0)

    out = {"input_ids": head(raw["input_ids"]), "loss_mask": head(raw["loss_mask"]).clone(),
           "target": head(raw["hidden_state"]), "hidden_state": head(raw["aux_hidden_state"])}
    if out["loss_mask"].numel():
        out["loss_mask"][0, -1] = 0
    out["attention_mask"] = torch.ones_like(out["loss_mask"], dtype=torch.long)
    return out


def collate_with_padding(features: List[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """data/utils.py:106-200 for sp_degree == 1: zero-pad every feature on the token axis to the longest sample, cat on
    the batch axis."""
    n = max(f["input_ids"].shape[1] for f in features)

    def pad(t: torch.Tensor) -> torch.Tensor:
        shape = list(t.shape)
        shape[1] = n - t.shape[1]
        return torch.cat((t, torch.zeros(shape, dtype=t.dtype)), dim=1)

    return {k: torch.cat([pad(f[k]) for f in features]) for k in ("input_ids", "attention_mask", "loss_mask", "hidden_state", "target")}


def normalize_offline_dflash_sample(raw: Dict[str, torch.Tensor], max_len: int) -> Dict[str, torch.Tensor]:
    """algorithms/common/dflash_family_data.py:37-70: first max_len tokens of input_ids / loss_mask / hidden_states
    ([seq, width] or [1, seq, width]), leading batch axis of 1, no renaming; two consecutive supervised tokens required."""
    hs = raw["hidden_states"]
    hs = hs.squeeze(0) if hs.dim() == 3 else hs
    out = {"input_ids": raw["input_ids"][:max_len].unsqueeze(0), "loss_mask": raw["loss_mask"][:max_len].unsqueeze(0),
           "hidden_states": hs[:max_len].unsqueeze(0)}
    lm = out["loss_mask"][0]
    if not bool(((lm[:-1] > 0) & (lm[1:] > 0)).any()):
        raise ValueError("offline DFlash-family samples require two consecutive supervised tokens")
    return out


def pad_and_concatenate(features: List[Dict[str, torch.Tensor]], keys=("input_ids", "loss_mask", "hidden_states")) -> Dict[str, torch.Tensor]:
    """algorithms/common/collation.py:24-70: zero-pad axis 1 of every key to the longest input_ids, concatenate on axis 0."""
    n = max(f["input_ids"].shape[-1] for f in features)
    out = {}
    for k in keys:
        parts = []
        for f in features:
            t = f[k]
            if t.shape[1] < n:
                shape = list(t.shape)
                shape[1] = n - t.shape[1]
                t = torch.cat([t, t.new_zeros(shape)], dim=1)
            parts.append(t)
        out[k] = torch.cat(parts, dim=0)
    return out
