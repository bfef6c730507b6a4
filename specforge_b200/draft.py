"""B200Eagle3DraftModel — the draft nn.Module behind the reference's `@register_draft` seam.

Mirrors LlamaForCausalLMEagle3 (specforge/modeling/draft/llama3_eagle.py:1653-1798) at the state-dict level:
identical parameter names and shapes, `t2d`/`d2t` buffers, frozen `embed_tokens`, the
`load_embedding` / `load_vocab_mapping` / `freeze_embedding` helpers of Eagle3DraftModel
(specforge/modeling/draft/base.py:128-206).  The trainable parameters are *views into one flat bf16 buffer*
owned by the engine (q/k/v and gate/up adjacent), so checkpoints/export see the reference layout while the
kernels see fused operands.  There is no eager PyTorch forward: compute goes through Eagle3Engine (C ABI).
"""
from __future__ import annotations

import json
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from .engine import DraftDims, Eagle3Engine, P_NAMES


def dims_from_config(config: Any) -> DraftDims:
    """Accepts a transformers LlamaConfig-like object or a dict (configs/*-eagle3.json)."""
    get = (lambda k, d=None: config.get(k, d)) if isinstance(config, dict) else (lambda k, d=None: getattr(config, k, d))
    H = get("hidden_size")
    nh = get("num_attention_heads")
    head_dim = get("head_dim") or H // nh
    rope_params = get("rope_parameters")
    theta = rope_params["rope_theta"] if rope_params else get("rope_theta", 10000.0)
    scaling = get("rope_scaling") or (rope_params if rope_params and rope_params.get("rope_type") not in (None, "default") else None)
    if scaling and scaling.get("rope_type", scaling.get("type")) == "mrope":
        raise NotImplementedError("mrope (three-axis multimodal positions) is not implemented on the CUDA path")
    return DraftDims(hidden_size=H, intermediate_size=get("intermediate_size"), num_heads=nh,
                     num_kv_heads=get("num_key_value_heads", nh), head_dim=head_dim, vocab_size=get("vocab_size"),
                     draft_vocab_size=get("draft_vocab_size"), target_hidden_size=get("target_hidden_size", H),
                     rms_norm_eps=get("rms_norm_eps", 1e-6), rope_theta=float(theta),
                     max_position_embeddings=get("max_position_embeddings", 2048), fc_norm=bool(get("fc_norm", False)),
                     norm_output=bool(get("norm_output", True)), rope_scaling=dict(scaling) if scaling else None)


try:  # the reference's registry requires `config_class` (registry.py:38-41); LlamaConfig is what its drafts use
    from transformers import LlamaConfig as _LlamaConfig
except Exception:  # pragma: no cover
    _LlamaConfig = None


class B200Eagle3DraftModel(nn.Module):
    architectures = ["LlamaForCausalLMEagle3"]  # what the draft JSON names; kept for export parity
    config_class = _LlamaConfig

    def __init__(self, config: Any, quant_config=None, attention_backend: str = "b200"):
        super().__init__()
        self.config = config
        self.dims = dims_from_config(config)
        self.attention_backend = attention_backend
        self.engine: Optional[Eagle3Engine] = None
        d = self.dims
        # placeholders with the reference shapes; re-pointed into the engine's flat buffer by bind_engine()
        self._param_names = []
        for name in P_NAMES:
            if name.startswith("fc_norm") and not d.fc_norm:
                continue
            self._param_names.append(name)
        self.embed_tokens_weight = nn.Parameter(torch.empty(0), requires_grad=False)
        self.register_buffer("t2d", torch.ones(d.vocab_size, dtype=torch.bool))
        self.register_buffer("d2t", torch.zeros(d.draft_vocab_size, dtype=torch.int64))
        self._flat_params: Dict[str, nn.Parameter] = {}
        self.vocab_mapping_loaded = False

    def state_dict_spec(self) -> Dict[str, tuple]:
        """name -> shape of everything state_dict() will hold (the reference LlamaForCausalLMEagle3 contract)."""
        d = self.dims
        H, I, hd, Ht = d.hidden_size, d.intermediate_size, d.head_dim, d.target_hidden_size
        spec = {
            "embed_tokens.weight": (d.vocab_size, H),
            "fc.weight": (H, 3 * Ht),
            "midlayer.self_attn.q_proj.weight": (d.num_heads * hd, 2 * H),
            "midlayer.self_attn.k_proj.weight": (d.num_kv_heads * hd, 2 * H),
            "midlayer.self_attn.v_proj.weight": (d.num_kv_heads * hd, 2 * H),
            "midlayer.self_attn.o_proj.weight": (H, d.num_heads * hd),
            "midlayer.mlp.gate_proj.weight": (I, H),
            "midlayer.mlp.up_proj.weight": (I, H),
            "midlayer.mlp.down_proj.weight": (H, I),
            "midlayer.hidden_norm.weight": (H,),
            "midlayer.input_layernorm.weight": (H,),
            "midlayer.post_attention_layernorm.weight": (H,),
            "norm.weight": (H,),
            "lm_head.weight": (d.draft_vocab_size, H),
            "t2d": (d.vocab_size,),
            "d2t": (d.draft_vocab_size,),
        }
        if d.fc_norm:
            for i in range(3):
                spec[f"fc_norm.{i}.weight"] = (Ht,)
        return spec

    # ---- engine binding -------------------------------------------------------------------------------
    def bind_engine(self, batch: int, seq_len: int, ttt_length: int, ploss_decay: float = 0.8, device=None,
                    init_std: float = 0.02, seed: int = 0) -> Eagle3Engine:
        eng = Eagle3Engine(self.dims, batch=batch, seq_len=seq_len, ttt_length=ttt_length, ploss_decay=ploss_decay, device=device)
        g = torch.Generator(device="cpu").manual_seed(seed)
        for name in self._param_names:
            view = eng.param_view(name)
            if view.dim() == 1:
                view.fill_(1.0)
            else:
                view.copy_((torch.randn(view.shape, generator=g) * init_std).to(torch.bfloat16))  # HF normal(0, 0.02)
            p = nn.Parameter(view, requires_grad=True)
            self._flat_params[name] = p
            self.register_parameter(name.replace(".", "__"), p)
        if self.embed_tokens_weight.numel() == 0:
            self.embed_tokens_weight = nn.Parameter(
                torch.zeros(self.dims.vocab_size, self.dims.hidden_size, dtype=torch.bfloat16, device=eng.device), requires_grad=False)
        self.t2d = self.t2d.to(eng.device)
        self.d2t = self.d2t.to(eng.device)
        self.engine = eng
        return eng

    def trainable_parameters(self):
        return [self._flat_params[n] for n in self._param_names]

    # ---- reference-named state dict ---------------------------------------------------------------------
    def state_dict(self, *args, **kwargs) -> Dict[str, torch.Tensor]:  # noqa: D401 - reference key layout
        out = {"embed_tokens.weight": self.embed_tokens_weight.detach()}
        for n in self._param_names:
            out[n] = self._flat_params[n].detach()
        out["t2d"], out["d2t"] = self.t2d, self.d2t
        return out

    def load_state_dict(self, state: Dict[str, torch.Tensor], strict: bool = True):
        missing = []
        with torch.no_grad():
            for n in self._param_names:
                if n in state:
                    self._flat_params[n].copy_(state[n].to(torch.bfloat16))
                else:
                    missing.append(n)
            if "embed_tokens.weight" in state:
                self.embed_tokens_weight.data = state["embed_tokens.weight"].to(self.embed_tokens_weight.device, torch.bfloat16).contiguous()
            for b in ("t2d", "d2t"):
                if b in state:
                    getattr(self, b).copy_(state[b])
        if strict and missing:
            raise KeyError(f"missing keys in draft state dict: {missing}")
        if self.engine is not None:
            self.engine.master = None  # fp32 masters are re-derived from the loaded weights
        return missing

    # ---- Eagle3DraftModel helpers -----------------------------------------------------------------------
    def freeze_embedding(self) -> None:
        self.embed_tokens_weight.requires_grad = False

    @torch.no_grad()
    def load_embedding(self, model_path: str, embedding_key: str = "model.embed_tokens.weight") -> None:
        import glob
        import os
        from safetensors import safe_open
        idx = glob.glob(os.path.join(model_path, "*.index.json"))
        if len(idx) > 1:
            raise FileNotFoundError(f"Multiple index.json files found in {model_path}")
        if idx:
            with open(idx[0]) as f:
                ckpt = json.load(f)["weight_map"][embedding_key]
            path = os.path.join(model_path, ckpt)
        else:
            path = os.path.join(model_path, "model.safetensors")
            if not os.path.exists(path):
                raise FileNotFoundError(f"No index.json or model.safetensors found in {model_path}")
        if path.endswith(".safetensors"):
            with safe_open(path, framework="pt") as f:
                w = f.get_tensor(embedding_key)
        else:
            w = torch.load(path, map_location="cpu")[embedding_key]
        self.embed_tokens_weight.data = w.to(self.embed_tokens_weight.device, torch.bfloat16).contiguous()

    def load_vocab_mapping(self, file_path: str) -> None:
        m = torch.load(file_path)
        self.t2d.copy_(m["t2d"])
        self.d2t.copy_(m["d2t"])
        self.vocab_mapping_loaded = True

    def sync_frozen(self, target_head_weight: torch.Tensor) -> None:
        """Hand the frozen tables (embedding, target head, vocab map) to the engine."""
        self.engine.set_frozen(embed_tokens=self.embed_tokens_weight.data, target_head=target_head_weight, t2d=self.t2d, d2t=self.d2t)

    def forward(self, *args, **kwargs):
        raise RuntimeError("B200Eagle3DraftModel has no eager forward: drive it through B200Eagle3TrainStrategy.forward_loss "
                           "(the CUDA path has no PyTorch fallback)")
