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
from .optimizer import REFERENCE_PARAM_ORDER


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
        # Three-axis multimodal RoPE (llama3_eagle.py:145-183,389-424).  For TEXT positions the three axes carry the same index and
        # the sectioned cos / sin equal the plain ones ("the text embedding rotary position embedding has no difference with modern
        # LLMs", :153-155), so the default tables serve; the strategy refuses batches whose position_ids axes differ (vision tokens).
        if float(scaling.get("factor", 1.0) or 1.0) != 1.0:
            raise NotImplementedError("mrope with a scaling factor != 1 is not implemented on the CUDA path")
        scaling = None
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
        # registered in the order the reference module yields its parameters (so `parameters()` / optimizer state line up
        # with LlamaForCausalLMEagle3); the physical order inside the flat buffer is the engine's (P_NAMES)
        self._param_names = [n for n in REFERENCE_PARAM_ORDER if n in P_NAMES and (d.fc_norm or not n.startswith("fc_norm"))]
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
                    init_std: float = 0.02, seed: int = 0, lk_loss_type: Optional[str] = None, kl_scale: float = 1.0,
                    kl_decay: float = 1.0) -> Eagle3Engine:
        eng = Eagle3Engine(self.dims, batch=batch, seq_len=seq_len, ttt_length=ttt_length, ploss_decay=ploss_decay,
                           lk_loss_type=lk_loss_type, kl_scale=kl_scale, kl_decay=kl_decay, device=device)
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
    def state_dict(self, *args, destination=None, prefix: str = "", keep_vars: bool = False):  # reference key layout
        out = destination if destination is not None else {}
        out[prefix + "t2d"], out[prefix + "d2t"] = self.t2d, self.d2t
        out[prefix + "embed_tokens.weight"] = self.embed_tokens_weight if keep_vars else self.embed_tokens_weight.detach()
        for n in self._param_names:
            out[prefix + n] = self._flat_params[n] if keep_vars else self._flat_params[n].detach()
        return out

    def load_state_dict(self, state: Dict[str, torch.Tensor], strict: bool = True, assign: bool = False):
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
        self._resync_frozen()
        spec = self.state_dict_spec()
        missing += [k for k in ("embed_tokens.weight", "t2d", "d2t") if k not in state]
        unexpected = [k for k in state if k not in spec and "rotary_emb" not in k]    # RoPE buffers are non-persistent upstream
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

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
        self._resync_frozen()

    def load_vocab_mapping(self, file_path: str) -> None:
        m = torch.load(file_path)
        self.t2d.copy_(m["t2d"])
        self.d2t.copy_(m["d2t"])
        self.vocab_mapping_loaded = True
        self._resync_frozen()

    def sync_frozen(self, target_head_weight: torch.Tensor) -> None:
        """Hand the frozen tables (embedding, target head, vocab map) to the engine."""
        self._target_head_weight = target_head_weight
        self.engine.set_frozen(embed_tokens=self.embed_tokens_weight.data, target_head=target_head_weight, t2d=self.t2d, d2t=self.d2t)

    def _resync_frozen(self) -> None:
        """The engine holds its own copies of the frozen tables (the vocab map as bytes): refresh them whenever a loader
        rebinds the embedding or rewrites t2d / d2t after the strategy was built (resume, load_embedding, load_vocab_mapping)."""
        head = getattr(self, "_target_head_weight", None)
        if self.engine is not None and head is not None:
            self.sync_frozen(head)

    # ---- Eagle3DraftModel forward surface (modeling/draft/base.py:45-109, llama3_eagle.py:1705-1798), forward-only -------------
    # Each method is a few calls of the exported C-ABI ops (sf_embedding_gather, sf_rmsnorm_fwd, sf_gemm_bf16, sf_rope,
    # sf_ttt_attention_fwd, sf_swiglu_fwd) on this module's weights: what eval utilities / export sanity checks call outside
    # the training step.  No autograd — training goes through B200Eagle3TrainStrategy.forward_loss.
    def _w(self, name: str) -> torch.Tensor:
        return self._flat_params[name].detach()

    @torch.no_grad()
    def embed_input_ids(self, input_ids: torch.Tensor) -> torch.Tensor:
        from . import ops
        ids = input_ids.to(self.embed_tokens_weight.device, torch.int64).contiguous()
        out = ops.embedding_gather(self.embed_tokens_weight.data, ids.view(-1))
        return out.view(*ids.shape, self.dims.hidden_size)

    @torch.no_grad()
    def project_hidden_states(self, hidden_states: torch.Tensor) -> torch.Tensor:
        from . import ops
        d = self.dims
        assert hidden_states.size(-1) == d.target_hidden_size * 3
        lead = hidden_states.shape[:-1]
        x = hidden_states.to(torch.bfloat16).reshape(-1, 3 * d.target_hidden_size).contiguous()
        if d.fc_norm:
            y = torch.empty_like(x)
            for i in range(3):
                sl = slice(i * d.target_hidden_size, (i + 1) * d.target_hidden_size)
                ops.rmsnorm_fwd(x[:, sl], self._w(f"fc_norm.{i}.weight"), d.rms_norm_eps, out=y[:, sl])
            x = y
        return ops.gemm(x, self._w("fc.weight")).view(*lead, d.hidden_size)

    @torch.no_grad()
    def compute_logits(self, hidden_states: torch.Tensor) -> torch.Tensor:
        from . import ops
        d = self.dims
        lead = hidden_states.shape[:-1]
        x = hidden_states.to(torch.bfloat16).reshape(-1, d.hidden_size).contiguous()
        if d.norm_output:
            x = ops.rmsnorm_fwd(x, self._w("norm.weight"), d.rms_norm_eps)
        return ops.gemm(x, self._w("lm_head.weight")).view(*lead, d.draft_vocab_size)

    @staticmethod
    def _key_mask(attention_mask, B: int, S: int, device) -> Optional[torch.Tensor]:
        """[B, S] byte key-padding mask from either the collator's [B, S] mask or the reference's additive [B, 1, S, S] decoder
        mask (prepare_decoder_attention_mask): its last query row leaves exactly the non-padded keys unmasked."""
        if attention_mask is None:
            return None
        if attention_mask.dim() == 4:
            km = attention_mask[:, 0, -1, :] == 0
        else:
            km = attention_mask.reshape(B, S) != 0
        return km.to(device=device, dtype=torch.uint8).contiguous()

    @torch.no_grad()
    def backbone(self, input_embeds: torch.Tensor, hidden_states: torch.Tensor, cache_hidden, attention_mask, position_ids,
                 past_key_values=None, use_cache: bool = True) -> torch.Tensor:
        """One decoder-layer pass at TTT step j = len(cache_hidden[0]) (llama3_eagle.py:1625-1650, attention :717-785).  The K / V
        appended to `cache_hidden` are views of this step's fused, RoPE-rotated [q;k;v] buffer, which later steps read back."""
        from . import ops
        d = self.dims
        B, S, H = hidden_states.shape
        dev = self.embed_tokens_weight.device
        if cache_hidden is None:
            cache_hidden = [[], []]
        j = len(cache_hidden[0])
        if position_ids is not None:
            pos = position_ids.reshape(-1, S)[0]
            if not torch.equal(pos.to(dev), torch.arange(S, device=dev)):
                raise NotImplementedError("backbone: only the default position_ids = arange(seq) is implemented on the CUDA path")
        nh, nkv, hd = d.num_heads, d.num_kv_heads, d.head_dim
        A, KV = nh * hd, nkv * hd
        emb = input_embeds.to(dev, torch.bfloat16).reshape(B * S, H).contiguous()
        h = hidden_states.to(dev, torch.bfloat16).reshape(B * S, H).contiguous()
        xcat = torch.empty(B * S, 2 * H, dtype=torch.bfloat16, device=dev)
        ops.rmsnorm_fwd(emb, self._w("midlayer.input_layernorm.weight"), d.rms_norm_eps, out=xcat[:, :H])
        ops.rmsnorm_fwd(h, self._w("midlayer.hidden_norm.weight"), d.rms_norm_eps, out=xcat[:, H:])
        o0 = self.engine.offsets["midlayer.self_attn.q_proj.weight"]
        wqkv = self.engine.params[o0:o0 + (A + 2 * KV) * 2 * H].view(A + 2 * KV, 2 * H)     # physically fused [q;k;v]
        qkv = ops.gemm(xcat, wqkv)
        ops.rope_(qkv, nh + nkv, hd, self.engine.rope_cos, self.engine.rope_sin, S, j, inverse=False)
        blocks = []
        for kview in cache_hidden[0]:
            base = kview._base if kview._base is not None else kview
            if base.dim() != 2 or base.shape != (B * S, A + 2 * KV):
                raise ValueError("backbone: cache_hidden entries must be the K/V views this method appended at earlier steps")
            blocks.append(base)
        blocks.append(qkv)
        cache_hidden[0].append(qkv.view(B, S, nh + 2 * nkv, hd)[:, :, nh:nh + nkv].permute(0, 2, 1, 3))
        cache_hidden[1].append(qkv.view(B, S, nh + 2 * nkv, hd)[:, :, nh + nkv:].permute(0, 2, 1, 3))
        attn, _ = ops.ttt_attention_fwd(blocks, B, S, nh, nkv, hd, key_mask=self._key_mask(attention_mask, B, S, dev))
        hmid = ops.gemm(attn, self._w("midlayer.self_attn.o_proj.weight"), residual=h, epi=ops.EPI_BF16_RESID)
        hn2 = ops.rmsnorm_fwd(hmid, self._w("midlayer.post_attention_layernorm.weight"), d.rms_norm_eps)
        og = self.engine.offsets["midlayer.mlp.gate_proj.weight"]
        wgu = self.engine.params[og:og + 2 * d.intermediate_size * H].view(2 * d.intermediate_size, H)   # fused [gate;up]
        act = ops.swiglu_fwd(ops.gemm(hn2, wgu))
        out = ops.gemm(act, self._w("midlayer.mlp.down_proj.weight"), residual=hmid, epi=ops.EPI_BF16_RESID)
        return out.view(B, S, H)

    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor, inputs_embeds: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                ttt_length: int = 1):
        """LlamaForCausalLMEagle3.forward (llama3_eagle.py:1705-1757): fc -> one decoder layer -> norm, forward-only."""
        from . import ops
        B, S, _ = hidden_states.shape
        h = self.project_hidden_states(hidden_states)
        h = self.backbone(inputs_embeds, h, None if ttt_length == 1 else [[], []], attention_mask, None)
        return ops.rmsnorm_fwd(h.reshape(B * S, -1), self._w("norm.weight"), self.dims.rms_norm_eps).view(B, S, -1)
