"""Host side of the DFlash block-parallel draft step (first correct CUDA version behind
`sf_dflash_*`).  Mirrors what `OnlineDFlashModel` owns on the host in the reference: anchor sampling
(`algorithms/common/dflash_family_model.py:179-210`, torch RNG), the `DFlashDraftModel` state-dict names and shapes
(`modeling/draft/dflash.py:336-375`), rotary tables (transformers `Qwen3RotaryEmbedding`, default rope).  Everything after the
anchors are drawn runs in the CUDA library; there is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import dataclasses
from ctypes import POINTER, Structure, c_float, c_int32, c_int64, c_size_t, c_uint32, c_void_p
from typing import Dict, List, Optional, Tuple

import torch

from ._lib import check, lib

PER_LAYER = ("self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight", "self_attn.o_proj.weight",
             "mlp.gate_proj.weight", "mlp.up_proj.weight", "mlp.down_proj.weight", "self_attn.q_norm.weight",
             "self_attn.k_norm.weight", "input_layernorm.weight", "post_attention_layernorm.weight")
GLOBALS = ("fc.weight", "hidden_norm.weight", "norm.weight")


class SfDflashConfig(Structure):
    _fields_ = [(n, c_int32) for n in ("batch", "seq_len", "num_blocks", "block_size", "hidden_size", "num_target_feats",
                                       "intermediate", "num_heads", "num_kv_heads", "head_dim", "num_layers", "vocab",
                                       "mask_token_id", "rope_rows")] + [("rms_eps", c_float), ("loss_decay_gamma", c_float),
                                                                          ("grad_of_numerator", c_int32), ("loss_type", c_int32),
                                                                          ("dpace_alpha", c_float), ("sliding_window", c_int32),
                                                                          ("sliding_layers", c_uint32)]


class SfDflashFrozen(Structure):
    _fields_ = [("embed_tokens", c_void_p), ("lm_head", c_void_p), ("rope_cos", c_void_p), ("rope_sin", c_void_p)]


class SfDflashBatch(Structure):
    _fields_ = [("input_ids", c_void_p), ("hidden_states", c_void_p), ("loss_mask", c_void_p), ("anchors", c_void_p),
                ("keep", c_void_p)]


@dataclasses.dataclass
class DFlashDims:
    hidden_size: int
    intermediate_size: int
    num_heads: int
    num_kv_heads: int
    head_dim: int
    num_layers: int
    num_target_feats: int
    vocab_size: int
    block_size: int = 16
    mask_token_id: int = 0
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1000000.0
    max_position_embeddings: int = 40960
    loss_decay_gamma: Optional[float] = None
    loss_type: str = "dflash"         # "dflash" | "dpace" | "dpace-cumulative-confidence-only" | "dpace-continuation-value-only"
    dpace_alpha: float = 0.5          # (dflash_family_model.py:29-31,245-279)
    layer_types: Optional[Tuple[str, ...]] = None   # per layer "full_attention" | "sliding_attention" (modeling/draft/dflash.py:24-68)
    sliding_window: Optional[int] = None            # window of the sliding layers (dflash_family_model.py:73-84)

    def sliding_layout(self) -> Tuple[int, int]:
        """(window, bit mask of the sliding layers) — the validation of `resolve_dflash_attention_layout` (dflash.py:38-68)."""
        if self.layer_types is None:
            return 0, 0
        types = tuple(self.layer_types)
        if len(types) != self.num_layers:
            raise ValueError(f"DFlash config.layer_types must contain exactly num_hidden_layers={self.num_layers} entries, got {len(types)}")
        invalid = set(types) - {"full_attention", "sliding_attention"}
        if invalid:
            raise ValueError(f"DFlash config.layer_types supports only full_attention and sliding_attention, got {sorted(invalid)}")
        if "sliding_attention" not in types:
            return 0, 0
        if self.sliding_window is None or self.sliding_window <= 0:
            raise ValueError("DFlash sliding_attention layers require use_sliding_window=true and a positive config.sliding_window")
        return int(self.sliding_window), sum(1 << i for i, t in enumerate(types) if t == "sliding_attention")


LOSS_TYPES = {"dflash": 0, "dpace": 1, "dpace-cumulative-confidence-only": 2, "dpace-continuation-value-only": 3}


def sample_anchor_positions(loss_mask: torch.Tensor, num_anchors: int, generator: Optional[torch.Generator] = None,
                            fixed_width: bool = False):
    """Anchors whose clean token and first target are supervised (dflash_family_model.py:179-210); returns
    (anchors [B, W] sorted, 0 where dropped; keep [B, W] bool).  Same tensor ops and RNG call as the reference, so with the
    same generator state it draws the same anchors.

    The reference narrows W to min(num_anchors, max candidates in the batch), which costs a device->host sync per step
    (`.item()`).  `fixed_width=True` keeps W = num_anchors and lets `keep` drop the surplus blocks instead: the kept anchors,
    and therefore loss, metrics and gradients, are identical (dropped blocks contribute nothing), with no sync."""
    B, S = loss_mask.shape
    dev = loss_mask.device
    n_cand = max(S - 1, 0)
    valid = (loss_mask[:, :n_cand] > 0.5) & (loss_mask[:, 1:n_cand + 1] > 0.5)
    counts = valid.sum(dim=1)
    if fixed_width:
        width = min(num_anchors, n_cand)
        if width == 0:
            raise ValueError("DFlash-family training requires two consecutive supervised tokens")
    else:
        width = min(num_anchors, int(counts.max().item()))
        if width == 0:
            raise ValueError("DFlash-family training requires two consecutive supervised tokens")
    r = torch.rand(valid.shape, device=dev, generator=generator)
    r.masked_fill_(~valid, 2.0)
    cand = r.argsort(dim=1)[:, :width]
    keep = torch.arange(width, device=dev).unsqueeze(0) < counts.clamp(max=width).unsqueeze(1)
    sentinel = valid.shape[1]
    anchors = torch.where(keep, cand, torch.full_like(cand, sentinel)).sort(dim=1).values
    keep = anchors < sentinel
    return torch.where(keep, anchors, 0), keep


def rope_tables(dims: DFlashDims, rows: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos/sin of cat(freqs, freqs) in fp32, cast to bf16 (what Qwen3RotaryEmbedding hands a bf16 model)."""
    d = dims.head_dim
    inv = 1.0 / (dims.rope_theta ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
    f = torch.arange(rows, dtype=torch.float32).unsqueeze(1) * inv.unsqueeze(0)
    emb = torch.cat((f, f), dim=-1)
    return emb.cos().to(device, torch.bfloat16).contiguous(), emb.sin().to(device, torch.bfloat16).contiguous()


def _declare(l) -> None:
    if getattr(l, "_sf_dflash_declared", False):
        return
    l.sf_dflash_num_params.argtypes = [POINTER(SfDflashConfig)]
    l.sf_dflash_param_layout.argtypes = [POINTER(SfDflashConfig), POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]
    l.sf_dflash_workspace_bytes.restype = c_size_t
    l.sf_dflash_workspace_bytes.argtypes = [POINTER(SfDflashConfig)]
    l.sf_dflash_forward.argtypes = [POINTER(SfDflashConfig), c_void_p, POINTER(SfDflashFrozen), POINTER(SfDflashBatch), c_void_p, c_size_t,
                                    c_void_p, c_void_p, ctypes.c_int, c_void_p]
    l.sf_dflash_backward.argtypes = [POINTER(SfDflashConfig), c_void_p, POINTER(SfDflashFrozen), POINTER(SfDflashBatch), c_void_p, c_size_t,
                                     c_void_p, ctypes.c_int, c_void_p]
    l._sf_dflash_declared = True


class DFlashEngine:
    """Flat bf16 parameters + fp32 gradient accumulators + workspace for `sf_dflash_forward/backward`, bound to a maximum
    (batch, seq_len, num_blocks); the number of blocks of a call may be smaller (the reference's anchor width varies)."""

    def __init__(self, dims: DFlashDims, batch: int, seq_len: int, num_blocks: int, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("DFlashEngine needs a CUDA device (no CPU / PyTorch fallback)")
        self.dims, self.B, self.S, self.N = dims, batch, seq_len, num_blocks
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        L = lib()
        _declare(L)
        rope_rows = max(dims.max_position_embeddings, seq_len + dims.block_size) + dims.block_size
        self.cfg = self._cfg(batch, seq_len, num_blocks, rope_rows)
        n = L.sf_dflash_num_params(self.cfg)
        if n < 0:
            check(n, "sf_dflash_num_params")
        offs, sizes, total = (c_int64 * n)(), (c_int64 * n)(), c_int64()
        check(L.sf_dflash_param_layout(self.cfg, offs, sizes, ctypes.byref(total)), "sf_dflash_param_layout")
        self.names: List[str] = [f"layers.{l}.{p}" for l in range(dims.num_layers) for p in PER_LAYER] + list(GLOBALS)
        self.offsets = {nm: int(offs[i]) for i, nm in enumerate(self.names)}
        self.sizes = {nm: int(sizes[i]) for i, nm in enumerate(self.names)}
        self.n_params = int(total.value)
        self.params = torch.zeros(self.n_params, dtype=torch.bfloat16, device=self.device)
        self.grads_f32 = torch.zeros(self.n_params, dtype=torch.float32, device=self.device)
        self.rope_cos, self.rope_sin = rope_tables(dims, rope_rows, self.device)
        ws = L.sf_dflash_workspace_bytes(self.cfg)
        if ws == 0:
            check(-22, "sf_dflash_workspace_bytes")
        self.workspace_bytes = int(ws)
        self.workspace = torch.empty(self.workspace_bytes + 1024, dtype=torch.uint8, device=self.device)
        self._ws_ptr = (self.workspace.data_ptr() + 1023) // 1024 * 1024
        self.metrics = torch.zeros(4, dtype=torch.float32, device=self.device)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._frozen = None
        self._batch = None
        # optimizer state (same members as Eagle3Engine, so B200TrainingBackend drives either engine)
        self.grads_bf16 = torch.zeros(self.n_params, dtype=torch.bfloat16, device=self.device)
        self.master = self.exp_avg = self.exp_avg_sq = None
        self.opt_step = 0
        self._scratch = torch.zeros(1024, dtype=torch.float32, device=self.device)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=self.device)

    def _cfg(self, B, S, N, rope_rows) -> SfDflashConfig:
        d = self.dims
        return SfDflashConfig(B, S, N, d.block_size, d.hidden_size, d.num_target_feats, d.intermediate_size, d.num_heads,
                              d.num_kv_heads, d.head_dim, d.num_layers, d.vocab_size, d.mask_token_id, rope_rows, d.rms_norm_eps,
                              float(d.loss_decay_gamma) if d.loss_decay_gamma else 0.0, int(getattr(self, "grad_of_numerator", 0)),
                              LOSS_TYPES[d.loss_type], float(d.dpace_alpha), *d.sliding_layout())

    def shape_of(self, name: str) -> Tuple[int, ...]:
        d = self.dims
        A, KV = d.num_heads * d.head_dim, d.num_kv_heads * d.head_dim
        tail = name.split(".", 2)[2] if name.startswith("layers.") else name
        return {"self_attn.q_proj.weight": (A, d.hidden_size), "self_attn.k_proj.weight": (KV, d.hidden_size),
                "self_attn.v_proj.weight": (KV, d.hidden_size), "self_attn.o_proj.weight": (d.hidden_size, A),
                "mlp.gate_proj.weight": (d.intermediate_size, d.hidden_size), "mlp.up_proj.weight": (d.intermediate_size, d.hidden_size),
                "mlp.down_proj.weight": (d.hidden_size, d.intermediate_size), "self_attn.q_norm.weight": (d.head_dim,),
                "self_attn.k_norm.weight": (d.head_dim,), "input_layernorm.weight": (d.hidden_size,),
                "post_attention_layernorm.weight": (d.hidden_size,), "fc.weight": (d.hidden_size, d.num_target_feats * d.hidden_size),
                "hidden_norm.weight": (d.hidden_size,), "norm.weight": (d.hidden_size,)}[tail]

    def param_view(self, name: str, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        buf = self.params if buf is None else buf
        o, n = self.offsets[name], self.sizes[name]
        return buf[o:o + n].view(self.shape_of(name))

    def load_params(self, state: Dict[str, torch.Tensor]) -> None:
        for name in self.names:
            self.param_view(name).copy_(state[name].to(self.device, torch.bfloat16))

    def set_frozen(self, *, embed_tokens: torch.Tensor, lm_head: torch.Tensor) -> None:
        d = self.dims
        e = embed_tokens.to(self.device, torch.bfloat16).contiguous()
        h = lm_head.to(self.device, torch.bfloat16).contiguous()
        assert e.shape == (d.vocab_size, d.hidden_size) and h.shape == (d.vocab_size, d.hidden_size), (e.shape, h.shape)
        self._frozen_keep = (e, h)
        self._frozen = SfDflashFrozen(e.data_ptr(), h.data_ptr(), self.rope_cos.data_ptr(), self.rope_sin.data_ptr())

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _bind(self, batch: Dict[str, torch.Tensor], anchors: torch.Tensor, keep: torch.Tensor) -> None:
        dev, d = self.device, self.dims
        t = {"input_ids": batch["input_ids"].to(dev, torch.int64).contiguous(),
             "hidden_states": batch["hidden_states"].to(dev, torch.bfloat16).contiguous(),
             "loss_mask": (batch["loss_mask"].to(dev) > 0.5).to(torch.int64).contiguous(),
             "anchors": anchors.to(dev, torch.int32).contiguous(), "keep": keep.to(dev).to(torch.uint8).contiguous()}
        B, S = (int(v) for v in t["input_ids"].shape)
        N = int(t["anchors"].shape[1])
        if t["hidden_states"].shape != (B, S, d.num_target_feats * d.hidden_size):
            raise ValueError(f"hidden_states must be [{B}, {S}, {d.num_target_feats * d.hidden_size}], got {tuple(t['hidden_states'].shape)}")
        if t["loss_mask"].shape != (B, S) or t["anchors"].shape != (B, N) or t["keep"].shape != (B, N):
            raise ValueError("loss_mask / anchors / keep shapes do not match the batch")
        cfg = self._cfg(B, S, N, self.cfg.rope_rows)
        need = lib().sf_dflash_workspace_bytes(cfg)
        if need == 0:
            check(-22, "sf_dflash_workspace_bytes")
        if need > self.workspace_bytes:
            raise ValueError(f"batch [{B}, {S}] x {N} blocks exceeds the shape the engine was bound with")
        self._call_cfg = cfg
        self._keep = t
        self._batch = SfDflashBatch(t["input_ids"].data_ptr(), t["hidden_states"].data_ptr(), t["loss_mask"].data_ptr(),
                                    t["anchors"].data_ptr(), t["keep"].data_ptr())

    def forward(self, batch: Dict[str, torch.Tensor], anchors: torch.Tensor, keep: torch.Tensor, need_grad: bool = True):
        """Returns (loss[1], metrics[4] = loss_num, loss_den, correct, accuracy_den) as device tensors (no host sync)."""
        if self._frozen is None:
            raise RuntimeError("set_frozen() must be called before forward()")
        self._bind(batch, anchors, keep)
        check(lib().sf_dflash_forward(self._call_cfg, self.params.data_ptr(), self._frozen, self._batch, self._ws_ptr, self.workspace_bytes,
                                      self.metrics.data_ptr(), self.loss.data_ptr(), int(need_grad), self._stream()), "sf_dflash_forward")
        return self.loss, self.metrics

    def backward(self, accumulate: bool = False, loss_scale: float = 1.0, on_ready=None) -> None:
        """`loss_scale` must be 1 (the backend scales on the way to bf16); `on_ready` is accepted for interface parity with
        Eagle3Engine and ignored: the backend then all-reduces the whole buffer in step()."""
        if loss_scale != 1.0:
            raise ValueError("DFlashEngine.backward: loss_scale != 1 is not supported")
        if self._batch is None:
            raise RuntimeError("backward() without a preceding forward(need_grad=True)")
        check(lib().sf_dflash_backward(self._call_cfg, self.params.data_ptr(), self._frozen, self._batch, self._ws_ptr, self.workspace_bytes,
                                       self.grads_f32.data_ptr(), int(accumulate), self._stream()), "sf_dflash_backward")


    def grads_to_bf16(self, scale: Optional[torch.Tensor] = None, first: int = 0, count: Optional[int] = None) -> torch.Tensor:
        """fp32 accumulators -> the bf16 gradient buffer (optionally times a DEVICE scalar, no host sync)."""
        L = lib()
        L.sf_grads_to_bf16.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
        sp = None
        if scale is not None:
            scale = scale.detach().to(self.device, torch.float32).reshape(1)
            self._scale_keep = scale
            sp = scale.data_ptr()
        count = self.n_params - first if count is None else count
        check(L.sf_grads_to_bf16(self.grads_f32.data_ptr() + 4 * first, self.grads_bf16.data_ptr() + 2 * first, count, sp, self._stream()),
              "sf_grads_to_bf16")
        return self.grads_bf16

    def ensure_optimizer_state(self) -> None:
        if self.master is None:
            self.master = self.params.float()
            self.exp_avg = torch.zeros_like(self.master)
            self.exp_avg_sq = torch.zeros_like(self.master)

    def optimizer_step(self, lr: float, *, grad_scale: float = 1.0, max_grad_norm: float = 0.5, betas=(0.9, 0.999), eps: float = 1e-8,
                       weight_decay: float = 0.0) -> torch.Tensor:
        """Fused clip + AdamW on the flat buffers (optimizer.py:95-168 semantics, see sf_optimizer_step)."""
        L = lib()
        L.sf_optimizer_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                                        c_float, c_float, c_float, ctypes.c_int, c_void_p, c_void_p, c_void_p]
        self.ensure_optimizer_state()
        self.opt_step += 1
        check(L.sf_optimizer_step(self.grads_bf16.data_ptr(), self.master.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                  self.params.data_ptr(), self.n_params, grad_scale, max_grad_norm, lr, betas[0], betas[1], eps, weight_decay,
                                  self.opt_step, self.grad_norm.data_ptr(), self._scratch.data_ptr(), self._stream()), "sf_optimizer_step")
        return self.grad_norm


# ------------------------------------------------------------------------------------------------ draft-registry seam
def dims_from_config(config) -> DFlashDims:
    """DFlashDims from a Qwen3Config-like object or the dict of a draft JSON (configs/qwen3-8b-dflash.json)."""
    get = (lambda k, d=None: config.get(k, d)) if isinstance(config, dict) else (lambda k, d=None: getattr(config, k, d))
    H, nh, L = get("hidden_size"), get("num_attention_heads"), get("num_hidden_layers")
    dcfg = get("dflash_config") or {}
    layer_ids = dcfg.get("target_layer_ids")
    if layer_ids is None:          # build_target_layer_ids (modeling/draft/dflash.py:271-281)
        nt = get("num_target_layers")
        layer_ids = [nt // 2] if L == 1 else [int(round(1 + i * (nt - 4) / (L - 1))) for i in range(L)]
    types = get("layer_types")
    rope_params = get("rope_parameters")
    theta = rope_params["rope_theta"] if rope_params else get("rope_theta", 1000000.0)
    for rp in (get("rope_scaling"), rope_params):       # transformers 5 mirrors rope_parameters into rope_scaling
        if rp and rp.get("rope_type", rp.get("type", "default")) not in (None, "default"):
            raise NotImplementedError("rope scaling is not implemented for the DFlash CUDA path")
    return DFlashDims(hidden_size=H, intermediate_size=get("intermediate_size"), num_heads=nh, num_kv_heads=get("num_key_value_heads", nh),
                      head_dim=get("head_dim") or H // nh, num_layers=L, num_target_feats=len(layer_ids), vocab_size=get("vocab_size"),
                      block_size=get("block_size", 16), mask_token_id=int(dcfg.get("mask_token_id") or 0),
                      rms_norm_eps=get("rms_norm_eps", 1e-6), rope_theta=float(theta),
                      max_position_embeddings=get("max_position_embeddings", 40960),
                      layer_types=tuple(types) if types else None, sliding_window=get("sliding_window"))


try:  # the reference's registry needs `config_class` (registry.py:38-41); its DFlash draft uses Qwen3Config
    from transformers.models.qwen3.modeling_qwen3 import Qwen3Config as _Qwen3Config
except Exception:  # pragma: no cover
    _Qwen3Config = None


class B200DFlashDraftModel(torch.nn.Module):
    """State-dict twin of `DFlashDraftModel` (modeling/draft/dflash.py:336-375) for the `@register_draft` seam: same
    parameter names and shapes, `target_layer_ids`, `block_size`, `mask_token_id`; the parameters are views into the
    DFlashEngine's flat bf16 buffer once `bind_engine()` ran.  No eager forward: compute goes through `sf_dflash_*`."""
    architectures = ["DFlashDraftModel"]
    config_class = _Qwen3Config

    def __init__(self, config, dflash_kernels=None):
        super().__init__()
        self.config = config
        self.dims = dims_from_config(config)
        get = (lambda k, d=None: config.get(k, d)) if isinstance(config, dict) else (lambda k, d=None: getattr(config, k, d))
        dcfg = get("dflash_config") or {}
        self.block_size = self.dims.block_size
        self.mask_token_id = dcfg.get("mask_token_id", None)
        self.target_layer_ids = dcfg.get("target_layer_ids")
        self.sliding_window = None
        self.engine: Optional[DFlashEngine] = None
        self._names = [f"layers.{l}.{p}" for l in range(self.dims.num_layers) for p in PER_LAYER] + list(GLOBALS)
        self._flat: Dict[str, torch.nn.Parameter] = {}

    def state_dict_spec(self) -> Dict[str, tuple]:
        d = self.dims
        A, KV, H, I = d.num_heads * d.head_dim, d.num_kv_heads * d.head_dim, d.hidden_size, d.intermediate_size
        per = {"self_attn.q_proj.weight": (A, H), "self_attn.k_proj.weight": (KV, H), "self_attn.v_proj.weight": (KV, H),
               "self_attn.o_proj.weight": (H, A), "mlp.gate_proj.weight": (I, H), "mlp.up_proj.weight": (I, H),
               "mlp.down_proj.weight": (H, I), "self_attn.q_norm.weight": (d.head_dim,), "self_attn.k_norm.weight": (d.head_dim,),
               "input_layernorm.weight": (H,), "post_attention_layernorm.weight": (H,)}
        spec = {f"layers.{l}.{k}": v for l in range(d.num_layers) for k, v in per.items()}
        spec.update({"fc.weight": (H, d.num_target_feats * H), "hidden_norm.weight": (H,), "norm.weight": (H,)})
        return spec

    def bind_engine(self, batch: int, seq_len: int, num_anchors: int = 512, device=None, init_std: float = 0.02, seed: int = 0) -> DFlashEngine:
        eng = DFlashEngine(self.dims, batch=batch, seq_len=seq_len, num_blocks=num_anchors, device=device)
        g = torch.Generator(device="cpu").manual_seed(seed)
        for name in self._names:
            view = eng.param_view(name)
            if view.dim() == 1:
                view.fill_(1.0)
            else:
                view.copy_((torch.randn(view.shape, generator=g) * init_std).to(torch.bfloat16))
            p = torch.nn.Parameter(view, requires_grad=True)
            self._flat[name] = p
            self.register_parameter(name.replace(".", "__"), p)
        self.engine = eng
        return eng

    def state_dict(self, *args, destination=None, prefix: str = "", keep_vars: bool = False):
        out = destination if destination is not None else {}
        for name in self._names:
            if name in self._flat:
                p = self._flat[name]
                out[prefix + name] = p if keep_vars else p.detach()
        return out

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        if self.engine is None:
            raise RuntimeError("bind_engine() must be called before load_state_dict()")
        missing = [n for n in self._names if n not in state_dict]
        unexpected = [k for k in state_dict if k not in self._names and "rotary_emb" not in k]
        if strict and (missing or unexpected):
            raise RuntimeError(f"state_dict mismatch: missing={missing} unexpected={unexpected}")
        self.engine.load_params({n: state_dict[n] for n in self._names if n in state_dict} | {n: self._flat[n] for n in missing})
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)


# ------------------------------------------------------------------------------------------------ strategy seam
class _DFlashStepFn(torch.autograd.Function):
    """(loss, loss_num) = f(flat params): forward / backward run in the CUDA library, the fp32 flat gradient stays in the engine.
    Exactly one of the two outputs may be backpropagated, fixed by the strategy's normalisation mode: `loss` = loss_num / loss_den
    when the engine differentiates the ratio ("local"), `loss_num` when it differentiates the numerator for the reference
    controller's loss_terms path ("global", controller.py:334-398)."""

    @staticmethod
    def forward(ctx, flat_params, strategy, batch_tensors, anchors, keep, need_grad):
        loss, metrics = strategy.engine.forward(batch_tensors, anchors, keep, need_grad=need_grad)
        ctx.strategy = strategy
        return loss.clone().reshape(()), metrics[0].clone().reshape(())

    @staticmethod
    def backward(ctx, grad_loss, grad_num):
        st = ctx.strategy
        want_num = st.normalization == "global"
        g = grad_num if want_num else grad_loss
        other = grad_loss if want_num else grad_num
        if other is not None and bool((other != 0).any()):
            raise RuntimeError(f"DFlash strategy in normalization={st.normalization!r} mode: backpropagate "
                               f"{'loss_terms[0] (the numerator)' if want_num else 'StepOutput.loss'}, not the other output")
        st._last_grad_out = g.detach()
        st.engine.backward(accumulate=st._micro_in_window > 0)
        st._micro_in_window += 1
        return None, None, None, None, None, None


class _DFlashTrainable(torch.nn.Module):
    def __init__(self, draft_model):
        super().__init__()
        self.draft_model = draft_model


class B200DFlashTrainStrategy:
    """`DFlashTrainStrategy` (training/strategies/base.py:415-452) on the CUDA step: same name, required features and
    `StepOutput` fields (`accuracy`, `accuracy_denom`, `ratio_metrics["acc"]`, `loss_terms`).  Anchors are drawn here with
    the reference's procedure (`OnlineDFlashModel._sample_anchor_positions`), on the batch's device, from `generator`."""
    name = "dflash"
    required_features = {"input_ids", "hidden_states", "loss_mask"}

    def __init__(self, draft_model: "B200DFlashDraftModel", *, target_embed_weight: torch.Tensor, target_head_weight: torch.Tensor,
                 num_anchors: int = 512, generator: Optional[torch.Generator] = None, sync_free_anchors: bool = True,
                 normalization: str = "local"):
        """normalization: "local" — the step's loss is loss_num / loss_den of THIS micro-batch and `loss_terms` is not reported, so the
        reference controller treats it like any scalar loss (mean over accumulation steps and ranks); "global" — the reference's
        DFlash contract: `loss_terms = (loss_num with grad, loss_den)`, the engine differentiates the numerator and the controller
        divides the synchronised gradients by the global denominator (training/strategies/base.py:415-452, controller.py:334-398)."""
        if draft_model.engine is None:
            raise RuntimeError("bind_engine() must be called on the draft model first")
        if normalization not in ("local", "global"):
            raise ValueError("normalization must be 'local' or 'global'")
        self.normalization = normalization
        self.draft_model, self.engine = draft_model, draft_model.engine
        self.engine.grad_of_numerator = int(normalization == "global")
        self.num_anchors, self.generator = num_anchors, generator
        self.sync_free_anchors = sync_free_anchors      # see sample_anchor_positions(fixed_width=...)
        self.engine.set_frozen(embed_tokens=target_embed_weight, lm_head=target_head_weight)
        self._module = _DFlashTrainable(draft_model)
        self._micro_in_window = 0
        self._last_grad_out: Optional[torch.Tensor] = None
        self._is_boundary = True
        self.grad_ready_hook = None
        self.return_autograd_grads = False

    def trainable_module(self) -> torch.nn.Module:
        return self._module

    def validate_batch(self, batch) -> None:
        missing = {f for f in self.required_features if f not in batch.tensors}
        if missing:
            raise ValueError(f"{self.name} batch missing required features {sorted(missing)}; present={sorted(batch.tensors)}")

    def forward_loss(self, batch, ctx=None):
        from .contracts import StepOutput
        self.validate_batch(batch)
        t = batch.tensors
        anchors, keep = sample_anchor_positions(t["loss_mask"].to(self.engine.device).float(), self.num_anchors, generator=self.generator,
                                                fixed_width=self.sync_free_anchors)
        need_grad = torch.is_grad_enabled()
        eng = self.engine
        flat = eng.params if not need_grad else eng.params.detach().requires_grad_(True)
        loss, loss_num = _DFlashStepFn.apply(flat, self, t, anchors, keep, need_grad)
        m = eng.metrics.clone()            # loss_num, loss_den, correct, accuracy_den (device, no host sync)
        if not self.sync_free_anchors and float(m[1]) <= 0:
            raise ValueError("DFlash batch has no supervised draft position (need two consecutive loss-mask tokens)")
        terms = (loss_num, m[1]) if self.normalization == "global" else None
        return StepOutput(loss=loss, metrics={"accuracy": m[2] / m[3].clamp_min(1e-6), "accuracy_denom": m[3]},
                          ratio_metrics={"acc": (m[2], m[3])}, loss_terms=terms)

    def checkpoint_state_filter(self, state_dict):
        return {k.replace("draft_model.", ""): v for k, v in state_dict.items() if "draft_model." in k}
