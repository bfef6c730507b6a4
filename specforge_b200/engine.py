"""Host-side driver of the C-ABI EAGLE3 step (device memory + streams come from PyTorch; compute is libspecforge_b200).

`Eagle3Engine` owns, for one draft head on one GPU:
  * the flat bf16 parameter buffer (the nn.Parameters of the draft module are views into it, so q/k/v and
    gate/up are physically adjacent and the reference state-dict names/shapes are preserved),
  * the flat fp32 gradient accumulator and the flat bf16 gradient buffer that is all-reduced,
  * fp32 AdamW state (masters, exp_avg, exp_avg_sq),
  * the step workspace (activations of all TTT steps; sized by sf_eagle3_workspace_bytes).
"""
from __future__ import annotations

import ctypes
import math
from ctypes import POINTER, Structure, c_float, c_int32, c_int64, c_size_t, c_uint8, c_void_p
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from ._lib import check, lib

P_NAMES = [
    "fc.weight",
    "midlayer.self_attn.q_proj.weight",
    "midlayer.self_attn.k_proj.weight",
    "midlayer.self_attn.v_proj.weight",
    "midlayer.self_attn.o_proj.weight",
    "midlayer.mlp.gate_proj.weight",
    "midlayer.mlp.up_proj.weight",
    "midlayer.mlp.down_proj.weight",
    "midlayer.hidden_norm.weight",
    "midlayer.input_layernorm.weight",
    "midlayer.post_attention_layernorm.weight",
    "norm.weight",
    "lm_head.weight",
    "fc_norm.0.weight",
    "fc_norm.1.weight",
    "fc_norm.2.weight",
]
P_COUNT = 16


class SfConfig(Structure):
    _fields_ = [
        ("batch", c_int32), ("seq_len", c_int32), ("ttt_length", c_int32), ("hidden_size", c_int32),
        ("target_hidden", c_int32), ("intermediate", c_int32), ("num_heads", c_int32), ("num_kv_heads", c_int32),
        ("head_dim", c_int32), ("vocab", c_int32), ("draft_vocab", c_int32), ("fc_norm", c_int32),
        ("norm_output", c_int32), ("rope_rows", c_int32), ("rms_eps", c_float), ("ploss_decay", c_float),
        ("lk_loss_type", c_int32), ("kl_scale", c_float), ("kl_decay", c_float),
    ]


class SfFrozen(Structure):
    _fields_ = [("embed_tokens", c_void_p), ("target_head", c_void_p), ("rope_cos", c_void_p), ("rope_sin", c_void_p),
                ("t2d", c_void_p), ("d2t", c_void_p)]


class SfBatch(Structure):
    _fields_ = [("input_ids", c_void_p), ("attention_mask", c_void_p), ("loss_mask", c_void_p),
                ("hidden_state", c_void_p), ("target", c_void_p)]


GRAD_READY_FN = ctypes.CFUNCTYPE(None, c_int32, c_int32, c_void_p)

_declared = False


def _declare():
    global _declared
    if _declared:
        return
    l = lib()
    l.sf_eagle3_param_layout.restype = ctypes.c_int
    l.sf_eagle3_param_layout.argtypes = [POINTER(SfConfig), POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]
    l.sf_eagle3_workspace_bytes.restype = c_size_t
    l.sf_eagle3_workspace_bytes.argtypes = [POINTER(SfConfig)]
    l.sf_eagle3_forward.restype = ctypes.c_int
    l.sf_eagle3_forward.argtypes = [POINTER(SfConfig), c_void_p, POINTER(SfFrozen), POINTER(SfBatch), c_void_p, c_size_t,
                                    c_void_p, c_void_p, ctypes.c_int, c_void_p]
    l.sf_eagle3_backward.restype = ctypes.c_int
    l.sf_eagle3_backward.argtypes = [POINTER(SfConfig), c_void_p, POINTER(SfFrozen), POINTER(SfBatch), c_void_p, c_size_t,
                                     c_float, c_void_p, ctypes.c_int, c_void_p]
    l.sf_eagle3_backward_ex.restype = ctypes.c_int
    l.sf_eagle3_backward_ex.argtypes = [POINTER(SfConfig), c_void_p, POINTER(SfFrozen), POINTER(SfBatch), c_void_p, c_size_t,
                                        c_float, c_void_p, ctypes.c_int, GRAD_READY_FN, c_void_p, c_void_p]
    l.sf_eagle3_workspace_view.restype = ctypes.c_int
    l.sf_eagle3_workspace_view.argtypes = [POINTER(SfConfig), ctypes.c_char_p, POINTER(c_int64), POINTER(c_int64)]
    l.sf_grads_to_bf16.restype = ctypes.c_int
    l.sf_grads_to_bf16.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
    l.sf_optimizer_step.restype = ctypes.c_int
    l.sf_optimizer_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float,
                                    c_float, c_float, c_float, c_float, c_int32, c_void_p, c_void_p, c_void_p]
    _declared = True


@dataclass
class DraftDims:
    """Shape-defining fields of the draft config (configs/*-eagle3.json in the reference)."""
    hidden_size: int
    intermediate_size: int
    num_heads: int
    num_kv_heads: int
    head_dim: int
    vocab_size: int
    draft_vocab_size: int
    target_hidden_size: Optional[int] = None
    rms_norm_eps: float = 1e-6
    rope_theta: float = 10000.0
    max_position_embeddings: int = 2048
    fc_norm: bool = False
    norm_output: bool = True
    rope_scaling: Optional[dict] = None

    def __post_init__(self):
        if self.target_hidden_size is None:
            self.target_hidden_size = self.hidden_size


def _rope_inv_freq_and_scale(dims: "DraftDims", rows: int):
    """inv_freq / position divisor / table multiplier for every RoPE flavour the reference supports through
    `rope_scaling` (llama3_eagle.py:218-536), restated: default, "linear", "dynamic" (NTK), "llama3", "yarn".
    "mrope" (three-axis multimodal positions) is not a table transform and is not implemented."""
    d, base = dims.head_dim, float(dims.rope_theta)
    idx = torch.arange(0, d, 2, dtype=torch.float32)
    inv_freq = 1.0 / (base ** (idx / d))
    t_div, mult = 1.0, 1.0
    sc = dims.rope_scaling
    if not sc:
        return inv_freq, t_div, mult
    kind = sc.get("rope_type", sc.get("type"))
    factor = sc.get("factor")
    if kind in (None, "default"):
        pass
    elif kind == "linear":
        if factor is None:
            raise ValueError("Linear RoPE scaling requires 'factor' in rope_scaling config.")
        t_div = float(factor)
    elif kind == "dynamic":
        if factor is None:
            raise ValueError("Dynamic RoPE scaling requires 'factor' in rope_scaling config.")
        if rows > dims.max_position_embeddings:      # NTK: the base grows with the cached length
            nb = base * ((factor * rows / dims.max_position_embeddings) - (factor - 1)) ** (d / (d - 2))
            inv_freq = 1.0 / (nb ** (idx / d))
    elif kind == "llama3":
        f = 1.0 if factor is None else float(factor)
        lo, hi, orig = sc.get("low_freq_factor"), sc.get("high_freq_factor"), sc.get("original_max_position_embeddings")
        if None not in (lo, hi, orig):
            wave = 2 * math.pi / inv_freq
            smooth = (orig / wave - lo) / (hi - lo) if lo != hi else 0
            mid = (1 - smooth) * inv_freq / f + smooth * inv_freq
            inv_freq = torch.where(wave < orig / hi, inv_freq, torch.where(wave > orig / lo, inv_freq / f, mid))
    elif kind == "yarn":
        f = float(factor)
        orig = sc.get("original_max_position_embeddings")
        bf, bs = sc.get("beta_fast") or 32, sc.get("beta_slow") or 1
        ms, msa = sc.get("mscale") or 1, sc.get("mscale_all_dim") or 0

        def corr_dim(rot):
            return (d * math.log(orig / (rot * 2 * math.pi))) / (2 * math.log(base))

        low, high = max(math.floor(corr_dim(bf)), 0), min(math.ceil(corr_dim(bs)), d - 1)
        if low == high:
            high += 0.001
        ramp = torch.clamp((torch.arange(d // 2, dtype=torch.float32) - low) / (high - low), 0, 1)
        keep = 1.0 - ramp
        inv_freq = (inv_freq / f) * (1 - keep) + inv_freq * keep

        def mscale(scale, m):
            return 1.0 if scale <= 1 else 0.1 * m * math.log(scale) + 1.0

        mult = float(mscale(f, ms) / mscale(f, msa))
    else:
        raise NotImplementedError(f"rope_scaling type {kind!r} is not implemented on the CUDA path")
    return inv_freq, t_div, mult


def rope_tables(dims: "DraftDims", rows: int, device) -> tuple:
    """cos/sin tables as the reference's rotary modules build them (fp32 math) followed by the module-wide
    .to(bfloat16) (algorithms/model_providers.py:112).  `rows` >= S + T."""
    inv_freq, t_div, mult = _rope_inv_freq_and_scale(dims, rows)
    t = torch.arange(rows, dtype=torch.float32) / t_div
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos() * mult, emb.sin() * mult
    return cos.to(torch.bfloat16).to(device).contiguous(), sin.to(torch.bfloat16).to(device).contiguous()


class Eagle3Engine:
    LK_TYPES = {None: 0, "lambda": 1, "alpha": 2}

    def __init__(self, dims: DraftDims, *, batch: int, seq_len: int, ttt_length: int = 7, ploss_decay: float = 0.8,
                 lk_loss_type: Optional[str] = None, kl_scale: float = 1.0, kl_decay: float = 1.0,
                 device: Optional[torch.device] = None):
        _declare()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if device.type != "cuda":
            raise RuntimeError("specforge_b200 has no CPU path: a CUDA (sm_100a) device is required")
        if lk_loss_type not in self.LK_TYPES:
            raise ValueError(f"Unknown lk loss type: {lk_loss_type}")
        self.dims, self.device = dims, device
        self.B, self.S, self.T = batch, seq_len, ttt_length
        self.ploss_decay = ploss_decay
        rope_rows = max(dims.max_position_embeddings + 20, seq_len + ttt_length)
        self.cfg = SfConfig(batch, seq_len, ttt_length, dims.hidden_size, dims.target_hidden_size, dims.intermediate_size,
                            dims.num_heads, dims.num_kv_heads, dims.head_dim, dims.vocab_size, dims.draft_vocab_size,
                            int(dims.fc_norm), int(dims.norm_output), rope_rows, dims.rms_norm_eps, ploss_decay,
                            self.LK_TYPES[lk_loss_type], kl_scale, kl_decay)
        offs = (c_int64 * P_COUNT)()
        sizes = (c_int64 * P_COUNT)()
        total = c_int64()
        check(lib().sf_eagle3_param_layout(self.cfg, offs, sizes, ctypes.byref(total)), "sf_eagle3_param_layout")
        self.offsets = {P_NAMES[i]: int(offs[i]) for i in range(P_COUNT) if sizes[i] > 0}
        self.sizes = {P_NAMES[i]: int(sizes[i]) for i in range(P_COUNT) if sizes[i] > 0}
        self.n_params = int(total.value)
        self.params = torch.zeros(self.n_params, dtype=torch.bfloat16, device=device)
        self.grads_f32 = torch.zeros(self.n_params, dtype=torch.float32, device=device)
        self.grads_bf16 = torch.zeros(self.n_params, dtype=torch.bfloat16, device=device)
        self.master = None
        self.exp_avg = None
        self.exp_avg_sq = None
        self.opt_step = 0
        self._scratch = torch.zeros(1024, dtype=torch.float32, device=device)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=device)
        self.rope_cos, self.rope_sin = rope_tables(dims, rope_rows, device)
        ws_bytes = lib().sf_eagle3_workspace_bytes(self.cfg)
        if ws_bytes == 0:
            check(-22, "sf_eagle3_workspace_bytes")
        self.workspace_bytes = int(ws_bytes)
        self.workspace = torch.empty(self.workspace_bytes + 1024, dtype=torch.uint8, device=device)
        self._ws_ptr = (self.workspace.data_ptr() + 1023) // 1024 * 1024
        self.metrics = torch.zeros(ttt_length, 8, dtype=torch.float32, device=device)
        self.loss = torch.zeros(1, dtype=torch.float32, device=device)
        self.frozen_tensors: Dict[str, torch.Tensor] = {}
        self._frozen = None
        self._batch = None
        self._call_cfg = self.cfg
        self._batch_keep = None
        self._grads_dirty = False

    # ------------------------------------------------------------------ parameters
    def shape_of(self, name: str):
        d = self.dims
        H, I, hd = d.hidden_size, d.intermediate_size, d.head_dim
        return {
            "fc.weight": (H, 3 * d.target_hidden_size),
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
            "fc_norm.0.weight": (d.target_hidden_size,),
            "fc_norm.1.weight": (d.target_hidden_size,),
            "fc_norm.2.weight": (d.target_hidden_size,),
        }[name]

    def param_view(self, name: str, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        buf = self.params if buf is None else buf
        o, n = self.offsets[name], self.sizes[name]
        return buf[o:o + n].view(self.shape_of(name))

    def load_params(self, state: Dict[str, torch.Tensor]) -> None:
        for name in self.offsets:
            self.param_view(name).copy_(state[name].to(self.device, torch.bfloat16))
        self.master = None

    def set_frozen(self, *, embed_tokens: torch.Tensor, target_head: torch.Tensor, t2d: torch.Tensor, d2t: torch.Tensor) -> None:
        dev = self.device
        ft = {
            "embed_tokens": embed_tokens.to(dev, torch.bfloat16).contiguous(),
            "target_head": target_head.to(dev, torch.bfloat16).contiguous(),
            "t2d": t2d.to(dev).to(torch.uint8).contiguous(),
            "d2t": d2t.to(dev, torch.int64).contiguous(),
        }
        d = self.dims
        assert ft["embed_tokens"].shape == (d.vocab_size, d.hidden_size), ft["embed_tokens"].shape
        assert ft["target_head"].shape == (d.vocab_size, d.target_hidden_size), ft["target_head"].shape
        assert ft["t2d"].shape == (d.vocab_size,) and ft["d2t"].shape == (d.draft_vocab_size,)
        if int(ft["t2d"].sum().item()) != d.draft_vocab_size:
            raise ValueError("t2d must select exactly draft_vocab_size target tokens")
        self.frozen_tensors = ft
        self._frozen = SfFrozen(ft["embed_tokens"].data_ptr(), ft["target_head"].data_ptr(), self.rope_cos.data_ptr(),
                                self.rope_sin.data_ptr(), ft["t2d"].data_ptr(), ft["d2t"].data_ptr())

    # ------------------------------------------------------------------ step
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _bind_batch(self, batch: Dict[str, torch.Tensor]) -> None:
        B, S, d = self.B, self.S, self.dims
        dev = self.device
        t = {
            "input_ids": batch["input_ids"].to(dev, torch.int64, non_blocking=True).contiguous(),
            "loss_mask": batch["loss_mask"].to(dev, torch.int64, non_blocking=True).contiguous(),
            "hidden_state": batch["hidden_state"].to(dev, torch.bfloat16, non_blocking=True).contiguous(),
            "target": batch["target"].to(dev, torch.bfloat16, non_blocking=True).contiguous(),
        }
        am = batch.get("attention_mask")
        t["attention_mask"] = None if am is None else am.to(dev, torch.int64, non_blocking=True).contiguous()
        pinned = [v for v in batch.values() if isinstance(v, torch.Tensor) and not v.is_cuda and v.is_pinned()]
        if pinned:   # asynchronous copies out of reusable pinned buffers: tell their producer when they have executed
            from .feed import mark_copied
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            for v in pinned:
                mark_copied(v, ev)
        if t["input_ids"].dim() != 2:
            raise ValueError(f"input_ids must be [batch, seq], got {tuple(t['input_ids'].shape)}")
        # The engine is bound to a MAXIMUM (batch, seq_len): the reference collator pads each batch to its own longest
        # sample (data/utils.py:122) and the loss is a mean over all B*S rows (core/loss.py), so the batch's own shape
        # has to reach the kernels.  The C ABI takes the shape per call; only the workspace bound is fixed.
        Bb, Sb = (int(v) for v in t["input_ids"].shape)
        if (Bb, Sb) == (B, S):
            self._call_cfg = self.cfg
        else:
            cfg = SfConfig.from_buffer_copy(self.cfg)
            cfg.batch, cfg.seq_len = Bb, Sb
            need = lib().sf_eagle3_workspace_bytes(cfg)
            if need == 0:
                check(-22, "sf_eagle3_workspace_bytes")
            if need > self.workspace_bytes or Sb + self.T > self.cfg.rope_rows:
                raise ValueError(f"batch [{Bb}, {Sb}] exceeds the shape the engine was bound with ([{B}, {S}])")
            self._call_cfg = cfg
        B, S = Bb, Sb
        if t["loss_mask"].numel() != B * S:
            raise ValueError(f"loss_mask must have {B * S} elements, got {tuple(t['loss_mask'].shape)}")
        if t["attention_mask"] is not None and t["attention_mask"].numel() != B * S:
            raise ValueError(f"attention_mask must have {B * S} elements, got {tuple(t['attention_mask'].shape)}")
        if t["hidden_state"].shape != (B, S, 3 * d.target_hidden_size):
            raise ValueError(f"hidden_state must be [{B}, {S}, {3 * d.target_hidden_size}], got {tuple(t['hidden_state'].shape)}")
        if t["target"].shape != (B, S, d.target_hidden_size):
            raise ValueError(f"target must be [{B}, {S}, {d.target_hidden_size}], got {tuple(t['target'].shape)}")
        self._batch_keep = t
        self._batch = SfBatch(t["input_ids"].data_ptr(), 0 if t["attention_mask"] is None else t["attention_mask"].data_ptr(),
                              t["loss_mask"].data_ptr(), t["hidden_state"].data_ptr(), t["target"].data_ptr())

    def forward(self, batch: Dict[str, torch.Tensor], need_grad: bool = True):
        """Teacher + TTT forward + loss.  Returns (loss[1], metrics[T, 8]) device tensors (no host sync)."""
        if self._frozen is None:
            raise RuntimeError("set_frozen() must be called before forward()")
        self._bind_batch(batch)
        check(lib().sf_eagle3_forward(self._call_cfg, self.params.data_ptr(), self._frozen, self._batch, self._ws_ptr,
                                      self.workspace_bytes, self.metrics.data_ptr(), self.loss.data_ptr(), int(need_grad),
                                      self._stream()), "sf_eagle3_forward")
        return self.loss, self.metrics

    _VIEW_DTYPES = {"teacher_stats": torch.float32, "teacher_ids": torch.int64, "position_mask": torch.int32}

    def workspace_view(self, name: str) -> torch.Tensor:
        """A named tensor of the LAST step inside the workspace (sf_eagle3_workspace_view), shaped for the batch that ran:
        "logits" [T, B*S, DV] (logits after forward(need_grad=False); d(loss)/d(logits) after need_grad=True), "h" [T+1, B*S, H],
        "qkv", "attn", "hf", "teacher_xg" [B, S+T, DV], "teacher_stats" [B, S+T, 4], "teacher_ids" [B, S+T], "position_mask" [B, S].
        A view, not a copy: it is overwritten by the next step."""
        off, size = c_int64(), c_int64()
        cfg = self._call_cfg
        check(lib().sf_eagle3_workspace_view(cfg, name.encode(), ctypes.byref(off), ctypes.byref(size)), "sf_eagle3_workspace_view")
        base = self._ws_ptr - self.workspace.data_ptr()
        raw = self.workspace[base + off.value: base + off.value + size.value]
        dt = self._VIEW_DTYPES.get(name, torch.bfloat16)
        t = raw.view(dt)
        d, B, S, T = self.dims, cfg.batch, cfg.seq_len, self.T
        M = B * S
        A, KV = d.num_heads * d.head_dim, d.num_kv_heads * d.head_dim
        shape = {"h": (T + 1, M, d.hidden_size), "qkv": (T, M, A + 2 * KV), "attn": (T, M, A), "hf": (T, M, d.hidden_size),
                 "logits": (T, M, d.draft_vocab_size), "teacher_xg": (B, S + T, d.draft_vocab_size), "teacher_stats": (B, S + T, 4),
                 "teacher_ids": (B, S + T), "position_mask": (B, S)}[name]
        return t.view(shape)

    def backward(self, loss_scale: float = 1.0, accumulate: bool = False, on_ready=None) -> None:
        """Full backward.  on_ready(first_elem, n_elems) is called (on the host, in stream order of the producing
        kernels) as soon as each contiguous slice [first_elem, first_elem + n_elems) of the flat gradient is complete."""
        if self._batch is None:
            raise RuntimeError("backward() without a preceding forward(need_grad=True)")
        if on_ready is None:
            check(lib().sf_eagle3_backward(self._call_cfg, self.params.data_ptr(), self._frozen, self._batch, self._ws_ptr,
                                           self.workspace_bytes, loss_scale, self.grads_f32.data_ptr(), int(accumulate),
                                           self._stream()), "sf_eagle3_backward")
        else:
            names = P_NAMES
            errors = []

            def _cb(first, n, _user):
                try:
                    lo = self.offsets[names[first]]
                    hi = self.offsets[names[first + n - 1]] + self.sizes[names[first + n - 1]]
                    on_ready(lo, hi - lo)
                except Exception as exc:  # never let an exception cross the C boundary
                    errors.append(exc)

            cb = GRAD_READY_FN(_cb)
            check(lib().sf_eagle3_backward_ex(self._call_cfg, self.params.data_ptr(), self._frozen, self._batch, self._ws_ptr,
                                              self.workspace_bytes, loss_scale, self.grads_f32.data_ptr(), int(accumulate),
                                              cb, None, self._stream()), "sf_eagle3_backward_ex")
            if errors:
                raise errors[0]
        self._grads_dirty = True

    def grads_to_bf16(self, scale: Optional[torch.Tensor] = None, first: int = 0, count: Optional[int] = None) -> torch.Tensor:
        """fp32 accumulators -> the bf16 gradient buffer (optionally times a DEVICE scalar, no host sync); `first`/`count`
        restrict the conversion to a slice (used when slices are all-reduced as they become ready)."""
        sp = 0
        if scale is not None:
            scale = scale.detach().to(self.device, torch.float32).reshape(1)
            self._scale_keep = scale
            sp = scale.data_ptr()
        count = self.n_params - first if count is None else count
        check(lib().sf_grads_to_bf16(self.grads_f32.data_ptr() + 4 * first, self.grads_bf16.data_ptr() + 2 * first, count, sp,
                                     self._stream()), "sf_grads_to_bf16")
        return self.grads_bf16

    def ensure_optimizer_state(self) -> None:
        """fp32 masters (cloned from the bf16 weights, optimizer.py:33-41) and zeroed AdamW moments, allocated on first use."""
        if self.master is None:
            self.master = self.params.float()
            self.exp_avg = torch.zeros_like(self.master)
            self.exp_avg_sq = torch.zeros_like(self.master)

    def optimizer_step(self, lr: float, *, grad_scale: float = 1.0, max_grad_norm: float = 0.5, betas=(0.9, 0.999),
                       eps: float = 1e-8, weight_decay: float = 0.0) -> torch.Tensor:
        self.ensure_optimizer_state()
        self.opt_step += 1
        check(lib().sf_optimizer_step(self.grads_bf16.data_ptr(), self.master.data_ptr(), self.exp_avg.data_ptr(),
                                      self.exp_avg_sq.data_ptr(), self.params.data_ptr(), self.n_params, grad_scale,
                                      max_grad_norm, lr, betas[0], betas[1], eps, weight_decay, self.opt_step,
                                      self.grad_norm.data_ptr(), self._scratch.data_ptr(), self._stream()),
              "sf_optimizer_step")
        return self.grad_norm

    # ------------------------------------------------------------------ FLOP model
    def flops_per_step(self) -> float:
        d = self.dims
        M = self.B * self.S
        H, I, Ht = d.hidden_size, d.intermediate_size, d.target_hidden_size
        A, KV = d.num_heads * d.head_dim, d.num_kv_heads * d.head_dim
        per_tok_step = 2 * (2 * H) * (A + 2 * KV) + 2 * A * H + 3 * 2 * H * I + 2 * H * d.draft_vocab_size
        attn = 4 * (self.S / 2) * d.num_heads * d.head_dim
        fwd = self.T * (per_tok_step + attn) + 2 * 3 * Ht * H
        train = 3 * self.T * (per_tok_step + attn) + 2 * (2 * 3 * Ht * H)
        teacher = 2 * Ht * d.vocab_size
        _ = fwd
        return float(M) * (train + teacher)
