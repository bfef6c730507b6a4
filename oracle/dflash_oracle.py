"""TEST INFRASTRUCTURE — CPU restatement of the reference's DFlash block-parallel draft training step (SURVEY §8f row 1,
BASELINE config 4).  Only tests/ (and, next round, smoke()/bench's CPU arm) may import this; no product path does.

Status: the CUDA path for this row is NOT built yet (round-1 scope is the EAGLE3 step); this oracle and its goldens
(`tests/golden/dflash_*.pt`, generated from the UNMODIFIED reference by `oracle/make_dflash_golden.py`) pin the
algorithm so the kernels of the next round have a checker from day one.

What is restated, with the reference lines it follows:
  * anchor sampling                `algorithms/common/dflash_family_model.py:179-210`  (needs torch's CPU RNG stream to
                                   reproduce the reference's draw; the step itself takes anchors as INPUTS)
  * noise ids / embedding          `:221-245`     (anchor token at each block start, mask token elsewhere)
  * position ids                   `:212-219`, `:289-293`
  * DFlash mask                    `:47-89` (dense form; context keys strictly before the anchor, own block bidirectional,
                                   dropped blocks attend to nothing)
  * draft backbone                 `modeling/draft/dflash.py:431-460` (fc + hidden_norm on the concatenated target
                                   layers, N x Qwen3 layer), attention `:97-215` (per-head q/k RMSNorm, K/V = [context ;
                                   noise], RoPE with q on the LAST q_len positions `:71-77`, rows without any allowed key
                                   give zeros `:80-94,203-213`), layer `:218-268`
  * objective                      `dflash_family_model.py:332-461` (same-position labels, weight mask = kept block x in
                                   bounds x not the anchor slot x loss mask, optional exp decay, loss = sum(w nll)/sum(w),
                                   accuracy on weighted positions)
RMSNorm / MLP / rotary follow transformers' Qwen3 modules the reference instantiates (`dflash_kernels.py:20-31`):
fp32 statistics, `weight * x.to(input_dtype)`; SwiGLU; rotary cos/sin computed in fp32 and cast to the activation dtype.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclasses.dataclass
class DFlashConfig:
    hidden_size: int = 64
    intermediate_size: int = 128
    num_heads: int = 4
    num_kv_heads: int = 2
    head_dim: int = 16
    num_layers: int = 2
    num_target_feats: int = 2          # len(target_layer_ids): the context feature is [S, num_target_feats * hidden]
    vocab_size: int = 256
    block_size: int = 4
    num_anchors: int = 6
    mask_token_id: int = 255
    rms_norm_eps: float = 1e-6
    rope_theta: float = 10000.0
    loss_decay_gamma: Optional[float] = None


def param_shapes(c: DFlashConfig) -> Dict[str, Tuple[int, ...]]:
    """State-dict names and shapes of `DFlashDraftModel` (dflash.py:336-375)."""
    H, A, KV, I, d = c.hidden_size, c.num_heads * c.head_dim, c.num_kv_heads * c.head_dim, c.intermediate_size, c.head_dim
    s: Dict[str, Tuple[int, ...]] = {}
    for i in range(c.num_layers):
        p = f"layers.{i}."
        s[p + "self_attn.q_proj.weight"] = (A, H)
        s[p + "self_attn.k_proj.weight"] = (KV, H)
        s[p + "self_attn.v_proj.weight"] = (KV, H)
        s[p + "self_attn.o_proj.weight"] = (H, A)
        s[p + "self_attn.q_norm.weight"] = (d,)
        s[p + "self_attn.k_norm.weight"] = (d,)
        s[p + "mlp.gate_proj.weight"] = (I, H)
        s[p + "mlp.up_proj.weight"] = (I, H)
        s[p + "mlp.down_proj.weight"] = (H, I)
        s[p + "input_layernorm.weight"] = (H,)
        s[p + "post_attention_layernorm.weight"] = (H,)
    s["norm.weight"] = (H,)
    s["fc.weight"] = (H, c.num_target_feats * H)
    s["hidden_norm.weight"] = (H,)
    return s


def init_params(c: DFlashConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    P = {}
    for n, shp in param_shapes(c).items():
        if len(shp) == 1:
            P[n] = (1.0 + 0.1 * torch.randn(shp, generator=g)).to(dtype)
        else:
            P[n] = (torch.randn(shp, generator=g) * 0.05).to(dtype)
    return P


# ------------------------------------------------------------------------------------------------ host-side sampling
def sample_anchor_positions(loss_mask: torch.Tensor, num_anchors: int, generator: Optional[torch.Generator] = None):
    """dflash_family_model.py:179-210.  A position is a candidate when it and its successor are supervised; up to
    `num_anchors` of them are drawn without replacement (argsort of uniforms), sorted; rows with fewer candidates get
    dropped blocks (keep=False, anchor 0)."""
    B, S = loss_mask.shape
    n_cand = max(S - 1, 0)
    valid = (loss_mask[:, :n_cand] > 0.5) & (loss_mask[:, 1:n_cand + 1] > 0.5)
    counts = valid.sum(dim=1)
    width = min(num_anchors, int(counts.max().item()))
    if width == 0:
        raise ValueError("DFlash-family training requires two consecutive supervised tokens")
    r = torch.rand(valid.shape, generator=generator)
    r.masked_fill_(~valid, 2.0)
    cand = r.argsort(dim=1)[:, :width]
    keep = torch.arange(width).unsqueeze(0) < counts.clamp(max=width).unsqueeze(1)
    sentinel = valid.shape[1]
    anchors = torch.where(keep, cand, torch.full_like(cand, sentinel)).sort(dim=1).values
    keep = anchors < sentinel
    return torch.where(keep, anchors, 0), keep


def noise_ids(input_ids: torch.Tensor, anchors: torch.Tensor, keep: torch.Tensor, c: DFlashConfig) -> torch.Tensor:
    """:221-245 — [B, N*bs] token ids: the anchor token opens each kept block, everything else is the mask token."""
    B, S = input_ids.shape
    N, bs = anchors.shape[1], c.block_size
    ids = torch.full((B, N * bs), c.mask_token_id, dtype=torch.long)
    tok = torch.gather(input_ids, 1, anchors.clamp(0, S - 1))
    ids[:, ::bs] = torch.where(keep, tok, torch.full_like(tok, c.mask_token_id))
    return ids


def dflash_mask(anchors: torch.Tensor, keep: torch.Tensor, S: int, bs: int) -> torch.Tensor:
    """:47-89 (no sliding window) — bool [B, N*bs, S + N*bs]."""
    B, N = anchors.shape
    q = torch.arange(N * bs)
    kv = torch.arange(S + N * bs)
    a = anchors.repeat_interleave(bs, dim=1).unsqueeze(-1)                       # [B, Q, 1]
    ctx = (kv < S).view(1, 1, -1) & (kv.view(1, 1, -1) < a)
    same = (kv >= S).view(1, 1, -1) & ((q // bs).view(1, -1, 1) == ((kv - S) // bs).view(1, 1, -1))
    return (ctx | same) & keep.repeat_interleave(bs, dim=1).unsqueeze(-1)


# ------------------------------------------------------------------------------------------------ modules
def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return w * xf.to(x.dtype)


def linear(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    return F.linear(x, w)


def rope_tables(pos: torch.Tensor, d: int, theta: float, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    """Qwen3RotaryEmbedding, default rope: fp32 angles, cos/sin cast to the activation dtype.  pos [B, L] -> [B, L, d]."""
    inv = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
    f = pos.float().unsqueeze(-1) * inv.view(1, 1, -1)
    emb = torch.cat((f, f), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def attention(P, pre: str, h: torch.Tensor, ctx: torch.Tensor, cos, sin, mask: torch.Tensor, c: DFlashConfig) -> torch.Tensor:
    """dflash.py:158-215.  h [B, Q, H] (normed noise stream), ctx [B, S, H]; mask bool [B, Q, S+Q]."""
    B, Q, _ = h.shape
    S = ctx.shape[1]
    nh, nkv, d = c.num_heads, c.num_kv_heads, c.head_dim
    q = linear(h, P[pre + "q_proj.weight"]).view(B, Q, nh, d)
    q = rmsnorm(q, P[pre + "q_norm.weight"], c.rms_norm_eps).transpose(1, 2)
    k = torch.cat([linear(ctx, P[pre + "k_proj.weight"]), linear(h, P[pre + "k_proj.weight"])], dim=1).view(B, S + Q, nkv, d)
    v = torch.cat([linear(ctx, P[pre + "v_proj.weight"]), linear(h, P[pre + "v_proj.weight"])], dim=1).view(B, S + Q, nkv, d)
    k = rmsnorm(k, P[pre + "k_norm.weight"], c.rms_norm_eps).transpose(1, 2)
    v = v.transpose(1, 2)
    cu, su = cos.unsqueeze(1), sin.unsqueeze(1)                                   # [B, 1, S+Q, d]
    q = q * cu[..., -Q:, :] + rotate_half(q) * su[..., -Q:, :]
    k = k * cu + rotate_half(k) * su
    g = nh // nkv
    k = k.repeat_interleave(g, dim=1)
    v = v.repeat_interleave(g, dim=1)
    scores = (q @ k.transpose(-1, -2)) * (d ** -0.5)
    add = torch.zeros(mask.shape, dtype=scores.dtype).masked_fill_(~mask, torch.finfo(scores.dtype).min)
    scores = scores + add.unsqueeze(1)
    w = torch.softmax(scores, dim=-1, dtype=torch.float32).to(q.dtype)
    o = (w @ v).transpose(1, 2).reshape(B, Q, nh * d)
    o = linear(o, P[pre + "o_proj.weight"])
    return o.masked_fill(~mask.any(dim=-1, keepdim=True), 0)                      # rows with no allowed key -> zeros


def draft_forward(P, c: DFlashConfig, noise_emb: torch.Tensor, target_hidden: torch.Tensor, pos_all: torch.Tensor,
                  mask: torch.Tensor) -> torch.Tensor:
    """dflash.py:431-460."""
    x = noise_emb
    ctx = rmsnorm(linear(target_hidden, P["fc.weight"]), P["hidden_norm.weight"], c.rms_norm_eps)
    cos, sin = rope_tables(pos_all, c.head_dim, c.rope_theta, x.dtype)
    for i in range(c.num_layers):
        p = f"layers.{i}."
        h = rmsnorm(x, P[p + "input_layernorm.weight"], c.rms_norm_eps)
        x = x + attention(P, p + "self_attn.", h, ctx, cos, sin, mask, c)
        h = rmsnorm(x, P[p + "post_attention_layernorm.weight"], c.rms_norm_eps)
        x = x + linear(F.silu(linear(h, P[p + "mlp.gate_proj.weight"])) * linear(h, P[p + "mlp.up_proj.weight"]),
                       P[p + "mlp.down_proj.weight"])
    return rmsnorm(x, P["norm.weight"], c.rms_norm_eps)


# ------------------------------------------------------------------------------------------------ the step
def labels_and_weights(input_ids: torch.Tensor, loss_mask: torch.Tensor, anchors: torch.Tensor, keep: torch.Tensor, bs: int):
    """:398-428 — target ids [B, N, bs] and the weight mask (kept block x in bounds x not slot 0 x loss mask)."""
    B, S = input_ids.shape
    idx = anchors.unsqueeze(-1) + torch.arange(bs).view(1, 1, -1)
    inb = idx < S
    safe = idx.clamp(max=S - 1)
    N = anchors.shape[1]
    tgt = torch.gather(input_ids.unsqueeze(1).expand(-1, N, -1), 2, safe)
    w = keep.unsqueeze(-1).expand(-1, -1, bs).float() * inb.float() * (torch.arange(bs).view(1, 1, -1) > 0).float()
    w = w * torch.gather(loss_mask.unsqueeze(1).expand(-1, N, -1), 2, safe).to(w.dtype)
    return tgt, w


def forward_loss(P, c: DFlashConfig, batch: Dict[str, torch.Tensor], anchors: torch.Tensor, keep: torch.Tensor,
                 embed_w: torch.Tensor, lm_head_w: torch.Tensor):
    """OnlineDFlashModel.forward with loss_type="dflash" (:385-461).  Returns (loss, accuracy, terms)."""
    ids, hs, lm = batch["input_ids"], batch["hidden_states"], batch["loss_mask"]
    B, S = ids.shape
    bs = c.block_size
    N = anchors.shape[1]
    emb = F.embedding(noise_ids(ids, anchors, keep, c), embed_w)
    pos_ctx = torch.arange(S).unsqueeze(0).expand(B, -1)
    pos_draft = (anchors.unsqueeze(-1) + torch.arange(bs).view(1, 1, -1)).view(B, -1)
    mask = dflash_mask(anchors, keep, S, bs)
    out = draft_forward(P, c, emb, hs, torch.cat([pos_ctx, pos_draft], dim=1), mask)
    tgt, w = labels_and_weights(ids, lm, anchors, keep, bs)
    logits = linear(out, lm_head_w).view(B, N, bs, -1)
    nll = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), tgt.reshape(-1), reduction="none").reshape_as(tgt)
    lw = w
    if c.loss_decay_gamma is not None and c.loss_decay_gamma > 0:
        lw = lw * torch.exp(-(torch.arange(bs).view(1, 1, -1) - 1).clamp(min=0).float() / c.loss_decay_gamma)
    loss_num, loss_den = (nll * lw).sum(), lw.sum()
    pred = logits.argmax(dim=-1)
    correct = ((pred == tgt) & (w > 0.5)).sum().float()
    acc_den = w.sum()
    return loss_num / loss_den, correct / acc_den, {"loss_num": loss_num, "loss_den": loss_den, "correct": correct, "acc_den": acc_den}


def train_step(P, c: DFlashConfig, batch, anchors, keep, embed_w, lm_head_w):
    """Forward + backward on leaf copies of P; returns (loss, accuracy, terms, grads)."""
    leaves = {n: p.detach().clone().requires_grad_(True) for n, p in P.items()}
    loss, acc, terms = forward_loss(leaves, c, batch, anchors, keep, embed_w, lm_head_w)
    loss.backward()
    return loss.detach(), acc.detach(), {k: v.detach() for k, v in terms.items()}, {n: p.grad for n, p in leaves.items()}
