"""Generate golden vectors by running the UNMODIFIED reference (sgl-project/SpecForge @ /root/reference)
on CPU — TEST INFRASTRUCTURE, build-container only (the GPU box has no /root/reference).

    TORCHDYNAMO_DISABLE=1 python oracle/make_golden.py

Two shims, the same the reference's own tests use:
  * TORCHDYNAMO_DISABLE=1 — Inductor's CPU compile is broken in this image;
  * specforge.algorithms.eagle3.model.LogSoftmaxLoss.apply -> specforge.core.loss._compute_loss
    (the reference's own torch twin of its Triton kernel; tests/test_utils/test_loss.py pins the pair to 1e-4).

The path exercised is the real one: Eagle3TrainStrategy.forward_loss -> TargetHead ->
OnlineEagle3Model(attention_backend="sdpa") -> LlamaForCausalLMEagle3 -> backward -> BF16Optimizer.step.
Outputs (tests/golden/eagle3_<case>.pt): inputs are regenerated from seeds by the oracle's make_batch /
init_params, so only results are stored: per-step plosses/acces/acceptance, loss, logits slice, gradients
(full for the tiny cases), updated weights after one optimizer step.
"""
import json
import os
import sys
import tempfile

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402

from oracle import eagle3_oracle as O  # noqa: E402

CASES = {
    # name: (cfg kwargs, B, S, pad_tail, lk_loss_type)
    "tiny": (dict(hidden_size=64, intermediate_size=128, num_heads=4, num_kv_heads=2, head_dim=16, vocab_size=256,
                  draft_vocab_size=64, rms_norm_eps=1e-5, rope_theta=10000.0, max_position_embeddings=512,
                  ttt_length=3), 2, 16, 0, None),
    "tiny_pad": (dict(hidden_size=64, intermediate_size=128, num_heads=4, num_kv_heads=2, head_dim=16, vocab_size=256,
                      draft_vocab_size=64, rms_norm_eps=1e-5, rope_theta=10000.0, max_position_embeddings=512,
                      ttt_length=4), 3, 24, 5, None),
    "tiny_lk_lambda": (dict(hidden_size=64, intermediate_size=128, num_heads=4, num_kv_heads=2, head_dim=16,
                            vocab_size=256, draft_vocab_size=64, rms_norm_eps=1e-5, rope_theta=10000.0,
                            max_position_embeddings=512, ttt_length=3), 2, 16, 0, "lambda"),
    "tiny_lk_alpha": (dict(hidden_size=64, intermediate_size=128, num_heads=4, num_kv_heads=2, head_dim=16,
                           vocab_size=256, draft_vocab_size=64, rms_norm_eps=1e-5, rope_theta=10000.0,
                           max_position_embeddings=512, ttt_length=3), 2, 16, 0, "alpha"),
    # GQA 4:1, head_dim 128 (the kernel's shape class), still seconds on CPU
    "small_d128": (dict(hidden_size=256, intermediate_size=512, num_heads=4, num_kv_heads=1, head_dim=128,
                        vocab_size=1024, draft_vocab_size=256, rms_norm_eps=1e-6, rope_theta=1000000.0,
                        max_position_embeddings=2048, ttt_length=7), 2, 160, 9, None),
    # head_dim 64 cases the CUDA path can run: LK objectives, EAGLE3.1 fc_norm (target hidden != draft hidden), norm_output=False
    "small_lk_lambda": (dict(hidden_size=128, intermediate_size=256, num_heads=2, num_kv_heads=1, head_dim=64, vocab_size=512,
                             draft_vocab_size=128, rms_norm_eps=1e-5, rope_theta=10000.0, max_position_embeddings=512,
                             ttt_length=3), 2, 96, 7, "lambda"),
    "small_lk_alpha": (dict(hidden_size=128, intermediate_size=256, num_heads=2, num_kv_heads=1, head_dim=64, vocab_size=512,
                            draft_vocab_size=128, rms_norm_eps=1e-5, rope_theta=10000.0, max_position_embeddings=512,
                            ttt_length=3), 2, 96, 7, "alpha"),
    "small_fcnorm": (dict(hidden_size=128, intermediate_size=256, num_heads=2, num_kv_heads=2, head_dim=64, vocab_size=512,
                          draft_vocab_size=128, rms_norm_eps=1e-6, rope_theta=10000.0, max_position_embeddings=512,
                          ttt_length=3, fc_norm=True, target_hidden_size=192), 2, 64, 0, None),
    "small_nonorm": (dict(hidden_size=128, intermediate_size=256, num_heads=2, num_kv_heads=2, head_dim=64, vocab_size=512,
                          draft_vocab_size=128, rms_norm_eps=1e-6, rope_theta=10000.0, max_position_embeddings=512,
                          ttt_length=2, norm_output=False), 2, 64, 0, None),
    # BASELINE config 1 shape: Qwen2.5-0.5B draft, TTT=3, bs=1, seq=128 (configs/qwen2.5-0.5b-eagle3.json)
    "qwen25_05b_cfg1": (dict(hidden_size=896, intermediate_size=4864, num_heads=14, num_kv_heads=2, head_dim=64,
                             vocab_size=151936, draft_vocab_size=16000, rms_norm_eps=1e-6, rope_theta=1000000.0,
                             max_position_embeddings=32768, ttt_length=3), 1, 128, 0, None),
    # ---- full model dims of the BASELINE configs at B=1 (the sizes the headline numbers are quoted on; minutes on CPU).
    # head_std = 1/sqrt(H_t): teacher logits ~ N(0, 1), i.e. a broad soft-label distribution instead of a one-hot.
    # BASELINE config 2: Qwen3-8B draft (configs/qwen3-8b-eagle3.json), S=2048, TTT=7
    "qwen3_8b_cfg2_b1": (dict(hidden_size=4096, intermediate_size=12288, num_heads=32, num_kv_heads=8, head_dim=128,
                              vocab_size=151936, draft_vocab_size=32000, rms_norm_eps=1e-6, rope_theta=1000000.0,
                              max_position_embeddings=40960, ttt_length=7), 1, 2048, 0, None, dict(head_std=1.0 / 64)),
    # BASELINE config 3: Llama3-8B draft (configs/llama3-8B-eagle3.json: I=14336, V=128256, eps 1e-5, default theta,
    # max_position_embeddings=2048 -> S+T rows fit in the initial 2068-row RoPE cache), padded batch
    "llama3_8b_cfg3_b1": (dict(hidden_size=4096, intermediate_size=14336, num_heads=32, num_kv_heads=8, head_dim=128,
                               vocab_size=128256, draft_vocab_size=32000, rms_norm_eps=1e-5, rope_theta=10000.0,
                               max_position_embeddings=2048, ttt_length=7), 1, 2048, 129, None, dict(head_std=1.0 / 64)),
    # BASELINE config 5: Qwen3-30B-A3B EAGLE3.1 draft (configs/qwen3-30B-A3B-eagle3.1.json: H=2048, nkv=4, fc_norm),
    # S=4096; max_position_embeddings raised from 2048 to 8192 on BOTH sides (the reference
    # rebuilds its RoPE cache from a bf16 arange past max_position_embeddings+20, which is garbage we do not reproduce)
    "qwen3_30b_a3b_cfg5_b1": (dict(hidden_size=2048, intermediate_size=12288, num_heads=32, num_kv_heads=4, head_dim=128,
                                   vocab_size=151936, draft_vocab_size=32000, rms_norm_eps=1e-6, rope_theta=1000000.0,
                                   max_position_embeddings=8192, ttt_length=7, fc_norm=True), 1, 4096, 0, None,
                              dict(head_std=2048 ** -0.5)),
}


def logits_sample_index(S: int, DV: int):
    """Rows / columns of the per-step logits kept in the big goldens: 12 rows (start, middle, end of the sequence) x
    (first 512 columns + every 61st column)."""
    rows = list(range(0, 4)) + list(range(S // 2, S // 2 + 4)) + list(range(S - 4, S))
    cols = list(range(0, min(512, DV))) + list(range(512, DV, 61))
    return torch.tensor(rows), torch.tensor(cols)


def build_reference(cfg: O.Eagle3Config, P, t2d, d2t, head_w, lk_loss_type, workdir):
    import specforge.algorithms.eagle3.model as ref_model
    from specforge.core.loss import _compute_loss

    class _TorchLogSoftmaxLoss:  # same shim as tests/test_algorithms/test_eagle3_position_ids.py:12-35
        @staticmethod
        def apply(logits, target_p, position_mask):
            return _compute_loss(logits, target_p, position_mask)

    ref_model.LogSoftmaxLoss = _TorchLogSoftmaxLoss
    from transformers import LlamaConfig

    from specforge.algorithms.eagle3.model import OnlineEagle3Model
    from specforge.modeling.draft.llama3_eagle import LlamaForCausalLMEagle3
    from specforge.modeling.target.target_head import TargetHead
    from specforge.training.strategies.base import Eagle3TrainStrategy

    hf = LlamaConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                     num_attention_heads=cfg.num_heads, num_key_value_heads=cfg.num_kv_heads,
                     num_hidden_layers=1, vocab_size=cfg.vocab_size, rms_norm_eps=cfg.rms_norm_eps,
                     max_position_embeddings=cfg.max_position_embeddings, hidden_act="silu",
                     tie_word_embeddings=False, pad_token_id=0, rope_theta=cfg.rope_theta)
    hf.head_dim = cfg.head_dim
    hf.draft_vocab_size = cfg.draft_vocab_size
    hf.rope_theta = cfg.rope_theta
    hf.fc_norm = cfg.fc_norm
    hf.norm_output = cfg.norm_output
    hf.target_hidden_size = cfg.target_hidden_size
    draft = LlamaForCausalLMEagle3(hf, attention_backend="sdpa")
    sd = {k: v.clone() for k, v in P.items()}
    sd["t2d"], sd["d2t"] = t2d, d2t
    missing, unexpected = draft.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in m or "inv_freq" in m for m in missing), missing
    draft = draft.to(torch.bfloat16)
    draft.freeze_embedding()
    model = OnlineEagle3Model(draft_model=draft, length=cfg.ttt_length, attention_backend="sdpa",
                              lk_loss_type=lk_loss_type)
    tdir = os.path.join(workdir, "target")
    os.makedirs(tdir, exist_ok=True)
    with open(os.path.join(tdir, "config.json"), "w") as f:
        json.dump({"architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_size": cfg.target_hidden_size,
                   "vocab_size": cfg.vocab_size, "num_hidden_layers": 1, "num_attention_heads": 4,
                   "intermediate_size": 128}, f)
    head = TargetHead(tdir)
    with torch.no_grad():
        head.fc.weight.copy_(head_w.float())
    head.freeze_weights()
    head = head.eval().to(torch.bfloat16)
    strategy = Eagle3TrainStrategy(model, target_head=head, ploss_decay=cfg.ploss_decay)
    return draft, model, strategy


def run_case(name):
    from specforge.optimizer import BF16Optimizer
    from specforge.runtime.contracts import TrainBatch

    kw, B, S, pad_tail, lk = CASES[name][:5]
    extra = CASES[name][5] if len(CASES[name]) > 5 else {}
    head_std = float(extra.get("head_std", 1.0))
    cfg = O.Eagle3Config(**kw)
    torch.manual_seed(0)
    P = O.init_params(cfg, seed=0)
    t2d, d2t = O.make_vocab_map(cfg.vocab_size, cfg.draft_vocab_size, seed=0)
    g = torch.Generator().manual_seed(1234)
    head_w = (torch.randn(cfg.vocab_size, cfg.target_hidden_size, generator=g) * head_std).to(torch.bfloat16)
    batch = O.make_batch(cfg, B, S, seed=0, pad_tail=pad_tail)
    with tempfile.TemporaryDirectory() as wd:
        draft, model, strategy = build_reference(cfg, P, t2d, d2t, head_w, lk, wd)
        # grab step logits through a forward hook on lm_head (reference module, unmodified)
        logits_seen = []
        hk = draft.lm_head.register_forward_hook(lambda m, i, o: logits_seen.append(o.detach().clone()))
        opt = BF16Optimizer(draft, lr=1e-3, max_grad_norm=0.5, total_steps=100, warmup_ratio=0.1)
        tb = TrainBatch(sample_ids=[str(i) for i in range(B)], strategy="eagle3",
                        tensors={k: v.clone() for k, v in batch.items()}, metadata={"target_repr": "hidden_state"})
        out = strategy.forward_loss(tb)
        out.loss.backward()
        hk.remove()
        grads = {n: p.grad.detach().clone() for n, p in draft.named_parameters() if p.grad is not None}
        lr_used = opt.get_learning_rate()
        gnorm = opt.step()
        new_w = {n: p.detach().clone() for n, p in draft.named_parameters() if p.requires_grad}
    big = cfg.hidden_size > 256
    gold = {
        "case": name, "cfg": kw, "B": B, "S": S, "pad_tail": pad_tail, "lk_loss_type": lk, "head_seed": 1234, "head_std": head_std,
        "loss": out.loss.detach().float(),
        "plosses": torch.stack([p.float() for p in out.metrics["plosses"]]),
        "acces": torch.stack([a.float() for a in out.metrics["acces"]]),
        "acceptance_rates": torch.stack([a.float() for a in out.metrics["acceptance_rates"]]),
        "acc_corrects": torch.stack([c.float() for c in out.metrics["acc_corrects"]]),
        "acc_denoms": torch.stack([c.float() for c in out.metrics["acc_denoms"]]),
        "logits_slice": torch.stack([l[:, :8, :64] for l in logits_seen]),
        "grad_norm": gnorm.float(), "lr_used": lr_used,
    }
    if cfg.hidden_size >= 2048:   # full-size cases: a spread sample of every step's logits [T, B, rows, cols]
        ri, ci = logits_sample_index(S, cfg.draft_vocab_size)
        gold["logits_rows"], gold["logits_cols"] = ri, ci
        gold["logits_sample"] = torch.stack([l[:, ri][:, :, ci] for l in logits_seen])
    if big:
        gold["grad_stats"] = {n: torch.stack([g_.float().norm(), g_.float().abs().max(), g_.float().flatten()[:16].norm()])
                              for n, g_ in grads.items()}
        gold["grad_slices"] = {n: g_.flatten()[:256].clone() for n, g_ in grads.items()}
        gold["new_w_slices"] = {n: w.flatten()[:256].clone() for n, w in new_w.items()}
    else:
        gold["grads"] = grads
        gold["new_w"] = new_w
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    path = os.path.join(ROOT, "tests", "golden", f"eagle3_{name}.pt")
    torch.save(gold, path)
    print(name, "loss", float(gold["loss"]), "plosses", gold["plosses"].tolist(), "acc", gold["acces"].tolist(),
          "accept", gold["acceptance_rates"].tolist(), "gnorm", float(gnorm), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    torch.set_num_threads(8)
    for n in (sys.argv[1:] or list(CASES)):
        run_case(n)
