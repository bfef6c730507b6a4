"""Generates tests/golden/dflash_*.pt by running the UNMODIFIED reference (`/root/reference`, build container only):
`DFlashDraftModel` + `OnlineDFlashModel.forward` + autograd, eager attention backend (the one that defines the
"dropped block -> zeros" behaviour on CPU, dflash.py:80-94,203-213).  Each file stores the inputs (parameters, batch,
frozen embedding / head, the anchors the reference drew) and the reference's outputs (loss, accuracy, loss terms,
gradients), so tests/test_dflash_oracle.py can pin oracle/dflash_oracle.py without the reference.

    python oracle/make_dflash_golden.py
"""
import os
import sys

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch                                                                     # noqa: E402
from transformers.models.qwen3.modeling_qwen3 import Qwen3Config                 # noqa: E402

from oracle import dflash_oracle as D                                            # noqa: E402
from specforge.algorithms.common.dflash_family_model import OnlineDFlashModel    # noqa: E402
from specforge.modeling.draft.dflash import DFlashDraftModel                     # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (config overrides, B, S, dtype, loss-mask layout, seed)
    "dflash_tiny_f32": (dict(), 2, 40, torch.float32, "prefix5", 0),
    "dflash_tiny_bf16": (dict(), 2, 40, torch.bfloat16, "prefix5", 0),
    # row 1 has only 3 candidate anchors -> dropped blocks; anchors near the end -> labels out of bounds
    "dflash_dropped_blocks_f32": (dict(num_anchors=8), 2, 24, torch.float32, "short_row", 1),
    "dflash_decay_f32": (dict(loss_decay_gamma=2.0, block_size=8, num_anchors=5), 3, 48, torch.float32, "holes", 2),
    "dflash_gqa_d32_bf16": (dict(hidden_size=128, intermediate_size=256, num_heads=4, num_kv_heads=1, head_dim=32, num_layers=3,
                                 num_target_feats=3, vocab_size=512, mask_token_id=511, block_size=16, num_anchors=4), 2, 96,
                            torch.bfloat16, "prefix5", 3),
    # many tokens are 7 or 8 and the frozen head is biased towards them -> non-trivial accuracy numerator
    "dflash_biased_head_f32": (dict(num_anchors=7), 2, 40, torch.float32, "biased", 4),
    # D-PACE objectives (dflash_family_model.py:247-279): detached confidence weights, loss normalised by the batch size
    "dflash_dpace_f32": (dict(loss_type="dpace", dpace_alpha=0.5, block_size=8, num_anchors=5), 2, 48, torch.float32, "holes", 7),
    "dflash_dpace_cumconf_f32": (dict(loss_type="dpace-cumulative-confidence-only", dpace_alpha=0.3, block_size=8, num_anchors=5), 2, 48,
                                 torch.float32, "holes", 8),
    "dflash_dpace_contval_f32": (dict(loss_type="dpace-continuation-value-only", dpace_alpha=0.7, block_size=8, num_anchors=5), 2, 48,
                                 torch.float32, "prefix5", 9),
    # shapes the tensor-core block attention covers (group 4 x block 16 = 64 rows per anchor block, head_dim 64 / 128);
    # written as dflashtc_* so the default GPU suite does not pick them up before that path has been run once
    "dflashtc_d64_bf16": (dict(hidden_size=256, intermediate_size=512, num_heads=4, num_kv_heads=1, head_dim=64, num_layers=2,
                               num_target_feats=2, vocab_size=512, mask_token_id=511, block_size=16, num_anchors=6), 2, 160,
                          torch.bfloat16, "prefix5", 5),
    "dflashtc_d128_drop_bf16": (dict(hidden_size=512, intermediate_size=512, num_heads=8, num_kv_heads=2, head_dim=128, num_layers=1,
                                     num_target_feats=2, vocab_size=512, mask_token_id=511, block_size=16, num_anchors=7), 2, 130,
                                torch.bfloat16, "short_row", 6),
    # sliding-window layers (dflash.py:24-68, dflash_family_model.py:73-84; configs/qwen3.6-27b-dflash.json mixes 4 sliding + 1 full)
    "dflash_sliding_f32": (dict(num_layers=3, layer_types=("sliding_attention", "full_attention", "sliding_attention"), sliding_window=6,
                                block_size=4, num_anchors=7), 2, 40, torch.float32, "prefix5", 10),
    "dflashtc_sliding_bf16": (dict(hidden_size=256, intermediate_size=512, num_heads=4, num_kv_heads=1, head_dim=64, num_layers=2,
                                   layer_types=("sliding_attention", "full_attention"), sliding_window=48,
                                   num_target_feats=2, vocab_size=512, mask_token_id=511, block_size=16, num_anchors=6), 2, 200,
                              torch.bfloat16, "prefix5", 11),
    "dflashtc_sliding_d128_bf16": (dict(hidden_size=512, intermediate_size=512, num_heads=8, num_kv_heads=2, head_dim=128, num_layers=2,
                                        layer_types=("sliding_attention", "sliding_attention"), sliding_window=100,
                                        num_target_feats=2, vocab_size=512, mask_token_id=511, block_size=16, num_anchors=7), 2, 230,
                                   torch.bfloat16, "short_row", 12),
}


def loss_mask_for(kind, B, S, g):
    lm = torch.ones(B, S)
    if kind == "prefix5":
        lm[:, :5] = 0
    elif kind == "short_row":
        lm[0, :3] = 0
        lm[1] = 0
        lm[1, S - 4:] = 1                      # 3 candidates: S-4, S-3, S-2 (the last two run past the end)
    elif kind == "holes":
        lm = (torch.rand(B, S, generator=g) > 0.3).float()
        lm[:, :4] = 0
    return lm


def main():
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    for name, (over, B, S, dtype, lmk, seed) in CASES.items():
        if only and name not in only:
            continue
        c = D.DFlashConfig(**over)
        qc = Qwen3Config(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size, num_attention_heads=c.num_heads,
                         num_key_value_heads=c.num_kv_heads, head_dim=c.head_dim, num_hidden_layers=c.num_layers,
                         vocab_size=c.vocab_size, rms_norm_eps=c.rms_norm_eps, rope_theta=c.rope_theta,
                         max_position_embeddings=1024, attention_bias=False,
                         layer_types=list(c.layer_types) if c.layer_types else ["full_attention"] * c.num_layers,
                         sliding_window=c.sliding_window, use_sliding_window=c.sliding_window is not None)
        qc.block_size = c.block_size
        qc.num_target_layers = 8
        qc.dflash_config = {"mask_token_id": c.mask_token_id, "target_layer_ids": list(range(c.num_target_feats))}
        qc._attn_implementation = "eager"
        g = torch.Generator().manual_seed(100 + seed)
        P = D.init_params(c, seed=seed, dtype=dtype)
        draft = DFlashDraftModel(qc).to(dtype)
        missing, unexpected = draft.load_state_dict(P, strict=False)
        assert not unexpected and all("rotary" in m for m in missing), (missing, unexpected)
        embed_w = (torch.randn(c.vocab_size, c.hidden_size, generator=g) * 0.5).to(dtype)
        head_w = (torch.randn(c.vocab_size, c.hidden_size, generator=g) * 0.2).to(dtype)
        emb = torch.nn.Embedding(c.vocab_size, c.hidden_size).to(dtype)
        head = torch.nn.Linear(c.hidden_size, c.vocab_size, bias=False).to(dtype)
        with torch.no_grad():
            emb.weight.copy_(embed_w)
            head.weight.copy_(head_w)
        emb.requires_grad_(False)
        head.requires_grad_(False)
        model = OnlineDFlashModel(draft, head, emb, mask_token_id=c.mask_token_id, block_size=c.block_size, attention_backend="eager",
                                  num_anchors=c.num_anchors, loss_decay_gamma=c.loss_decay_gamma, objective_chunk_blocks=0,
                                  loss_type=c.loss_type, dpace_alpha=c.dpace_alpha)
        if lmk == "biased":
            head_w[7] *= 6.0
            head_w[8] = -head_w[7]          # the hidden states of all slots are similar: one of the two always wins
            with torch.no_grad():
                head.weight.copy_(head_w)
        batch = {"input_ids": torch.randint(0, c.vocab_size - 2, (B, S), generator=g),
                 "hidden_states": torch.randn(B, S, c.num_target_feats * c.hidden_size, generator=g).to(dtype),
                 "loss_mask": loss_mask_for(lmk, B, S, g)}
        if lmk == "biased":
            batch["input_ids"][torch.rand(B, S, generator=g) < 0.35] = 7
            batch["input_ids"][torch.rand(B, S, generator=g) < 0.35] = 8
        # reproduce the reference's anchor draw: same global RNG state for its torch.rand and for the oracle's restatement
        torch.manual_seed(7 + seed)
        state = torch.get_rng_state()
        anchors, keep = model._sample_anchor_positions(S, batch["loss_mask"], torch.device("cpu"))
        torch.set_rng_state(state)
        o_anchors, o_keep = D.sample_anchor_positions(batch["loss_mask"], c.num_anchors)
        assert torch.equal(anchors, o_anchors) and torch.equal(keep, o_keep), "oracle anchor sampling differs from the reference"
        torch.set_rng_state(state)
        loss, acc, metrics = model(batch["input_ids"], batch["hidden_states"], batch["loss_mask"])
        loss.backward()
        grads = {n: p.grad.detach().clone() for n, p in draft.named_parameters()}
        ln, ld = metrics["loss_terms"]
        cn, ad = metrics["ratio_metrics"]["acc"]
        torch.save({"config": c.__dict__, "dtype": str(dtype), "params": P, "batch": batch, "embed_w": embed_w, "lm_head_w": head_w,
                    "anchors": anchors, "keep": keep, "loss": loss.detach(), "accuracy": acc.detach(), "loss_num": ln.detach(),
                    "loss_den": ld.detach(), "correct": cn, "acc_den": ad, "grads": grads}, os.path.join(OUT, name + ".pt"))
        print(f"{name}: loss {float(loss):.6f} acc {float(acc):.4f} anchors {tuple(anchors.shape)} kept {int(keep.sum())}/{keep.numel()}")


if __name__ == "__main__":
    main()
