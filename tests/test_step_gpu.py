"""Full-step parity (GPU): sf_eagle3_forward/backward/optimizer through the C ABI against
  (a) golden vectors produced by the UNMODIFIED reference (tests/golden, see oracle/make_golden.py) and
  (b) the fp32 oracle on the same seeded inputs.
Tolerances (SURVEY §8c calibration): loss / per-step ploss rel <= 1e-3; acceptance abs 1e-4 + rel 2e-3; top-1
counts within ties; weight-gradient cosine >= 0.999 and norm ratio within 1 %."""
import glob
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD_DIR = os.path.join(os.path.dirname(__file__), "golden")


def _engine_for(gold):
    from oracle import eagle3_oracle as O
    from specforge_b200.engine import DraftDims, Eagle3Engine
    kw = gold["cfg"]
    cfg = O.Eagle3Config(**kw)
    dims = DraftDims(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size, num_heads=cfg.num_heads,
                     num_kv_heads=cfg.num_kv_heads, head_dim=cfg.head_dim, vocab_size=cfg.vocab_size,
                     draft_vocab_size=cfg.draft_vocab_size, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta,
                     max_position_embeddings=cfg.max_position_embeddings, target_hidden_size=cfg.target_hidden_size,
                     fc_norm=cfg.fc_norm, norm_output=cfg.norm_output)
    eng = Eagle3Engine(dims, batch=gold["B"], seq_len=gold["S"], ttt_length=cfg.ttt_length,
                       lk_loss_type=gold.get("lk_loss_type"))
    P = O.init_params(cfg, seed=0)
    t2d, d2t = O.make_vocab_map(cfg.vocab_size, cfg.draft_vocab_size, seed=0)
    g = torch.Generator().manual_seed(gold["head_seed"])
    head_w = (torch.randn(cfg.vocab_size, cfg.target_hidden_size, generator=g) * gold.get("head_std", 1.0)).to(torch.bfloat16)
    eng.load_params(P)
    eng.set_frozen(embed_tokens=P["embed_tokens.weight"], target_head=head_w, t2d=t2d, d2t=d2t)
    batch = O.make_batch(cfg, gold["B"], gold["S"], seed=0, pad_tail=gold["pad_tail"])
    return eng, cfg, P, batch, head_w, t2d, d2t


def _logit_stats(got, ref):
    ulp = 2.0 ** -7 * ref.abs().clamp_min(1.0)          # bf16 spacing at |x| in [1, 2) is 2^-7; smooth lower bound above that
    err = (got - ref).abs() / ulp
    return {"max_ulp": err.max().item(), "frac_le_2ulp": (err <= 2.0).float().mean().item(),
            "per_step_max_ulp": [err[j].max().item() for j in range(err.shape[0])],
            "per_step_cos": [torch.nn.functional.cosine_similarity(got[j].flatten(), ref[j].flatten(), dim=0).item()
                             for j in range(ref.shape[0])]}


def _check_logits(eng, gold, batch):
    """The draft logits of every TTT step (llama3_eagle.py:1772-1777, captured on the reference's lm_head by a forward hook)
    against ours after a forward-only pass.  Two references, two stated tolerances:
      * `logits_sample` / `logits_slice`: the UNMODIFIED reference on its CPU `sdpa` (eager) path.  That path rounds the pre-softmax
        scores and every RoPE op to bf16 (llama3_eagle.py:133-142,749-769); at H = 4096 the scores reach |s| ~ 10, so that rounding
        alone moves the logits by several bf16 ulps (the reference's own flex / flash backends do not round there; its inter-backend
        test tolerance is 1e-2, tests/test_utils/test_flex_attention.py:124-132).  Tolerance: cosine >= 0.9997 per step, no element
        off by more than 24 ulps of max(|x|, 1) (ulp = 2^-7); at the small dims (|s| << 1) the strict bound below applies instead.
        Measured on B200 at config 2: cosine 0.99996 (step 0) .. 0.99980 (step 6), max 16 ulps.
      * `logits_sample_fused` (full-size cases): the oracle — bit-identical to the reference on the eager schedule, asserted when
        the sample was made — run with the rounding schedule of the reference's GPU backends (fp32 scores, one-rounding RoPE;
        oracle/eagle3_oracle.py NUMERICS).  Tolerance: cosine >= 0.9999 per step (the SURVEY section 8c target), >= 85 % of the
        sampled elements within 2 ulps, none beyond 12.  Measured at config 2: cosine 0.999986 .. 0.999918, max 3.5 .. 8 ulps.
        What is left is the flash-style softmax itself: P is rounded to bf16 relative to a running maximum, the materialised
        softmax rounds the normalised P — with peaked attention (|s| ~ 10) those two roundings do not average out.
    Losses, acceptance rates and gradients — the quantities the north star bounds — meet 1e-3 / 2e-3 / cosine 0.999 at every size."""
    eng.forward(batch, need_grad=False)
    torch.cuda.synchronize()
    B, S = gold["B"], gold["S"]
    lg = eng.workspace_view("logits").view(eng.T, B, S, -1)
    if "logits_sample" in gold:
        ri, ci = gold["logits_rows"].to(lg.device), gold["logits_cols"].to(lg.device)
        got = lg[:, :, ri][:, :, :, ci].float().cpu()
        ref = gold["logits_sample"].float()
    else:
        ref = gold["logits_slice"].float()                       # [T, B, 8, 64]
        got = lg[:, :, :ref.shape[2], :ref.shape[3]].float().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    stats = {"case": gold.get("case"), "vs_reference_eager": _logit_stats(got, ref)}
    if "logits_sample_fused" in gold:
        stats["vs_fused_schedule"] = _logit_stats(got, gold["logits_sample_fused"].float())
    try:                                               # evidence for profiles/ (best effort)
        import json
        out = os.path.join(os.path.dirname(GOLD_DIR), "..", "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "logits_parity.jsonl"), "a") as f:
            f.write(json.dumps(stats) + "\n")
    except OSError:
        pass
    e = stats["vs_reference_eager"]
    assert min(e["per_step_cos"]) >= 0.9997 and e["max_ulp"] <= 24.0, stats
    if "vs_fused_schedule" in stats:
        f = stats["vs_fused_schedule"]
        assert f["frac_le_2ulp"] >= 0.85 and f["max_ulp"] <= 12.0 and min(f["per_step_cos"]) >= 0.9999, stats
    else:
        assert e["frac_le_2ulp"] >= 0.995 and e["max_ulp"] <= 4.0 and min(e["per_step_cos"]) >= 0.9999, stats


@pytest.mark.parametrize("case", ["small_d128", "qwen25_05b_cfg1", "small_lk_lambda", "small_lk_alpha", "small_fcnorm",
                                  "small_nonorm", "qwen3_8b_cfg2_b1", "llama3_8b_cfg3_b1", "qwen3_30b_a3b_cfg5_b1"])
def test_step_matches_reference_golden(case):
    """Includes the three BASELINE configurations at their full model dimensions (B = 1): config 2 (Qwen3-8B, S = 2048, the
    shape the headline number is quoted on), config 3 (Llama3-8B dims, padded batch), config 5 (Qwen3-30B-A3B EAGLE3.1 draft,
    S = 4096, fc_norm)."""
    path = os.path.join(GOLD_DIR, f"eagle3_{case}.pt")
    if not os.path.exists(path):
        pytest.fail(f"golden {path} missing: run oracle/make_golden.py {case} in the build container")
    gold = torch.load(path)
    eng, cfg, P, batch, *_ = _engine_for(gold)
    _check_logits(eng, gold, batch)
    loss, metrics = eng.forward(batch, need_grad=True)
    eng.backward()
    torch.cuda.synchronize()
    m = metrics.cpu()
    torch.testing.assert_close(m[:, 0], gold["plosses"], rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(loss.cpu()[0], gold["loss"], rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(m[:, 3], gold["acceptance_rates"], rtol=2e-3, atol=1e-4)
    torch.testing.assert_close(m[:, 2], gold["acc_denoms"], rtol=0, atol=0)
    assert (m[:, 1] - gold["acc_corrects"]).abs().max() <= 2     # argmax ties at bf16 resolution
    g32 = eng.grads_f32
    if "grads" in gold:
        for n, ref in gold["grads"].items():
            got = eng.param_view(n, g32).float().cpu()
            ref = ref.float()
            cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
            ratio = (got.norm() / ref.norm()).item()
            assert cos >= 0.999 and abs(ratio - 1) <= 1e-2, (n, cos, ratio)
    else:
        for n, st in gold["grad_stats"].items():
            got = eng.param_view(n, g32).float().cpu()
            assert abs((got.norm() / st[0]).item() - 1) <= 1e-2, (n, got.norm().item(), st[0].item())
            ref = gold["grad_slices"][n].float()
            cos = torch.nn.functional.cosine_similarity(got.flatten()[:256], ref, dim=0).item()
            assert cos >= 0.995, (n, cos)
    # optimizer: one BF16Optimizer.step (optimizer.py:140-168) from our gradients vs the reference's new weights
    eng.grads_to_bf16()
    gn = eng.optimizer_step(lr=gold["lr_used"], max_grad_norm=0.5)
    torch.cuda.synchronize()
    assert abs(gn.item() / gold["grad_norm"].item() - 1) <= 1e-2
    key = "new_w" if "new_w" in gold else "new_w_slices"
    for n, ref in gold[key].items():
        got = eng.param_view(n).float().cpu().flatten()[: ref.numel()]
        ref = ref.float().flatten()
        # AdamW's first step moves every weight by ~lr*sign(g): compare the update direction
        assert (got - ref).abs().max().item() <= 2.5 * gold["lr_used"] + 2 ** -8 * ref.abs().max().item(), n


def test_step_matches_oracle_on_fresh_inputs():
    """Independent of the goldens: the bf16 oracle (CPU, same rounding points as the reference) vs the CUDA path on a
    padded batch, d=64, GQA 2:1.  (An fp32 oracle is NOT comparable at 1e-3: the reference rounds the teacher logits
    to bf16, which alone moves the loss by ~1 % at these tiny dims.)"""
    from oracle import eagle3_oracle as O
    gold = {"cfg": dict(hidden_size=128, intermediate_size=384, num_heads=4, num_kv_heads=2, head_dim=64, vocab_size=512,
                        draft_vocab_size=128, rms_norm_eps=1e-5, rope_theta=10000.0, max_position_embeddings=256,
                        ttt_length=4), "B": 3, "S": 96, "pad_tail": 17, "head_seed": 7}
    eng, cfg, P, batch, head_w, t2d, d2t = _engine_for(gold)
    loss, metrics = eng.forward(batch, need_grad=True)
    eng.backward()
    torch.cuda.synchronize()
    res, grads = O.train_step(P, cfg, batch, head_w, t2d, d2t)
    ref_pl = torch.stack([p.detach().float() for p in res.plosses])
    torch.testing.assert_close(metrics[:, 0].cpu(), ref_pl, rtol=1e-3, atol=1e-5)
    for n, ref in grads.items():
        got = eng.param_view(n, eng.grads_f32).float().cpu()
        cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.float().flatten(), dim=0).item()
        assert cos >= 0.999, (n, cos)


def test_eval_forward_and_errors():
    gold = torch.load(os.path.join(GOLD_DIR, "eagle3_small_d128.pt"))
    eng, cfg, P, batch, *_ = _engine_for(gold)
    loss, metrics = eng.forward(batch, need_grad=False)
    torch.cuda.synchronize()
    torch.testing.assert_close(metrics[:, 0].cpu(), gold["plosses"], rtol=1e-3, atol=1e-5)
    bad = dict(batch)
    bad["hidden_state"] = batch["hidden_state"][:, :-1]
    with pytest.raises(ValueError):
        eng.forward(bad)
