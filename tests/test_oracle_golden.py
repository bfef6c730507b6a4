"""Pins oracle/eagle3_oracle.py against outputs of the UNMODIFIED reference (tests/golden/*.pt,
generated in the build container by oracle/make_golden.py).  CPU only."""
import glob
import os

import pytest
import torch

from oracle import eagle3_oracle as O

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "eagle3_*.pt")))


def _setup(gold):
    cfg = O.Eagle3Config(**gold["cfg"])
    P = O.init_params(cfg, seed=0)
    t2d, d2t = O.make_vocab_map(cfg.vocab_size, cfg.draft_vocab_size, seed=0)
    g = torch.Generator().manual_seed(gold["head_seed"])
    head_w = torch.randn(cfg.vocab_size, cfg.target_hidden_size, generator=g).to(torch.bfloat16)
    batch = O.make_batch(cfg, gold["B"], gold["S"], seed=0, pad_tail=gold["pad_tail"])
    return cfg, P, t2d, d2t, head_w, batch


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[7:-3] for p in GOLD])
def test_oracle_matches_reference(path):
    gold = torch.load(path)
    if gold["cfg"]["hidden_size"] > 256 and os.environ.get("SF_FULL_CPU_TESTS") != "1":
        pytest.skip("config-1 sized case: set SF_FULL_CPU_TESTS=1 (takes ~1 min on 8 cores)")
    cfg, P, t2d, d2t, head_w, batch = _setup(gold)
    res, grads = O.train_step(P, cfg, batch, head_w, t2d, d2t, lk_loss_type=gold["lk_loss_type"], keep=True)
    # same ops in the same order as the reference's eager CPU path -> tight tolerances
    torch.testing.assert_close(torch.stack([p.detach().float() for p in res.plosses]), gold["plosses"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(res.loss.detach().float(), gold["loss"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(torch.stack([a.float() for a in res.acces]), gold["acces"], rtol=0, atol=1e-6)
    torch.testing.assert_close(torch.stack([a.float() for a in res.acceptance_rates]), gold["acceptance_rates"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(torch.stack([c.float() for c in res.acc_corrects]), gold["acc_corrects"], rtol=0, atol=0)
    logits = torch.stack([l[:, :8, :64] for l in res.logits])
    torch.testing.assert_close(logits.float(), gold["logits_slice"].float(), rtol=0, atol=0)
    if "grads" in gold:
        for n, g in gold["grads"].items():
            torch.testing.assert_close(grads[n].float(), g.float(), rtol=2e-2, atol=1e-5, msg=lambda m: f"{n}: {m}")
            cos = torch.nn.functional.cosine_similarity(grads[n].float().flatten(), g.float().flatten(), dim=0)
            assert cos > 0.9999, (n, float(cos))
    else:
        for n, g in gold["grad_slices"].items():
            torch.testing.assert_close(grads[n].flatten()[:256].float(), g.float(), rtol=2e-2, atol=1e-5)


@pytest.mark.parametrize("path", [p for p in GOLD if "tiny.pt" in p or "tiny_pad" in p])
def test_oracle_optimizer_matches_reference(path):
    """BF16Optimizer.step (optimizer.py:140-168) restated: new weights after one step from the golden grads."""
    gold = torch.load(path)
    cfg, P, *_ = _setup(gold)
    names = [n for n in O.PARAM_NAMES if n in gold["grads"]]
    params = [P[n].clone() for n in names]
    masters = [p.float().clone() for p in params]
    ea = [torch.zeros_like(m) for m in masters]
    es = [torch.zeros_like(m) for m in masters]
    gnorm = O.adamw_clip_step(params, masters, ea, es, [gold["grads"][n] for n in names], step=1, lr=gold["lr_used"])
    torch.testing.assert_close(gnorm, gold["grad_norm"], rtol=1e-5, atol=1e-7)
    for n, p in zip(names, params):
        torch.testing.assert_close(p.float(), gold["new_w"][n].float(), rtol=0, atol=1e-2 * 2 ** -7)
        assert (p != gold["new_w"][n]).float().mean() < 1e-3, n
