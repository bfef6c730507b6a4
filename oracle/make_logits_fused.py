"""Adds `logits_sample_fused` (+ `plosses_fused`) to the full-size goldens: the oracle's forward in the "fused" rounding schedule
(oracle/eagle3_oracle.py NUMERICS: fp32 RoPE with one rounding, fp32 attention scores — what the reference's GPU backends compute)
on the same seeded inputs.  The goldens' own `logits_sample` come from the UNMODIFIED reference on its CPU `sdpa` (eager) path,
whose bf16 pre-softmax scores move the logits by several bf16 ulps at H = 4096; this second sample isolates that effect.
TEST INFRASTRUCTURE, build-container only.     python oracle/make_logits_fused.py [case ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import eagle3_oracle as O  # noqa: E402
from oracle.make_golden import CASES, logits_sample_index  # noqa: E402


def run(name):
    path = os.path.join(ROOT, "tests", "golden", f"eagle3_{name}.pt")
    gold = torch.load(path)
    cfg = O.Eagle3Config(**gold["cfg"])
    P = O.init_params(cfg, seed=0)
    t2d, d2t = O.make_vocab_map(cfg.vocab_size, cfg.draft_vocab_size, seed=0)
    g = torch.Generator().manual_seed(gold["head_seed"])
    head_w = (torch.randn(cfg.vocab_size, cfg.target_hidden_size, generator=g) * gold.get("head_std", 1.0)).to(torch.bfloat16)
    batch = O.make_batch(cfg, gold["B"], gold["S"], seed=0, pad_tail=gold["pad_tail"])
    ri, ci = logits_sample_index(gold["S"], cfg.draft_vocab_size)
    with torch.no_grad():
        # sanity: the eager schedule reproduces the reference's own logits sample bit for bit
        res = O.forward_loss(P, cfg, batch, head_w, t2d, d2t, lk_loss_type=gold.get("lk_loss_type"), keep=True)
        eager = torch.stack([l[:, ri][:, :, ci] for l in res.logits])
        assert torch.equal(eager, gold["logits_sample"]), "oracle (eager schedule) != reference logits"
        with O.numerics("fused"):
            res = O.forward_loss(P, cfg, batch, head_w, t2d, d2t, lk_loss_type=gold.get("lk_loss_type"), keep=True)
    gold["logits_sample_fused"] = torch.stack([l[:, ri][:, :, ci] for l in res.logits])
    gold["plosses_fused"] = torch.stack([p.float() for p in res.plosses])
    torch.save(gold, path)
    d = (gold["logits_sample_fused"].float() - gold["logits_sample"].float()).abs()
    print(name, "fused-vs-eager logits: max", float(d.max()), "plosses eager", gold["plosses"].tolist(), "fused", gold["plosses_fused"].tolist())


if __name__ == "__main__":
    torch.set_num_threads(8)
    for n in (sys.argv[1:] or [k for k, v in CASES.items() if len(v) > 5]):
        run(n)
