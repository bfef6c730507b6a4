"""DFlash — oracle pinned against outputs of the unmodified reference (tests/golden/dflash_*.pt, made by
oracle/make_dflash_golden.py).  CPU only.  The CUDA path of this row is next round's work; these goldens are its checker."""
import glob
import os

import pytest
import torch

from oracle import dflash_oracle as D

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "dflash*.pt")))   # dflash_* and dflashtc_*


def test_goldens_present():
    assert len(GOLDEN) >= 8


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-3])
def test_oracle_matches_reference(path):
    g = torch.load(path, weights_only=False)
    c = D.DFlashConfig(**g["config"])
    f32 = g["dtype"] == "torch.float32"
    loss, acc, terms, grads = D.train_step(g["params"], c, g["batch"], g["anchors"], g["keep"], g["embed_w"], g["lm_head_w"])
    rt = 1e-5 if f32 else 2e-2
    torch.testing.assert_close(loss.float(), g["loss"].float(), rtol=rt, atol=1e-6)
    torch.testing.assert_close(terms["loss_num"].float(), g["loss_num"].float(), rtol=rt, atol=1e-5)
    torch.testing.assert_close(terms["loss_den"].float(), g["loss_den"].float(), rtol=1e-6, atol=0)
    assert float(terms["acc_den"]) == float(g["acc_den"])
    if f32:
        assert float(terms["correct"]) == float(g["correct"])
        torch.testing.assert_close(acc, g["accuracy"])
    else:
        assert abs(float(terms["correct"]) - float(g["correct"])) <= 1      # a bf16 near-tie may flip one argmax
    assert set(grads) == set(g["grads"])
    for n, ref in g["grads"].items():
        got = grads[n]
        assert got is not None, n
        if f32:
            torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-6, msg=lambda m: f"{n}: {m}")
        else:
            cos = torch.nn.functional.cosine_similarity(got.float().flatten(), ref.float().flatten(), dim=0).item()
            assert cos >= 0.99, (n, cos)


def test_anchor_sampling_properties():
    torch.manual_seed(0)
    lm = (torch.rand(4, 64) > 0.4).float()
    lm[3] = 0
    lm[3, 10:13] = 1
    a, keep = D.sample_anchor_positions(lm, 16)
    assert a.shape == keep.shape and a.shape[1] <= 16
    for b in range(4):
        kept = a[b][keep[b]]
        assert torch.equal(kept, kept.sort().values) and kept.unique().numel() == kept.numel()      # sorted, no repeats
        assert all(lm[b, i] > 0.5 and lm[b, i + 1] > 0.5 for i in kept.tolist())                   # both tokens supervised
    assert int(keep[3].sum()) == 2 and not bool((a[3][~keep[3]] != 0).any())                        # 2 candidates, rest dropped
    with pytest.raises(ValueError):
        D.sample_anchor_positions(torch.zeros(2, 8), 4)


def test_mask_semantics():
    anchors = torch.tensor([[2, 5]])
    keep = torch.tensor([[True, False]])
    m = D.dflash_mask(anchors, keep, S=8, bs=2)
    assert m.shape == (1, 4, 12)
    assert m[0, 0].tolist() == [True, True] + [False] * 6 + [True, True, False, False]      # context < anchor, own block
    assert m[0, 1].tolist() == m[0, 0].tolist()                                              # bidirectional inside a block
    assert not m[0, 2:].any()                                                                # dropped block sees nothing


def test_product_anchor_sampler_equals_oracle_under_same_rng():
    """specforge_b200.dflash.sample_anchor_positions is the host code that ships; same draws as the pinned oracle."""
    from specforge_b200.dflash import sample_anchor_positions
    lm = (torch.rand(5, 70, generator=torch.Generator().manual_seed(3)) > 0.35).float()
    torch.manual_seed(11)
    a1, k1 = sample_anchor_positions(lm, 9)
    torch.manual_seed(11)
    a2, k2 = D.sample_anchor_positions(lm, 9)
    assert torch.equal(a1, a2) and torch.equal(k1, k2)


def test_dflash_strategy_host_logic_with_a_fake_engine():
    """B200DFlashTrainStrategy: StepOutput contract of the reference's DFlashTrainStrategy (strategies/base.py:415-452) and the
    micro-batch accumulation protocol the backend relies on — checked on CPU with a stand-in engine."""
    from specforge_b200.contracts import TrainBatch
    from specforge_b200.dflash import B200DFlashTrainStrategy

    class FakeEngine:
        device = torch.device("cpu")

        def __init__(self):
            self.params = torch.zeros(8, dtype=torch.bfloat16)
            self.metrics = torch.tensor([6.0, 3.0, 2.0, 4.0])
            self.loss = torch.tensor([2.0])
            self.calls = []

        def set_frozen(self, **kw):
            self.frozen = kw

        def forward(self, batch, anchors, keep, need_grad=True):
            self.calls.append(("fwd", tuple(anchors.shape), need_grad))
            return self.loss, self.metrics

        def backward(self, accumulate=False):
            self.calls.append(("bwd", accumulate))

    class FakeDraft:
        engine = FakeEngine()

    st = B200DFlashTrainStrategy(FakeDraft(), target_embed_weight=torch.zeros(4, 4), target_head_weight=torch.zeros(4, 4), num_anchors=5,
                                 generator=torch.Generator().manual_seed(0))
    assert st.name == "dflash" and st.required_features == {"input_ids", "hidden_states", "loss_mask"}
    batch = TrainBatch(sample_ids=["0", "1"], strategy="dflash",
                       tensors={"input_ids": torch.zeros(2, 12, dtype=torch.long), "hidden_states": torch.zeros(2, 12, 8),
                                "loss_mask": torch.ones(2, 12)}, metadata={})
    out = st.forward_loss(batch)
    assert float(out.loss) == 2.0 and out.loss.requires_grad
    assert float(out.metrics["accuracy"]) == 0.5 and float(out.metrics["accuracy_denom"]) == 4.0
    assert [float(v) for v in out.ratio_metrics["acc"]] == [2.0, 4.0]
    assert out.loss_terms is None                      # "local" normalisation: a plain scalar loss for the controller
    (out.loss / 2).backward()
    out2 = st.forward_loss(batch)
    (out2.loss / 2).backward()
    eng = FakeDraft.engine
    assert eng.calls == [("fwd", (2, 5), True), ("bwd", False), ("fwd", (2, 5), True), ("bwd", True)]
    assert float(st._last_grad_out) == 0.5
    with torch.no_grad():
        st.forward_loss(batch)
    assert eng.calls[-1] == ("fwd", (2, 5), False)
    # "global": the reference's loss_terms contract — the numerator carries the gradient (dflash_family_model.py:459,
    # controller.py:334-351), the engine is told to differentiate it, and backpropagating the ratio instead is refused
    sg = B200DFlashTrainStrategy(FakeDraft(), target_embed_weight=torch.zeros(4, 4), target_head_weight=torch.zeros(4, 4), num_anchors=5,
                                 generator=torch.Generator().manual_seed(0), normalization="global")
    assert eng.grad_of_numerator == 1
    og = sg.forward_loss(batch)
    num, den = og.loss_terms
    assert float(num) == 6.0 and float(den) == 3.0 and num.requires_grad and not den.requires_grad
    n0 = len(eng.calls)
    (num / 4).backward()
    assert eng.calls[n0:] == [("bwd", False)] and float(sg._last_grad_out) == 0.25
    with pytest.raises(RuntimeError, match="numerator"):
        sg.forward_loss(batch).loss.backward()
    with pytest.raises(ValueError):
        st.forward_loss(TrainBatch(sample_ids=["0"], strategy="dflash", tensors={"input_ids": torch.zeros(1, 4)}, metadata={}))


def test_sync_free_anchor_sampling_keeps_the_same_anchors():
    """fixed_width=True pads with dropped blocks instead of narrowing the anchor table: kept anchors are the reference's."""
    from specforge_b200.dflash import sample_anchor_positions
    lm = (torch.rand(4, 50, generator=torch.Generator().manual_seed(5)) > 0.3).float()
    lm[1, 6:] = 0
    lm[2, 4:] = 0                         # every row has fewer than 40 candidates -> the reference narrows the table
    torch.manual_seed(21)
    a_ref, k_ref = D.sample_anchor_positions(lm, 40)
    torch.manual_seed(21)
    a_fix, k_fix = sample_anchor_positions(lm, 40, fixed_width=True)
    assert a_fix.shape[1] == 40 and a_ref.shape[1] < 40
    for b in range(4):
        assert torch.equal(a_fix[b][k_fix[b]], a_ref[b][k_ref[b]])
        assert not bool(a_fix[b][~k_fix[b]].any())
    # and the objective is unchanged by the extra dropped blocks
    c = D.DFlashConfig(num_anchors=40)
    P = D.init_params(c, seed=2)
    g = torch.Generator().manual_seed(9)
    batch = {"input_ids": torch.randint(0, 250, (4, 50), generator=g), "hidden_states": torch.randn(4, 50, 128, generator=g), "loss_mask": lm}
    emb, head = torch.randn(256, 64, generator=g) * 0.5, torch.randn(256, 64, generator=g) * 0.2
    l_ref, _, t_ref = D.forward_loss(P, c, batch, a_ref, k_ref, emb, head)
    l_fix, _, t_fix = D.forward_loss(P, c, batch, a_fix, k_fix, emb, head)
    torch.testing.assert_close(l_fix, l_ref, rtol=1e-6, atol=1e-7)
    assert float(t_fix["acc_den"]) == float(t_ref["acc_den"]) and float(t_fix["correct"]) == float(t_ref["correct"])


@pytest.mark.parametrize("g,bs", [(4, 16), (8, 16), (8, 8), (4, 32)])
def test_tensor_core_tiling_reproduces_the_dflash_mask(g, bs):
    """Index arithmetic of csrc/sf_dflash_attn_tc*.cu restated in Python (unit = 128 rows = 128/R blocks x g heads x bs slots,
    CTA = 2 units; context tiles with a per-row key limit, then one own-keys tile with a block-diagonal mask; the context
    dK/dV kernel iterates the contiguous range of kept blocks beyond the tile start) against oracle.dflash_mask."""
    torch.manual_seed(g * 100 + bs)
    B, S, N = 2, 200, 11
    lm = (torch.rand(B, S) > 0.15).float()
    lm[1, 40:] = 0                                   # row 1 has few candidates -> dropped blocks at the end
    anchors, keep = D.sample_anchor_positions(lm, N)
    N = anchors.shape[1]
    want = D.dflash_mask(anchors, keep, S, bs)       # [B, N*bs, S + N*bs]
    R, BKV = g * bs, 64
    BPU = 128 // R
    BPC = 2 * BPU
    assert 256 // g <= BKV
    got = torch.zeros_like(want)
    for b in range(B):
        for cta in range((N + BPC - 1) // BPC):
            n0 = cta * BPC
            kept_anchors = [int(anchors[b, n]) for n in range(n0, min(n0 + BPC, N)) if keep[b, n]]
            n_ctx = (max(kept_anchors, default=0) + BKV - 1) // BKV
            for x in range(2):
                for r in range(128):
                    blk = x * BPU + r // R
                    n = n0 + blk
                    hg, o = (r % R) // bs, r % bs
                    if n >= N or hg != 0:            # the mask does not depend on the head: check head 0 of the group
                        continue
                    kept = bool(keep[b, n])
                    a_r = int(anchors[b, n]) if kept else 0
                    q = n * bs + o
                    for t in range(n_ctx):           # context tiles
                        for cc in range(BKV):
                            key = t * BKV + cc
                            if kept and key < a_r and key < S:
                                got[b, q, key] = True
                    for cc in range(BKV):            # own tile: key column cc is draft row n0*bs + cc
                        if kept and cc // bs == blk and n0 * bs + cc < N * bs:
                            got[b, q, S + n0 * bs + cc] = True
    assert torch.equal(got, want)
    # context-key-stationary backward: per 128-key tile, the kept blocks with anchor > k0 form [n_lo, n_hi)
    for b in range(B):
        for k0 in range(0, S, 128):
            n_hi = n_lo = 0
            for n in range(N):
                if not keep[b, n]:
                    break
                n_hi = n + 1
                if int(anchors[b, n]) <= k0:
                    n_lo = n + 1
            attends = [n for n in range(N) if keep[b, n] and bool(want[b, n * bs, k0:min(k0 + 128, S)].any())]
            assert attends == list(range(n_lo, n_hi))


# ------------------------------------------------------------------------------------------------ sliding-window layers
def test_sliding_mask_semantics():
    """dflash_family_model.py:73-84 — slot o of a block anchored at a sees context keys [a + o - (W - 1), a) and own slots <= o."""
    anchors = torch.tensor([[5, 9]])
    keep = torch.tensor([[True, True]])
    S, bs, W = 12, 4, 3
    m = D.dflash_mask(anchors, keep, S, bs, W)[0]                   # [N*bs, S + N*bs]
    for n, a in enumerate((5, 9)):
        for o in range(bs):
            row = m[n * bs + o]
            ctx = [k for k in range(S) if row[k]]
            assert ctx == [k for k in range(max(0, a + o - (W - 1)), a)], (n, o, ctx)
            own = [k for k in range(bs) if row[S + n * bs + k]]
            assert own == list(range(o + 1)), (n, o, own)
            other = (1 - n) * bs
            assert not row[S + other:S + other + bs].any()
    full = D.dflash_mask(anchors, keep, S, bs)[0]
    assert (m & ~full).sum() == 0 and m.sum() < full.sum()          # a sliding mask only removes keys
    assert torch.equal(D.dflash_mask(anchors, keep, S, bs, S + bs + 1)[0][:, :S], full[:, :S])   # a window wider than the context: same context keys


def test_sliding_mask_equals_the_reference_builder():
    import os
    import sys
    if not os.path.isdir("/root/reference/specforge"):
        pytest.skip("reference checkout not present")
    sys.path.insert(0, "/root/reference")
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    from specforge.algorithms.common.dflash_family_model import create_dflash_sdpa_mask
    gen = torch.Generator().manual_seed(3)
    lm = (torch.rand(3, 70, generator=gen) > 0.2).float()
    lm[2, 6:] = 0
    anchors, keep = D.sample_anchor_positions(lm, 9, generator=gen)
    for W in (None, 1, 5, 33, 200):
        want = create_dflash_sdpa_mask(anchors, keep, 70, 8, torch.device("cpu"), sliding_window=W)[:, 0]
        assert torch.equal(D.dflash_mask(anchors, keep, 70, 8, W), want), W


def test_sliding_layout_validation_and_the_checked_in_hybrid_config():
    """modeling/draft/dflash.py:38-68 and configs/qwen3.6-27b-dflash.json (4 sliding + 1 full layer, window 2048)."""
    from specforge_b200.dflash import DFlashDims, dims_from_config
    cfg = {"hidden_size": 5120, "intermediate_size": 17408, "num_attention_heads": 32, "num_key_value_heads": 8, "head_dim": 128,
           "num_hidden_layers": 5, "num_target_layers": 64, "vocab_size": 248320, "block_size": 16, "rope_theta": 10000000,
           "dflash_config": {"mask_token_id": 248070, "target_layer_ids": [1, 16, 31, 46, 61]}, "sliding_window": 2048,
           "layer_types": ["sliding_attention"] * 4 + ["full_attention"], "use_sliding_window": True, "rope_scaling": None}
    dims = dims_from_config(cfg)
    assert dims.sliding_layout() == (2048, 0b01111)
    ref_json = "/root/reference/configs/qwen3.6-27b-dflash.json"
    if os.path.exists(ref_json):                       # the checked-in file itself, where the reference checkout is present
        import json
        with open(ref_json) as f:
            real = dims_from_config(json.load(f))
        assert real.sliding_layout() == (2048, 0b01111) and real.num_layers == 5 and real.hidden_size == 5120
        assert real.num_target_feats == 5 and real.mask_token_id == 248070 and real.block_size == 16
    base = dict(hidden_size=64, intermediate_size=128, num_heads=4, num_kv_heads=2, head_dim=16, num_layers=2, num_target_feats=2, vocab_size=256)
    assert DFlashDims(**base).sliding_layout() == (0, 0)
    assert DFlashDims(**base, layer_types=("full_attention", "full_attention"), sliding_window=7).sliding_layout() == (0, 0)
    assert DFlashDims(**base, layer_types=("full_attention", "sliding_attention"), sliding_window=7).sliding_layout() == (7, 0b10)
    for types, w in ((("full_attention",), None), (("full_attention", "unknown"), None), (("sliding_attention", "full_attention"), None),
                     (("sliding_attention", "full_attention"), 0), (("sliding_attention", "full_attention"), -1)):
        with pytest.raises(ValueError):
            DFlashDims(**base, layer_types=types, sliding_window=w).sliding_layout()
        with pytest.raises(ValueError):
            D.layer_windows(D.DFlashConfig(num_layers=2, layer_types=types, sliding_window=w))
