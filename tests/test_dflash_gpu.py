"""DFlash step on the GPU (first correct CUDA version, `sf_dflash_*`) against the reference goldens and the bf16 oracle."""
import glob
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "dflash_*.pt")))


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-3])
def test_dflash_step_matches_reference_and_oracle(path):
    from oracle import dflash_oracle as D
    from specforge_b200.dflash import DFlashDims, DFlashEngine
    g = torch.load(path, weights_only=False)
    c = D.DFlashConfig(**g["config"])
    dims = DFlashDims(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size, num_heads=c.num_heads, num_kv_heads=c.num_kv_heads,
                      head_dim=c.head_dim, num_layers=c.num_layers, num_target_feats=c.num_target_feats, vocab_size=c.vocab_size,
                      block_size=c.block_size, mask_token_id=c.mask_token_id, rms_norm_eps=c.rms_norm_eps, rope_theta=c.rope_theta,
                      max_position_embeddings=1024, loss_decay_gamma=c.loss_decay_gamma, loss_type=c.loss_type, dpace_alpha=c.dpace_alpha,
                      layer_types=c.layer_types, sliding_window=c.sliding_window)
    B, S = g["batch"]["input_ids"].shape
    N = g["anchors"].shape[1]
    eng = DFlashEngine(dims, batch=B, seq_len=S, num_blocks=N)
    P16 = {k: v.bfloat16() for k, v in g["params"].items()}
    eng.load_params(P16)
    eng.set_frozen(embed_tokens=g["embed_w"], lm_head=g["lm_head_w"])
    loss, metrics = eng.forward(g["batch"], g["anchors"], g["keep"])
    eng.backward()
    torch.cuda.synchronize()
    loss, metrics = float(loss), metrics.cpu()
    # --- the bf16 oracle on the same (bf16-rounded) inputs
    b16 = dict(g["batch"]); b16["hidden_states"] = b16["hidden_states"].bfloat16()
    o_loss, o_acc, o_terms, o_grads = D.train_step(P16, c, b16, g["anchors"], g["keep"], g["embed_w"].bfloat16(), g["lm_head_w"].bfloat16())
    assert abs(loss - float(o_loss)) <= 5e-3 * abs(float(o_loss)), (loss, float(o_loss))
    assert float(metrics[1]) == pytest.approx(float(o_terms["loss_den"]), rel=1e-6)
    assert float(metrics[3]) == float(o_terms["acc_den"])
    assert abs(float(metrics[2]) - float(o_terms["correct"])) <= 1
    for n in eng.names:
        got = eng.param_view(n, eng.grads_f32).cpu()
        ref = o_grads[n].float()
        cos = _cos(got, ref)
        assert cos >= (0.98 if got.dim() == 1 else 0.995), (n, cos)
        assert got.norm().item() == pytest.approx(ref.norm().item(), rel=5e-2), n
    # --- the unmodified reference's own numbers when the golden was produced in bf16
    if g["dtype"] == "torch.bfloat16":
        assert abs(loss - float(g["loss"])) <= 2e-2 * abs(float(g["loss"]))
        for n in eng.names:
            assert _cos(eng.param_view(n, eng.grads_f32).cpu(), g["grads"][n]) >= 0.99, n


def test_dflash_deterministic_accumulate_and_eval():
    from oracle import dflash_oracle as D
    from specforge_b200.dflash import DFlashDims, DFlashEngine, sample_anchor_positions
    c = D.DFlashConfig(hidden_size=128, intermediate_size=256, num_heads=4, num_kv_heads=2, head_dim=32, num_layers=2, num_target_feats=2,
                       vocab_size=512, block_size=8, num_anchors=12, mask_token_id=511)
    dims = DFlashDims(hidden_size=128, intermediate_size=256, num_heads=4, num_kv_heads=2, head_dim=32, num_layers=2, num_target_feats=2,
                      vocab_size=512, block_size=8, mask_token_id=511, rope_theta=10000.0, max_position_embeddings=1024)
    B, S = 3, 160
    gen = torch.Generator().manual_seed(0)
    batch = {"input_ids": torch.randint(0, 500, (B, S), generator=gen), "hidden_states": torch.randn(B, S, 256, generator=gen).bfloat16(),
             "loss_mask": (torch.rand(B, S, generator=gen) > 0.2).float()}
    batch["loss_mask"][2, 9:] = 0                                     # at most 8 candidates in the last row -> dropped blocks
    anchors, keep = sample_anchor_positions(batch["loss_mask"], 12, generator=gen)
    assert not bool(keep.all()) and bool(keep.any())
    eng = DFlashEngine(dims, batch=B, seq_len=S, num_blocks=12)
    eng.load_params(D.init_params(c, seed=1))
    eng.set_frozen(embed_tokens=torch.randn(512, 128, generator=gen) * 0.5, lm_head=torch.randn(512, 128, generator=gen) * 0.2)
    eng.forward(batch, anchors, keep); eng.backward()
    l1, m1, g1 = eng.loss.clone(), eng.metrics.clone(), eng.grads_f32.clone()
    eng.forward(batch, anchors, keep); eng.backward()
    assert torch.equal(eng.loss, l1) and torch.equal(eng.metrics, m1) and torch.equal(eng.grads_f32, g1)   # no atomics anywhere
    eng.forward(batch, anchors, keep); eng.backward(accumulate=True)
    torch.testing.assert_close(eng.grads_f32, 2 * g1, rtol=1e-5, atol=1e-8)
    eng.forward(batch, anchors, keep, need_grad=False)
    assert torch.equal(eng.loss, l1)
    assert torch.isfinite(g1).all() and float(g1.abs().sum()) > 0


TC_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "dflashtc_*.pt")))


@pytest.mark.parametrize("tc", [-1, 0], ids=["cuda_core", "tcgen05"])   # option value: -1 forces CUDA cores, 0 = default (tcgen05)
@pytest.mark.parametrize("path", TC_GOLDEN, ids=lambda p: os.path.basename(p)[:-3])
def test_dflash_tc_shapes(path, tc):
    import ctypes
    from specforge_b200._lib import lib
    L = lib()
    L.sf_debug_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    assert L.sf_debug_option(b"dflash_attn_tc", tc) == 0
    try:
        test_dflash_step_matches_reference_and_oracle(path)
    finally:
        L.sf_debug_option(b"dflash_attn_tc", 0)
