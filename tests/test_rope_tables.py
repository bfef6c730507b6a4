"""RoPE table builders (every rope_scaling flavour) vs rows produced by the reference's rotary modules
(tests/golden/rope_tables.pt, oracle/make_rope_golden.py).  CPU only."""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "rope_tables.pt")


@pytest.mark.parametrize("name", ["default", "linear", "dynamic", "llama3", "yarn"])
def test_rope_tables_match_reference(name):
    from specforge_b200.engine import DraftDims, rope_tables
    g = torch.load(GOLD)
    case = g["cases"][name]
    dims = DraftDims(hidden_size=512, intermediate_size=1024, num_heads=4, num_kv_heads=4, head_dim=g["head_dim"], vocab_size=1024,
                     draft_vocab_size=256, rope_theta=g["rope_theta"], max_position_embeddings=g["max_position_embeddings"],
                     rope_scaling=case["scaling"])
    cos, sin = rope_tables(dims, case["rows"], "cpu")
    pos = g["positions"]
    # bf16 tables from fp32 math: allow one bf16 ulp where fp32 op order differs
    torch.testing.assert_close(cos[pos].float(), case["cos"].float(), rtol=0, atol=2 ** -7)
    torch.testing.assert_close(sin[pos].float(), case["sin"].float(), rtol=0, atol=2 ** -7)
    assert (cos[pos] != case["cos"]).float().mean() < 0.02 and (sin[pos] != case["sin"]).float().mean() < 0.02


def test_mrope_with_a_scaling_factor_is_rejected():
    from specforge_b200.draft import dims_from_config
    with pytest.raises(NotImplementedError):
        dims_from_config({"hidden_size": 256, "num_attention_heads": 4, "intermediate_size": 512, "vocab_size": 1024,
                          "draft_vocab_size": 256, "rope_scaling": {"type": "mrope", "mrope_section": [16, 24, 24], "factor": 2.0}})


def test_mrope_text_positions_equal_the_default_tables():
    """rope_scaling type "mrope" (llama3_eagle.py:145-183,389-424) with text positions (three equal axes): the sectioned cos / sin
    the reference builds on the fly equal our default tables row for row, so the CUDA path serves text-only mrope drafts."""
    import os
    import sys
    import pytest
    if not os.path.isdir("/root/reference/specforge") and not os.path.isdir(os.path.join(os.path.dirname(__file__), "..", "baseline", "_ref", "specforge")):
        pytest.skip("reference not importable")
    sys.path.insert(0, "/root/reference" if os.path.isdir("/root/reference/specforge") else os.path.join(os.path.dirname(__file__), "..", "baseline", "_ref"))
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    import torch
    from specforge.modeling.draft.llama3_eagle import LlamaMutiRotaryEmbedding
    from specforge_b200.draft import dims_from_config
    from specforge_b200.engine import rope_tables
    cfg = dict(hidden_size=256, intermediate_size=512, num_attention_heads=4, num_key_value_heads=2, head_dim=64, vocab_size=1024,
               draft_vocab_size=256, rope_theta=1e6, max_position_embeddings=512,
               rope_scaling={"rope_type": "mrope", "mrope_section": [8, 12, 12]})
    dims = dims_from_config(cfg)
    assert dims.rope_scaling is None
    cos, sin = rope_tables(dims, 300, "cpu")
    emb = LlamaMutiRotaryEmbedding(64, max_position_embeddings=512, base=1e6)
    pos = torch.arange(300).view(1, 1, 300).expand(3, 1, 300)
    c3, s3 = emb(torch.zeros(1, dtype=torch.bfloat16), pos)            # [3, 1, 300, 64]
    sec = [8, 12, 12] * 2
    rc = torch.cat([m[i % 3] for i, m in enumerate(c3.split(sec, dim=-1))], dim=-1)[0]
    rs = torch.cat([m[i % 3] for i, m in enumerate(s3.split(sec, dim=-1))], dim=-1)[0]
    assert torch.equal(rc, cos) and torch.equal(rs, sin)
    from specforge_b200.strategy import B200Eagle3TrainStrategy
    ids = torch.zeros(2, 300, dtype=torch.long)
    B200Eagle3TrainStrategy._check_position_ids(pos.expand(3, 2, 300), ids)
    with pytest.raises(NotImplementedError):
        B200Eagle3TrainStrategy._check_position_ids(torch.stack([pos[0].expand(2, 300), pos[0].expand(2, 300) + 1, pos[0].expand(2, 300)]), ids)
