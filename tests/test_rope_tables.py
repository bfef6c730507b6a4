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


def test_mrope_is_rejected():
    from specforge_b200.draft import dims_from_config
    with pytest.raises(NotImplementedError):
        dims_from_config({"hidden_size": 256, "num_attention_heads": 4, "intermediate_size": 512, "vocab_size": 1024,
                          "draft_vocab_size": 256, "rope_scaling": {"type": "mrope", "mrope_section": [16, 24, 24]}})
