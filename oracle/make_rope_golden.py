"""Golden cos/sin rows from the reference's rotary modules (llama3_eagle.py:218-536) for every supported rope_scaling
type — build-container only (imports /root/reference).  Output: tests/golden/rope_tables.pt"""
import os
import sys

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402

from specforge.modeling.draft.llama3_eagle import LlamaAttention  # noqa: E402

CASES = {
    "default": None,
    "linear": {"rope_type": "linear", "factor": 4.0},
    "dynamic": {"rope_type": "dynamic", "factor": 2.0},
    "llama3": {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
               "original_max_position_embeddings": 8192},
    "yarn": {"rope_type": "yarn", "factor": 40.0, "original_max_position_embeddings": 4096, "beta_fast": 32, "beta_slow": 1,
             "mscale": 1.0, "mscale_all_dim": 1.0},
}
POS = [0, 1, 17, 511, 1000, 2047]
out = {"positions": POS, "head_dim": 128, "rope_theta": 500000.0, "max_position_embeddings": 2048, "cases": {}}
for name, sc in CASES.items():
    from types import SimpleNamespace
    cfg = SimpleNamespace(hidden_size=512, num_attention_heads=4, num_key_value_heads=4, max_position_embeddings=2048,
                          head_dim=128, rope_theta=500000.0, rope_scaling=sc)   # v4-style attrs (get_rope_config fallback)
    attn = LlamaAttention(cfg)
    rot = attn.rotary_emb
    cos = rot.cos_cached[0, 0].to(torch.bfloat16)   # the trainer casts the module to bf16 (model_providers.py:112)
    sin = rot.sin_cached[0, 0].to(torch.bfloat16)
    out["cases"][name] = {"scaling": sc, "rows": cos.shape[0], "cos": cos[POS].clone(), "sin": sin[POS].clone()}
    print(name, cos.shape, float(cos[POS].float().abs().sum()))
torch.save(out, os.path.join(ROOT, "tests", "golden", "rope_tables.pt"))
