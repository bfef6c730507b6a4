import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from specforge_b200 import ops
B, S, nh, nkv, hd = 8, 2048, 32, 8, 128
qkv = [(torch.randn(B * S, (nh + 2 * nkv) * hd, device="cuda") * 0.5).bfloat16()]
for _ in range(3): ops.ttt_attention_fwd(qkv, B, S, nh, nkv, hd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(20): ops.ttt_attention_fwd(qkv, B, S, nh, nkv, hd)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("SF_ATTN_DEBUG", "0"), "fwd ms", e0.elapsed_time(e1) / 20)
