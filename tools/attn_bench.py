"""Times the TTT attention kernels at config-2 shapes (B=8, S=2048, nh=32, nkv=8, d=128)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from specforge_b200 import ops

B, S, nh, nkv, hd = 8, 2048, 32, 8, 128
dev = "cuda"
torch.manual_seed(0)
res = {}
for J in (0, 6):
    qkv = [(torch.randn(B * S, (nh + 2 * nkv) * hd, device=dev) * 0.5).bfloat16() for _ in range(J + 1)]
    out, lse = ops.ttt_attention_fwd(qkv, B, S, nh, nkv, hd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        out, lse = ops.ttt_attention_fwd(qkv, B, S, nh, nkv, hd)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 4 * (S / 2) * S * nh * hd * B
    res[f"fwd_J{J}"] = {"ms": ms, "tflops": fl / ms / 1e9}
    dout = torch.randn_like(out)
    ops.ttt_attention_bwd(qkv, out, dout, lse, B, S, nh, nkv, hd)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        ops.ttt_attention_bwd(qkv, out, dout, lse, B, S, nh, nkv, hd)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res[f"bwd_J{J}"] = {"ms": ms, "tflops_algorithmic": 2.5 * fl / ms / 1e9}
print(json.dumps(res))
