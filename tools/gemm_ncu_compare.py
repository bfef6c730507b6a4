"""Launches the tcgen05 GEMM and torch.matmul (cuBLAS) on the same operands for a few step shapes — meant to run under
`ncu --set full --clock-control none -k regex:'gemm_kernel|nvjet|cutlass|sm100|xmma'` so the two can be compared metric by
metric (tensor-pipe active %, L2/DRAM throughput, cycles).  Diagnostic only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from specforge_b200 import ops            # noqa: E402

SHAPES = [("o_proj", 16384, 4096, 4096), ("down", 16384, 4096, 12288), ("lm_head", 16384, 32000, 4096)]
dev = "cuda"
torch.manual_seed(0)
for name, m, n, k in SHAPES:
    a = (torch.randn(m, k, device=dev) * 0.05).bfloat16()
    b = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
    out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        ops.gemm(a, b, out=out)
        torch.matmul(a, b.t(), out=out)
    torch.cuda.synchronize()
    del a, b, out
print("done")
