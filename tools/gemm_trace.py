"""Where the 512 x 256 GEMM tiling spends its time per tile: clock64() stamps written by cluster 0 (sf_debug_gemm_trace).
    python tools/gemm_trace.py [M N K]
Columns (cycles relative to the tile's start in the MMA issuer): full0 = first operand stage landed, acc0 = half-0 accumulator free,
acc1 = half-1 accumulator free, mid = half the k-blocks issued, end = all MMAs issued; epilogue (warp 4): f0w/f0 = wait start / half 0
complete, d0 = half 0 drained, f1w/f1/d1 likewise."""
import ctypes
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from specforge_b200 import ops            # noqa: E402
from specforge_b200._lib import lib       # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (16384, 4096, 4096)
L = lib()
a = (torch.randn(M, K, device="cuda") * 0.05).bfloat16()
b = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
for _ in range(3):
    ops.gemm(a, b, out=out)
L.sf_debug_gemm_trace.argtypes = [ctypes.c_void_p]
L.sf_debug_gemm_trace.restype = None
L.sf_debug_gemm_trace(ctypes.c_void_p(buf.data_ptr()))
ops.gemm(a, b, out=out)
torch.cuda.synchronize()
L.sf_debug_gemm_trace(None)
t = buf.cpu().view(32, 16)
print(f"shape {M}x{N}x{K}; k-blocks {K // 64}; ideal MMA time per tile {K // 64 * 1024} clk")
print("tile  start(abs-prev)  full0  acc0  acc1  mid  end | epi: f0w  f0  d0  f1w  f1  d1")
prev = None
for i in range(32):
    r = t[i]
    if int(r[0]) == 0:
        break
    s0 = int(r[0])
    rel = lambda j: int(r[j]) - s0 if int(r[j]) else -1
    print(f"{i:3d}  {'' if prev is None else s0 - prev:>12}  {rel(1):6d} {rel(2):6d} {rel(3):6d} {rel(4):7d} {rel(5):7d} | {rel(8):7d} {rel(9):7d} {rel(10):7d} {rel(11):7d} {rel(12):7d} {rel(13):7d}")
    prev = s0
