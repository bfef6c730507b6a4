"""Sustained throughput of the step's GEMM shapes: libspecforge_b200's tcgen05 kernel vs torch.matmul (cuBLAS) on the SAME
box, alternated in blocks of ~`--ms` milliseconds each so the power/clock state is comparable (the pool's B200s differ by
+-8 % and every number on a 1 kW-capped part is a clock statement).  Diagnostic, not a bench value.
    python tools/gemm_vs_cublas.py [--ms 600] [--rounds 3] [--out gpurun_out/gemm_vs_cublas.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from specforge_b200 import ops            # noqa: E402

M, H, I, QKV, A, DV, V, T = 16384, 4096, 12288, 6144, 4096, 32000, 151936, 7
SHAPES = [
    # name, (a_major, b_major), M, N, K, epi
    ("fwd qkv        [M,2H]x[QKV,2H]", (0, 0), M, QKV, 2 * H, 0),
    ("fwd o_proj     [M,A]x[H,A]", (0, 0), M, H, A, 0),
    ("fwd gate_up    [M,H]x[2I,H]", (0, 0), M, 2 * I, H, 0),
    ("fwd down       [M,I]x[H,I]", (0, 0), M, H, I, 0),
    ("fwd lm_head    [M,H]x[DV,H]", (0, 0), M, DV, H, 0),
    ("fwd teacher/2  [M/2,H]x[V,H]", (0, 0), M // 2, V, H, 0),
    ("dgrad lm_head  [M,DV]x[DV,H]", (0, 1), M, H, DV, 0),
    ("dgrad down     [M,H]x[H,I]", (0, 1), M, I, H, 0),
    ("dgrad gate_up  [M,2I]x[2I,H]", (0, 1), M, H, 2 * I, 0),
    ("wgrad lm_head  [TM,DV]^T[TM,H]", (1, 1), DV, H, T * M, 2),
    ("wgrad gate_up  [TM,2I]^T[TM,H]", (1, 1), 2 * I, H, T * M, 2),
    ("wgrad down     [TM,H]^T[TM,I]", (1, 1), H, I, T * M, 2),
]


def bench_block(fn, target_ms):
    """Run fn back to back for about target_ms; return (TFLOP-agnostic) ms per call measured with CUDA events."""
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    fn(); fn()
    torch.cuda.synchronize()
    s.record(); fn(); e.record(); torch.cuda.synchronize()
    one = max(s.elapsed_time(e), 1e-3)
    n = max(3, int(target_ms / one))
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ms", type=float, default=600.0)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/gemm_vs_cublas.json")
    ap.add_argument("--only", default=None)
    ap.add_argument("--variants", nargs="*", default=[], help="sf_debug_option settings to time as extra arms, e.g. gemm_stages=7")
    a = ap.parse_args()
    import ctypes
    from specforge_b200._lib import lib
    L = lib()
    L.sf_debug_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    dev = "cuda"
    torch.manual_seed(0)
    res = []
    for name, (am, bm), m, n, k, epi in SHAPES:
        if a.only and a.only not in name:
            continue
        # operands as the step holds them: K-major [rows, K] or MN-major [K, rows]
        Aop = (torch.randn((m, k) if am == 0 else (k, m), device=dev) * 0.05).bfloat16()
        Bop = (torch.randn((n, k) if bm == 0 else (k, n), device=dev) * 0.05).bfloat16()
        out = torch.empty(m, n, device=dev, dtype=torch.float32 if epi == 2 else torch.bfloat16)
        ref_out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
        Aref = Aop if am == 0 else Aop.t()
        Bref = Bop.t() if bm == 0 else Bop

        def ours():
            ops.gemm(Aop, Bop, a_major=am, b_major=bm, out=out, epi=epi)

        def cublas():
            torch.matmul(Aref, Bref, out=ref_out)    # bf16 output (cuBLAS has no fp32-out bf16 GEMM through torch)

        t_ours, t_cub = [], []
        t_var = {v: [] for v in a.variants}
        for _ in range(a.rounds):
            t_ours.append(bench_block(ours, a.ms))
            for v in a.variants:
                nm, val = v.split("=")
                assert L.sf_debug_option(nm.encode(), int(val)) == 0, v
                t_var[v].append(bench_block(ours, a.ms))
                L.sf_debug_option(nm.encode(), 0)
            t_cub.append(bench_block(cublas, a.ms))
        fl = 2.0 * m * n * k / 1e9
        r = {"shape": name, "M": m, "N": n, "K": k, "ours_tflops": [round(fl / t, 1) for t in t_ours],
             "cublas_tflops": [round(fl / t, 1) for t in t_cub],
             "ratio_median": round(sorted(t_cub)[len(t_cub) // 2] / sorted(t_ours)[len(t_ours) // 2], 4)}
        for v in a.variants:
            r[v] = [round(fl / t, 1) for t in t_var[v]]
        print(json.dumps(r), flush=True)
        res.append(r)
        del Aop, Bop, out, ref_out
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
