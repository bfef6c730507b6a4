"""Torch-tensor convenience wrappers over the C-ABI (device pointers + sizes only cross the boundary)."""
from __future__ import annotations

import torch

from ._lib import check, lib

MAJOR_K, MAJOR_MN = 0, 1
EPI_BF16, EPI_BF16_RESID, EPI_F32, EPI_F32_ACCUM = 0, 1, 2, 3


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_major=MAJOR_K, b_major=MAJOR_K, out=None, residual=None,
         epi=EPI_BF16, cta_group=0) -> torch.Tensor:
    """D[m,n] = sum_k A(m,k) B(n,k).  K-major operand: [rows, K]; MN-major operand: [K, rows]."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_cuda and b.is_cuda
    assert a.stride(-1) == 1 and b.stride(-1) == 1
    M, K = (a.shape if a_major == MAJOR_K else (a.shape[1], a.shape[0]))
    N, Kb = (b.shape if b_major == MAJOR_K else (b.shape[1], b.shape[0]))
    assert K == Kb, (a.shape, b.shape)
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32 if epi >= EPI_F32 else torch.bfloat16)
    check(lib().sf_gemm_bf16(a.data_ptr(), a.stride(0), a_major, b.data_ptr(), b.stride(0), b_major,
                             out.data_ptr(), out.stride(0), _ptr(residual),
                             0 if residual is None else residual.stride(0), M, N, K, epi, cta_group, _stream()),
          "sf_gemm_bf16")
    return out
