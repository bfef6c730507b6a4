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


def gemm_rope(x: torch.Tensor, w: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, S: int, pos_offset: int, head_dim: int,
              rope_cols: int) -> torch.Tensor:
    """qkv = RoPE(x @ w.T): projection with the rotary embedding fused into the GEMM epilogue (columns < rope_cols rotated)."""
    import ctypes
    l = lib()
    l.sf_gemm_bf16_rope.restype = ctypes.c_int
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty(M, N, dtype=torch.bfloat16, device=x.device)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    check(l.sf_gemm_bf16_rope(P(x), ctypes.c_int64(x.stride(0)), P(w), ctypes.c_int64(w.stride(0)), P(out), ctypes.c_int64(N), M, N, K,
                              P(cos), P(sin), S, pos_offset, head_dim, rope_cols, ctypes.c_void_p(_stream())), "sf_gemm_bf16_rope")
    return out


def gemm_swiglu(x: torch.Tensor, w_gate_up: torch.Tensor):
    """(gu, act) = fused gate/up projection + SwiGLU (EPI_SWIGLU): x [M, K], w_gate_up [2I, K] -> gu [M, 2I], act [M, I]."""
    import ctypes
    l = lib()
    M, K = x.shape
    I = w_gate_up.shape[0] // 2
    gu = torch.empty(M, 2 * I, dtype=torch.bfloat16, device=x.device)
    act = torch.empty(M, I, dtype=torch.bfloat16, device=x.device)
    l.sf_gemm_bf16_ex.restype = ctypes.c_int
    check(l.sf_gemm_bf16_ex(ctypes.c_void_p(x.data_ptr()), ctypes.c_int64(x.stride(0)), 0, ctypes.c_void_p(w_gate_up.data_ptr()),
                            ctypes.c_int64(w_gate_up.stride(0)), 0, ctypes.c_void_p(gu.data_ptr()), ctypes.c_int64(2 * I), None,
                            ctypes.c_int64(0), ctypes.c_void_p(act.data_ptr()), ctypes.c_int64(I), I, M, 2 * I, K, 4, 0,
                            ctypes.c_void_p(_stream())), "sf_gemm_bf16_ex")
    return gu, act


def gemm_swiglu_bwd(dy: torch.Tensor, w_down: torch.Tensor, gu: torch.Tensor):
    """d(gu) [M, 2I] = SwiGLU backward of d(act) = dy [M, H] @ w_down [H, I] (EPI_SWIGLU_BWD; d(act) never reaches HBM)."""
    import ctypes
    l = lib()
    M, H = dy.shape
    I = w_down.shape[1]
    dgu = torch.empty(M, 2 * I, dtype=torch.bfloat16, device=dy.device)
    l.sf_gemm_bf16_ex.restype = ctypes.c_int
    check(l.sf_gemm_bf16_ex(ctypes.c_void_p(dy.data_ptr()), ctypes.c_int64(dy.stride(0)), 0, ctypes.c_void_p(w_down.data_ptr()),
                            ctypes.c_int64(w_down.stride(0)), 1, ctypes.c_void_p(dgu.data_ptr()), ctypes.c_int64(2 * I),
                            ctypes.c_void_p(gu.data_ptr()), ctypes.c_int64(2 * I), None, ctypes.c_int64(0), I, M, I, H, 5, 0,
                            ctypes.c_void_p(_stream())), "sf_gemm_bf16_ex")
    return dgu


def _declare_ops():
    import ctypes
    from ctypes import POINTER, c_float, c_int, c_int64, c_void_p
    l = lib()
    if getattr(l, "_sf_ops_declared", False):
        return l
    l.sf_rmsnorm_fwd.restype = c_int
    l.sf_rmsnorm_fwd.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_void_p]
    l.sf_rmsnorm_bwd.restype = c_int
    l.sf_rmsnorm_bwd.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_int, c_float, c_void_p]
    l.sf_rmsnorm_bwd_scratch_bytes.restype = c_int64
    l.sf_rmsnorm_bwd_scratch_bytes.argtypes = [c_int]
    l.sf_swiglu_fwd.restype = c_int
    l.sf_swiglu_fwd.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p]
    l.sf_swiglu_bwd.restype = c_int
    l.sf_swiglu_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]
    l.sf_rope.restype = c_int
    l.sf_rope.argtypes = [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_void_p]
    l.sf_ttt_attention_fwd.restype = c_int
    l.sf_ttt_attention_fwd.argtypes = [POINTER(c_void_p), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                       c_int, c_int, c_int, c_void_p]
    l.sf_ttt_attention_bwd.restype = c_int
    l.sf_ttt_attention_bwd.argtypes = [POINTER(c_void_p), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       POINTER(c_void_p), POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                       c_int, c_int, c_int, c_int, c_void_p]
    l._sf_ops_declared = True
    return l


def embedding_gather(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """out[i] = table[ids[i]]  (frozen-embedding lookup, llama3_eagle.py:1759-1760)."""
    import ctypes
    l = lib()
    l.sf_embedding_gather.restype = ctypes.c_int
    out = torch.empty(ids.numel(), table.shape[1], dtype=table.dtype, device=table.device)
    check(l.sf_embedding_gather(ctypes.c_void_p(table.data_ptr()), ctypes.c_int64(table.shape[0]), table.shape[1],
                                ctypes.c_void_p(ids.data_ptr()), ctypes.c_int64(ids.numel()), ctypes.c_void_p(out.data_ptr()),
                                ctypes.c_void_p(_stream())), "sf_embedding_gather")
    return out


def rmsnorm_fwd(x, w, eps, out=None):
    l = _declare_ops()
    if out is None:
        out = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    check(l.sf_rmsnorm_fwd(x.data_ptr(), x.stride(0), w.data_ptr(), out.data_ptr(), out.stride(0), x.shape[0], x.shape[1],
                           eps, _stream()), "sf_rmsnorm_fwd")
    return out


def rmsnorm_bwd(x, w, dy, eps, add=None):
    l = _declare_ops()
    dx = torch.empty_like(x)
    dw = torch.zeros(x.shape[1], dtype=torch.float32, device=x.device)
    scratch = torch.empty(l.sf_rmsnorm_bwd_scratch_bytes(x.shape[1]) // 4, dtype=torch.float32, device=x.device)
    check(l.sf_rmsnorm_bwd(x.data_ptr(), x.stride(0), w.data_ptr(), dy.data_ptr(), dy.stride(0), _ptr(add), dx.data_ptr(),
                           dw.data_ptr(), scratch.data_ptr(), x.shape[0], x.shape[1], eps, _stream()), "sf_rmsnorm_bwd")
    return dx, dw


def swiglu_fwd(gu):
    l = _declare_ops()
    M, I2 = gu.shape
    act = torch.empty(M, I2 // 2, dtype=gu.dtype, device=gu.device)
    check(l.sf_swiglu_fwd(gu.data_ptr(), act.data_ptr(), M, I2 // 2, _stream()), "sf_swiglu_fwd")
    return act


def swiglu_bwd(gu, dact):
    l = _declare_ops()
    dgu = torch.empty_like(gu)
    check(l.sf_swiglu_bwd(gu.data_ptr(), dact.data_ptr(), dgu.data_ptr(), gu.shape[0], gu.shape[1] // 2, _stream()),
          "sf_swiglu_bwd")
    return dgu


def rope_(x, n_heads, head_dim, cos, sin, S, pos_offset, inverse=False):
    l = _declare_ops()
    check(l.sf_rope(x.data_ptr(), x.stride(0), n_heads, head_dim, cos.data_ptr(), sin.data_ptr(), S, pos_offset, x.shape[0],
                    int(inverse), _stream()), "sf_rope")
    return x


def _ptr_array(tensors):
    import ctypes
    arr = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


def ttt_attention_fwd(qkv_list, B, S, nh, nkv, hd, key_mask=None):
    """qkv_list[i]: [B*S, (nh+2nkv)*hd] fused buffers of TTT blocks 0..J (RoPE already applied); queries = block J."""
    l = _declare_ops()
    J = len(qkv_list) - 1
    dev = qkv_list[0].device
    out = torch.empty(B * S, nh * hd, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, nh, S, dtype=torch.float32, device=dev)
    sd = torch.empty(B * nh * S * max(J, 1), dtype=torch.float32, device=dev)
    kvl = torch.zeros(2 * B, dtype=torch.int32, device=dev)
    check(l.sf_ttt_attention_fwd(_ptr_array(qkv_list), J, out.data_ptr(), lse.data_ptr(), sd.data_ptr(), _ptr(key_mask),
                                 kvl.data_ptr(), B, S, nh, nkv, hd, _stream()), "sf_ttt_attention_fwd")
    return out, lse


def ttt_attention_bwd(qkv_list, out, dout, lse, B, S, nh, nkv, hd, key_mask=None):
    l = _declare_ops()
    J = len(qkv_list) - 1
    dev = out.device
    dk = [torch.zeros(B * S, nkv * hd, dtype=torch.float32, device=dev) for _ in range(J + 1)]
    dv = [torch.zeros(B * S, nkv * hd, dtype=torch.float32, device=dev) for _ in range(J + 1)]
    dq = torch.empty(B * S, nh * hd, dtype=torch.bfloat16, device=dev)
    sd = torch.empty(B * nh * S * max(J, 1), dtype=torch.float32, device=dev)
    delta = torch.empty(B * nh * S, dtype=torch.float32, device=dev)
    dqd = torch.empty(B * S, nh * hd, dtype=torch.float32, device=dev)
    kvl = torch.zeros(2 * B, dtype=torch.int32, device=dev)
    check(l.sf_ttt_attention_bwd(_ptr_array(qkv_list), J, out.data_ptr(), dout.data_ptr(), lse.data_ptr(), sd.data_ptr(),
                                 _ptr(key_mask), _ptr_array(dk), _ptr_array(dv), dq.data_ptr(), delta.data_ptr(),
                                 dqd.data_ptr(), kvl.data_ptr(), B, S, nh, nkv, hd, _stream()), "sf_ttt_attention_bwd")
    return dq, dk, dv
