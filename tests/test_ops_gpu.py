"""Kernel-level parity (GPU): each CUDA op, called through the C ABI, against the oracle / an fp32 torch
restatement of the same op on the same seeded inputs.  Tolerances are bf16-level and stated per test."""
import ctypes
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


def _trunc_normal(shape, std=0.02, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * std).clamp_(-2 * std, 2 * std)


@pytest.mark.parametrize("G", [1, 2])
@pytest.mark.parametrize("am,bm", [(0, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K,epi", [(512, 512, 256, 0), (384, 320, 200, 2), (1000, 776, 1096, 1), (136, 72, 64, 3)])
def test_gemm(G, am, bm, M, N, K, epi):
    from specforge_b200 import ops
    torch.manual_seed(0)
    dev = _dev()
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    a = A if am == 0 else A.t().contiguous()
    b = B if bm == 0 else B.t().contiguous()
    ref = A.float() @ B.float().t()
    R = out0 = out = None
    if epi == 1:
        R = torch.randn(M, N, device=dev).bfloat16()
        ref = ref.bfloat16().float() + R.float()
    if epi == 3:
        out0 = torch.randn(M, N, device=dev)
        ref = ref + out0
        out = out0.clone()
    got = ops.gemm(a, b, a_major=am, b_major=bm, out=out, residual=R, epi=epi, cta_group=G).float()
    tol = 1e-5 if epi >= 2 else 2 ** -7     # fp32 out: accumulation-order noise only; bf16 out: 1 rounding (+1 for resid)
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() <= tol * scale * (2 if epi == 1 else 1) + 1e-4


def _set_opt(name, value):
    from specforge_b200._lib import lib
    L = lib()
    L.sf_debug_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    assert L.sf_debug_option(name, value) == 0


@pytest.mark.parametrize("am,bm", [(0, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K,epi", [(1024, 768, 320, 0), (1000, 776, 1096, 1), (520, 264, 200, 3), (2048, 512, 4160, 2)])
def test_gemm_wide_tiling(am, bm, M, N, K, epi):
    """The 512 x 256 tiling (sf_gemm_wide.cuh; forced with gemm_wide = 2) against the fp32 reference and, bit for bit, against
    the 256 x 256 tiling: every output element accumulates the same K sequence in the same order in both."""
    from specforge_b200 import ops
    torch.manual_seed(1)
    dev = _dev()
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    a = A if am == 0 else A.t().contiguous()
    b = B if bm == 0 else B.t().contiguous()
    ref = A.float() @ B.float().t()
    R = out0 = None
    if epi == 1:
        R = torch.randn(M, N, device=dev).bfloat16()
        ref = ref.bfloat16().float() + R.float()
    if epi == 3:
        out0 = torch.randn(M, N, device=dev)
        ref = ref + out0
    outs = {}
    try:
        for wide in (-1, 2):
            _set_opt(b"gemm_wide", wide)
            outs[wide] = ops.gemm(a, b, a_major=am, b_major=bm, out=None if out0 is None else out0.clone(), residual=R, epi=epi)
            torch.cuda.synchronize()
    finally:
        _set_opt(b"gemm_wide", 0)
    got = outs[2].float()
    tol = 1e-5 if epi >= 2 else 2 ** -7
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() <= tol * scale * (2 if epi == 1 else 1) + 1e-4
    assert torch.equal(outs[2], outs[-1])


def test_gemm_wide_fused_epilogues():
    """SwiGLU forward / backward epilogues and the row-statistics epilogue under the 512 x 256 tiling == the 256 x 256 tiling."""
    from specforge_b200 import ops
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(3)
    M, H, I = 1536, 512, 1280
    x = torch.randn(M, H, device=dev, generator=g).bfloat16()
    w = (torch.randn(2 * I, H, device=dev, generator=g) * H ** -0.5).bfloat16()
    dy = (torch.randn(M, H, device=dev, generator=g) * 0.1).bfloat16()
    wd = (torch.randn(H, I, device=dev, generator=g) * I ** -0.5).bfloat16()
    res = {}
    try:
        for wide in (-1, 2):
            _set_opt(b"gemm_wide", wide)
            gu, act = ops.gemm_swiglu(x, w)
            res[wide] = (gu, act, ops.gemm_swiglu_bwd(dy, wd, gu))
            torch.cuda.synchronize()
    finally:
        _set_opt(b"gemm_wide", 0)
    for t2, t1 in zip(res[2], res[-1]):
        assert torch.equal(t2, t1)


@pytest.mark.parametrize("M,N,K,epi", [(1000, 776, 1096, 0), (1000, 776, 1096, 1), (300, 264, 64, 1)])
def test_gemm_staged_epilogue_narrow(M, N, K, epi):
    """The warp-staged (coalesced) bf16 epilogue of the 256 x 256 tiling (gemm_epi_staged = 1) == the direct per-thread stores,
    ragged M / N included; plus the fused SwiGLU epilogues under staging."""
    from specforge_b200 import ops
    torch.manual_seed(2)
    dev = _dev()
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    b = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    R = torch.randn(M, N, device=dev).bfloat16() if epi == 1 else None
    x = (torch.randn(1000, 512, device=dev) * 0.5).bfloat16()
    w = (torch.randn(2 * 1280, 512, device=dev) * 0.05).bfloat16()
    dy = (torch.randn(1000, 512, device=dev) * 0.1).bfloat16()
    wd = (torch.randn(512, 1280, device=dev) * 0.03).bfloat16()
    outs = {}
    try:
        for st in (-1, 1):
            _set_opt(b"gemm_wide", -1)
            _set_opt(b"gemm_epi_staged", st)
            outs[st] = ops.gemm(a, b, residual=R, epi=epi)
            if M == 1000 and epi == 0:
                gu, act = ops.gemm_swiglu(x, w)
                outs[(st, "sw")] = (gu, act, ops.gemm_swiglu_bwd(dy, wd, gu))
            torch.cuda.synchronize()
    finally:
        _set_opt(b"gemm_wide", 0)
        _set_opt(b"gemm_epi_staged", 0)
    assert torch.equal(outs[1], outs[-1])
    if (1, "sw") in outs:
        for t2, t1 in zip(outs[(1, "sw")], outs[(-1, "sw")]):
            assert torch.equal(t2, t1)


@pytest.mark.parametrize("hd,nh,nkv,M,S", [(128, 4, 2, 640, 160), (64, 6, 2, 600, 100), (128, 2, 1, 96, 96)])
def test_gemm_rope_epilogue(hd, nh, nkv, M, S):
    """RoPE fused into the [q;k;v] projection's epilogue (EPI_BF16_ROPE) == the plain GEMM followed by the rope kernel, bit for bit,
    in both tilings (M >= 512: 512 x 256 warp-staged; M = 96: the 1-CTA kernel), positions wrapping at sequence boundaries."""
    from oracle import eagle3_oracle as O
    from specforge_b200 import ops
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(4)
    cfg = O.Eagle3Config(hidden_size=256, intermediate_size=512, num_heads=nh, num_kv_heads=nkv, head_dim=hd, vocab_size=1024,
                         draft_vocab_size=256, rope_theta=1e4, max_position_embeddings=512)
    cos, sin = O.rope_tables(cfg, torch.bfloat16)
    cos, sin = cos.to(dev).contiguous(), sin.to(dev).contiguous()
    K, N = 320, (nh + 2 * nkv) * hd
    x = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    for wide in (-1, 0):
        _set_opt(b"gemm_wide", wide)
        try:
            fused = ops.gemm_rope(x, w, cos, sin, S, 3, hd, (nh + nkv) * hd)
            ref = ops.gemm(x, w)
        finally:
            _set_opt(b"gemm_wide", 0)
        ops.rope_(ref, nh + nkv, hd, cos, sin, S, 3, inverse=False)
        torch.cuda.synchronize()
        # same arithmetic (fp32 on the bf16-rounded projection, one rounding); the two kernels may contract x1*c - x2*s into FMAs
        # differently, so allow a last-bit difference on a vanishing fraction of the elements
        assert (fused != ref).float().mean().item() < 1e-3, wide
        assert (fused.float() - ref.float()).abs().max().item() <= 2 ** -7 * ref.float().abs().max().item(), wide
        assert torch.equal(fused[:, (nh + nkv) * hd:], ref[:, (nh + nkv) * hd:])       # v columns pass through untouched


def test_gemm_step_shapes():
    """The shapes the headline step runs that the small cases above do not reach: the K = T*M = 114 688 weight-gradient
    contraction (MN-major x MN-major, fp32 accumulate in TMEM, EPI_F32_ACCUM) and an N = 151 936 target-head row block."""
    from specforge_b200 import ops
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(0)
    # wgrad: dW [4096, 4096] += dY^T X over 114 688 tokens (o_proj's weight gradient at config 2)
    K, M, N = 114688, 4096, 4096
    a = (torch.randn(K, M, device=dev, generator=g) * 0.05).bfloat16()       # dY, MN-major operand [K, rows]
    b = (torch.randn(K, N, device=dev, generator=g) * 0.5).bfloat16()        # X
    out0 = torch.randn(M, N, device=dev, generator=g)
    out = out0.clone()
    ops.gemm(a, b, a_major=1, b_major=1, out=out, epi=3)
    ref = out0.double()
    for k0 in range(0, K, 16384):                                            # fp64 reference in K chunks
        ref += a[k0:k0 + 16384].double().t() @ b[k0:k0 + 16384].double()
    err = (out.double() - ref).abs().max().item()
    # fp32 accumulation of 1.1e5 products in TMEM (7168 dependent UMMA adds per element): stated bound 5e-4 of the largest entry
    print(f"wgrad K=114688 fp32-accumulate: max|err| {err:.3e} of max|ref| {ref.abs().max().item():.3e}")
    assert err <= 5e-4 * ref.abs().max().item(), (err, ref.abs().max().item())
    del a, b, out, out0, ref
    # target head: [2048, 4096] x [151936, 4096]^T, bf16 out
    M, N, K = 2048, 151936, 4096
    x = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) * (K ** -0.5)).bfloat16()
    y = ops.gemm(x, w)
    ref = x.float() @ w.float().t()
    scale = ref.abs().max().item()
    assert (y.float() - ref).abs().max().item() <= 2 ** -7 * scale
    # bit-level agreement with the bf16 rounding of the exact product except where fp32 accumulation order moves a tie
    assert (y != ref.bfloat16()).float().mean().item() < 2e-2


def test_gemm_fused_swiglu_epilogues():
    """EPI_SWIGLU / EPI_SWIGLU_BWD directly (a13: LlamaMLP, llama3_eagle.py:1518-1549) against the unfused kernels and an fp32
    restatement: act = bf16(bf16(silu(g)) * u) from the bf16-rounded gate/up outputs; d(gu) from d(act) = dY W_down."""
    from specforge_b200 import ops
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(1)
    M, H, I = 1024, 512, 1280                      # I % 128 == 0; 5 n-blocks of 256 with a ragged last block
    x = torch.randn(M, H, device=dev, generator=g).bfloat16()
    w = (torch.randn(2 * I, H, device=dev, generator=g) * H ** -0.5).bfloat16()
    gu, act = ops.gemm_swiglu(x, w)
    gu_ref = ops.gemm(x, w)
    assert torch.equal(gu, gu_ref)                                           # same accumulation, same rounding
    assert torch.equal(act, ops.swiglu_fwd(gu_ref))                          # the unfused kernel on the same gu
    ref = torch.nn.functional.silu(gu.float()[:, :I]).bfloat16().float() * gu.float()[:, I:]
    assert (act.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    dy = (torch.randn(M, H, device=dev, generator=g) * 0.1).bfloat16()
    wd = (torch.randn(H, I, device=dev, generator=g) * I ** -0.5).bfloat16()
    dgu = ops.gemm_swiglu_bwd(dy, wd, gu)
    dact = ops.gemm(dy, wd, b_major=1)                                       # bf16 d(act) as the unfused path stores it
    dgu_ref = ops.swiglu_bwd(gu, dact)
    assert (dgu.float() - dgu_ref.float()).abs().max().item() <= 2 ** -6 * dgu_ref.float().abs().max().item()


def test_rmsnorm_fwd_bwd():
    from oracle import eagle3_oracle as O
    from specforge_b200 import ops
    dev = _dev()
    M, H = 300, 896
    x = torch.randn(M, H, device=dev).bfloat16()
    w = (1 + 0.1 * torch.randn(H, device=dev)).bfloat16()
    dy = torch.randn(M, H, device=dev).bfloat16()
    add = torch.randn(M, H, device=dev).bfloat16()
    y = ops.rmsnorm_fwd(x, w, 1e-6)
    ref = O.rms_norm(x, w, 1e-6)
    assert (y.float() - ref.float()).abs().max().item() <= 2 ** -6 * ref.float().abs().max().item()
    # backward vs autograd of the fp32 restatement
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    yf = wf * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6))
    yf.backward(dy.float())
    dx, dw = ops.rmsnorm_bwd(x, w, dy, 1e-6, add=add)
    ref_dx = xf.grad + add.float()
    assert (dx.float() - ref_dx).abs().max().item() <= 2 ** -6 * ref_dx.abs().max().item()
    assert (dw - wf.grad).abs().max().item() <= 2e-2 * wf.grad.abs().max().item()


def test_swiglu_and_rope():
    from oracle import eagle3_oracle as O
    from specforge_b200 import ops
    dev = _dev()
    M, I = 257, 4864
    gu = torch.randn(M, 2 * I, device=dev).bfloat16()
    act = ops.swiglu_fwd(gu)
    ref = torch.nn.functional.silu(gu[:, :I]) * gu[:, I:]
    assert (act.float() - ref.float()).abs().max().item() <= 2 ** -7 * ref.float().abs().max().item()
    g = gu[:, :I].float().requires_grad_(True)
    u = gu[:, I:].float().requires_grad_(True)
    d = torch.randn(M, I, device=dev).bfloat16()
    (torch.nn.functional.silu(g) * u).backward(d.float())
    dgu = ops.swiglu_bwd(gu, d)
    refd = torch.cat([g.grad, u.grad], dim=1)
    assert (dgu.float() - refd).abs().max().item() <= 2 ** -6 * refd.abs().max().item()
    # rope at offset 3 vs the oracle formula evaluated in fp32 with the bf16 tables
    cfg = O.Eagle3Config(hidden_size=256, intermediate_size=512, num_heads=4, num_kv_heads=1, head_dim=128, vocab_size=1024,
                         draft_vocab_size=256, rope_theta=1e6, max_position_embeddings=512)
    cos, sin = O.rope_tables(cfg, torch.bfloat16)
    cos, sin = cos.to(dev), sin.to(dev)
    B, S, nh, hd = 2, 70, 5, 128
    x = torch.randn(B * S, nh * hd, device=dev).bfloat16()
    pos = (torch.arange(S, device=dev) + 3).repeat(B)
    xr = x.float().view(B * S, nh, hd)
    c, s_ = cos[pos].float()[:, None], sin[pos].float()[:, None]
    ref = xr * c + O.rotate_half(xr) * s_
    y = x.clone()
    ops.rope_(y, nh, hd, cos, sin, S, 3, inverse=False)
    assert (y.float().view_as(ref) - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    ops.rope_(y, nh, hd, cos, sin, S, 3, inverse=True)   # transpose rotation: R^T R x = (cos^2+sin^2) x ~ x
    assert (y.float() - x.float()).abs().max().item() <= 0.05 * x.float().abs().max().item()


@pytest.mark.parametrize("hd,nh,nkv,S,J,pad", [(128, 4, 1, 160, 0, 0), (128, 4, 1, 160, 3, 9), (64, 14, 2, 128, 2, 0),
                                               (128, 8, 2, 333, 6, 40), (64, 4, 4, 64, 1, 0),
                                               (128, 32, 8, 2048, 6, 200)])   # config-2 head geometry, full S, padded
def test_ttt_attention_fwd_bwd(hd, nh, nkv, S, J, pad):
    """vs the oracle's eager TTT attention (llama3_eagle.py:717-785) in fp32 with autograd."""
    from oracle import eagle3_oracle as O
    from specforge_b200 import ops
    dev = _dev()
    torch.manual_seed(J + S)
    B = 2
    A, KV = nh * hd, nkv * hd
    qkv = [(torch.randn(B * S, A + 2 * KV, device=dev) * 0.7).bfloat16() for _ in range(J + 1)]
    am = torch.ones(B, S, dtype=torch.long, device=dev)
    if pad:
        am[-1, S - pad:] = 0
    key_mask = am.to(torch.uint8) if pad else None
    out, lse = ops.ttt_attention_fwd(qkv, B, S, nh, nkv, hd, key_mask=key_mask)
    # reference (fp32, autograd)
    leaves = [t.float().requires_grad_(True) for t in qkv]
    cache_k, cache_v = [], []
    mask = O.decoder_attention_mask(am.cpu(), S, torch.float32).to(dev)
    ref = None
    for i in range(J + 1):
        q = leaves[i][:, :A].view(B, S, nh, hd).transpose(1, 2)
        k = leaves[i][:, A:A + KV].view(B, S, nkv, hd).transpose(1, 2)
        v = leaves[i][:, A + KV:].view(B, S, nkv, hd).transpose(1, 2)
        ref = O.ttt_attention(q, k, v, cache_k, cache_v, mask, nh // nkv)
    ref = ref.reshape(B * S, A)
    valid = am.bool().view(-1)            # padded QUERY rows see a fully-masked block 0 in the reference (-inf quirk)
    err = (out.float() - ref)[valid].abs().max().item()
    assert err <= 2e-2 * ref[valid].abs().max().item() + 1e-3, err
    dout = (torch.randn(B * S, A, device=dev) * 0.5).bfloat16()
    dout[~valid] = 0
    ref.backward(dout.float())
    dq, dk, dv = ops.ttt_attention_bwd(qkv, out, dout, lse, B, S, nh, nkv, hd, key_mask=key_mask)
    gq = leaves[J].grad[:, :A]
    assert torch.nn.functional.cosine_similarity(dq.float().flatten(), gq.flatten(), dim=0) > 0.999
    assert (dq.float() - gq).abs().max().item() <= 3e-2 * gq.abs().max().item() + 1e-3
    for i in range(J + 1):
        # grads w.r.t. block i's K/V from THIS step only (the query grad of block i<J is zero here)
        gk = leaves[i].grad[:, A:A + KV]
        gv = leaves[i].grad[:, A + KV:]
        assert (dk[i] - gk).abs().max().item() <= 3e-2 * gk.abs().max().item() + 1e-3, i
        assert (dv[i] - gv).abs().max().item() <= 3e-2 * gv.abs().max().item() + 1e-3, i
