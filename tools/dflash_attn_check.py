"""Op-level check of the DFlash block attention (GPU): CUDA-core kernels vs the tcgen05 kernels vs a dense fp32 PyTorch
evaluation of the same masked attention (the oracle's mask).  Prints per-tensor max error / cosine for forward and backward.
    python tools/dflash_attn_check.py [--d 128] [--g 4] [--bs 16] [--S 300] [--N 9] [--impl 0 1]
First tool to run when bringing up csrc/sf_dflash_attn_tc*.cu (written without GPU time)."""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dflash_oracle as D                        # noqa: E402   (checker only)
from specforge_b200._lib import check, lib                   # noqa: E402


def dense_reference(q, kn, vn, kc, vc, anchors, keep, B, S, N, bs, nh, nkv, d, dout, window=None):
    """fp32 autograd reference on the GPU: scores over [context ; noise] keys with the oracle's boolean mask."""
    g = nh // nkv
    leaves = [t.float().detach().clone().requires_grad_(True) for t in (q, kn, vn, kc, vc)]
    qf, knf, vnf, kcf, vcf = leaves
    Q = N * bs
    mask = D.dflash_mask(anchors.cpu().long(), keep.cpu().bool(), S, bs, window).to(q.device)                     # [B, Q, S+Q]
    qh = qf.view(B, Q, nh, d).transpose(1, 2)
    k = torch.cat([kcf.view(B, S, nkv, d), knf.view(B, Q, nkv, d)], dim=1).transpose(1, 2).repeat_interleave(g, dim=1)
    v = torch.cat([vcf.view(B, S, nkv, d), vnf.view(B, Q, nkv, d)], dim=1).transpose(1, 2).repeat_interleave(g, dim=1)
    s = (qh @ k.transpose(-1, -2)) * d ** -0.5
    row_ok = mask.any(dim=-1).unsqueeze(1).unsqueeze(-1)                                                   # dropped blocks see nothing
    s = torch.where(row_ok, s.masked_fill(~mask.unsqueeze(1), float("-inf")), torch.zeros_like(s))         # (no NaNs in autograd)
    p = torch.softmax(s, dim=-1) * row_ok
    o = (p @ v).transpose(1, 2).reshape(B * Q, nh * d)
    o.backward(dout.float())
    lse = torch.logsumexp(s.masked_fill(~mask.unsqueeze(1), float("-inf")), dim=-1).transpose(1, 2).reshape(B * Q, nh)
    return o.detach(), lse.detach(), [t.grad for t in leaves]


def run(L, impl, q, kn, vn, kc, vc, anchors, keep, dout, dims):
    B, S, N, bs, nh, nkv, d = dims
    Mq, Mc = B * N * bs, B * S
    dev = q.device
    out = torch.empty(Mq, nh * d, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(Mq, nh, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    check(L.sf_dflash_attention_fwd(P(q), P(kn), P(vn), P(kc), P(vc), P(out), P(lse), P(anchors), P(keep), B, S, N, bs, nh, nkv, d, impl,
                                    ctypes.c_void_p(st)), "sf_dflash_attention_fwd")
    dq = torch.empty_like(q); dkn = torch.empty_like(kn); dvn = torch.empty_like(vn); dkc = torch.empty_like(kc); dvc = torch.empty_like(vc)
    delta = torch.empty(Mq, nh, dtype=torch.float32, device=dev)
    check(L.sf_dflash_attention_bwd(P(q), P(kn), P(vn), P(kc), P(vc), P(out), P(lse), P(dout), P(anchors), P(keep), P(dq), P(dkn), P(dvn),
                                    P(dkc), P(dvc), P(delta), B, S, N, bs, nh, nkv, d, impl, ctypes.c_void_p(st)), "sf_dflash_attention_bwd")
    torch.cuda.synchronize()
    return out, lse, [dq, dkn, dvn, dkc, dvc]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, default=128); ap.add_argument("--g", type=int, default=4); ap.add_argument("--nkv", type=int, default=2)
    ap.add_argument("--bs", type=int, default=16); ap.add_argument("--S", type=int, default=300); ap.add_argument("--N", type=int, default=9)
    ap.add_argument("--B", type=int, default=2); ap.add_argument("--impl", type=int, nargs="+", default=[0, 1])
    ap.add_argument("--window", type=int, default=0, help="sliding-window layer: keys [a + o - (W - 1), a) + own slots <= o (0 = full)")
    a = ap.parse_args()
    L = lib()
    L.sf_debug_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    check(L.sf_debug_option(b"dflash_attn_window", a.window), "sf_debug_option")
    for f in (L.sf_dflash_attention_fwd, L.sf_dflash_attention_bwd):
        f.restype = ctypes.c_int
    dev = torch.device("cuda", 0)
    B, S, N, bs, nkv, d = a.B, a.S, a.N, a.bs, a.nkv, a.d
    nh = nkv * a.g
    gen = torch.Generator().manual_seed(0)
    lm = (torch.rand(B, S, generator=gen) > 0.1).float()
    lm[-1, max(4, N // 2):] = 0                              # the last sequence has fewer candidates than N -> dropped blocks
    anchors, keep = D.sample_anchor_positions(lm, N, generator=gen)
    N = anchors.shape[1]
    Mq, Mc = B * N * bs, B * S
    mk = lambda rows, cols: (torch.randn(rows, cols, generator=gen) * 0.7).bfloat16().to(dev)
    q, kn, vn, kc, vc, dout = mk(Mq, nh * d), mk(Mq, nkv * d), mk(Mq, nkv * d), mk(Mc, nkv * d), mk(Mc, nkv * d), mk(Mq, nh * d)
    anc, kp = anchors.to(dev, torch.int32).contiguous(), keep.to(dev).to(torch.uint8).contiguous()
    ref_o, ref_lse, ref_g = dense_reference(q, kn, vn, kc, vc, anc, kp, B, S, N, bs, nh, nkv, d, dout, a.window or None)
    kept_rows = keep.repeat_interleave(bs, dim=1).reshape(-1).to(dev)
    names = ["dq", "dkn", "dvn", "dkc", "dvc"]
    for impl in a.impl:
        try:
            out, lse, grads = run(L, impl, q, kn, vn, kc, vc, anc, kp, dout, (B, S, N, bs, nh, nkv, d))
        except Exception as exc:                             # e.g. shape not covered by the tcgen05 path
            print(f"impl {impl}: {exc}")
            continue
        cos = lambda x, y: torch.nn.functional.cosine_similarity(x.float().flatten(), y.float().flatten(), dim=0).item()
        print(f"window {a.window} impl {impl}: out max|err| {float((out.float() - ref_o).abs().max()):.3e} cos {cos(out, ref_o):.6f}   "
              f"lse max|err| (kept rows) {float((lse - ref_lse)[kept_rows].abs().max()):.3e}")
        for n, g_, r_ in zip(names, grads, ref_g):
            print(f"          {n:4s} max|err| {float((g_.float() - r_).abs().max()):.3e} (ref max {float(r_.abs().max()):.3e}) cos {cos(g_, r_):.6f}")


if __name__ == "__main__":
    main()
