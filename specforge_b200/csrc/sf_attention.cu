// sf_attention.cu — EAGLE3 "TTT" attention, forward and backward.
//
// Semantics (reference: specforge/modeling/draft/llama3_eagle.py:717-785, eager cache branch; the
// decomposition follows _FlashCachedMergeFunc, :1024-1151): at TTT step j, query row t attends to
//   * block 0: keys K_0[0..t] (causal, optional key-padding mask), values V_0
//   * one diagonal key per later block i = 1..j: K_i[t] / V_i[t]   (never masked)
// with ONE softmax over the concatenation.  GQA: q head h reads kv head h / (nh / nkv).
//
// Implementation: block 0 runs on the tcgen05 kernels of sf_attention_tc.cu / sf_attention_tc_bwd.cu (flash-style, running
// (max, sum) state seeded with the diagonal scores, diagonal P*V terms added in the epilogue; backward with the global
// log-sum-exp, dK/dV with the KV tile stationary and dQ with the Q tile stationary — no atomics, deterministic).  This file
// holds the rank-1 diagonal terms as vectorised row kernels (scores, delta = rowsum(dO*O), backward of blocks 1..j), the
// key-padding prefix scan, argument validation and the dispatch.
#include "sf_host.h"
#include "sf_ptx.cuh"
#include <cstdlib>

namespace sf {

struct AttnParams {
    const __nv_bfloat16* q; int64_t ldq;        // [B*S, nh*D] view (row stride ldq)
    const __nv_bfloat16* k0; const __nv_bfloat16* v0; int64_t ldkv;  // block-0 K/V: [B*S, nkv*D] views
    __nv_bfloat16* out; int64_t ldo;            // [B*S, nh*D]
    float* lse;                                 // [B, nh, S]   log2-domain LSE of the scaled scores
    float* sd;                            // [B, nh, S, J] diagonal scores (already * scale * log2e), or null
    const __nv_bfloat16* vdiag[8];              // V_i bases for i = 1..J (same ldkv)
    const __nv_bfloat16* kdiag[8];              // K_i bases for i = 1..J
    const uint8_t* key_mask;                    // [B, S] 1 = attend, or null
    int B, S, nh, nkv, J;
    float scale_log2;                           // (1/sqrt(D)) * log2(e)
    // backward
    const __nv_bfloat16* dout; int64_t lddo;    // [B*S, nh*D]
    float* delta;                         // [B, nh, S]  rowsum(dO * O)
    float* dk0_acc; float* dv0_acc; int64_t ldacc;  // fp32 accumulators for block 0: [B*S, nkv*D]
    float* dkdiag[8]; float* dvdiag[8];         // fp32 accumulators for blocks 1..J
    __nv_bfloat16* dq; int64_t lddq;            // [B*S, nh*D] view
    float* dq_diag;                             // [B*S, nh*D] fp32 (diagonal-term dq), or null
};

// ------------------------------------------------------------------ diagonal scores
// Row kernels of the TTT diagonal terms.  A head's D elements are spread over LPH = D/8 lanes, 16 bytes (8 bf16) each,
// so one warp serves 32/LPH (row, head) items with fully vectorised loads and log2(LPH) shuffle steps per dot product.
__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
        f[2 * i] = __bfloat162float(b.x);
        f[2 * i + 1] = __bfloat162float(b.y);
    }
}
template <int LPH>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LPH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float dot8(const float (&a)[8], const float (&b)[8]) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += a[e] * b[e];
    return s;
}

// sd[b,h,t,i-1] = (q[b,t,h,:] . K_i[b,t,h/g,:]) * scale * log2e
template <int D>
__global__ void __launch_bounds__(256) diag_scores_kernel(AttnParams p) {
    constexpr int LPH = D / 8, HPW = 32 / LPH;
    const int g = p.nh / p.nkv;
    const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, sub = lane / LPH, l = lane % LPH;
    const int64_t total = (int64_t)p.B * p.S * p.nh;
    const int64_t item = warp_global * HPW + sub;
    const bool ok = item < total;
    const int h = ok ? (int)(item % p.nh) : 0;
    const int64_t r = ok ? item / p.nh : 0;  // b*S + t
    const int b = (int)(r / p.S), t = (int)(r % p.S);
    float qf[8];
    unpack8f(*reinterpret_cast<const uint4*>(p.q + r * p.ldq + h * D + l * 8), qf);
    for (int i = 0; i < p.J; ++i) {
        float kf[8];
        unpack8f(*reinterpret_cast<const uint4*>(p.kdiag[i] + r * p.ldkv + (h / g) * D + l * 8), kf);
        const float acc = group_sum<LPH>(dot8(qf, kf));
        if (ok && l == 0) p.sd[(((int64_t)b * p.nh + h) * p.S + t) * p.J + i] = acc * p.scale_log2;
    }
}

// ------------------------------------------------------------------ backward prep: delta = rowsum(dO * O)
template <int D>
__global__ void __launch_bounds__(256) attn_delta_kernel(AttnParams p) {
    constexpr int LPH = D / 8, HPW = 32 / LPH;
    const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, sub = lane / LPH, l = lane % LPH;
    const int64_t total = (int64_t)p.B * p.S * p.nh;
    const int64_t item = warp_global * HPW + sub;
    const bool ok = item < total;
    const int h = ok ? (int)(item % p.nh) : 0;
    const int64_t r = ok ? item / p.nh : 0;
    float of[8], df[8];
    unpack8f(*reinterpret_cast<const uint4*>(p.out + r * p.ldo + h * D + l * 8), of);
    unpack8f(*reinterpret_cast<const uint4*>(p.dout + r * p.lddo + h * D + l * 8), df);
    const float acc = group_sum<LPH>(dot8(of, df));
    if (ok && l == 0) p.delta[((r / p.S) * p.nh + h) * p.S + (r % p.S)] = acc;
}

// ------------------------------------------------------------------ backward, diagonal terms (row kernel)
// One warp per (row, kv head).  The g query heads of the group are processed 2*HPW at a time (HPW of them side by side in
// the warp's lane groups), the diagonal blocks in the outer loop, so dK_i/dV_i of a row are read-modify-written once per
// head chunk and need no atomics.
template <int D>
__global__ void __launch_bounds__(256) attn_bwd_diag_kernel(AttnParams p) {
    constexpr int LPH = D / 8, HPW = 32 / LPH, CH = 2 * HPW;
    const int g = p.nh / p.nkv;
    const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, sub = lane / LPH, l = lane % LPH;
    const int64_t total = (int64_t)p.B * p.S * p.nkv;
    if (warp_global >= total) return;
    const int kvh = (int)(warp_global % p.nkv);
    const int64_t r = warp_global / p.nkv;
    const int b = (int)(r / p.S), t = (int)(r % p.S);
    const float scale = p.scale_log2 * 0.6931471805599453f;  // back to 1/sqrt(D)
    for (int h0 = 0; h0 < g; h0 += CH) {
        float qf[2][8], dof[2][8], dq[2][8], lse[2], dl[2];
        bool okh[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int hg = h0 + hh * HPW + sub;
            okh[hh] = hg < g;
            const int h = kvh * g + (okh[hh] ? hg : 0);
            unpack8f(*reinterpret_cast<const uint4*>(p.q + r * p.ldq + h * D + l * 8), qf[hh]);
            unpack8f(*reinterpret_cast<const uint4*>(p.dout + r * p.lddo + h * D + l * 8), dof[hh]);
#pragma unroll
            for (int e = 0; e < 8; ++e) dq[hh][e] = 0.f;
            lse[hh] = p.lse[((int64_t)b * p.nh + h) * p.S + t];
            dl[hh] = p.delta[((int64_t)b * p.nh + h) * p.S + t];
        }
        // k_{i+1} / v_{i+1} and the fp32 accumulators this iteration updates are requested before iteration i's arithmetic and
        // shuffles, so every warp keeps ~100 B per lane in flight instead of one dependent round trip at a time
        uint4 kn = make_uint4(0, 0, 0, 0), vn = kn;
        if (p.J > 0) {
            kn = __ldg(reinterpret_cast<const uint4*>(p.kdiag[0] + r * p.ldkv + kvh * D + l * 8));
            vn = __ldg(reinterpret_cast<const uint4*>(p.vdiag[0] + r * p.ldkv + kvh * D + l * 8));
        }
        for (int i = 0; i < p.J; ++i) {
            float kf[8], vf[8], dk[8], dv[8];
            unpack8f(kn, kf);
            unpack8f(vn, vf);
            float* acc_dst = (sub == 0 ? p.dkdiag[i] : p.dvdiag[i]) + r * p.ldacc + kvh * D + l * 8;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            if (sub < 2) { a0 = reinterpret_cast<const float4*>(acc_dst)[0]; a1 = reinterpret_cast<const float4*>(acc_dst)[1]; }
            if (i + 1 < p.J) {
                kn = __ldg(reinterpret_cast<const uint4*>(p.kdiag[i + 1] + r * p.ldkv + kvh * D + l * 8));
                vn = __ldg(reinterpret_cast<const uint4*>(p.vdiag[i + 1] + r * p.ldkv + kvh * D + l * 8));
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) { dk[e] = 0.f; dv[e] = 0.f; }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const float sdot = group_sum<LPH>(dot8(qf[hh], kf));
                const float pdot = group_sum<LPH>(dot8(dof[hh], vf));
                const float pr = okh[hh] ? exp2f(sdot * p.scale_log2 - lse[hh]) : 0.f;
                const float ds = pr * (pdot - dl[hh]) * scale;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    dq[hh][e] += ds * kf[e];
                    dk[e] += ds * qf[hh][e];
                    dv[e] += pr * dof[hh][e];
                }
            }
            // sum the lane groups (different heads, same elements); group 0 updates dK, group 1 dV
#pragma unroll
            for (int o = LPH; o < 32; o <<= 1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    dk[e] += __shfl_xor_sync(0xffffffffu, dk[e], o);
                    dv[e] += __shfl_xor_sync(0xffffffffu, dv[e], o);
                }
            }
            if (sub < 2) {
                float* dst = acc_dst;
                const float* src = sub == 0 ? dk : dv;
                a0.x += src[0]; a0.y += src[1]; a0.z += src[2]; a0.w += src[3];
                a1.x += src[4]; a1.y += src[5]; a1.z += src[6]; a1.w += src[7];
                reinterpret_cast<float4*>(dst)[0] = a0;
                reinterpret_cast<float4*>(dst)[1] = a1;
            }
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            if (!okh[hh]) continue;
            const int h = kvh * g + h0 + hh * HPW + sub;
            float4* dst = reinterpret_cast<float4*>(p.dq_diag + r * (int64_t)(p.nh * D) + h * D + l * 8);
            dst[0] = make_float4(dq[hh][0], dq[hh][1], dq[hh][2], dq[hh][3]);
            dst[1] = make_float4(dq[hh][4], dq[hh][5], dq[hh][6], dq[hh][7]);
        }
    }
}

// ------------------------------------------------------------------ host side
int attn_validate(const AttnDesc& a) {
    if (a.head_dim != 64 && a.head_dim != 128) return set_error(-22, "attention: head_dim=%d unsupported (64 or 128)", a.head_dim);
    if (a.nh % a.nkv) return set_error(-22, "attention: nh=%d not a multiple of nkv=%d", a.nh, a.nkv);
    if (a.nh / a.nkv > 8) return set_error(-22, "attention: GQA group %d > 8 unsupported", a.nh / a.nkv);
    if (a.J < 0 || a.J > 8) return set_error(-22, "attention: %d diagonal blocks unsupported (max 8, TTT length <= 9)", a.J);
    if (a.ldq % 8 || a.ldkv % 8 || a.ldo % 8) return set_error(-22, "attention: leading dims must be multiples of 8");
    return 0;
}
static void fill(AttnParams& p, const AttnDesc& a) {
    p.q = (const __nv_bfloat16*)a.q; p.ldq = a.ldq;
    p.k0 = (const __nv_bfloat16*)a.k[0]; p.v0 = (const __nv_bfloat16*)a.v[0]; p.ldkv = a.ldkv;
    p.out = (__nv_bfloat16*)a.out; p.ldo = a.ldo; p.lse = a.lse; p.sd = a.sd_ws;
    for (int i = 0; i < 8; ++i) {
        p.kdiag[i] = (i < a.J) ? (const __nv_bfloat16*)a.k[i + 1] : nullptr;
        p.vdiag[i] = (i < a.J) ? (const __nv_bfloat16*)a.v[i + 1] : nullptr;
        p.dkdiag[i] = (i < a.J) ? a.dk_acc[i + 1] : nullptr;
        p.dvdiag[i] = (i < a.J) ? a.dv_acc[i + 1] : nullptr;
    }
    p.key_mask = a.key_mask; p.B = a.B; p.S = a.S; p.nh = a.nh; p.nkv = a.nkv; p.J = a.J;
    p.scale_log2 = (1.0f / sqrtf((float)a.head_dim)) * 1.4426950408889634f;
    p.dout = (const __nv_bfloat16*)a.dout; p.lddo = a.lddo; p.delta = a.delta_ws;
    p.dk0_acc = a.dk_acc[0]; p.dv0_acc = a.dv_acc[0]; p.ldacc = a.ldacc;
    p.dq = (__nv_bfloat16*)a.dq; p.lddq = a.lddq; p.dq_diag = (a.J > 0) ? a.dq_diag_ws : nullptr;
}
// kvlen[b] = number of leading 1s of key_mask[b, :]; nonprefix[b] = 1 if a 1 follows a 0 (one warp per row)
__global__ void mask_prefix_kernel(const uint8_t* __restrict__ key_mask, int B, int S, int* __restrict__ kvlen,
                                   int* __restrict__ nonprefix) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    int first_zero = S, last_one = -1;
    for (int s = lane; s < S; s += 32) {
        if (key_mask[(int64_t)b * S + s]) last_one = max(last_one, s);
        else first_zero = min(first_zero, s);
    }
    for (int o = 16; o > 0; o >>= 1) {
        first_zero = min(first_zero, __shfl_xor_sync(0xffffffffu, first_zero, o));
        last_one = max(last_one, __shfl_xor_sync(0xffffffffu, last_one, o));
    }
    if (lane == 0) { kvlen[b] = first_zero; nonprefix[b] = last_one > first_zero ? 1 : 0; }
}
int mask_prefix(const uint8_t* key_mask, int B, int S, int* kvlen, int* nonprefix, cudaStream_t st) {
    mask_prefix_kernel<<<(B + 7) / 8, 256, 0, st>>>(key_mask, B, S, kvlen, nonprefix);
    SF_CUDA_CHECK_LAUNCH("mask_prefix");
    return 0;
}

int attn_fwd(const AttnDesc& a, cudaStream_t st) {
    if (int rc = attn_validate(a)) return rc;
    if (a.key_mask && !a.kvlen) return set_error(-22, "attention: key_mask given without kvlen/nonprefix (call mask_prefix)");
    AttnParams p; fill(p, a);
    // diagonal scores by the row kernel, block 0 on tcgen05
    if (p.J > 0) {
        const int ipw = 256 / a.head_dim;   // (row, head) items per warp
        const int64_t warps = ((int64_t)p.B * p.S * p.nh + ipw - 1) / ipw;
        if (a.head_dim == 128) diag_scores_kernel<128><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(p);
        else diag_scores_kernel<64><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(p);
        SF_CUDA_CHECK_LAUNCH("diag_scores");
    }
    return attn_fwd_tc(a, st);
}
template <int D>
static int attn_bwd_prep_t(AttnParams& p, cudaStream_t st) {
    {
        const int64_t warps = ((int64_t)p.B * p.S * p.nh + (256 / D) - 1) / (256 / D);
        attn_delta_kernel<D><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(p);
        SF_CUDA_CHECK_LAUNCH("attn_delta");
    }
    if (p.J > 0) {
        const int64_t warps = (int64_t)p.B * p.S * p.nkv;
        attn_bwd_diag_kernel<D><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(p);
        SF_CUDA_CHECK_LAUNCH("attn_bwd_diag");
    }
    return 0;
}

int attn_bwd(const AttnDesc& a, cudaStream_t st) {
    if (int rc = attn_validate(a)) return rc;
    if (a.key_mask && !a.kvlen) return set_error(-22, "attention: key_mask given without kvlen/nonprefix (call mask_prefix)");
    AttnParams p; fill(p, a);
    // delta + diagonal terms by row kernels, block 0 (dK/dV and dQ) on tcgen05
    if (int rc = (a.head_dim == 128 ? attn_bwd_prep_t<128>(p, st) : attn_bwd_prep_t<64>(p, st))) return rc;
    return attn_bwd_tc(a, st);
}

}  // namespace sf
