// sf_attention.cu — EAGLE3 "TTT" attention, forward and backward.
//
// Semantics (reference: specforge/modeling/draft/llama3_eagle.py:717-785, eager cache branch; the
// decomposition follows _FlashCachedMergeFunc, :1024-1151): at TTT step j, query row t attends to
//   * block 0: keys K_0[0..t] (causal, optional key-padding mask), values V_0
//   * one diagonal key per later block i = 1..j: K_i[t] / V_i[t]   (never masked)
// with ONE softmax over the concatenation.  GQA: q head h reads kv head h / (nh / nkv).
//
// Implementation: a flash-style kernel over block 0 whose running (max, sum) state is seeded with the
// diagonal scores (computed by a small row kernel) and whose epilogue adds the diagonal P*V terms;
// backward = flash backward over block 0 using the global log-sum-exp (two kernels: dK/dV with the KV tile
// stationary, dQ with the Q tile stationary — no atomics, deterministic) + a row kernel for the diagonal
// terms.  Tensor-core path here is mma.sync m16n8k16 (bf16 -> fp32); attention is ~2.4 % of step FLOPs.
// TODO(round 2): move QK^T / PV to tcgen05 with S/P in TMEM.
#include "sf_host.h"
#include "sf_ptx.cuh"
#include <cstdlib>

namespace sf {

// ------------------------------------------------------------------ small PTX helpers
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Row-major [rows][D] bf16 tile in smem, 16-byte chunks XOR-swizzled by (row & 7).
template <int D>
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
    return (uint32_t)(row * (D * 2) + ((chunk ^ (row & 7)) << 4));
}
// Cooperative async load of `rows` rows x D cols (bf16) from global (row stride ld) into a swizzled tile.
// Rows >= rows_valid are zero-filled.
template <int D, int kThreads>
__device__ __forceinline__ void load_tile(uint32_t smem_tile, const __nv_bfloat16* g, int64_t ld, int rows, int rows_valid) {
    constexpr int CH = D / 8;
    for (int i = threadIdx.x; i < rows * CH; i += kThreads) {
        const int r = i / CH, c = i % CH;
        const bool ok = r < rows_valid;
        cp_async16(smem_tile + swz<D>(r, c), g + (ok ? (int64_t)r * ld + c * 8 : 0), ok);
    }
}

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

// ------------------------------------------------------------------ forward
template <int D>
__global__ void __launch_bounds__(256, 1) attn_fwd_kernel(AttnParams p) {
    constexpr int BR = 128, BC = 64, CH = D / 8;
    constexpr int Q_BYTES = BR * D * 2, KV_BYTES = BC * D * 2;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sK = sQ + Q_BYTES;           // 2 buffers
    const uint32_t sV = sK + 2 * KV_BYTES;      // 2 buffers

    const int qb = gridDim.x - 1 - blockIdx.x;  // heavy (late) blocks first
    const int h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / (p.nh / p.nkv);
    const int q0 = qb * BR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;

    const __nv_bfloat16* qg = p.q + ((int64_t)b * p.S + q0) * p.ldq + h * D;
    const __nv_bfloat16* kg = p.k0 + (int64_t)b * p.S * p.ldkv + kvh * D;
    const __nv_bfloat16* vg = p.v0 + (int64_t)b * p.S * p.ldkv + kvh * D;
    const int q_valid = min(BR, p.S - q0);
    int n_kv = (min(q0 + BR, p.S) + BC - 1) / BC;

    load_tile<D, 256>(sQ, qg, p.ldq, BR, q_valid);
    load_tile<D, 256>(sK, kg, p.ldkv, BC, min(BC, p.S));
    load_tile<D, 256>(sV, vg, p.ldkv, BC, min(BC, p.S));
    cp_async_commit();

    const int row_a = q0 + warp * 16 + gid;  // this thread's two query rows
    const int row_b = row_a + 8;
    float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;
    // seed the running softmax state with the diagonal scores (llama3_eagle.py:756-764)
    if (p.J > 0) {
        const float* sa = p.sd + (((int64_t)b * p.nh + h) * p.S + min(row_a, p.S - 1)) * p.J;
        const float* sb = p.sd + (((int64_t)b * p.nh + h) * p.S + min(row_b, p.S - 1)) * p.J;
        for (int i = 0; i < p.J; ++i) { m_a = fmaxf(m_a, sa[i]); m_b = fmaxf(m_b, sb[i]); }
        // l is kept as per-lane partial sums (quad-reduced at the end): seed it in one lane of the quad only
        if (tig == 0)
            for (int i = 0; i < p.J; ++i) { l_a += exp2f(sa[i] - m_a); l_b += exp2f(sb[i] - m_b); }
    }
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }

    for (int kb = 0; kb < n_kv; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < n_kv) {
            const int k1 = (kb + 1) * BC;
            load_tile<D, 256>(sK + (buf ^ 1) * KV_BYTES, kg + (int64_t)k1 * p.ldkv, p.ldkv, BC, min(BC, p.S - k1));
            load_tile<D, 256>(sV + (buf ^ 1) * KV_BYTES, vg + (int64_t)k1 * p.ldkv, p.ldkv, BC, min(BC, p.S - k1));
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint32_t sKb = sK + buf * KV_BYTES, sVb = sV + buf * KV_BYTES;

        // ---- S = Q K^T  (16 rows x 64 keys per warp)
        float s[BC / 8][4];
#pragma unroll
        for (int i = 0; i < BC / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t a0, a1, a2, a3;
            {
                const int mi = lane >> 3;
                const int row = warp * 16 + (mi & 1) * 8 + (lane & 7);
                ldsm_x4(sQ + swz<D>(row, ks * 2 + (mi >> 1)), a0, a1, a2, a3);
            }
#pragma unroll
            for (int nt = 0; nt < BC / 8; nt += 2) {
                uint32_t b0, b1, b2, b3;
                const int mi = lane >> 3;
                const int krow = nt * 8 + (mi >> 1) * 8 + (lane & 7);
                ldsm_x4(sKb + swz<D>(krow, ks * 2 + (mi & 1)), b0, b1, b2, b3);
                mma16816(s[nt], a0, a1, a2, a3, b0, b1);
                mma16816(s[nt + 1], a0, a1, a2, a3, b2, b3);
            }
        }
        // ---- scale, mask
        const int k_base = kb * BC;
        const bool need_mask = (k_base + BC - 1 > q0 + warp * 16) || (k_base + BC > p.S) || p.key_mask != nullptr;
#pragma unroll
        for (int nt = 0; nt < BC / 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v = s[nt][e] * p.scale_log2;
                if (need_mask) {
                    const int key = k_base + nt * 8 + tig * 2 + (e & 1);
                    const int row = (e < 2) ? row_a : row_b;
                    bool ok = key <= row && key < p.S;
                    if (ok && p.key_mask) ok = p.key_mask[(int64_t)b * p.S + key] != 0;
                    if (!ok) v = -INFINITY;
                }
                s[nt][e] = v;
            }
        }
        // ---- online softmax
        float mx_a = m_a, mx_b = m_b;
#pragma unroll
        for (int nt = 0; nt < BC / 8; ++nt) {
            mx_a = fmaxf(mx_a, fmaxf(s[nt][0], s[nt][1]));
            mx_b = fmaxf(mx_b, fmaxf(s[nt][2], s[nt][3]));
        }
        mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
        mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
        mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
        mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
        const float base_a = (mx_a == -INFINITY) ? 0.f : mx_a;
        const float base_b = (mx_b == -INFINITY) ? 0.f : mx_b;
        const float corr_a = exp2f(m_a - base_a), corr_b = exp2f(m_b - base_b);  // m=-inf -> 0
        m_a = mx_a; m_b = mx_b;
        float rs_a = 0.f, rs_b = 0.f;
#pragma unroll
        for (int nt = 0; nt < BC / 8; ++nt) {
            s[nt][0] = exp2f(s[nt][0] - base_a); s[nt][1] = exp2f(s[nt][1] - base_a);
            s[nt][2] = exp2f(s[nt][2] - base_b); s[nt][3] = exp2f(s[nt][3] - base_b);
            rs_a += s[nt][0] + s[nt][1];
            rs_b += s[nt][2] + s[nt][3];
        }
        l_a = l_a * corr_a + rs_a;  // per-thread partial sums; quad-reduced at the end
        l_b = l_b * corr_b + rs_b;
#pragma unroll
        for (int i = 0; i < D / 8; ++i) { o[i][0] *= corr_a; o[i][1] *= corr_a; o[i][2] *= corr_b; o[i][3] *= corr_b; }
        // ---- O += P V
#pragma unroll
        for (int kk = 0; kk < BC / 16; ++kk) {
            const uint32_t a0 = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
            const uint32_t a1 = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
            const uint32_t a2 = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            const uint32_t a3 = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int nt = 0; nt < D / 8; nt += 2) {
                uint32_t b0, b1, b2, b3;
                const int mi = lane >> 3;
                const int vrow = kk * 16 + (mi & 1) * 8 + (lane & 7);
                ldsm_x4_t(sVb + swz<D>(vrow, nt + (mi >> 1)), b0, b1, b2, b3);
                mma16816(o[nt], a0, a1, a2, a3, b0, b1);
                mma16816(o[nt + 1], a0, a1, a2, a3, b2, b3);
            }
        }
        __syncthreads();
    }
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 1); l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
    l_b += __shfl_xor_sync(0xffffffffu, l_b, 1); l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);

    // ---- epilogue: diagonal P*V terms, normalise, write
    const float inv_a = (l_a > 0.f) ? 1.f / l_a : 0.f;
    const float inv_b = (l_b > 0.f) ? 1.f / l_b : 0.f;
    for (int i = 0; i < p.J; ++i) {
        const int ra = min(row_a, p.S - 1), rb = min(row_b, p.S - 1);
        const float wa = exp2f(p.sd[(((int64_t)b * p.nh + h) * p.S + ra) * p.J + i] - m_a);
        const float wb = exp2f(p.sd[(((int64_t)b * p.nh + h) * p.S + rb) * p.J + i] - m_b);
        const __nv_bfloat16* va = p.vdiag[i] + ((int64_t)b * p.S + ra) * p.ldkv + kvh * D + tig * 2;
        const __nv_bfloat16* vb = p.vdiag[i] + ((int64_t)b * p.S + rb) * p.ldkv + kvh * D + tig * 2;
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt) {
            const __nv_bfloat162 xa = *reinterpret_cast<const __nv_bfloat162*>(va + nt * 8);
            const __nv_bfloat162 xb = *reinterpret_cast<const __nv_bfloat162*>(vb + nt * 8);
            o[nt][0] += wa * __bfloat162float(xa.x); o[nt][1] += wa * __bfloat162float(xa.y);
            o[nt][2] += wb * __bfloat162float(xb.x); o[nt][3] += wb * __bfloat162float(xb.y);
        }
    }
    // stage through smem (Q tile region) for coalesced 16-byte stores
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
        const int ra = warp * 16 + gid, rb = ra + 8;
        *reinterpret_cast<uint32_t*>(smem + swz<D>(ra, nt) + tig * 4) = pack_bf16x2(o[nt][0] * inv_a, o[nt][1] * inv_a);
        *reinterpret_cast<uint32_t*>(smem + swz<D>(rb, nt) + tig * 4) = pack_bf16x2(o[nt][2] * inv_b, o[nt][3] * inv_b);
    }
    if (tig == 0) {
        if (row_a < p.S) p.lse[((int64_t)b * p.nh + h) * p.S + row_a] = m_a + log2f(l_a);
        if (row_b < p.S) p.lse[((int64_t)b * p.nh + h) * p.S + row_b] = m_b + log2f(l_b);
    }
    __syncthreads();
    __nv_bfloat16* og = p.out + ((int64_t)b * p.S + q0) * p.ldo + h * D;
    for (int i = threadIdx.x; i < BR * CH; i += 256) {
        const int r = i / CH, c = i % CH;
        if (r < q_valid) *reinterpret_cast<uint4*>(og + (int64_t)r * p.ldo + c * 8) = *reinterpret_cast<const uint4*>(smem + swz<D>(r, c));
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
        for (int i = 0; i < p.J; ++i) {
            float kf[8], vf[8], dk[8], dv[8];
            unpack8f(*reinterpret_cast<const uint4*>(p.kdiag[i] + r * p.ldkv + kvh * D + l * 8), kf);
            unpack8f(*reinterpret_cast<const uint4*>(p.vdiag[i] + r * p.ldkv + kvh * D + l * 8), vf);
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
                float* dst = (sub == 0 ? p.dkdiag[i] : p.dvdiag[i]) + r * p.ldacc + kvh * D + l * 8;
                float4 a0 = reinterpret_cast<float4*>(dst)[0], a1 = reinterpret_cast<float4*>(dst)[1];
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

template <int D>
__global__ void __launch_bounds__(256, 1) attn_bwd_dkv_kernel(AttnParams p) {
    constexpr int BC = 128, BR = 32, CH = D / 8;
    constexpr int KV_BYTES = BC * D * 2, Q_BYTES = BR * D * 2;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sK = smem_u32(smem);
    const uint32_t sV = sK + KV_BYTES;
    const uint32_t sQ = sV + KV_BYTES;         // 2 buffers
    const uint32_t sdO = sQ + 2 * Q_BYTES;     // 2 buffers
    float* sL = reinterpret_cast<float*>(smem + 2 * KV_BYTES + 4 * Q_BYTES);   // [2][BR] lse
    float* sDl = sL + 2 * BR;                                                    // [2][BR] delta

    const int kb = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
    const int g = p.nh / p.nkv;
    const int k0 = kb * BC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const int k_valid = min(BC, p.S - k0);
    const float scale = p.scale_log2 * 0.6931471805599453f;

    load_tile<D, 256>(sK, p.k0 + ((int64_t)b * p.S + k0) * p.ldkv + kvh * D, p.ldkv, BC, k_valid);
    load_tile<D, 256>(sV, p.v0 + ((int64_t)b * p.S + k0) * p.ldkv + kvh * D, p.ldkv, BC, k_valid);
    cp_async_commit();

    const int first_qt = k0 / BR;                       // first query tile that can see this kv block
    const int n_qt = (p.S + BR - 1) / BR;
    const int tiles_per_head = n_qt - first_qt;
    const int total_iters = tiles_per_head * g;

    auto issue = [&](int it, int buf) {
        const int hh = it / tiles_per_head, qt = first_qt + it % tiles_per_head;
        const int h = kvh * g + hh;
        const int q0 = qt * BR;
        const int qv = min(BR, p.S - q0);
        load_tile<D, 256>(sQ + buf * Q_BYTES, p.q + ((int64_t)b * p.S + q0) * p.ldq + h * D, p.ldq, BR, qv);
        load_tile<D, 256>(sdO + buf * Q_BYTES, p.dout + ((int64_t)b * p.S + q0) * p.lddo + h * D, p.lddo, BR, qv);
        if (threadIdx.x < BR) {
            const int row = q0 + threadIdx.x;
            const bool ok = row < p.S;
            sL[buf * BR + threadIdx.x] = ok ? p.lse[((int64_t)b * p.nh + h) * p.S + row] : INFINITY;
            sDl[buf * BR + threadIdx.x] = ok ? p.delta[((int64_t)b * p.nh + h) * p.S + row] : 0.f;
        }
        cp_async_commit();
    };

    float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }

    const int key_a = k0 + warp * 16 + gid, key_b = key_a + 8;  // this thread's two key rows
    bool mask_a = key_a < p.S, mask_b = key_b < p.S;
    if (p.key_mask) {
        if (mask_a) mask_a = p.key_mask[(int64_t)b * p.S + key_a] != 0;
        if (mask_b) mask_b = p.key_mask[(int64_t)b * p.S + key_b] != 0;
    }

    if (total_iters > 0) issue(0, 0);
    for (int it = 0; it < total_iters; ++it) {
        const int buf = it & 1;
        if (it + 1 < total_iters) { issue(it + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const int qt = first_qt + it % tiles_per_head;
        const int q0 = qt * BR;
        const uint32_t sQb = sQ + buf * Q_BYTES, sdOb = sdO + buf * Q_BYTES;

        // ---- S^T = K Q^T and dP^T = V dO^T   (16 keys x 32 queries per warp)
        float st[BR / 8][4], dpt[BR / 8][4];
#pragma unroll
        for (int i = 0; i < BR / 8; ++i) { st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f; dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t ka0, ka1, ka2, ka3, va0, va1, va2, va3;
            {
                const int mi = lane >> 3;
                const int row = warp * 16 + (mi & 1) * 8 + (lane & 7);
                ldsm_x4(sK + swz<D>(row, ks * 2 + (mi >> 1)), ka0, ka1, ka2, ka3);
                ldsm_x4(sV + swz<D>(row, ks * 2 + (mi >> 1)), va0, va1, va2, va3);
            }
#pragma unroll
            for (int nt = 0; nt < BR / 8; nt += 2) {
                uint32_t b0, b1, b2, b3;
                const int mi = lane >> 3;
                const int qrow = nt * 8 + (mi >> 1) * 8 + (lane & 7);
                ldsm_x4(sQb + swz<D>(qrow, ks * 2 + (mi & 1)), b0, b1, b2, b3);
                mma16816(st[nt], ka0, ka1, ka2, ka3, b0, b1);
                mma16816(st[nt + 1], ka0, ka1, ka2, ka3, b2, b3);
                ldsm_x4(sdOb + swz<D>(qrow, ks * 2 + (mi & 1)), b0, b1, b2, b3);
                mma16816(dpt[nt], va0, va1, va2, va3, b0, b1);
                mma16816(dpt[nt + 1], va0, va1, va2, va3, b2, b3);
            }
        }
        // ---- P^T = exp2(S^T*c - L[q]) (causal/key mask), dS^T = P^T (dP^T - delta[q]) * scale
        uint32_t pa[BR / 16][4], dsa[BR / 16][4];
#pragma unroll
        for (int nt = 0; nt < BR / 8; ++nt) {
            float pv[4], dsv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ql = nt * 8 + tig * 2 + (e & 1);
                const int qrow = q0 + ql;
                const int key = (e < 2) ? key_a : key_b;
                const bool ok = ((e < 2) ? mask_a : mask_b) && key <= qrow;
                const float pr = ok ? exp2f(st[nt][e] * p.scale_log2 - sL[buf * BR + ql]) : 0.f;
                pv[e] = pr;
                dsv[e] = pr * (dpt[nt][e] - sDl[buf * BR + ql]) * scale;
            }
            // accumulator (16 keys x 8 queries) -> A-operand fragment (16 keys x 16 queries) halves
            pa[nt / 2][(nt & 1) * 2 + 0] = pack_bf16x2(pv[0], pv[1]);
            pa[nt / 2][(nt & 1) * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
            dsa[nt / 2][(nt & 1) * 2 + 0] = pack_bf16x2(dsv[0], dsv[1]);
            dsa[nt / 2][(nt & 1) * 2 + 1] = pack_bf16x2(dsv[2], dsv[3]);
        }
        // ---- dV += P^T dO ; dK += dS^T Q     (k = queries, n = head dim)
#pragma unroll
        for (int kk = 0; kk < BR / 16; ++kk) {
#pragma unroll
            for (int nt = 0; nt < D / 8; nt += 2) {
                uint32_t b0, b1, b2, b3;
                const int mi = lane >> 3;
                const int qrow = kk * 16 + (mi & 1) * 8 + (lane & 7);
                ldsm_x4_t(sdOb + swz<D>(qrow, nt + (mi >> 1)), b0, b1, b2, b3);
                mma16816(dv[nt], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b0, b1);
                mma16816(dv[nt + 1], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b2, b3);
                ldsm_x4_t(sQb + swz<D>(qrow, nt + (mi >> 1)), b0, b1, b2, b3);
                mma16816(dk[nt], dsa[kk][0], dsa[kk][1], dsa[kk][2], dsa[kk][3], b0, b1);
                mma16816(dk[nt + 1], dsa[kk][0], dsa[kk][1], dsa[kk][2], dsa[kk][3], b2, b3);
            }
        }
        __syncthreads();
    }
    // ---- accumulate into the fp32 block-0 gradient buffers (this CTA owns its tile: plain RMW)
    const int ra = k0 + warp * 16 + gid, rb = ra + 8;
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
        const int col = kvh * D + nt * 8 + tig * 2;
        if (ra < p.S) {
            float2* pk = reinterpret_cast<float2*>(p.dk0_acc + ((int64_t)b * p.S + ra) * p.ldacc + col);
            float2* pv = reinterpret_cast<float2*>(p.dv0_acc + ((int64_t)b * p.S + ra) * p.ldacc + col);
            float2 a = *pk, c = *pv;
            a.x += dk[nt][0]; a.y += dk[nt][1]; c.x += dv[nt][0]; c.y += dv[nt][1];
            *pk = a; *pv = c;
        }
        if (rb < p.S) {
            float2* pk = reinterpret_cast<float2*>(p.dk0_acc + ((int64_t)b * p.S + rb) * p.ldacc + col);
            float2* pv = reinterpret_cast<float2*>(p.dv0_acc + ((int64_t)b * p.S + rb) * p.ldacc + col);
            float2 a = *pk, c = *pv;
            a.x += dk[nt][2]; a.y += dk[nt][3]; c.x += dv[nt][2]; c.y += dv[nt][3];
            *pk = a; *pv = c;
        }
    }
}

// ------------------------------------------------------------------ backward, block 0: dQ (Q tile stationary)
template <int D>
__global__ void __launch_bounds__(256, 1) attn_bwd_dq_kernel(AttnParams p) {
    constexpr int BR = 128, BC = 64, CH = D / 8;
    constexpr int Q_BYTES = BR * D * 2, KV_BYTES = BC * D * 2;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sdO = sQ + Q_BYTES;
    const uint32_t sK = sdO + Q_BYTES;        // 2 buffers
    const uint32_t sV = sK + 2 * KV_BYTES;    // 2 buffers

    const int qb = gridDim.x - 1 - blockIdx.x;
    const int h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / (p.nh / p.nkv);
    const int q0 = qb * BR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const int q_valid = min(BR, p.S - q0);
    const int n_kv = (min(q0 + BR, p.S) + BC - 1) / BC;
    const float scale = p.scale_log2 * 0.6931471805599453f;
    const __nv_bfloat16* kg = p.k0 + (int64_t)b * p.S * p.ldkv + kvh * D;
    const __nv_bfloat16* vg = p.v0 + (int64_t)b * p.S * p.ldkv + kvh * D;

    load_tile<D, 256>(sQ, p.q + ((int64_t)b * p.S + q0) * p.ldq + h * D, p.ldq, BR, q_valid);
    load_tile<D, 256>(sdO, p.dout + ((int64_t)b * p.S + q0) * p.lddo + h * D, p.lddo, BR, q_valid);
    load_tile<D, 256>(sK, kg, p.ldkv, BC, min(BC, p.S));
    load_tile<D, 256>(sV, vg, p.ldkv, BC, min(BC, p.S));
    cp_async_commit();

    const int row_a = q0 + warp * 16 + gid, row_b = row_a + 8;
    const float L_a = (row_a < p.S) ? p.lse[((int64_t)b * p.nh + h) * p.S + row_a] : INFINITY;
    const float L_b = (row_b < p.S) ? p.lse[((int64_t)b * p.nh + h) * p.S + row_b] : INFINITY;
    const float dl_a = (row_a < p.S) ? p.delta[((int64_t)b * p.nh + h) * p.S + row_a] : 0.f;
    const float dl_b = (row_b < p.S) ? p.delta[((int64_t)b * p.nh + h) * p.S + row_b] : 0.f;

    float dq[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }

    for (int kb = 0; kb < n_kv; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < n_kv) {
            const int k1 = (kb + 1) * BC;
            load_tile<D, 256>(sK + (buf ^ 1) * KV_BYTES, kg + (int64_t)k1 * p.ldkv, p.ldkv, BC, min(BC, p.S - k1));
            load_tile<D, 256>(sV + (buf ^ 1) * KV_BYTES, vg + (int64_t)k1 * p.ldkv, p.ldkv, BC, min(BC, p.S - k1));
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint32_t sKb = sK + buf * KV_BYTES, sVb = sV + buf * KV_BYTES;
        float s[BC / 8][4], dp[BC / 8][4];
#pragma unroll
        for (int i = 0; i < BC / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t qa0, qa1, qa2, qa3, da0, da1, da2, da3;
            {
                const int mi = lane >> 3;
                const int row = warp * 16 + (mi & 1) * 8 + (lane & 7);
                ldsm_x4(sQ + swz<D>(row, ks * 2 + (mi >> 1)), qa0, qa1, qa2, qa3);
                ldsm_x4(sdO + swz<D>(row, ks * 2 + (mi >> 1)), da0, da1, da2, da3);
            }
#pragma unroll
            for (int nt = 0; nt < BC / 8; nt += 2) {
                uint32_t b0, b1, b2, b3;
                const int mi = lane >> 3;
                const int krow = nt * 8 + (mi >> 1) * 8 + (lane & 7);
                ldsm_x4(sKb + swz<D>(krow, ks * 2 + (mi & 1)), b0, b1, b2, b3);
                mma16816(s[nt], qa0, qa1, qa2, qa3, b0, b1);
                mma16816(s[nt + 1], qa0, qa1, qa2, qa3, b2, b3);
                ldsm_x4(sVb + swz<D>(krow, ks * 2 + (mi & 1)), b0, b1, b2, b3);
                mma16816(dp[nt], da0, da1, da2, da3, b0, b1);
                mma16816(dp[nt + 1], da0, da1, da2, da3, b2, b3);
            }
        }
        const int k_base = kb * BC;
        uint32_t dsa[BC / 16][4];
#pragma unroll
        for (int nt = 0; nt < BC / 8; ++nt) {
            float dsv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = k_base + nt * 8 + tig * 2 + (e & 1);
                const int row = (e < 2) ? row_a : row_b;
                bool ok = key <= row && key < p.S;
                if (ok && p.key_mask) ok = p.key_mask[(int64_t)b * p.S + key] != 0;
                const float pr = ok ? exp2f(s[nt][e] * p.scale_log2 - ((e < 2) ? L_a : L_b)) : 0.f;
                dsv[e] = pr * (dp[nt][e] - ((e < 2) ? dl_a : dl_b)) * scale;
            }
            dsa[nt / 2][(nt & 1) * 2 + 0] = pack_bf16x2(dsv[0], dsv[1]);
            dsa[nt / 2][(nt & 1) * 2 + 1] = pack_bf16x2(dsv[2], dsv[3]);
        }
        // dQ += dS K   (k = keys, n = head dim)
#pragma unroll
        for (int kk = 0; kk < BC / 16; ++kk) {
#pragma unroll
            for (int nt = 0; nt < D / 8; nt += 2) {
                uint32_t b0, b1, b2, b3;
                const int mi = lane >> 3;
                const int krow = kk * 16 + (mi & 1) * 8 + (lane & 7);
                ldsm_x4_t(sKb + swz<D>(krow, nt + (mi >> 1)), b0, b1, b2, b3);
                mma16816(dq[nt], dsa[kk][0], dsa[kk][1], dsa[kk][2], dsa[kk][3], b0, b1);
                mma16816(dq[nt + 1], dsa[kk][0], dsa[kk][1], dsa[kk][2], dsa[kk][3], b2, b3);
            }
        }
        __syncthreads();
    }
    // epilogue: add diagonal-term dq, write bf16
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
        const int col = h * D + nt * 8 + tig * 2;
        if (row_a < p.S) {
            float x = dq[nt][0], y = dq[nt][1];
            if (p.dq_diag) { const float2 d2 = *reinterpret_cast<const float2*>(p.dq_diag + ((int64_t)b * p.S + row_a) * (int64_t)(p.nh * D) + col); x += d2.x; y += d2.y; }
            *reinterpret_cast<uint32_t*>(p.dq + ((int64_t)b * p.S + row_a) * p.lddq + col) = pack_bf16x2(x, y);
        }
        if (row_b < p.S) {
            float x = dq[nt][2], y = dq[nt][3];
            if (p.dq_diag) { const float2 d2 = *reinterpret_cast<const float2*>(p.dq_diag + ((int64_t)b * p.S + row_b) * (int64_t)(p.nh * D) + col); x += d2.x; y += d2.y; }
            *reinterpret_cast<uint32_t*>(p.dq + ((int64_t)b * p.S + row_b) * p.lddq + col) = pack_bf16x2(x, y);
        }
    }
}

// ------------------------------------------------------------------ host side
template <int D>
static int attn_fwd_t(AttnParams& p, cudaStream_t st) {
    if (p.J > 0) {
        const int64_t warps = ((int64_t)p.B * p.S * p.nh + (256 / D) - 1) / (256 / D);   // 256/D (row, head) items per warp
        diag_scores_kernel<D><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(p);
        SF_CUDA_CHECK_LAUNCH("diag_scores");
    }
    constexpr int smem = 128 * D * 2 + 4 * 64 * D * 2;
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); set = true; }
    dim3 grid((p.S + 127) / 128, p.nh, p.B);
    attn_fwd_kernel<D><<<grid, 256, smem, st>>>(p);
    SF_CUDA_CHECK_LAUNCH("attn_fwd");
    return 0;
}
template <int D>
static int attn_bwd_t(AttnParams& p, cudaStream_t st) {
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
    {
        constexpr int smem = 2 * 128 * D * 2 + 4 * 32 * D * 2 + 4 * 32 * 4;
        static bool set = false;
        if (!set) { cudaFuncSetAttribute(attn_bwd_dkv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); set = true; }
        dim3 grid((p.S + 127) / 128, p.nkv, p.B);
        attn_bwd_dkv_kernel<D><<<grid, 256, smem, st>>>(p);
        SF_CUDA_CHECK_LAUNCH("attn_bwd_dkv");
    }
    {
        constexpr int smem = 2 * 128 * D * 2 + 4 * 64 * D * 2;
        static bool set = false;
        if (!set) { cudaFuncSetAttribute(attn_bwd_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); set = true; }
        dim3 grid((p.S + 127) / 128, p.nh, p.B);
        attn_bwd_dq_kernel<D><<<grid, 256, smem, st>>>(p);
        SF_CUDA_CHECK_LAUNCH("attn_bwd_dq");
    }
    return 0;
}

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

static bool use_legacy_attention() { return opt(OPT_ATTN_LEGACY) == 1; }

int attn_fwd(const AttnDesc& a, cudaStream_t st) {
    if (int rc = attn_validate(a)) return rc;
    if (a.key_mask && !a.kvlen) return set_error(-22, "attention: key_mask given without kvlen/nonprefix (call mask_prefix)");
    AttnParams p; fill(p, a);
    if (!use_legacy_attention()) {
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
    return a.head_dim == 128 ? attn_fwd_t<128>(p, st) : attn_fwd_t<64>(p, st);
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
    if (!use_legacy_attention()) {
        // delta + diagonal terms by row kernels, block 0 (dK/dV and dQ) on tcgen05
        if (int rc = (a.head_dim == 128 ? attn_bwd_prep_t<128>(p, st) : attn_bwd_prep_t<64>(p, st))) return rc;
        return attn_bwd_tc(a, st);
    }
    return a.head_dim == 128 ? attn_bwd_t<128>(p, st) : attn_bwd_t<64>(p, st);
}

}  // namespace sf
