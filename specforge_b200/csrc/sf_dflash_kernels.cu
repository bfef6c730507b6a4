// sf_dflash_kernels.cu — row / tile kernels of the DFlash block-parallel draft step.
//
// FIRST CORRECT VERSION.  The projections run on the tcgen05 GEMM (sf_gemm.cuh); the kernels in this file are written
// for parity first: the block attention uses CUDA-core tiles (one CTA per anchor block x kv head), not tcgen05 yet —
// DESIGN.md §10 has the plan for the tensor-core variant.  Reference semantics, each kernel cites its lines:
//   noise ids / embedding      algorithms/common/dflash_family_model.py:221-245
//   labels, weights            :398-428, objective :332-383
//   per-head q/k RMSNorm+RoPE  modeling/draft/dflash.py:158-177 (Qwen3RMSNorm over head_dim, rotary on explicit positions)
//   DFlash mask                dflash_family_model.py:47-89 (context keys strictly before the anchor + the block's own
//                              noise keys, bidirectional; a dropped block attends to nothing and yields zeros, dflash.py:203-213)
#include "sf_host.h"
#include "sf_ptx.cuh"
#include "sf_dflash.h"

#include <cfloat>

namespace sf {
namespace dflash {

__device__ __forceinline__ float bf(const __nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float rbf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// ------------------------------------------------------------------ draft rows: positions, labels, weights
// One thread per draft row (b, n, o).  pos = anchor + o; tgt = input_ids[b, min(pos, S-1)];
// w = keep * (pos < S) * (o > 0) * loss_mask[b, min(pos, S-1)];  lw = w * exp(-max(o-1, 0) / gamma) (gamma > 0).
__global__ void rows_kernel(const int64_t* __restrict__ input_ids, const int64_t* __restrict__ loss_mask,
                            const int32_t* __restrict__ anchors, const uint8_t* __restrict__ keep, int B, int S, int N, int bs,
                            float gamma, int mask_id, int32_t* __restrict__ pos, int32_t* __restrict__ tgt,
                            int32_t* __restrict__ noise_id, float* __restrict__ w, float* __restrict__ lw) {
    const int64_t total = (int64_t)B * N * bs;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < total; r += (int64_t)gridDim.x * blockDim.x) {
        const int o = (int)(r % bs);
        const int n = (int)((r / bs) % N);
        const int b = (int)(r / ((int64_t)bs * N));
        const int a = anchors[b * N + n];
        const int kp = keep[b * N + n] ? 1 : 0;
        const int p = a + o;
        const int sp = min(p, S - 1);
        pos[r] = p;
        tgt[r] = (int32_t)input_ids[(int64_t)b * S + sp];
        const int at = (int)input_ids[(int64_t)b * S + min(max(a, 0), S - 1)];
        noise_id[r] = (o == 0 && kp) ? at : mask_id;
        const float wv = (kp && p < S && o > 0 && loss_mask[(int64_t)b * S + sp] != 0) ? 1.f : 0.f;
        w[r] = wv;
        lw[r] = (gamma > 0.f) ? wv * __expf(-(float)max(o - 1, 0) / gamma) : wv;
    }
}

// Deterministic sum of up to 4 float arrays of length n into out[0..3] (single block).
__global__ void __launch_bounds__(1024) sum4_kernel(const float* a0, const float* a1, const float* a2, const float* a3, int64_t n,
                                                    float* __restrict__ out) {
    __shared__ float red[4][32];
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const float* a[4] = {a0, a1, a2, a3};
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (a[k]) s[k] += a[k][i];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 4; ++k) { s[k] = warp_sum(s[k]); if (lane == 0) red[k][wid] = s[k]; }
    __syncthreads();
    if (wid == 0)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float t = warp_sum(lane < (int)(blockDim.x >> 5) ? red[k][lane] : 0.f);
            if (lane == 0 && a[k]) out[k] = t;
        }
}

// embedding gather of the noise ids: out[r, :] = embed[noise_id[r], :]
__global__ void __launch_bounds__(128) gather_rows_kernel(const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ ids,
                                                          __nv_bfloat16* __restrict__ out, int H) {
    const int64_t r = blockIdx.x;
    const uint4* src = reinterpret_cast<const uint4*>(table + (int64_t)ids[r] * H);
    uint4* dst = reinterpret_cast<uint4*>(out + r * H);
    for (int c = threadIdx.x; c < H / 8; c += blockDim.x) dst[c] = __ldg(src + c);
}

// ------------------------------------------------------------------ per-head RMSNorm + RoPE
// One warp per (row, head).  y = w * bf16(x * rstd) (bf16), out = bf16(bf16(y*cos) + bf16(rot(y)*sin)), rot = rotate_half.
// pos: explicit positions [M] (draft rows) or null -> row % S (context rows).
__global__ void __launch_bounds__(256)
headnorm_rope_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int n_heads, int d, const __nv_bfloat16* __restrict__ w,
                         const __nv_bfloat16* __restrict__ cos_t, const __nv_bfloat16* __restrict__ sin_t,
                         const int32_t* __restrict__ pos, int S, float eps, __nv_bfloat16* __restrict__ out, int64_t ldo, int64_t M) {
    const int64_t item = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= M * n_heads) return;
    const int h = (int)(item % n_heads);
    const int64_t r = item / n_heads;
    const int p = pos ? pos[r] : (int)(r % S);
    const __nv_bfloat16* xr = x + r * ldx + h * d;
    float ss = 0.f;
    for (int i = lane; i < d; i += 32) { const float v = bf(xr[i]); ss += v * v; }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)d + eps);
    const int half = d / 2;
    for (int i = lane; i < d; i += 32) {
        const int j = (i < half) ? i + half : i - half;
        const float yi = rbf(bf(w[i]) * rbf(bf(xr[i]) * rstd));
        const float yj = rbf(bf(w[j]) * rbf(bf(xr[j]) * rstd));
        const float rot = (i < half) ? -yj : yj;
        const float c = bf(cos_t[(int64_t)p * d + i]), s = bf(sin_t[(int64_t)p * d + i]);
        out[r * ldo + h * d + i] = __float2bfloat16_rn(rbf(yi * c) + rbf(rot * s));
    }
}

// Backward: g = d(out).  dy_i = g_i cos_i + (i < half ? g_{i+half} sin_{i+half} : -g_{i-half} sin_{i-half});
// dx = rstd (dy w - xhat mean(dy w xhat));  dw += dy * bf16(xhat) summed over rows and heads (per-block partials).
// Persistent grid; partial[blockIdx][d] is reduced by colsum_kernel in block order (deterministic).
__global__ void __launch_bounds__(256)
headnorm_rope_bwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int n_heads, int d, const __nv_bfloat16* __restrict__ w,
                         const __nv_bfloat16* __restrict__ cos_t, const __nv_bfloat16* __restrict__ sin_t,
                         const int32_t* __restrict__ pos, int S, float eps, const __nv_bfloat16* __restrict__ g, int64_t ldg,
                         __nv_bfloat16* __restrict__ dx, int64_t lddx, float* __restrict__ partial, int64_t M) {
    extern __shared__ float dw_s[];   // [8 warps][d]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* dw = dw_s + wid * d;
    for (int i = lane; i < d; i += 32) dw[i] = 0.f;
    const int64_t items = M * n_heads;
    const int half = d / 2;
    for (int64_t item = (int64_t)blockIdx.x * 8 + wid; item < items; item += (int64_t)gridDim.x * 8) {
        const int h = (int)(item % n_heads);
        const int64_t r = item / n_heads;
        const int p = pos ? pos[r] : (int)(r % S);
        const __nv_bfloat16* xr = x + r * ldx + h * d;
        const __nv_bfloat16* gr = g + r * ldg + h * d;
        float ss = 0.f;
        for (int i = lane; i < d; i += 32) { const float v = bf(xr[i]); ss += v * v; }
        ss = warp_sum(ss);
        const float rstd = rsqrtf(ss / (float)d + eps);
        float dot = 0.f;
        for (int i = lane; i < d; i += 32) {
            const int j = (i < half) ? i + half : i - half;
            const float sj = bf(sin_t[(int64_t)p * d + j]);
            const float dy = bf(gr[i]) * bf(cos_t[(int64_t)p * d + i]) + ((i < half) ? bf(gr[j]) * sj : -bf(gr[j]) * sj);
            const float xh = bf(xr[i]) * rstd;
            dot += dy * bf(w[i]) * xh;
            dw[i] += dy * rbf(xh);
        }
        dot = warp_sum(dot) / (float)d;
        for (int i = lane; i < d; i += 32) {
            const int j = (i < half) ? i + half : i - half;
            const float sj = bf(sin_t[(int64_t)p * d + j]);
            const float dy = bf(gr[i]) * bf(cos_t[(int64_t)p * d + i]) + ((i < half) ? bf(gr[j]) * sj : -bf(gr[j]) * sj);
            const float xh = bf(xr[i]) * rstd;
            dx[r * lddx + h * d + i] = __float2bfloat16_rn(rstd * (dy * bf(w[i]) - xh * dot));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += dw_s[k * d + i];
        partial[(int64_t)blockIdx.x * d + i] = t;
    }
}
// ---- the same two kernels for head dims that are a multiple of 32 (every real config: d = 64 / 128): lane l owns the EPL = d / 32
// contiguous elements [EPL*l, EPL*l + EPL) — one vector load per tensor per item instead of EPL two-byte loads, values stay in
// registers across the passes, and the rotate_half partner (i +- d/2) is always lane ^ 16.  Same formulas and rounding points.
template <int EPL> struct BfVec;
template <> struct BfVec<1> { using T = uint16_t; };
template <> struct BfVec<2> { using T = uint32_t; };
template <> struct BfVec<4> { using T = uint2; };
template <> struct BfVec<8> { using T = uint4; };
template <int EPL>
__device__ __forceinline__ void ld_bf(const __nv_bfloat16* p, float (&o)[EPL]) {
    typename BfVec<EPL>::T raw = __ldg(reinterpret_cast<const typename BfVec<EPL>::T*>(p));
    const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
    for (int i = 0; i < EPL; ++i) o[i] = __bfloat162float(e[i]);
}
template <int EPL>
__device__ __forceinline__ void st_bf(__nv_bfloat16* p, const float (&v)[EPL]) {
    typename BfVec<EPL>::T raw;
    __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(&raw);
#pragma unroll
    for (int i = 0; i < EPL; ++i) e[i] = __float2bfloat16_rn(v[i]);
    *reinterpret_cast<typename BfVec<EPL>::T*>(p) = raw;
}

template <int EPL>
__global__ void __launch_bounds__(256)
headnorm_rope_fwd_vec_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int n_heads, const __nv_bfloat16* __restrict__ w,
                             const __nv_bfloat16* __restrict__ cos_t, const __nv_bfloat16* __restrict__ sin_t,
                             const int32_t* __restrict__ pos, int S, float eps, __nv_bfloat16* __restrict__ out, int64_t ldo, int64_t M) {
    constexpr int d = EPL * 32;
    const int64_t item = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= M * n_heads) return;
    const int h = (int)(item % n_heads);
    const int64_t r = item / n_heads;
    const int p = pos ? pos[r] : (int)(r % S);
    float xv[EPL], wv[EPL], cv[EPL], sv[EPL], y[EPL], o[EPL];
    ld_bf<EPL>(x + r * ldx + h * d + lane * EPL, xv);
    ld_bf<EPL>(w + lane * EPL, wv);
    ld_bf<EPL>(cos_t + (int64_t)p * d + lane * EPL, cv);
    ld_bf<EPL>(sin_t + (int64_t)p * d + lane * EPL, sv);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) ss += xv[i] * xv[i];
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)d + eps);
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        y[i] = rbf(wv[i] * rbf(xv[i] * rstd));
        const float yj = __shfl_xor_sync(0xffffffffu, y[i], 16);
        const float rot = (lane < 16) ? -yj : yj;
        o[i] = rbf(y[i] * cv[i]) + rbf(rot * sv[i]);
    }
    st_bf<EPL>(out + r * ldo + h * d + lane * EPL, o);
}

template <int EPL>
__global__ void __launch_bounds__(256)
headnorm_rope_bwd_vec_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int n_heads, const __nv_bfloat16* __restrict__ w,
                             const __nv_bfloat16* __restrict__ cos_t, const __nv_bfloat16* __restrict__ sin_t,
                             const int32_t* __restrict__ pos, int S, float eps, const __nv_bfloat16* __restrict__ g, int64_t ldg,
                             __nv_bfloat16* __restrict__ dx, int64_t lddx, float* __restrict__ partial, int64_t M) {
    constexpr int d = EPL * 32;
    extern __shared__ float dw_s[];   // [8 warps][d]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float dw[EPL], wv[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) dw[i] = 0.f;
    ld_bf<EPL>(w + lane * EPL, wv);
    const int64_t items = M * n_heads;
    const int64_t stride = (int64_t)gridDim.x * 8;
    int64_t item = (int64_t)blockIdx.x * 8 + wid;
    // the next item's four vectors are requested before the current item's reductions (two items in flight per warp)
    float xn[EPL], gn[EPL], cn[EPL], sn[EPL];
    auto fetch = [&](int64_t it) {
        const int h = (int)(it % n_heads);
        const int64_t r = it / n_heads;
        const int p = pos ? pos[r] : (int)(r % S);
        ld_bf<EPL>(x + r * ldx + h * d + lane * EPL, xn);
        ld_bf<EPL>(g + r * ldg + h * d + lane * EPL, gn);
        ld_bf<EPL>(cos_t + (int64_t)p * d + lane * EPL, cn);
        ld_bf<EPL>(sin_t + (int64_t)p * d + lane * EPL, sn);
    };
    if (item < items) fetch(item);
    for (; item < items; item += stride) {
        float xv[EPL], gv[EPL], cv[EPL], sv[EPL], dy[EPL], xh[EPL], o[EPL];
#pragma unroll
        for (int i = 0; i < EPL; ++i) { xv[i] = xn[i]; gv[i] = gn[i]; cv[i] = cn[i]; sv[i] = sn[i]; }
        if (item + stride < items) fetch(item + stride);
        const int h = (int)(item % n_heads);
        const int64_t r = item / n_heads;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < EPL; ++i) ss += xv[i] * xv[i];
        ss = warp_sum(ss);
        const float rstd = rsqrtf(ss / (float)d + eps);
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const float gs = __shfl_xor_sync(0xffffffffu, gv[i] * sv[i], 16);      // g_j * sin_j of the rotate_half partner
            dy[i] = gv[i] * cv[i] + ((lane < 16) ? gs : -gs);
            xh[i] = xv[i] * rstd;
            dot += dy[i] * wv[i] * xh[i];
            dw[i] += dy[i] * rbf(xh[i]);
        }
        dot = warp_sum(dot) / (float)d;
#pragma unroll
        for (int i = 0; i < EPL; ++i) o[i] = rstd * (dy[i] * wv[i] - xh[i] * dot);
        st_bf<EPL>(dx + r * lddx + h * d + lane * EPL, o);
    }
#pragma unroll
    for (int i = 0; i < EPL; ++i) dw_s[wid * d + lane * EPL + i] = dw[i];
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += dw_s[k * d + i];
        partial[(int64_t)blockIdx.x * d + i] = t;
    }
}

// dst[c] (+)= sum_b partial[b][c], fixed order
__global__ void colsum_kernel(const float* __restrict__ partial, int nblocks, int d, float* __restrict__ dst, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    float t = 0.f;
    for (int b = 0; b < nblocks; ++b) t += partial[(int64_t)b * d + c];
    dst[c] = accumulate ? dst[c] + t : t;
}

// ------------------------------------------------------------------ block attention (CUDA-core tiles)
// CTA = (anchor block n, kv head, batch b): R = g*bs query rows (the g query heads of the group x the block's bs tokens),
// keys = context [0, anchor) then the block's own bs noise keys.  All R rows share that key set, so tiles need no per-row
// mask.  256 threads, TPR = 256/R threads per row; thread (r, p) owns score columns k = p, p+TPR, .. and output columns
// c = p, p+TPR, ..  Row index r = hg*bs + o  (hg = head within the group, o = token within the block).
constexpr int kTK = 32;
constexpr int kMaxCols = 64;   // d / TPR <= 64  (R = 128, d = 128)


__device__ __forceinline__ void load_rows_to_smem(float* dst, int stride, const __nv_bfloat16* src, int64_t ld, int rows, int d) {
    for (int i = threadIdx.x; i < rows * d; i += blockDim.x) {
        const int r = i / d, c = i % d;
        dst[r * stride + c] = bf(src[(int64_t)r * ld + c]);
    }
}

__global__ void __launch_bounds__(256) attn_fwd_kernel(AttnArgs a) {
    extern __shared__ float sm[];
    const int n = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
    const int d = a.d, bs = a.bs, g = a.nh / a.nkv, R = g * bs, TPR = 256 / R, ds = d + 1;
    float* Qs = sm;                       // [R][ds]
    float* Ks = Qs + R * ds;              // [kTK][ds]
    float* Vs = Ks + kTK * ds;            // [kTK][d]
    float* Ss = Vs + kTK * d;             // [R][kTK+1]
    float* m_s = Ss + R * (kTK + 1);      // [R]
    float* l_s = m_s + R;
    float* sc_s = l_s + R;
    const int t = threadIdx.x, r = t / TPR, p = t % TPR;
    const int hg = r / bs, o = r % bs;
    const int64_t qrow0 = ((int64_t)b * a.N + n) * bs;        // first draft row of the block
    const int anchor = a.anchors[b * a.N + n];
    const bool kept = a.keep[b * a.N + n] != 0;
    const int64_t orow = qrow0 + o;
    const int head = kvh * g + hg;
    if (!kept) {                                              // dropped block: zeros
        for (int c = p; c < d; c += TPR) a.out[orow * a.ldo + head * d + c] = __float2bfloat16_rn(0.f);
        if (p == 0) a.lse[orow * a.nh + head] = 0.f;
        return;
    }
    for (int i = t; i < R * d; i += 256) {
        const int rr = i / d, c = i % d;
        Qs[rr * ds + c] = bf(a.q[(qrow0 + rr % bs) * a.ldq + (kvh * g + rr / bs) * d + c]);
    }
    if (p == 0) { m_s[r] = -FLT_MAX; l_s[r] = 0.f; }
    float acc[kMaxCols];
#pragma unroll
    for (int i = 0; i < kMaxCols; ++i) acc[i] = 0.f;
    const int n_ctx_tiles = (anchor + kTK - 1) / kTK;
    // sliding-window layer: row o sees context keys >= lo_r and own slots <= o; tiles below the block's window are skipped
    const int W = a.window, lo_r = W > 0 ? anchor + o - (W - 1) : 0;
    const int tile0 = W > 0 ? min(n_ctx_tiles, max(0, anchor - (W - 1)) / kTK) : 0;
    for (int tile = tile0; tile <= n_ctx_tiles; ++tile) {
        const bool own = tile == n_ctx_tiles;
        const int k0 = own ? 0 : tile * kTK;
        const int nk = own ? bs : min(kTK, anchor - k0);
        __syncthreads();
        if (own) {
            load_rows_to_smem(Ks, ds, a.kn + qrow0 * a.ldkn + kvh * d, a.ldkn, nk, d);
            load_rows_to_smem(Vs, d, a.vn + qrow0 * a.ldvn + kvh * d, a.ldvn, nk, d);
        } else {
            load_rows_to_smem(Ks, ds, a.kc + ((int64_t)b * a.S + k0) * a.ldkc + kvh * d, a.ldkc, nk, d);
            load_rows_to_smem(Vs, d, a.vc + ((int64_t)b * a.S + k0) * a.ldvc + kvh * d, a.ldvc, nk, d);
        }
        __syncthreads();
        for (int k = p; k < nk; k += TPR) {
            float s = 0.f;
            for (int c = 0; c < d; ++c) s += Qs[r * ds + c] * Ks[k * ds + c];
            const bool allowed = W == 0 || (own ? k <= o : k0 + k >= lo_r);
            Ss[r * (kTK + 1) + k] = allowed ? s * a.scale : -FLT_MAX;      // -FLT_MAX marks a masked key (exp -> 0 below)
        }
        __syncthreads();
        if (p == 0) {
            float mx = m_s[r];
            for (int k = 0; k < nk; ++k) mx = fmaxf(mx, Ss[r * (kTK + 1) + k]);
            const float sc = __expf(m_s[r] - mx);
            float l = l_s[r] * sc;
            for (int k = 0; k < nk; ++k) {
                const float sv = Ss[r * (kTK + 1) + k];
                const float e = sv == -FLT_MAX ? 0.f : __expf(sv - mx);
                Ss[r * (kTK + 1) + k] = e; l += e;
            }
            m_s[r] = mx; l_s[r] = l; sc_s[r] = sc;
        }
        __syncthreads();
        const float sc = sc_s[r];
#pragma unroll
        for (int i = 0; i < kMaxCols; ++i) {
            const int c = p + i * TPR;
            if (c < d) {
                float v = acc[i] * sc;
                for (int k = 0; k < nk; ++k) v += Ss[r * (kTK + 1) + k] * Vs[k * d + c];
                acc[i] = v;
            }
        }
    }
    __syncthreads();
    const float inv = 1.f / l_s[r];
#pragma unroll
    for (int i = 0; i < kMaxCols; ++i) {
        const int c = p + i * TPR;
        if (c < d) a.out[orow * a.ldo + head * d + c] = __float2bfloat16_rn(acc[i] * inv);
    }
    if (p == 0) a.lse[orow * a.nh + head] = m_s[r] + __logf(l_s[r]);
}

// Backward, query-stationary: dQ for the block's rows over all its keys, dK/dV of the block's own noise keys, delta.
__global__ void __launch_bounds__(256) attn_bwd_q_kernel(AttnArgs a) {
    extern __shared__ float sm[];
    const int n = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
    const int d = a.d, bs = a.bs, g = a.nh / a.nkv, R = g * bs, TPR = 256 / R, ds = d + 1, ts = kTK + 1;
    float* Qs = sm;                       // [R][ds]
    float* Gs = Qs + R * ds;              // dO [R][ds]
    float* Ks = Gs + R * ds;              // [kTK][ds]
    float* Vs = Ks + kTK * ds;            // [kTK][ds]
    float* Ps = Vs + kTK * ds;            // P  [R][ts]
    float* Ds = Ps + R * ts;              // dS [R][ts]
    float* lse_s = Ds + R * ts;           // [R]
    float* del_s = lse_s + R;             // [R]
    const int t = threadIdx.x, r = t / TPR, p = t % TPR;
    const int hg = r / bs, o = r % bs;
    const int64_t qrow0 = ((int64_t)b * a.N + n) * bs;
    const int anchor = a.anchors[b * a.N + n];
    const bool kept = a.keep[b * a.N + n] != 0;
    const int64_t orow = qrow0 + o;
    const int head = kvh * g + hg;
    if (!kept) {
        for (int c = p; c < d; c += TPR) a.dq[orow * a.lddq + head * d + c] = __float2bfloat16_rn(0.f);
        if (p == 0) a.delta[orow * a.nh + head] = 0.f;
        for (int i = t; i < bs * d; i += 256) {
            const int k = i / d, c = i % d;
            a.dkn[(qrow0 + k) * a.lddkn + kvh * d + c] = __float2bfloat16_rn(0.f);
            a.dvn[(qrow0 + k) * a.lddvn + kvh * d + c] = __float2bfloat16_rn(0.f);
        }
        return;
    }
    for (int i = t; i < R * d; i += 256) {
        const int rr = i / d, c = i % d;
        const int64_t row = qrow0 + rr % bs;
        const int hh = kvh * g + rr / bs;
        Qs[rr * ds + c] = bf(a.q[row * a.ldq + hh * d + c]);
        Gs[rr * ds + c] = bf(a.dout[row * a.lddo + hh * d + c]);
    }
    {   // delta = rowsum(dO * O): TPR consecutive lanes share a row
        float dl = 0.f;
        for (int c = p; c < d; c += TPR) dl += bf(a.dout[orow * a.lddo + head * d + c]) * bf(a.out[orow * a.ldo + head * d + c]);
        for (int off = TPR / 2; off > 0; off >>= 1) dl += __shfl_xor_sync(0xffffffffu, dl, off);
        if (p == 0) { del_s[r] = dl; lse_s[r] = a.lse[orow * a.nh + head]; a.delta[orow * a.nh + head] = dl; }
    }
    float acc[kMaxCols];
#pragma unroll
    for (int i = 0; i < kMaxCols; ++i) acc[i] = 0.f;
    const int n_ctx_tiles = (anchor + kTK - 1) / kTK;
    const int W = a.window, lo_r = W > 0 ? anchor + o - (W - 1) : 0;          // sliding-window layer, as in the forward
    const int tile0 = W > 0 ? min(n_ctx_tiles, max(0, anchor - (W - 1)) / kTK) : 0;
    for (int tile = tile0; tile <= n_ctx_tiles; ++tile) {
        const bool own = tile == n_ctx_tiles;
        const int k0 = own ? 0 : tile * kTK;
        const int nk = own ? bs : min(kTK, anchor - k0);
        __syncthreads();
        if (own) {
            load_rows_to_smem(Ks, ds, a.kn + qrow0 * a.ldkn + kvh * d, a.ldkn, nk, d);
            load_rows_to_smem(Vs, ds, a.vn + qrow0 * a.ldvn + kvh * d, a.ldvn, nk, d);
        } else {
            load_rows_to_smem(Ks, ds, a.kc + ((int64_t)b * a.S + k0) * a.ldkc + kvh * d, a.ldkc, nk, d);
            load_rows_to_smem(Vs, ds, a.vc + ((int64_t)b * a.S + k0) * a.ldvc + kvh * d, a.ldvc, nk, d);
        }
        __syncthreads();
        for (int k = p; k < nk; k += TPR) {
            float s = 0.f, dp = 0.f;
            for (int c = 0; c < d; ++c) { s += Qs[r * ds + c] * Ks[k * ds + c]; dp += Gs[r * ds + c] * Vs[k * ds + c]; }
            const bool allowed = W == 0 || (own ? k <= o : k0 + k >= lo_r);
            const float pr = allowed ? __expf(s * a.scale - lse_s[r]) : 0.f;
            Ps[r * ts + k] = pr;
            Ds[r * ts + k] = pr * (dp - del_s[r]) * a.scale;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kMaxCols; ++i) {
            const int c = p + i * TPR;
            if (c < d) {
                float v = acc[i];
                for (int k = 0; k < nk; ++k) v += Ds[r * ts + k] * Ks[k * ds + c];
                acc[i] = v;
            }
        }
        if (own) {   // this CTA is the only reader of its block's noise keys for this kv head: plain stores
            for (int i = t; i < bs * d; i += 256) {
                const int k = i / d, c = i % d;
                float dk = 0.f, dv = 0.f;
                for (int rr = 0; rr < R; ++rr) { dk += Ds[rr * ts + k] * Qs[rr * ds + c]; dv += Ps[rr * ts + k] * Gs[rr * ds + c]; }
                a.dkn[(qrow0 + k) * a.lddkn + kvh * d + c] = __float2bfloat16_rn(dk);
                a.dvn[(qrow0 + k) * a.lddvn + kvh * d + c] = __float2bfloat16_rn(dv);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kMaxCols; ++i) {
        const int c = p + i * TPR;
        if (c < d) a.dq[orow * a.lddq + head * d + c] = __float2bfloat16_rn(acc[i]);
    }
}

// Backward, context-key-stationary: dK/dV of a tile of kTK context keys, summed over every kept block whose anchor lies
// beyond the tile start (no atomics: one CTA owns the tile).  Needs lse and delta from the two kernels above.
__global__ void __launch_bounds__(256) attn_bwd_ctx_kernel(AttnArgs a) {
    extern __shared__ float sm[];
    const int tile = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
    const int d = a.d, bs = a.bs, g = a.nh / a.nkv, R = g * bs, TPR = 256 / R, ds = d + 1, ts = kTK + 1;
    float* Qs = sm;                       // [R][ds]
    float* Gs = Qs + R * ds;              // [R][ds]
    float* Ks = Gs + R * ds;              // [kTK][ds]
    float* Vs = Ks + kTK * ds;            // [kTK][ds]
    float* Ps = Vs + kTK * ds;            // [R][ts]
    float* Ds = Ps + R * ts;              // [R][ts]
    const int t = threadIdx.x, r = t / TPR, p = t % TPR;
    const int k0 = tile * kTK;
    const int nk_tile = min(kTK, a.S - k0);
    load_rows_to_smem(Ks, ds, a.kc + ((int64_t)b * a.S + k0) * a.ldkc + kvh * d, a.ldkc, nk_tile, d);
    load_rows_to_smem(Vs, ds, a.vc + ((int64_t)b * a.S + k0) * a.ldvc + kvh * d, a.ldvc, nk_tile, d);
    // thread owns key kk = t / 8 and columns c = (t % 8) + 8*i  (kTK * 8 = 256 threads)
    const int kk = t / 8, c0 = t % 8;
    float dk[16], dv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { dk[i] = 0.f; dv[i] = 0.f; }
    for (int n = 0; n < a.N; ++n) {
        const int anchor = a.anchors[b * a.N + n];
        if (!a.keep[b * a.N + n] || anchor <= k0) continue;      // uniform across the CTA
        if (a.window > 0 && k0 + nk_tile - 1 < anchor - (a.window - 1)) continue;   // the whole tile lies below the block's window
        const int nk = min(nk_tile, anchor - k0);
        const int64_t qrow0 = ((int64_t)b * a.N + n) * bs;
        __syncthreads();
        for (int i = t; i < R * d; i += 256) {
            const int rr = i / d, c = i % d;
            const int64_t row = qrow0 + rr % bs;
            const int hh = kvh * g + rr / bs;
            Qs[rr * ds + c] = bf(a.q[row * a.ldq + hh * d + c]);
            Gs[rr * ds + c] = bf(a.dout[row * a.lddo + hh * d + c]);
        }
        __syncthreads();
        {
            const int64_t orow = qrow0 + r % bs;
            const int head = kvh * g + r / bs;
            const float lse = a.lse[orow * a.nh + head], del = a.delta[orow * a.nh + head];
            for (int k = p; k < kTK; k += TPR) {
                float pr = 0.f, dsv = 0.f;
                if (k < nk && (a.window == 0 || k0 + k >= anchor + r % bs - (a.window - 1))) {
                    float s = 0.f, dp = 0.f;
                    for (int c = 0; c < d; ++c) { s += Qs[r * ds + c] * Ks[k * ds + c]; dp += Gs[r * ds + c] * Vs[k * ds + c]; }
                    pr = __expf(s * a.scale - lse);
                    dsv = pr * (dp - del) * a.scale;
                }
                Ps[r * ts + k] = pr;
                Ds[r * ts + k] = dsv;
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int c = c0 + 8 * i;
            if (c < d) {
                float x = dk[i], y = dv[i];
                for (int rr = 0; rr < R; ++rr) { x += Ds[rr * ts + kk] * Qs[rr * ds + c]; y += Ps[rr * ts + kk] * Gs[rr * ds + c]; }
                dk[i] = x; dv[i] = y;
            }
        }
    }
    if (kk < nk_tile) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int c = c0 + 8 * i;
            if (c < d) {
                a.dkc[((int64_t)b * a.S + k0 + kk) * a.lddkc + kvh * d + c] = __float2bfloat16_rn(dk[i]);
                a.dvc[((int64_t)b * a.S + k0 + kk) * a.lddvc + kvh * d + c] = __float2bfloat16_rn(dv[i]);
            }
        }
    }
}

// ------------------------------------------------------------------ hard-label cross entropy (objective :332-383)
// One block per draft row: online (max, argmax, sum-exp) over the V logits, nll = lse - x[tgt];
// row_loss = nll * lw, row_correct = (argmax == tgt && w > 0.5); in place d(logits) = lw / loss_den * (softmax - onehot).
__global__ void __launch_bounds__(512) ce_kernel(__nv_bfloat16* __restrict__ logits, int64_t ld, int V, const int32_t* __restrict__ tgt,
                                                 const float* __restrict__ w, const float* __restrict__ lw,
                                                 const float* __restrict__ sums /* [0] = loss_den */, int write_grad,
                                                 float* __restrict__ row_loss, float* __restrict__ row_correct, int grad_of_numerator,
                                                 int mode, float* __restrict__ row_state /* [M][2] = {max, 1/sum-exp} */) {
    // mode 0: statistics + loss + gradient in one launch (loss_type "dflash": the row's weight lw is known up front).
    // mode 1: statistics only — row_loss = the raw -log q(target), row_state = {max, 1/sum-exp}: the D-PACE objectives need every
    //         slot's q before any weight exists (dflash_family_model.py:245-279,360-369).   mode 2: gradient only, from row_state.
    __shared__ float redv[16], redd[16];
    __shared__ int redi[16];
    const int64_t r = blockIdx.x;
    const float lwr = (mode == 1) ? 1.f : lw[r], wr = w[r];
    __nv_bfloat16* row = logits + r * ld;
    if (mode == 2) {
        const float coef = grad_of_numerator ? lwr : lwr / sums[0];
        if (coef == 0.f) {
            for (int c = threadIdx.x; c < V; c += 512) row[c] = __float2bfloat16_rn(0.f);
            return;
        }
        const float m2 = row_state[2 * r], inv2 = row_state[2 * r + 1];
        const int tg2 = tgt[r];
        for (int c = threadIdx.x; c < V; c += 512) {
            const float pr = __expf(bf(row[c]) - m2) * inv2;
            row[c] = __float2bfloat16_rn(coef * (pr - (c == tg2 ? 1.f : 0.f)));
        }
        return;
    }
    if (mode == 0 && lwr == 0.f && wr == 0.f) {
        if (write_grad)
            for (int c = threadIdx.x; c < V; c += 512) row[c] = __float2bfloat16_rn(0.f);
        if (threadIdx.x == 0) { row_loss[r] = 0.f; row_correct[r] = 0.f; }
        return;
    }
    float m = -FLT_MAX, dsum = 0.f;
    int am = 0x7fffffff;
    for (int c = threadIdx.x; c < V; c += 512) {
        const float v = bf(row[c]);
        if (v > m) { dsum *= __expf(m - v); m = v; am = c; }
        dsum += __expf(v - m);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    {
        float wm = m; int wi = am;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, wm, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (om > wm || (om == wm && oi < wi)) { wm = om; wi = oi; }
        }
        const float wd = warp_sum(dsum * __expf(m - wm));
        if (lane == 0) { redv[wid] = wm; redi[wid] = wi; redd[wid] = wd; }
        __syncthreads();
        float tm = lane < 16 ? redv[lane] : -FLT_MAX;
        int ti = lane < 16 ? redi[lane] : 0x7fffffff;
        const float td = lane < 16 ? redd[lane] : 0.f;
        float bm = tm; int bi = ti;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, bm, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }
        }
        dsum = warp_sum(td * __expf(tm - bm));
        m = bm; am = bi;
    }
    const int tg = tgt[r];
    const float lse = m + __logf(dsum);
    if (threadIdx.x == 0) {
        row_loss[r] = (lse - bf(row[tg])) * lwr;
        row_correct[r] = (am == tg && wr > 0.5f) ? 1.f : 0.f;
        if (mode == 1) { row_state[2 * r] = m; row_state[2 * r + 1] = 1.f / dsum; }
    }
    if (write_grad && mode == 0) {
        const float coef = grad_of_numerator ? lwr : lwr / sums[0];
        const float inv = 1.f / dsum;
        __syncthreads();   // row[tg] was read above by thread 0
        for (int c = threadIdx.x; c < V; c += 512) {
            const float pr = __expf(bf(row[c]) - m) * inv;
            row[c] = __float2bfloat16_rn(coef * (pr - (c == tg ? 1.f : 0.f)));
        }
    }
}

// D-PACE position weights (dflash_family_model.py:245-279): one thread per anchor block of bs slots.  nll[] holds -log q(target);
// smooth_k = (1-alpha) q_k + alpha on supervised slots (w > 0), 1 elsewhere; prefix_k = prod_{i<=k} smooth_i;
// type 2 (cumulative-confidence-only): weight = prefix;  type 1 (dpace): weight_k = sum_{i>=k} prefix_i w_i;
// type 3 (continuation-value-only): that suffix sum / max(prefix_k, tiny).  Writes lw = w * weight (detached) and
// row_loss = nll * lw in place.
__global__ void __launch_bounds__(256) dpace_weights_kernel(float* __restrict__ nll_loss, const float* __restrict__ w, float* __restrict__ lw,
                                                            int64_t blocks, int bs, int type, float alpha) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= blocks) return;
    float* nl = nll_loss + b * bs;
    const float* wb = w + b * bs;
    float* lwb = lw + b * bs;
    float prefix = 1.f;
    for (int k = 0; k < bs; ++k) {                       // forward: prefix products, parked in lw
        const float wk = wb[k];
        const float smooth = wk > 0.f ? (1.f - alpha) * __expf(-nl[k]) + alpha : 1.f;
        prefix *= smooth;
        lwb[k] = prefix;
    }
    float suffix = 0.f;
    for (int k = bs - 1; k >= 0; --k) {                  // backward: suffix sums of prefix * w
        const float pk = lwb[k], wk = wb[k];
        suffix += pk * wk;
        float dw = pk;
        if (type == 1) dw = suffix;
        else if (type == 3) dw = suffix / fmaxf(pk, 1.17549435e-38f);
        lwb[k] = wk * dw;
        nl[k] = nl[k] * wk * dw;
    }
}
__global__ void set_den_kernel(float* sums, float v) { sums[0] = v; }

// metrics = {loss_num, loss_den, correct, acc_den}; loss = loss_num / loss_den   (sums: see sf_dflash.h)
__global__ void finalize_kernel(const float* __restrict__ sums, float* __restrict__ metrics, float* __restrict__ loss) {
    metrics[0] = sums[2]; metrics[1] = sums[0]; metrics[2] = sums[3]; metrics[3] = sums[1];
    loss[0] = sums[0] > 0.f ? sums[2] / sums[0] : 0.f;    // no supervised position: 0 rather than 0/0
}

// ------------------------------------------------------------------ host wrappers
static int64_t attn_smem_fwd(int R, int d) { return (int64_t)(R * (d + 1) + kTK * (d + 1) + kTK * d + R * (kTK + 1) + 3 * R) * 4; }
static int64_t attn_smem_bwd(int R, int d) { return (int64_t)(2 * R * (d + 1) + 2 * kTK * (d + 1) + 2 * R * (kTK + 1) + 2 * R) * 4; }

int rows(const int64_t* input_ids, const int64_t* loss_mask, const int32_t* anchors, const uint8_t* keep, int B, int S, int N, int bs,
         float gamma, int mask_id, int32_t* pos, int32_t* tgt, int32_t* noise_id, float* w, float* lw, float* sums, cudaStream_t st) {
    const int64_t n = (int64_t)B * N * bs;
    rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(input_ids, loss_mask, anchors, keep, B, S, N, bs, gamma, mask_id, pos, tgt,
                                                              noise_id, w, lw);
    SF_CUDA_CHECK_LAUNCH("dflash rows");
    sum4_kernel<<<1, 1024, 0, st>>>(lw, w, nullptr, nullptr, n, sums);    // sums[0] = loss_den, sums[1] = acc_den
    SF_CUDA_CHECK_LAUNCH("dflash sums");
    return 0;
}
int gather_rows(const void* table, const int32_t* ids, void* out, int64_t M, int H, cudaStream_t st) {
    gather_rows_kernel<<<(unsigned)M, 128, 0, st>>>((const __nv_bfloat16*)table, ids, (__nv_bfloat16*)out, H);
    SF_CUDA_CHECK_LAUNCH("dflash gather");
    return 0;
}
int headnorm_rope_fwd(const void* x, int64_t ldx, int n_heads, int d, const void* w, const void* cos_t, const void* sin_t,
                      const int32_t* pos, int S, float eps, void* out, int64_t ldo, int64_t M, cudaStream_t st) {
    const int64_t warps = M * n_heads;
    const unsigned nblk = (unsigned)((warps * 32 + 255) / 256);
    const bool vec = (ldx % 8 == 0) && (ldo % 8 == 0);       // vector loads need the rows aligned to the widest vector (16 B)
#define SF_HN_FWD(EPL)                                                                                                       \
    headnorm_rope_fwd_vec_kernel<EPL><<<nblk, 256, 0, st>>>((const __nv_bfloat16*)x, ldx, n_heads, (const __nv_bfloat16*)w,  \
        (const __nv_bfloat16*)cos_t, (const __nv_bfloat16*)sin_t, pos, S, eps, (__nv_bfloat16*)out, ldo, M)
    if (vec && (d == 32 || d == 64 || d == 128 || d == 256)) {
        if (d == 32) SF_HN_FWD(1); else if (d == 64) SF_HN_FWD(2); else if (d == 128) SF_HN_FWD(4); else SF_HN_FWD(8);
        SF_CUDA_CHECK_LAUNCH("dflash headnorm_rope_fwd");
        return 0;
    }
#undef SF_HN_FWD
    headnorm_rope_fwd_kernel<<<nblk, 256, 0, st>>>(
        (const __nv_bfloat16*)x, ldx, n_heads, d, (const __nv_bfloat16*)w, (const __nv_bfloat16*)cos_t, (const __nv_bfloat16*)sin_t, pos, S,
        eps, (__nv_bfloat16*)out, ldo, M);
    SF_CUDA_CHECK_LAUNCH("dflash headnorm_rope_fwd");
    return 0;
}
int headnorm_rope_bwd(const void* x, int64_t ldx, int n_heads, int d, const void* w, const void* cos_t, const void* sin_t,
                      const int32_t* pos, int S, float eps, const void* g, int64_t ldg, void* dx, int64_t lddx, float* dw,
                      int accumulate, float* partial_ws, int64_t M, cudaStream_t st) {
    const bool vec = (ldx % 8 == 0) && (ldg % 8 == 0) && (lddx % 8 == 0);
#define SF_HN_BWD(EPL)                                                                                                       \
    headnorm_rope_bwd_vec_kernel<EPL><<<kNormBlocks, 256, 8 * d * 4, st>>>((const __nv_bfloat16*)x, ldx, n_heads,           \
        (const __nv_bfloat16*)w, (const __nv_bfloat16*)cos_t, (const __nv_bfloat16*)sin_t, pos, S, eps, (const __nv_bfloat16*)g, ldg, \
        (__nv_bfloat16*)dx, lddx, partial_ws, M)
    if (vec && (d == 32 || d == 64 || d == 128 || d == 256)) {
        if (d == 32) SF_HN_BWD(1); else if (d == 64) SF_HN_BWD(2); else if (d == 128) SF_HN_BWD(4); else SF_HN_BWD(8);
    } else {
        headnorm_rope_bwd_kernel<<<kNormBlocks, 256, 8 * d * 4, st>>>(
            (const __nv_bfloat16*)x, ldx, n_heads, d, (const __nv_bfloat16*)w, (const __nv_bfloat16*)cos_t, (const __nv_bfloat16*)sin_t, pos, S,
            eps, (const __nv_bfloat16*)g, ldg, (__nv_bfloat16*)dx, lddx, partial_ws, M);
    }
#undef SF_HN_BWD
    SF_CUDA_CHECK_LAUNCH("dflash headnorm_rope_bwd");
    colsum_kernel<<<(d + 127) / 128, 128, 0, st>>>(partial_ws, kNormBlocks, d, dw, accumulate);
    SF_CUDA_CHECK_LAUNCH("dflash colsum");
    return 0;
}
static int attn_check(const AttnArgs& a) {
    const int g = a.nh / a.nkv, R = g * a.bs;
    if (a.nh % a.nkv) return set_error(-22, "dflash attention: nh %% nkv != 0");
    if (R < 8 || R > 128 || (R & (R - 1))) return set_error(-22, "dflash attention: group*block_size = %d must be a power of two in [8, 128]", R);
    if (a.d > 128 || a.d % 2) return set_error(-22, "dflash attention: head_dim %d unsupported (even, <= 128)", a.d);
    if ((a.d + 256 / R - 1) / (256 / R) > kMaxCols) return set_error(-22, "dflash attention: head_dim %d too large for %d rows", a.d, R);
    return 0;
}
// The tcgen05 kernels (sf_dflash_attn_tc*.cu) are the default wherever their tiling covers the shape; the CUDA-core kernels
// below remain for the other shapes and as the A/B baseline (sf_debug_option("dflash_attn_tc", -1) / SF_DFLASH_ATTN_TC=-1).
int attn_fwd(const AttnArgs& a, cudaStream_t st) {
    if (opt(OPT_DFLASH_ATTN_TC) >= 0 && attn_tc_supported(a)) return attn_fwd_tc(a, st);
    return attn_fwd_cc(a, st);
}
int attn_bwd(const AttnArgs& a, cudaStream_t st) {
    if (opt(OPT_DFLASH_ATTN_TC) >= 0 && attn_tc_bwd_supported(a)) return attn_bwd_tc(a, st);
    return attn_bwd_cc(a, st);
}
int attn_fwd_cc(const AttnArgs& a, cudaStream_t st) {
    if (int rc = attn_check(a)) return rc;
    const int R = (a.nh / a.nkv) * a.bs;
    const int smem = (int)attn_smem_fwd(R, a.d);
    static int set = 0;
    if (smem > set) { cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); set = smem; }
    attn_fwd_kernel<<<dim3(a.N, a.nkv, a.B), 256, smem, st>>>(a);
    SF_CUDA_CHECK_LAUNCH("dflash attn_fwd");
    return 0;
}
int attn_bwd_cc(const AttnArgs& a, cudaStream_t st) {
    if (int rc = attn_check(a)) return rc;
    const int R = (a.nh / a.nkv) * a.bs;
    const int smem = (int)attn_smem_bwd(R, a.d);
    static int set = 0;
    if (smem > set) {
        cudaFuncSetAttribute(attn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(attn_bwd_ctx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        set = smem;
    }
    attn_bwd_q_kernel<<<dim3(a.N, a.nkv, a.B), 256, smem, st>>>(a);
    SF_CUDA_CHECK_LAUNCH("dflash attn_bwd_q");
    attn_bwd_ctx_kernel<<<dim3((a.S + kTK - 1) / kTK, a.nkv, a.B), 256, smem, st>>>(a);
    SF_CUDA_CHECK_LAUNCH("dflash attn_bwd_ctx");
    return 0;
}
int ce(void* logits, int64_t ld, int V, const int32_t* tgt, const float* w, float* lw, float* sums, int write_grad,
       float* row_loss, float* row_correct, int64_t M, int grad_of_numerator, int loss_type, float dpace_alpha, int bs, int batch,
       float* row_state, cudaStream_t st) {
    if (loss_type == 0) {
        ce_kernel<<<(unsigned)M, 512, 0, st>>>((__nv_bfloat16*)logits, ld, V, tgt, w, lw, sums, write_grad, row_loss, row_correct,
                                              grad_of_numerator, 0, nullptr);
        SF_CUDA_CHECK_LAUNCH("dflash ce");
    } else {
        // D-PACE: statistics (raw -log q per slot) -> per-block weights -> gradient; loss = sum(nll w weight) / batch size
        ce_kernel<<<(unsigned)M, 512, 0, st>>>((__nv_bfloat16*)logits, ld, V, tgt, w, lw, sums, 0, row_loss, row_correct, grad_of_numerator,
                                              1, row_state);
        SF_CUDA_CHECK_LAUNCH("dflash ce stats");
        const int64_t blocks = M / bs;
        dpace_weights_kernel<<<(unsigned)((blocks + 255) / 256), 256, 0, st>>>(row_loss, w, lw, blocks, bs, loss_type, dpace_alpha);
        SF_CUDA_CHECK_LAUNCH("dflash dpace weights");
        set_den_kernel<<<1, 1, 0, st>>>(sums, (float)batch);
        SF_CUDA_CHECK_LAUNCH("dflash dpace den");
        if (write_grad) {
            ce_kernel<<<(unsigned)M, 512, 0, st>>>((__nv_bfloat16*)logits, ld, V, tgt, w, lw, sums, 1, row_loss, row_correct,
                                                  grad_of_numerator, 2, row_state);
            SF_CUDA_CHECK_LAUNCH("dflash ce grad");
        }
    }
    sum4_kernel<<<1, 1024, 0, st>>>(nullptr, nullptr, row_loss, row_correct, M, sums);   // sums[2] = loss_num, sums[3] = correct
    SF_CUDA_CHECK_LAUNCH("dflash ce sums");
    return 0;
}
int finalize_loss(const float* sums, float* metrics, float* loss, cudaStream_t st) {
    finalize_kernel<<<1, 1, 0, st>>>(sums, metrics, loss);
    SF_CUDA_CHECK_LAUNCH("dflash finalize");
    return 0;
}
}  // namespace dflash
}  // namespace sf
