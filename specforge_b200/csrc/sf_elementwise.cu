// sf_elementwise.cu — HBM-bound row kernels of the EAGLE3 draft step: (gather+)RMSNorm fwd/bwd,
// RoPE fwd/bwd, SwiGLU fwd/bwd, sequence shift.  All are one-pass, 16-byte vectorised, fp32 math
// with a single rounding to bf16 (matches the reference's compiled GPU path).
#include "sf_host.h"
#include "sf_ptx.cuh"

namespace sf {

constexpr int kRowThreads = 256;
constexpr int kMaxChunks = 4;  // 256 threads * 4 chunks * 8 elems = 8192 columns max

struct bf16x8 {
    uint4 u;
    __device__ __forceinline__ void unpack(float (&f)[8]) const {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
            f[2 * i] = __bfloat162float(b.x);
            f[2 * i + 1] = __bfloat162float(b.y);
        }
    }
    __device__ __forceinline__ void pack(const float (&f)[8]) {
        u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
        u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    }
};
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < (blockDim.x >> 5)) ? red[l] : 0.f;
    t = warp_sum(t);
    return t;
}

// ------------------------------------------------------------------ RMSNorm forward (+ optional gather)
// out[r, :] = w * bf16(x[r] * rsqrt(mean(x^2) + eps))          (llama3_eagle.py:1561-1567)
// gather mode: x[r] = table[ids[b, s + shift]] (token 0 past the end; eagle3/model.py:382,428-432)
__global__ void __launch_bounds__(kRowThreads)
rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const int64_t* __restrict__ ids, int S, int shift,
                   const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ out, int64_t ldo, int H, float eps,
                   float* __restrict__ rstd_out) {
    __shared__ float red[32];
    const int64_t r = blockIdx.x;
    const __nv_bfloat16* xr;
    if (ids) {
        const int64_t b = r / S, s = r % S;
        const int64_t tok = (s + shift < S) ? ids[b * S + s + shift] : 0;
        xr = x + tok * ldx;
    } else {
        xr = x + r * ldx;
    }
    const int nchunks = H / 8;
    float v[kMaxChunks][8];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
        const int ch = threadIdx.x + c * kRowThreads;
        if (ch < nchunks) {
            bf16x8 t; t.u = __ldg(reinterpret_cast<const uint4*>(xr) + ch);
            t.unpack(v[c]);
#pragma unroll
            for (int i = 0; i < 8; ++i) ss += v[c][i] * v[c][i];
        }
    }
    ss = block_sum(ss, red);
    const float rstd = rsqrtf(ss / (float)H + eps);
    if (rstd_out && threadIdx.x == 0) rstd_out[r] = rstd;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
        const int ch = threadIdx.x + c * kRowThreads;
        if (ch < nchunks) {
            bf16x8 wt; wt.u = __ldg(reinterpret_cast<const uint4*>(w) + ch);
            float wf[8], o[8];
            wt.unpack(wf);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = wf[i] * bf16_round(v[c][i] * rstd);
            bf16x8 ot; ot.pack(o);
            reinterpret_cast<uint4*>(out + r * ldo)[ch] = ot.u;
        }
    }
}

// ------------------------------------------------------------------ RMSNorm backward
// y = w * xhat, xhat = x * rstd.  dx = rstd * (g - xhat * mean(g * xhat)), g = w * dy.
// dx_out = dx + add1 + add2 (optional residual-path gradients), dw[c] += sum_r dy * xhat (fp32 atomics).
// gather mode (ids != null): x rows come from the frozen embedding table, only dw is produced.
// One block per row (looping), kChunks 16-byte chunks per thread, ONE block reduction per row
// (sum x^2 and sum g*x together), so several blocks stay resident per SM and hide the HBM latency.
constexpr int kRmsStages = 3;
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int kChunks, int kThreads>
__global__ void __launch_bounds__(kThreads, kChunks == 1 ? 1024 / kThreads : 1)   // 64 registers: two 512-thread blocks per SM
rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const int64_t* __restrict__ ids, int S, int shift,
                   const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ dy, int64_t lddy,
                   const __nv_bfloat16* __restrict__ add1, const __nv_bfloat16* __restrict__ add2,
                   __nv_bfloat16* __restrict__ dx, float* __restrict__ dw_partial, int64_t M, int H, float eps) {
    __shared__ float2 red[32];
    const int nchunks = H / 8;
    float dwacc[kChunks][8];
    float wf[kChunks][8];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const int ch = threadIdx.x + c * kThreads;
#pragma unroll
        for (int i = 0; i < 8; ++i) { dwacc[c][i] = 0.f; wf[c][i] = 0.f; }
        if (ch < nchunks) {
            bf16x8 wt; wt.u = __ldg(reinterpret_cast<const uint4*>(w) + ch);
            wt.unpack(wf[c]);
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Row pipeline: every thread streams ITS OWN 16-byte chunks of the next kRmsStages - 1 rows (x, dy and the residual-gradient
    // addends) into private shared-memory slots with cp.async — no registers held for loads in flight, no cross-thread
    // synchronisation (a thread only reads back what it copied).  One row at a time left two dependent round trips per row and
    // ~0.45 of HBM bandwidth; register prefetch of one row reached 0.50 (profiles/r02_ncu_step_nongemm.csv).
    extern __shared__ uint4 rms_stage[];                 // [kRmsStages][4 streams][kChunks][kThreads]
    const bool has_add1 = dx && add1, has_add2 = dx && add2;
    auto slot = [&](int stage, int stream, int c) { return rms_stage + ((stage * 4 + stream) * kChunks + c) * kThreads + threadIdx.x; };
    auto issue = [&](int64_t r, int stage) {
        if (r < M) {
            const __nv_bfloat16* xr;
            if (ids) {
                const int64_t b = r / S, s = r % S;
                const int64_t tok = (s + shift < S) ? ids[b * S + s + shift] : 0;
                xr = x + tok * ldx;
            } else {
                xr = x + r * ldx;
            }
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const int ch = threadIdx.x + c * kThreads;
                if (ch < nchunks) {
                    cp_async_16(slot(stage, 0, c), reinterpret_cast<const uint4*>(xr) + ch);
                    cp_async_16(slot(stage, 1, c), reinterpret_cast<const uint4*>(dy + r * lddy) + ch);
                    if (has_add1) cp_async_16(slot(stage, 2, c), reinterpret_cast<const uint4*>(add1 + r * (int64_t)H) + ch);
                    if (has_add2) cp_async_16(slot(stage, 3, c), reinterpret_cast<const uint4*>(add2 + r * (int64_t)H) + ch);
                }
            }
        }
        cp_async_commit();                               // one group per row, empty past the end: the wait below stays uniform
    };
#pragma unroll
    for (int s = 0; s < kRmsStages - 1; ++s) issue(blockIdx.x + (int64_t)s * gridDim.x, s);
    int stage = 0;
    for (int64_t r = blockIdx.x; r < M; r += gridDim.x) {
        {
            int nstage = stage + kRmsStages - 1;
            if (nstage >= kRmsStages) nstage -= kRmsStages;
            issue(r + (int64_t)(kRmsStages - 1) * gridDim.x, nstage);     // the slot row r - gridDim.x used: consumed last iteration
        }
        cp_async_wait<kRmsStages - 1>();                 // row r's group has landed
        float xv[kChunks][8], gv[kChunks][8];
        uint4 a1c[kChunks], a2c[kChunks];
        float ss = 0.f, gx = 0.f;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const int ch = threadIdx.x + c * kThreads;
            if (ch < nchunks) {
                bf16x8 t; t.u = *slot(stage, 0, c);
                t.unpack(xv[c]);
                bf16x8 d; d.u = *slot(stage, 1, c);
                d.unpack(gv[c]);
                if (has_add1) a1c[c] = *slot(stage, 2, c);
                if (has_add2) a2c[c] = *slot(stage, 3, c);
#pragma unroll
                for (int i = 0; i < 8; ++i) { ss += xv[c][i] * xv[c][i]; gx += gv[c][i] * wf[c][i] * xv[c][i]; }
            }
        }
        if (++stage == kRmsStages) stage = 0;
        ss = warp_sum(ss); gx = warp_sum(gx);
        __syncthreads();
        if (lane == 0) red[warp] = make_float2(ss, gx);
        __syncthreads();
        {
            float2 t = (lane < kThreads / 32) ? red[lane] : make_float2(0.f, 0.f);
            ss = warp_sum(t.x); gx = warp_sum(t.y);
        }
        const float rstd = rsqrtf(ss / (float)H + eps);
        const float mean = gx * rstd / (float)H;   // mean(g * xhat)
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const int ch = threadIdx.x + c * kThreads;
            if (ch < nchunks) {
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float xh = xv[c][i] * rstd;
                    dwacc[c][i] += gv[c][i] * bf16_round(xh);
                    o[i] = rstd * (gv[c][i] * wf[c][i] - xh * mean);
                }
                if (dx) {
                    if (has_add1) {
                        bf16x8 a; a.u = a1c[c];
                        float af[8]; a.unpack(af);
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] += af[i];
                    }
                    if (has_add2) {
                        bf16x8 a; a.u = a2c[c];
                        float af[8]; a.unpack(af);
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] += af[i];
                    }
                    bf16x8 ot; ot.pack(o);
                    reinterpret_cast<uint4*>(dx + r * (int64_t)H)[ch] = ot.u;
                }
            }
        }
    }
    // per-block partial column sums; a second kernel adds them in block order (deterministic, no atomics)
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const int ch = threadIdx.x + c * kThreads;
        if (ch < nchunks) {
            float4* dst = reinterpret_cast<float4*>(dw_partial + (int64_t)blockIdx.x * H + ch * 8);
            dst[0] = make_float4(dwacc[c][0], dwacc[c][1], dwacc[c][2], dwacc[c][3]);
            dst[1] = make_float4(dwacc[c][4], dwacc[c][5], dwacc[c][6], dwacc[c][7]);
        }
    }
}
// dw[c] += sum_b partial[b][c]   (fixed summation order: 8 interleaved row groups, then an ordered smem add)
__global__ void __launch_bounds__(256) colsum_partials_kernel(const float* __restrict__ partial, int nblocks, int H, float* __restrict__ dw) {
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    float acc = 0.f;
    if (c < H)
        for (int b = ry; b < nblocks; b += 8) acc += partial[(int64_t)b * H + c];
    red[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && c < H) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][cx];
        dw[c] += t;
    }
}

// ------------------------------------------------------------------ RoPE (in place on the fused qkv buffer)
// x' = x*cos + rotate_half(x)*sin at position s + pos_offset; cos/sin are the reference's bf16 tables
// (llama3_eagle.py:133-142,296-301).  inverse=1 applies the transpose rotation (gradient).
// Optionally converts fp32 gradient accumulators (src32) to bf16 on the way (used for dK).
__global__ void __launch_bounds__(256)
rope_kernel(__nv_bfloat16* __restrict__ x, const float* __restrict__ src32, int64_t ld, int64_t ld32, int n_heads,
            int head_dim, const __nv_bfloat16* __restrict__ cos_t, const __nv_bfloat16* __restrict__ sin_t, int S,
            int pos_offset, int64_t M, int inverse) {
    const int half = head_dim / 2;
    const int64_t total = M * n_heads * (half / 8);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = i % (half / 8);
        const int h = (i / (half / 8)) % n_heads;
        const int64_t r = i / ((int64_t)(half / 8) * n_heads);
        const int pos = (int)(r % S) + pos_offset;
        float x1[8], x2[8], c[8], s[8];
        __nv_bfloat16* p1 = x + r * ld + h * head_dim + c8 * 8;
        __nv_bfloat16* p2 = p1 + half;
        if (src32) {
            const float* q1 = src32 + r * ld32 + h * head_dim + c8 * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) { x1[k] = q1[k]; x2[k] = q1[half + k]; }
        } else {
            bf16x8 a; a.u = *reinterpret_cast<const uint4*>(p1); a.unpack(x1);
            bf16x8 b; b.u = *reinterpret_cast<const uint4*>(p2); b.unpack(x2);
        }
        bf16x8 ct; ct.u = __ldg(reinterpret_cast<const uint4*>(cos_t + (int64_t)pos * head_dim + c8 * 8)); ct.unpack(c);
        bf16x8 st; st.u = __ldg(reinterpret_cast<const uint4*>(sin_t + (int64_t)pos * head_dim + c8 * 8)); st.unpack(s);
        float o1[8], o2[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (!inverse) { o1[k] = x1[k] * c[k] - x2[k] * s[k]; o2[k] = x2[k] * c[k] + x1[k] * s[k]; }
            else          { o1[k] = x1[k] * c[k] + x2[k] * s[k]; o2[k] = x2[k] * c[k] - x1[k] * s[k]; }
        }
        bf16x8 a, b; a.pack(o1); b.pack(o2);
        *reinterpret_cast<uint4*>(p1) = a.u;
        *reinterpret_cast<uint4*>(p2) = b.u;
    }
}

// ------------------------------------------------------------------ fp32 -> bf16 strided convert (dV accumulators)
__global__ void __launch_bounds__(256)
cvt_f32_bf16_kernel(const float* __restrict__ src, int64_t lds, __nv_bfloat16* __restrict__ dst, int64_t ldd, int64_t M,
                    int cols, float scale) {
    const int c8n = cols / 8;
    const int64_t total = M * c8n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / c8n;
        const int c = (int)(i % c8n) * 8;
        const float4 a = *reinterpret_cast<const float4*>(src + r * lds + c);
        const float4 b = *reinterpret_cast<const float4*>(src + r * lds + c + 4);
        float f[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b.x * scale, b.y * scale, b.z * scale, b.w * scale};
        bf16x8 o; o.pack(f);
        *reinterpret_cast<uint4*>(dst + r * ldd + c) = o.u;
    }
}

// ------------------------------------------------------------------ SwiGLU  (llama3_eagle.py:1547)
// act = bf16(bf16(silu(g)) * u) with g,u the bf16 gate/up GEMM outputs, gu = [M, 2I] (gate | up).
__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ gu, __nv_bfloat16* __restrict__ act, int64_t M, int I) {
    const int c8n = I / 8;
    const int64_t total = M * c8n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / c8n;
        const int c = (int)(i % c8n) * 8;
        bf16x8 g, u; float gf[8], uf[8], o[8];
        g.u = __ldg(reinterpret_cast<const uint4*>(gu + r * 2 * I + c)); g.unpack(gf);
        u.u = __ldg(reinterpret_cast<const uint4*>(gu + r * 2 * I + I + c)); u.unpack(uf);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = bf16_round(gf[k] / (1.f + __expf(-gf[k]))) * uf[k];
        bf16x8 ot; ot.pack(o);
        *reinterpret_cast<uint4*>(act + r * I + c) = ot.u;
    }
}
// d_gu = [d_act * u * silu'(g) | d_act * silu(g)]
__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ gu, const __nv_bfloat16* __restrict__ dact,
                  __nv_bfloat16* __restrict__ dgu, int64_t M, int I) {
    const int c8n = I / 8;
    const int64_t total = M * c8n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / c8n;
        const int c = (int)(i % c8n) * 8;
        bf16x8 g, u, d; float gf[8], uf[8], df[8], og[8], ou[8];
        g.u = __ldg(reinterpret_cast<const uint4*>(gu + r * 2 * I + c)); g.unpack(gf);
        u.u = __ldg(reinterpret_cast<const uint4*>(gu + r * 2 * I + I + c)); u.unpack(uf);
        d.u = __ldg(reinterpret_cast<const uint4*>(dact + r * I + c)); d.unpack(df);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float sg = 1.f / (1.f + __expf(-gf[k]));
            const float silu = gf[k] * sg;
            og[k] = df[k] * uf[k] * (sg * (1.f + gf[k] * (1.f - sg)));
            ou[k] = df[k] * silu;
        }
        bf16x8 a, b; a.pack(og); b.pack(ou);
        *reinterpret_cast<uint4*>(dgu + r * 2 * I + c) = a.u;
        *reinterpret_cast<uint4*>(dgu + r * 2 * I + I + c) = b.u;
    }
}

// ------------------------------------------------------------------ sequence left-shift (TargetHead.preprocess,
// modeling/target/target_head.py:103-108 / utils.py:128-135): dst[b, s] = src[b, s+1], dst[b, S-1] = 0
__global__ void __launch_bounds__(256)
shift_left_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t B, int S, int H) {
    const int c8n = H / 8;
    const int64_t total = B * S * c8n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / c8n;
        const int c = (int)(i % c8n);
        const int s = (int)(r % S);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (s + 1 < S) v = __ldg(reinterpret_cast<const uint4*>(src + (r + 1) * H) + c);
        reinterpret_cast<uint4*>(dst + r * H)[c] = v;
    }
}

// out = a (+ b)   (bf16, n % 8 == 0)
__global__ void __launch_bounds__(256)
add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ out, int64_t n8) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        bf16x8 x; x.u = __ldg(reinterpret_cast<const uint4*>(a) + i);
        if (b) {
            bf16x8 y; y.u = __ldg(reinterpret_cast<const uint4*>(b) + i);
            float xf[8], yf[8]; x.unpack(xf); y.unpack(yf);
#pragma unroll
            for (int k = 0; k < 8; ++k) xf[k] += yf[k];
            x.pack(xf);
        }
        reinterpret_cast<uint4*>(out)[i] = x.u;
    }
}

// out[i, :] = table[ids[i], :]   (frozen embedding lookup; ids outside [0, V) give zeros)
__global__ void __launch_bounds__(256)
embedding_gather_kernel(const __nv_bfloat16* __restrict__ table, int64_t V, int H, const int64_t* __restrict__ ids, int64_t n,
                        __nv_bfloat16* __restrict__ out) {
    const int c8n = H / 8;
    const int64_t total = n * c8n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / c8n;
        const int c = (int)(i % c8n);
        const int64_t tok = ids[r];
        uint4 v = make_uint4(0, 0, 0, 0);
        if (tok >= 0 && tok < V) v = __ldg(reinterpret_cast<const uint4*>(table + tok * H) + c);
        reinterpret_cast<uint4*>(out + r * H)[c] = v;
    }
}

static inline int grid_for(int64_t total, int threads = 256, int max_blocks = 148 * 16) {
    int64_t b = (total + threads - 1) / threads;
    if (b > max_blocks) b = max_blocks;
    if (b < 1) b = 1;
    return (int)b;
}

int rmsnorm_fwd(const void* x, int64_t ldx, const int64_t* ids, int S, int shift, const void* w, void* out, int64_t ldo,
                int64_t M, int H, float eps, float* rstd, cudaStream_t st) {
    if (H % 8 || H > kRowThreads * kMaxChunks * 8) return set_error(-22, "rmsnorm: H=%d must be a multiple of 8 and <= 8192", H);
    rmsnorm_fwd_kernel<<<(unsigned)M, kRowThreads, 0, st>>>((const __nv_bfloat16*)x, ldx, ids, S, shift,
                                                           (const __nv_bfloat16*)w, (__nv_bfloat16*)out, ldo, H, eps, rstd);
    SF_CUDA_CHECK_LAUNCH("rmsnorm_fwd");
    return 0;
}
int64_t rmsnorm_bwd_ws_bytes(int H) { return (int64_t)148 * 12 * H * 4; }

int rmsnorm_bwd(const void* x, int64_t ldx, const int64_t* ids, int S, int shift, const void* w, const void* dy,
                int64_t lddy, const void* add1, const void* add2, void* dx, float* dw, float* partial_ws, int64_t M, int H,
                float eps, cudaStream_t st) {
    if (!partial_ws) return set_error(-22, "rmsnorm_bwd: partial-sum workspace missing");
    int launched_blocks = 0;
    if (H % 8 || H > kRowThreads * kMaxChunks * 8) return set_error(-22, "rmsnorm: H=%d must be a multiple of 8 and <= 8192", H);
    const int nch = H / 8;
#define SF_RMS_BWD(CH, TH, BPS)                                                                                        \
    do {                                                                                                               \
        int blocks = 148 * BPS;                  /* one resident wave: BPS blocks per SM fit (registers and staging) */  \
        if (blocks > M) blocks = (int)M;                                                                               \
        launched_blocks = blocks;                                                                                      \
        const int smem = kRmsStages * 4 * CH * TH * 16;                                                                \
        static bool attr_set = false;                                                                                  \
        if (!attr_set) {                                                                                               \
            cudaError_t e = cudaFuncSetAttribute(rmsnorm_bwd_kernel<CH, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
            if (e != cudaSuccess) return set_error(-22, "rmsnorm_bwd smem attr: %s", cudaGetErrorString(e));           \
            attr_set = true;                                                                                           \
        }                                                                                                              \
        rmsnorm_bwd_kernel<CH, TH><<<blocks, TH, smem, st>>>((const __nv_bfloat16*)x, ldx, ids, S, shift,              \
            (const __nv_bfloat16*)w, (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)add1,                        \
            (const __nv_bfloat16*)add2, (__nv_bfloat16*)dx, partial_ws, M, H, eps);                                    \
    } while (0)
    if (nch <= 128) SF_RMS_BWD(1, 128, 8);
    else if (nch <= 256) SF_RMS_BWD(1, 256, 4);
    else if (nch <= 512) SF_RMS_BWD(1, 512, 2);
    else SF_RMS_BWD(2, 512, 1);
#undef SF_RMS_BWD
    SF_CUDA_CHECK_LAUNCH("rmsnorm_bwd");
    colsum_partials_kernel<<<(H + 31) / 32, 256, 0, st>>>(partial_ws, launched_blocks, H, dw);
    SF_CUDA_CHECK_LAUNCH("rmsnorm_bwd_colsum");
    return 0;
}
int rope(void* x, const float* src32, int64_t ld, int64_t ld32, int n_heads, int head_dim, const void* cos_t,
         const void* sin_t, int S, int pos_offset, int64_t M, int inverse, cudaStream_t st) {
    if (head_dim % 16) return set_error(-22, "rope: head_dim=%d must be a multiple of 16", head_dim);
    const int64_t total = M * n_heads * (head_dim / 16);
    rope_kernel<<<grid_for(total), 256, 0, st>>>((__nv_bfloat16*)x, src32, ld, ld32, n_heads, head_dim,
                                                (const __nv_bfloat16*)cos_t, (const __nv_bfloat16*)sin_t, S, pos_offset, M,
                                                inverse);
    SF_CUDA_CHECK_LAUNCH("rope");
    return 0;
}
int cvt_f32_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t M, int cols, float scale, cudaStream_t st) {
    if (cols % 8) return set_error(-22, "cvt: cols=%d must be a multiple of 8", cols);
    cvt_f32_bf16_kernel<<<grid_for(M * (cols / 8)), 256, 0, st>>>(src, lds, (__nv_bfloat16*)dst, ldd, M, cols, scale);
    SF_CUDA_CHECK_LAUNCH("cvt_f32_bf16");
    return 0;
}
int swiglu_fwd(const void* gu, void* act, int64_t M, int I, cudaStream_t st) {
    if (I % 8) return set_error(-22, "swiglu: I=%d must be a multiple of 8", I);
    swiglu_fwd_kernel<<<grid_for(M * (I / 8)), 256, 0, st>>>((const __nv_bfloat16*)gu, (__nv_bfloat16*)act, M, I);
    SF_CUDA_CHECK_LAUNCH("swiglu_fwd");
    return 0;
}
int swiglu_bwd(const void* gu, const void* dact, void* dgu, int64_t M, int I, cudaStream_t st) {
    if (I % 8) return set_error(-22, "swiglu: I=%d must be a multiple of 8", I);
    swiglu_bwd_kernel<<<grid_for(M * (I / 8)), 256, 0, st>>>((const __nv_bfloat16*)gu, (const __nv_bfloat16*)dact,
                                                            (__nv_bfloat16*)dgu, M, I);
    SF_CUDA_CHECK_LAUNCH("swiglu_bwd");
    return 0;
}
int embedding_gather(const void* table, int64_t V, int H, const int64_t* ids, int64_t n, void* out, cudaStream_t st) {
    if (H % 8) return set_error(-22, "embedding: H=%d must be a multiple of 8", H);
    if (n <= 0) return 0;
    embedding_gather_kernel<<<grid_for(n * (H / 8)), 256, 0, st>>>((const __nv_bfloat16*)table, V, H, ids, n, (__nv_bfloat16*)out);
    SF_CUDA_CHECK_LAUNCH("embedding_gather");
    return 0;
}
int add_bf16(const void* a, const void* b, void* out, int64_t n, cudaStream_t st) {
    if (n % 8) return set_error(-22, "add: n=%lld must be a multiple of 8", (long long)n);
    add_bf16_kernel<<<grid_for(n / 8), 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)out, n / 8);
    SF_CUDA_CHECK_LAUNCH("add_bf16");
    return 0;
}
int shift_left(const void* src, void* dst, int64_t B, int S, int H, cudaStream_t st) {
    if (H % 8) return set_error(-22, "shift: H=%d must be a multiple of 8", H);
    shift_left_kernel<<<grid_for(B * S * (H / 8)), 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, B, S, H);
    SF_CUDA_CHECK_LAUNCH("shift_left");
    return 0;
}

}  // namespace sf
