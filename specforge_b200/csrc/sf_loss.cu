// sf_loss.cu — teacher distribution, fused soft-label CE / acceptance / top-1 loss (+ in-place gradient),
// metric reduction, and the fused clip + AdamW optimizer step.  All HBM-bound row kernels.
#include "sf_host.h"
#include "sf_ptx.cuh"

namespace sf {

struct MaxIdx { float v; int i; };
__device__ __forceinline__ MaxIdx better(MaxIdx a, MaxIdx b) {
    // first index wins ties (torch.argmax semantics)
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ MaxIdx warp_argmax(MaxIdx a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MaxIdx b{__shfl_xor_sync(0xffffffffu, a.v, o), __shfl_xor_sync(0xffffffffu, a.i, o)};
        a = better(a, b);
    }
    return a;
}
__device__ __forceinline__ MaxIdx block_argmax(MaxIdx a, float* redv, int* redi) {
    a = warp_argmax(a);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) { redv[w] = a.v; redi[w] = a.i; }
    __syncthreads();
    MaxIdx t{(l < nw) ? redv[l] : -INFINITY, (l < nw) ? redi[l] : 0x7fffffff};
    return warp_argmax(t);
}
__device__ __forceinline__ float block_sum_f(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < nw) ? red[l] : 0.f;
    return warp_sum(t);
}
__device__ __forceinline__ float block_max_f(float v, float* red) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < nw) ? red[l] : -INFINITY;
    return warp_max(t);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
        f[2 * i] = __bfloat162float(b.x);
        f[2 * i + 1] = __bfloat162float(b.y);
    }
}

// ------------------------------------------------------------------ teacher
// Reference: algorithms/eagle3/model.py:487-501 (_compute_target_p) + :445-484 (padding).
// Per (b, s) row: argmax / logsumexp over the full target vocab, gather of the draft-vocab logits, softmax statistics
// over the draft vocab.  The reference materialises two fp32 [B, S, DV] tensors (target_p and p_on_draft); both are
// pure functions of the gathered bf16 logits x and three row scalars, so this kernel stores only
//     xg[row, i]   = x_i (bf16, lossless)                         tstats[row] = {md, inv, cs, 0}
//     target_p_i   = exp(x_i - md) * inv          (softmax over the draft vocab, fp32, as the reference computes it)
//     p_on_draft_i = exp(x_i - lse_full) = target_p_i * cs,   cs = dd * exp(md - lse_full)
// and the loss kernel re-evaluates them: 64 KB of HBM per row instead of 256 KB.  Rows live in the [B, S+T, DV] padded
// layout so that TTT step j reads row (b, s + j) without the reference's per-step `.contiguous()` copies.
//
// Persistent, one 512-thread block per SM, no dynamic smem and <= 72 registers/thread: the block fits NEXT TO a resident
// tcgen05 GEMM CTA (256 threads x 112 registers, 197 KB smem), which is what lets it run on the side stream under the
// draft's first GEMMs instead of serialising with them.
__global__ void __launch_bounds__(512, 2)
teacher_kernel(const __nv_bfloat16* __restrict__ tl, int64_t ld, const int* __restrict__ d2t_idx,
               const uint8_t* __restrict__ t2d, const int* __restrict__ loss_mask, __nv_bfloat16* __restrict__ xg,
               float4* __restrict__ tstats, int64_t* __restrict__ ids, int* __restrict__ position_mask, int S, int T, int V,
               int DV, int64_t row0, int64_t nrows) {
    __shared__ float redv[32];
    __shared__ int redi[32];
    const int nch = V / 8;
    for (int64_t r = row0 + blockIdx.x; r < row0 + nrows; r += gridDim.x) {
        const int b = (int)(r / S), s = (int)(r % S);
        const __nv_bfloat16* row = tl + r * ld;
        // pass 1: one streaming pass over the full vocab with an online max / sum-exp (and first-index argmax)
        MaxIdx mi{-INFINITY, 0x7fffffff};
        float d = 0.f;
#pragma unroll 4
        for (int c = threadIdx.x; c < nch; c += 512) {
            float f[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(row) + c), f);
            float cm = f[0]; int ci = 0;
#pragma unroll
            for (int e = 1; e < 8; ++e)
                if (f[e] > cm) { cm = f[e]; ci = e; }
            if (cm > mi.v) { d *= __expf(mi.v - cm); mi.v = cm; mi.i = c * 8 + ci; }
#pragma unroll
            for (int e = 0; e < 8; ++e) d += __expf(f[e] - mi.v);
        }
        for (int e = nch * 8 + threadIdx.x; e < V; e += 512) {
            const float f = __bfloat162float(row[e]);
            if (f > mi.v) { d *= __expf(mi.v - f); mi.v = f; mi.i = e; }
            d += __expf(f - mi.v);
        }
        const float my_m = mi.v;
        mi = block_argmax(mi, redv, redi);
        const float m = mi.v;
        d = block_sum_f(my_m == -INFINITY ? 0.f : d * __expf(my_m - m), redv);
        const float lse = m + logf(d);
        // pass 2: gather the draft-vocab logits (row is L2-hot), store them, online max / sum-exp over the draft vocab
        const int64_t orow = (int64_t)b * (S + T) + s;
        __nv_bfloat16* xo = xg + orow * DV;
        float md = -INFINITY, dd = 0.f;
#pragma unroll 4
        for (int i = 2 * threadIdx.x; i < DV; i += 1024) {
            const __nv_bfloat16 a = row[__ldg(d2t_idx + i)];
            const float fa = __bfloat162float(a);
            if (i + 1 < DV) {
                const __nv_bfloat16 bq = row[__ldg(d2t_idx + i + 1)];
                const float fb = __bfloat162float(bq);
                __nv_bfloat162 pr; pr.x = a; pr.y = bq;
                if ((DV & 1) == 0) *reinterpret_cast<__nv_bfloat162*>(xo + i) = pr;
                else { xo[i] = a; xo[i + 1] = bq; }
                const float cm = fmaxf(fa, fb);
                if (cm > md) { dd *= __expf(md - cm); md = cm; }
                dd += __expf(fa - md) + __expf(fb - md);
            } else {
                xo[i] = a;
                if (fa > md) { dd *= __expf(md - fa); md = fa; }
                dd += __expf(fa - md);
            }
        }
        const float my_md = md;
        md = block_max_f(md, redv);
        dd = block_sum_f(my_md == -INFINITY ? 0.f : dd * __expf(my_md - md), redv);
        if (threadIdx.x == 0) {
            tstats[orow] = make_float4(md, 1.f / dd, dd * __expf(md - lse), 0.f);
            ids[orow] = mi.i;
            position_mask[r] = (t2d[mi.i] ? 1 : 0) * loss_mask[r];
        }
    }
}

// padded tail rows: target_p = 1/DV, p_on_draft = 0, ids = 0   (eagle3/model.py:459-477)
__global__ void __launch_bounds__(256)
teacher_pad_kernel(__nv_bfloat16* __restrict__ xg, float4* __restrict__ tstats, int64_t* __restrict__ ids, int B, int S,
                   int T, int DV) {
    const int64_t total = (int64_t)B * T * DV;
    const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pr = i / DV;
        const int c = (int)(i % DV);
        const int b = (int)(pr / T), t = (int)(pr % T);
        const int64_t orow = (int64_t)b * (S + T) + S + t;
        xg[orow * DV + c] = z;
        if (c == 0) { ids[orow] = 0; tstats[orow] = make_float4(0.f, 1.0f / (float)DV, 0.f, 0.f); }
    }
}

// ---- teacher statistics fused into the target-head GEMM epilogue (EPI_TEACHER, sf_gemm.cuh) ----
// t2d (uint8 [V]) -> bit words + exclusive prefix counts per 32-column word: the epilogue appends the flagged columns of a
// chunk at xg[row, prefix[word] ...] (vocabulary order = the order of the reference's boolean-mask gather, model.py:492).
__global__ void __launch_bounds__(1024) t2d_index_kernel(const uint8_t* __restrict__ t2d, int V, uint32_t* __restrict__ bits,
                                                         int* __restrict__ prefix) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int nwords = (V + 31) / 32;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nwords; base += 1024) {
        const int w = base + threadIdx.x;
        uint32_t b = 0;
        if (w < nwords)
            for (int e = 0; e < 32; ++e) { const int c = w * 32 + e; if (c < V && t2d[c]) b |= 1u << e; }
        const int cnt = __popc(b);
        int incl = cnt;                                  // inclusive scan inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        int woff = 0;
        for (int i = 0; i < (int)(threadIdx.x >> 5); ++i) woff += wsum[i];
        const int carry = carry_s;
        if (w < nwords) { bits[w] = b; prefix[w] = carry + woff + incl - cnt; }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + woff + incl;
        __syncthreads();
    }
}
// One WARP per row: lane l combines the per-n-block partials l, l + 32, ... in ascending order (strict > keeps the first index), the
// lanes are merged with the (max, smaller index) rule — torch.argmax's first-index tie-break — and lane 0 writes what teacher_kernel
// writes: tstats = {md, 1/dd, dd * exp(md - lse_full), 0}, ids, position_mask.  (One thread per row walked ~1200 partials x 5 planes
// serially with only 512 warps on the whole GPU: 0.67 ms.)
__global__ void __launch_bounds__(256)
teacher_merge_kernel(const float* __restrict__ stats, int nb, int64_t M, const uint8_t* __restrict__ t2d,
                     const int* __restrict__ loss_mask, float4* __restrict__ tstats, int64_t* __restrict__ ids,
                     int* __restrict__ position_mask, int S, int T) {
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= M) return;                                  // warp-uniform
    const size_t plane = (size_t)nb * M;
    MaxIdx mi{-INFINITY, 0x7fffffff};
    float d = 0.f, md = -INFINITY, dd = 0.f;
    for (int b = lane; b < nb; b += 32) {
        const float* sp = stats + (size_t)r * nb + b;
        const float pm = __ldg(sp), pd = __ldg(sp + plane), pmd = __ldg(sp + 3 * plane), pdd = __ldg(sp + 4 * plane);
        if (pm > mi.v) { d = d * __expf(mi.v - pm) + pd; mi.v = pm; mi.i = __float_as_int(__ldg(sp + 2 * plane)); }
        else if (pm > -INFINITY) d += pd * __expf(pm - mi.v);
        if (pmd > md) { dd = dd * __expf(md - pmd) + pdd; md = pmd; }
        else if (pmd > -INFINITY) dd += pdd * __expf(pmd - md);
    }
    const float my_m = mi.v, my_md = md;
    mi = warp_argmax(mi);
    d = warp_sum(my_m == -INFINITY ? 0.f : d * __expf(my_m - mi.v));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) md = fmaxf(md, __shfl_xor_sync(0xffffffffu, md, o));
    dd = warp_sum(my_md == -INFINITY ? 0.f : dd * __expf(my_md - md));
    if (lane == 0) {
        const float lse = mi.v + logf(d);
        const int64_t orow = r + (r / S) * T;
        tstats[orow] = make_float4(md, 1.f / dd, dd * __expf(md - lse), 0.f);
        ids[orow] = mi.i;
        position_mask[r] = (t2d[mi.i] ? 1 : 0) * loss_mask[r];
    }
}

// ------------------------------------------------------------------ fused loss / metrics / gradient
// Reference: core/loss.py:15-21,49-170 (soft-label CE, mean over ALL rows, in-place backward),
// core/lk_loss.py:43-80 (acceptance = sum_v min(p_on_draft, softmax)), eagle3/model.py:161-173 (top-1).
// One block per row:
//   A  (max, sum-exp, first-index argmax) of the row — normally merged from the partials the lm_head GEMM epilogue left
//      (EPI_BF16_STATS: one partial per thread, one block reduction, the row is not read); without them an online pass over the row
//   B  logits row + the gathered bf16 teacher logits: target_p re-evaluated (see teacher_kernel),
//      soft-label CE sum, acceptance sum, and the gradient written over the logits row
//          d loss / d x_v = coef * (softmax(x)_v * sum_v(target_p) - target_p_v),   sum_v(target_p) = 1
// 192 KB of HBM traffic per 32000-wide row (64 read + 64 read + 64 written); no dynamic smem, <= 40 registers so three
// blocks fit per SM and the row-serial latency is hidden.
struct LossParams {
    __nv_bfloat16* logits; int64_t ld;
    const __nv_bfloat16* xg; const float4* tstats; const int64_t* tgt_ids;  // padded [B, S+T, ...]
    const int* position_mask; const int* loss_mask;                         // [B, S] (step-0 masks)
    const int64_t* d2t;
    int S, T, DV, step;
    int64_t M;
    float grad_coef;      // ploss_decay^step * upstream / M
    int write_grad;
    float* row_loss; float* row_accept; float* row_correct;  // [M]
    const float* stats; int stats_nb;   // optional pass-A partials from the lm_head GEMM epilogue (EPI_BF16_STATS): [3][M][nb]
};

__global__ void __launch_bounds__(512, 3) loss_kernel(LossParams p) {
    __shared__ float redv[32];
    __shared__ float redd[32];
    __shared__ int redi[32];
    const int nch = p.DV / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = blockIdx.x;
    const int b = (int)(r / p.S), s = (int)(r % p.S);
    const bool in_range = s + p.step < p.S;
    const int pm = in_range ? p.position_mask[(int64_t)b * p.S + s + p.step] : 0;
    const int lm = in_range ? p.loss_mask[(int64_t)b * p.S + s + p.step] : 0;
    uint4* grow = reinterpret_cast<uint4*>(p.logits + r * p.ld);
    if (pm == 0 && lm == 0) {  // nothing to measure: zero gradient row
        if (p.write_grad)
            for (int c = threadIdx.x; c < nch; c += 512) grow[c] = make_uint4(0, 0, 0, 0);
        if (threadIdx.x == 0) { p.row_loss[r] = 0.f; p.row_accept[r] = 0.f; p.row_correct[r] = 0.f; }
        return;
    }
    // pass A: online max / sum-exp, first-index argmax — either merged from the partials the lm_head GEMM epilogue left
    // (no read of the logits row at all) or computed here from the row
    MaxIdx mi{-INFINITY, 0x7fffffff};
    float d = 0.f;
    if (p.stats) {
        // every thread takes (at most) one of the row's partial blocks; the block reduction below merges them — a single warp
        // walking the 125-250 partials serially cost ~1.5 k cycles per row before the other 15 warps could start pass B
        for (int nb = threadIdx.x; nb < p.stats_nb; nb += 512) {
            const size_t plane = (size_t)p.stats_nb * p.M;
            const float* sp = p.stats + (size_t)r * p.stats_nb + nb;                // a row's partials are contiguous
            const float pm = __ldg(sp), pd = __ldg(sp + plane);
            if (pm > mi.v) { d = d * __expf(mi.v - pm) + pd; mi.v = pm; mi.i = __float_as_int(__ldg(sp + 2 * plane)); }
            else if (pm > -INFINITY) d += pd * __expf(pm - mi.v);
        }
    } else {
#pragma unroll 4
    for (int c = threadIdx.x; c < nch; c += 512) {
        float f[8];
        unpack8(grow[c], f);
        float cm = f[0]; int ci = 0;
#pragma unroll
        for (int e = 1; e < 8; ++e)
            if (f[e] > cm) { cm = f[e]; ci = e; }
        if (cm > mi.v) { d *= __expf(mi.v - cm); mi.v = cm; mi.i = c * 8 + ci; }
#pragma unroll
        for (int e = 0; e < 8; ++e) d += __expf(f[e] - mi.v);
    }
    }
    {   // one combined block reduction of (max, argmax, sum-exp); equal maxima: the smaller column index wins
        const float my_m = mi.v;
        MaxIdx wm = warp_argmax(mi);
        float wd = warp_sum(my_m == -INFINITY ? 0.f : d * __expf(my_m - wm.v));
        if (lane == 0) { redv[warp] = wm.v; redi[warp] = wm.i; redd[warp] = wd; }
        __syncthreads();
        MaxIdx t{(lane < 16) ? redv[lane] : -INFINITY, (lane < 16) ? redi[lane] : 0x7fffffff};
        const float td = (lane < 16) ? redd[lane] : 0.f;
        mi = warp_argmax(t);
        d = warp_sum(t.v == -INFINITY ? 0.f : td * __expf(t.v - mi.v));
        __syncthreads();   // red* are reused below
    }
    const float m = mi.v;
    const float inv_d = 1.f / d;
    const int64_t trow = (int64_t)b * (p.S + p.T) + s + p.step;
    float sum_px = 0.f, sum_min = 0.f;
    if (pm) {
        // pass B: soft-label CE sums, acceptance, gradient
        const uint4* tx = reinterpret_cast<const uint4*>(p.xg + trow * p.DV);
        const float4 ts = __ldg(p.tstats + trow);
        const float a = inv_d * p.grad_coef;
        const float tcoef = ts.y * p.grad_coef;
#pragma unroll 2
        for (int c = threadIdx.x; c < nch; c += 512) {
            float f[8], g[8], t[8];
            unpack8(grow[c], f);
            unpack8(__ldg(tx + c), g);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float ex = __expf(f[e] - m);
                const float et = __expf(g[e] - ts.x);
                const float tv = et * ts.y;
                sum_px += tv * f[e];
                sum_min += fminf(tv * ts.z, ex * inv_d);
                t[e] = ex * a - et * tcoef;
            }
            if (p.write_grad) {
                uint4 o;
                o.x = pack_bf16x2(t[0], t[1]); o.y = pack_bf16x2(t[2], t[3]);
                o.z = pack_bf16x2(t[4], t[5]); o.w = pack_bf16x2(t[6], t[7]);
                grow[c] = o;
            }
        }
        sum_px = warp_sum(sum_px);
        sum_min = warp_sum(sum_min);
        if (lane == 0) { redv[warp] = sum_px; redd[warp] = sum_min; }
        __syncthreads();
        if (warp == 0) {
            sum_px = warp_sum(lane < 16 ? redv[lane] : 0.f);
            sum_min = warp_sum(lane < 16 ? redd[lane] : 0.f);
        }
    } else if (p.write_grad) {
        for (int c = threadIdx.x; c < nch; c += 512) grow[c] = make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x == 0) {
        p.row_loss[r] = pm ? -(sum_px - (m + logf(d))) : 0.f;
        p.row_accept[r] = pm ? sum_min : 0.f;
        const int64_t pred = mi.i;
        p.row_correct[r] = (lm && (pred + p.d2t[pred] == p.tgt_ids[trow])) ? (float)lm : 0.f;
    }
}

// Deterministic reduction of the per-row values into the step's metric slots.
// metrics[step][0..7] = {ploss, acc_correct, acc_denom, acceptance_rate, accept_num, accept_den, loss_denom, kl_weight}
// ploss follows core/lk_loss.py:83-99: lk_type 0: the KL term; 1 ("lambda"): w*kl + (1-w)*(1-a), w = kl_scale*exp(-kl_decay*a);
// 2 ("alpha"): -mean_masked(log a_r).
__global__ void __launch_bounds__(1024)
metrics_reduce_kernel(const float* __restrict__ row_loss, const float* __restrict__ row_accept,
                      const float* __restrict__ row_correct, const int* __restrict__ position_mask,
                      const int* __restrict__ loss_mask, int B, int S, int step, int lk_type, float kl_scale, float kl_decay,
                      float* __restrict__ metrics) {
    __shared__ float red[32];
    const int64_t M = (int64_t)B * S;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f;
    for (int64_t r = threadIdx.x; r < M; r += blockDim.x) {
        const int s = (int)(r % S);
        a0 += row_loss[r];
        const float acc = row_accept[r];
        a1 += acc;
        a2 += row_correct[r];
        if (s + step < S) {
            const int pm = position_mask[r + step];
            a3 += (float)pm; a4 += (float)loss_mask[r + step];
            if (pm && acc > 0.f) a5 += logf(acc);
        }
    }
    a0 = block_sum_f(a0, red); a1 = block_sum_f(a1, red); a2 = block_sum_f(a2, red);
    a3 = block_sum_f(a3, red); a4 = block_sum_f(a4, red); a5 = block_sum_f(a5, red);
    if (threadIdx.x == 0) {
        float* o = metrics + step * 8;
        const float kl = a0 / (float)M;
        const float den = fmaxf(a3, 1e-8f);
        const float rate = a1 / den;
        float ploss = kl, w = 0.f;
        if (lk_type == 1) { w = kl_scale * expf(-kl_decay * rate); ploss = w * kl + (1.f - w) * (1.f - rate); }
        else if (lk_type == 2) { ploss = -(a5 / den); }
        o[0] = ploss;
        o[1] = a2;
        o[2] = fmaxf(a4, 1e-6f);
        o[3] = rate;
        o[4] = a1;
        o[5] = den;
        o[6] = (float)M;
        o[7] = w;
    }
}

// LK-loss gradient (second pass, only when lk_loss_type is set): d ploss / d logits written in place.
//   a_r = sum_v min(pod_v, q_v),  d a_r / d x_v = q_v (1[q_v < pod_v] - s_r),  s_r = sum_u 1[q_u < pod_u] q_u
//   lambda: g = w * dKL + -(1-w) * pm/den * d a_r        alpha: g = -pm / (den * a_r) * d a_r
__global__ void __launch_bounds__(512) lk_grad_kernel(LossParams p, int lk_type, float step_weight,
                                                      const float* __restrict__ metrics) {
    __shared__ float redv[32];
    const int64_t r = blockIdx.x;
    const int b = (int)(r / p.S), s = (int)(r % p.S);
    const bool in_range = s + p.step < p.S;
    const int pm = in_range ? p.position_mask[(int64_t)b * p.S + s + p.step] : 0;
    uint4* grow = reinterpret_cast<uint4*>(p.logits + r * p.ld);
    const int nch = p.DV / 8;
    if (pm == 0) {
        for (int c = threadIdx.x; c < nch; c += blockDim.x) grow[c] = make_uint4(0, 0, 0, 0);
        return;
    }
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < nch; c += blockDim.x) {
        float f[8];
        unpack8(grow[c], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) mx = fmaxf(mx, f[e]);
    }
    const float m = block_max_f(mx, redv);
    float d = 0.f;
    for (int c = threadIdx.x; c < nch; c += blockDim.x) {
        float f[8];
        unpack8(grow[c], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) d += __expf(f[e] - m);
    }
    d = block_sum_f(d, redv);
    const float inv_d = 1.f / d;
    const int64_t trow = (int64_t)b * (p.S + p.T) + s + p.step;
    const uint4* tx = reinterpret_cast<const uint4*>(p.xg + trow * p.DV);
    const float4 ts = __ldg(p.tstats + trow);
    float s_r = 0.f;
    for (int c = threadIdx.x; c < nch; c += blockDim.x) {
        float f[8], g[8];
        unpack8(grow[c], f);
        unpack8(__ldg(tx + c), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float q = __expf(f[e] - m) * inv_d;
            const float pod = __expf(g[e] - ts.x) * ts.y * ts.z;
            if (q < pod) s_r += q;
        }
    }
    s_r = block_sum_f(s_r, redv);
    const float* mt = metrics + p.step * 8;
    const float den = mt[5], w = mt[7];
    const float a_r = p.row_accept[r];
    float ck, ca;
    if (lk_type == 1) { ck = w * step_weight / (float)p.M; ca = -(1.f - w) * step_weight / den; }
    else { ck = 0.f; ca = (a_r > 0.f) ? -step_weight / (den * a_r) : 0.f; }
    for (int c = threadIdx.x; c < nch; c += blockDim.x) {
        float f[8], g[8], o[8];
        unpack8(grow[c], f);
        unpack8(__ldg(tx + c), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float q = __expf(f[e] - m) * inv_d;
            const float tv = __expf(g[e] - ts.x) * ts.y;
            o[e] = ck * (q - tv) + ca * q * ((q < tv * ts.z ? 1.f : 0.f) - s_r);
        }
        uint4 ou;
        ou.x = pack_bf16x2(o[0], o[1]); ou.y = pack_bf16x2(o[2], o[3]); ou.z = pack_bf16x2(o[4], o[5]); ou.w = pack_bf16x2(o[6], o[7]);
        grow[c] = ou;
    }
}

// ------------------------------------------------------------------ optimizer (optimizer.py:95-168)
// ||g||^2 over the flat bf16 gradient buffer (fp32 accumulation), two deterministic stages.
__global__ void __launch_bounds__(512)
sqnorm_partial_kernel(const __nv_bfloat16* __restrict__ g, int64_t n, float gscale, float* __restrict__ partials) {
    __shared__ float red[32];
    float acc = 0.f;
    const int64_t nch = n / 8;
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nch; c += (int64_t)gridDim.x * blockDim.x) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(g) + c), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float x = __bfloat162float(__float2bfloat16_rn(f[e] * gscale)); acc += x * x; }
    }
    if (blockIdx.x == 0)
        for (int64_t e = nch * 8 + threadIdx.x; e < n; e += blockDim.x) {
            const float x = __bfloat162float(__float2bfloat16_rn(__bfloat162float(g[e]) * gscale));
            acc += x * x;
        }
    acc = block_sum_f(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(1024) sqnorm_final_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
    acc = block_sum_f(acc, red);
    if (threadIdx.x == 0) out[0] = sqrtf(acc);
}
// clip coefficient min(1, max_norm / (||g|| + 1e-6)) read on the device (no host sync), then torch.optim.AdamW
// on the fp32 masters and a bf16 write-back.  gscale = 1/world (DDP averages gradients, backend.py:233-253).
__global__ void __launch_bounds__(256)
adamw_kernel(const __nv_bfloat16* __restrict__ g, float* __restrict__ master, float* __restrict__ m1, float* __restrict__ m2,
             __nv_bfloat16* __restrict__ param, int64_t n, const float* __restrict__ gnorm, float max_norm, float gscale,
             float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt) {
    const float clip = (max_norm > 0.f) ? fminf(1.0f, max_norm / (gnorm[0] + 1e-6f)) : 1.0f;
    const float step_size = lr / bc1, decay = 1.f - lr * wd;
    auto upd = [&](float gr_raw, float& w, float& a, float& v) {
        const float gr = __bfloat162float(__float2bfloat16_rn(gr_raw * gscale)) * clip;
        w *= decay;
        a = a + (gr - a) * (1.f - beta1);                   // lerp_
        v = v * beta2 + (1.f - beta2) * gr * gr;
        w -= step_size * (a / (sqrtf(v) / bc2_sqrt + eps));
    };
    const int64_t n4 = n / 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const uint2 gu = __ldg(reinterpret_cast<const uint2*>(g) + i);
        const __nv_bfloat162 g01 = *reinterpret_cast<const __nv_bfloat162*>(&gu.x);
        const __nv_bfloat162 g23 = *reinterpret_cast<const __nv_bfloat162*>(&gu.y);
        float4 w = reinterpret_cast<float4*>(master)[i];
        float4 a = reinterpret_cast<float4*>(m1)[i];
        float4 v = reinterpret_cast<float4*>(m2)[i];
        upd(__bfloat162float(g01.x), w.x, a.x, v.x);
        upd(__bfloat162float(g01.y), w.y, a.y, v.y);
        upd(__bfloat162float(g23.x), w.z, a.z, v.z);
        upd(__bfloat162float(g23.y), w.w, a.w, v.w);
        reinterpret_cast<float4*>(master)[i] = w;
        reinterpret_cast<float4*>(m1)[i] = a;
        reinterpret_cast<float4*>(m2)[i] = v;
        uint2 o; o.x = pack_bf16x2(w.x, w.y); o.y = pack_bf16x2(w.z, w.w);
        reinterpret_cast<uint2*>(param)[i] = o;
    }
    if (blockIdx.x == 0)
        for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
            float w = master[i], a = m1[i], v = m2[i];
            upd(__bfloat162float(g[i]), w, a, v);
            master[i] = w; m1[i] = a; m2[i] = v; param[i] = __float2bfloat16_rn(w);
        }
}
__global__ void __launch_bounds__(256) cvt_flat_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                               int64_t n, const float* __restrict__ scale_dev) {
    const float s = scale_dev ? scale_dev[0] : 1.0f;
    const int64_t n4 = n / 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        uint2 o;
        o.x = pack_bf16x2(v.x * s, v.y * s);
        o.y = pack_bf16x2(v.z * s, v.w * s);
        reinterpret_cast<uint2*>(dst)[i] = o;
    }
    if (blockIdx.x == 0)
        for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) dst[i] = __float2bfloat16_rn(src[i] * s);
}

// ------------------------------------------------------------------ host wrappers
static int sm_count() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148; }
    return n;
}

// rows [row0, row0 + nrows) of the [B*S, V] teacher logits; `lean` = one block per SM (fits beside a GEMM CTA).
int teacher(const void* tl, int64_t ld, const int* d2t_idx, const uint8_t* t2d, const int* loss_mask, void* xg, float* tstats,
            int64_t* ids, int* position_mask, int B, int S, int T, int V, int DV, int64_t row0, int64_t nrows, int pad,
            int lean, cudaStream_t st) {
    if (V % 8) return set_error(-22, "teacher: target vocab %d must be a multiple of 8", V);
    if (nrows > 0) {
        int64_t grid = (int64_t)sm_count() * (lean ? 1 : 2);
        if (grid > nrows) grid = nrows;
        teacher_kernel<<<(unsigned)grid, 512, 0, st>>>((const __nv_bfloat16*)tl, ld, d2t_idx, t2d, loss_mask, (__nv_bfloat16*)xg,
                                                     reinterpret_cast<float4*>(tstats), ids, position_mask, S, T, V, DV, row0, nrows);
        SF_CUDA_CHECK_LAUNCH("teacher");
    }
    if (pad && T > 0) {
        teacher_pad_kernel<<<sm_count(), 256, 0, st>>>((__nv_bfloat16*)xg, reinterpret_cast<float4*>(tstats), ids, B, S, T, DV);
        SF_CUDA_CHECK_LAUNCH("teacher_pad");
    }
    return 0;
}

int t2d_index(const uint8_t* t2d, int V, uint32_t* bits, int* prefix, cudaStream_t st) {
    t2d_index_kernel<<<1, 1024, 0, st>>>(t2d, V, bits, prefix);
    SF_CUDA_CHECK_LAUNCH("t2d_index");
    return 0;
}
int teacher_merge(const float* stats, int nb, int64_t M, const uint8_t* t2d, const int* loss_mask, float* tstats, int64_t* ids,
                  int* position_mask, void* xg, int B, int S, int T, int DV, cudaStream_t st) {
    teacher_merge_kernel<<<(unsigned)((M * 32 + 255) / 256), 256, 0, st>>>(stats, nb, M, t2d, loss_mask, reinterpret_cast<float4*>(tstats),
                                                                    ids, position_mask, S, T);
    SF_CUDA_CHECK_LAUNCH("teacher_merge");
    if (T > 0) {
        teacher_pad_kernel<<<sm_count(), 256, 0, st>>>((__nv_bfloat16*)xg, reinterpret_cast<float4*>(tstats), ids, B, S, T, DV);
        SF_CUDA_CHECK_LAUNCH("teacher_pad");
    }
    return 0;
}

int loss_step(void* logits, int64_t ld, const void* xg, const float* tstats, const int64_t* tgt_ids,
              const int* position_mask, const int* loss_mask, const int64_t* d2t, int B, int S, int T, int DV, int step,
              float step_weight, int write_grad, int lk_type, float kl_scale, float kl_decay, float* row_ws, float* metrics,
              const float* stats, int stats_nb, cudaStream_t st) {
    if (DV % 8 || ld % 8) return set_error(-22, "loss: draft vocab %d / ld must be multiples of 8", DV);
    LossParams p;
    p.logits = (__nv_bfloat16*)logits; p.ld = ld; p.xg = (const __nv_bfloat16*)xg; p.tstats = reinterpret_cast<const float4*>(tstats);
    p.tgt_ids = tgt_ids;
    p.position_mask = position_mask; p.loss_mask = loss_mask; p.d2t = d2t; p.S = S; p.T = T; p.DV = DV; p.step = step;
    const int64_t M = (int64_t)B * S;
    p.M = M;
    p.grad_coef = step_weight / (float)M; p.write_grad = (lk_type == 0) ? write_grad : 0;
    p.row_loss = row_ws; p.row_accept = row_ws + M; p.row_correct = row_ws + 2 * M;
    p.stats = stats; p.stats_nb = stats_nb;
    loss_kernel<<<(unsigned)M, 512, 0, st>>>(p);
    SF_CUDA_CHECK_LAUNCH("loss");
    metrics_reduce_kernel<<<1, 1024, 0, st>>>(p.row_loss, p.row_accept, p.row_correct, position_mask, loss_mask, B, S, step,
                                             lk_type, kl_scale, kl_decay, metrics);
    SF_CUDA_CHECK_LAUNCH("metrics_reduce");
    if (lk_type != 0 && write_grad) {
        lk_grad_kernel<<<(unsigned)M, 512, 0, st>>>(p, lk_type, step_weight, metrics);
        SF_CUDA_CHECK_LAUNCH("lk_grad");
    }
    return 0;
}

int grad_norm(const void* g, int64_t n, float gscale, float* partials_ws, float* out, cudaStream_t st) {
    const int blocks = 148 * 4;
    sqnorm_partial_kernel<<<blocks, 512, 0, st>>>((const __nv_bfloat16*)g, n, gscale, partials_ws);
    SF_CUDA_CHECK_LAUNCH("sqnorm_partial");
    sqnorm_final_kernel<<<1, 1024, 0, st>>>(partials_ws, blocks, out);
    SF_CUDA_CHECK_LAUNCH("sqnorm_final");
    return 0;
}
int adamw(const void* g, float* master, float* m1, float* m2, void* param, int64_t n, const float* gnorm, float max_norm,
          float gscale, float lr, float beta1, float beta2, float eps, float wd, int step, cudaStream_t st) {
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2 = 1.f - powf(beta2, (float)step);
    adamw_kernel<<<148 * 8, 256, 0, st>>>((const __nv_bfloat16*)g, master, m1, m2, (__nv_bfloat16*)param, n, gnorm, max_norm,
                                         gscale, lr, beta1, beta2, eps, wd, bc1, sqrtf(bc2));
    SF_CUDA_CHECK_LAUNCH("adamw");
    return 0;
}
int cvt_flat_f32_bf16(const float* src, void* dst, int64_t n, const float* scale_dev, cudaStream_t st) {
    cvt_flat_f32_bf16_kernel<<<148 * 8, 256, 0, st>>>(src, (__nv_bfloat16*)dst, n, scale_dev);
    SF_CUDA_CHECK_LAUNCH("cvt_flat");
    return 0;
}

}  // namespace sf
