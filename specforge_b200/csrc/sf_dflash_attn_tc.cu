// sf_dflash_attn_tc.cu — DFlash block attention on tcgen05 / TMEM / TMA, forward.
//
// STATUS: default DFlash attention wherever attn_tc_supported() covers the shape (round 2: first run on a B200 matched the
// dense fp32 reference and the CUDA-core kernels at d = 64 / 128, with dropped blocks, S up to 1100 — profiles/
// r02_dflash_attn_tc_check.txt; goldens tests/golden/dflashtc_*.pt).  sf_debug_option("dflash_attn_tc", -1) or
// SF_DFLASH_ATTN_TC=-1 falls back to the CUDA-core kernels of sf_dflash_kernels.cu for A/B runs.
//
// Derived from attn_fwd_tc_kernel (sf_attention_tc.cu) — same warp roles, rings, TMEM plan and lazy-rescale softmax.
// What changes is the tiling of the problem.  A DFlash query row (block n, slot o, head h) may see the context keys
// [0, anchor_n) and the bs noise keys of its own block (dflash_family_model.py:47-89); all g*bs rows of (block, kv head)
// share that key set.  With R = g*bs rows per block:
//   unit  = 128 query rows = 128/R consecutive anchor blocks x the g heads of one kv head   (TMEM lane = row)
//   CTA   = 2 units (the two "heads" of the TTT kernel) = 256/R consecutive blocks, one kv head, one sequence
//   keys  = context tiles [0, max anchor of the CTA) with the per-row limit key < anchor(row's block), then ONE own tile:
//           the CTA's own 256/g noise keys with a block-diagonal mask.  Anchors are sorted, so neighbouring blocks differ
//           by a few positions and almost nothing of a context tile is masked away.
// Q rows are gathered by TMA straight from the [Mq, nh*d] layout with boxes of bs rows (bs % 8 == 0 keeps every box on a
// swizzle-atom boundary), so nothing is re-laid-out in memory.
#include "sf_gemm.cuh"
#include "sf_host.h"
#include "sf_dflash.h"

namespace sf {

int make_tmap_3d_bf16(CUtensorMap* tm, const void* base, int64_t cols, int64_t rows, int64_t batches, int64_t ld, int box_rows);

namespace dflash {

template <int kRegs> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
__device__ __forceinline__ uint32_t sw128(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

struct TcParams {
    __nv_bfloat16* out; int64_t ldo;
    float* lse;                    // [Mq, nh], natural log of sum exp(scaled score)
    const int32_t* anchors; const uint8_t* keep;
    int B, S, N, bs, nh, nkv, g;
    int window;                    // sliding-window layer (0 = full attention)
    float scale_log2;              // d^-0.5 * log2(e)
};

template <int D>
struct TcCfg {
    static constexpr int BQ = 128, BKV = 64, NB = D / 64;
    static constexpr int Q_BYTES = BQ * D * 2;
    static constexpr int KV_BYTES = BKV * D * 2;
    static constexpr int P_BYTES = BQ * BKV * 2;
    static constexpr int kStages = 4;
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = 2 * Q_BYTES;
    static constexpr int OFF_V = OFF_K + kStages * KV_BYTES;
    static constexpr int OFF_P = OFF_V + kStages * KV_BYTES;
    static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
    static constexpr int SMEM = OFF_BAR + 512 + 1024;
    static constexpr int TM_S = 0;
    static constexpr int TM_O = 256;
    static constexpr float kRescaleThreshold = 8.0f;
};

template <int D>
__global__ void __launch_bounds__(384, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kc,
                   const __grid_constant__ CUtensorMap tm_vc, const __grid_constant__ CUtensorMap tm_kn,
                   const __grid_constant__ CUtensorMap tm_vn, const TcParams p) {
    using C = TcCfg<D>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t bar0 = sbase + C::OFF_BAR;
    const uint32_t b_qfull = bar0;
    auto b_kfull = [&](int s) { return bar0 + 8u * (1 + s); };
    auto b_kempty = [&](int s) { return bar0 + 8u * (5 + s); };
    auto b_vfull = [&](int s) { return bar0 + 8u * (9 + s); };
    auto b_vempty = [&](int s) { return bar0 + 8u * (13 + s); };
    auto b_sfull = [&](int x, int u) { return bar0 + 8u * (17 + x * 2 + u); };
    auto b_pfull = [&](int x) { return bar0 + 8u * (21 + x); };
    auto b_pvdone = [&](int x) { return bar0 + 8u * (23 + x); };
    const uint32_t tmem_slot = bar0 + 8u * 25;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kvh = blockIdx.y, b = blockIdx.z;
    const int R = p.g * p.bs;                 // query rows per anchor block (64 for Qwen3-8B: 4 heads x 16 slots)
    const int BPU = C::BQ / R;                // blocks per unit
    const int BPC = 2 * BPU;                  // blocks per CTA
    const int n0 = blockIdx.x * BPC;
    constexpr int nx = 2;
    int amax = 0, amin = 0x7fffffff;          // largest / smallest anchor of a kept block of this CTA (uniform)
    for (int i = 0; i < BPC; ++i) {
        const int n = n0 + i;
        if (n < p.N && p.keep[b * p.N + n]) { amax = max(amax, p.anchors[b * p.N + n]); amin = min(amin, p.anchors[b * p.N + n]); }
    }
    // sliding-window layer (window > 0): slot o of a block anchored at a sees context keys [a + o - (W - 1), a) and own slots <= o
    // (dflash_family_model.py:73-84); context tiles that end below every kept block's window are not visited at all
    const int t_lo = (p.window > 0 && amax > 0) ? max(0, amin - (p.window - 1)) / C::BKV : 0;
    const int n_ctx = (amax + C::BKV - 1) / C::BKV - t_lo;
    const int n_kv = n_ctx + 1;               // + the own-keys tile

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kc); tma_prefetch_desc(&tm_vc); tma_prefetch_desc(&tm_kn); tma_prefetch_desc(&tm_vn); }
    if (warp == 1 && lane == 0) {
        mbar_init(b_qfull, 1);
        for (int s = 0; s < C::kStages; ++s) { mbar_init(b_kfull(s), 1); mbar_init(b_kempty(s), nx); mbar_init(b_vfull(s), 1); mbar_init(b_vempty(s), nx); }
        for (int x = 0; x < 2; ++x) {
            for (int u = 0; u < 2; ++u) mbar_init(b_sfull(x, u), 1);
            mbar_init(b_pfull(x), 4); mbar_init(b_pvdone(x), 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t*>(sgen + C::OFF_BAR + 8 * 25);

    if (warp < 4) {
        reg_dec<40>();
        if (warp == 0 && lane == 0) {
            // ================= TMA producer: Q (gathered by block / head), K ring =================
            mbar_expect_tx(b_qfull, C::Q_BYTES * nx);
            for (int x = 0; x < nx; ++x)
                for (int blk = 0; blk < BPU; ++blk)
                    for (int hg = 0; hg < p.g; ++hg)
                        for (int kb = 0; kb < C::NB; ++kb)
                            tma_load_3d(sbase + C::OFF_Q + x * C::Q_BYTES + kb * (C::BQ * 128) + (blk * R + hg * p.bs) * 128, &tm_q, b_qfull,
                                        (kvh * p.g + hg) * D + kb * 64, (n0 + x * BPU + blk) * p.bs, b);
            for (int t = 0; t < n_kv; ++t) {
                const int s = t % C::kStages;
                mbar_wait(b_kempty(s), ((t / C::kStages) & 1) ^ 1u, 11);
                mbar_expect_tx(b_kfull(s), C::KV_BYTES);
                for (int kb = 0; kb < C::NB; ++kb) {
                    const uint32_t dst = sbase + C::OFF_K + s * C::KV_BYTES + kb * (C::BKV * 128);
                    if (t < n_ctx) tma_load_3d(dst, &tm_kc, b_kfull(s), kvh * D + kb * 64, (t_lo + t) * C::BKV, b);
                    else           tma_load_3d(dst, &tm_kn, b_kfull(s), kvh * D + kb * 64, n0 * p.bs, b);
                }
            }
        } else if (warp == 3 && lane == 0) {
            // ================= V producer =================
            for (int t = 0; t < n_kv; ++t) {
                const int s = t % C::kStages;
                mbar_wait(b_vempty(s), ((t / C::kStages) & 1) ^ 1u, 12);
                mbar_expect_tx(b_vfull(s), C::KV_BYTES);
                for (int kb = 0; kb < C::NB; ++kb) {
                    const uint32_t dst = sbase + C::OFF_V + s * C::KV_BYTES + kb * (C::BKV * 128);
                    if (t < n_ctx) tma_load_3d(dst, &tm_vc, b_vfull(s), kvh * D + kb * 64, (t_lo + t) * C::BKV, b);
                    else           tma_load_3d(dst, &tm_vn, b_vfull(s), kvh * D + kb * 64, n0 * p.bs, b);
                }
            }
        } else if ((warp == 1 || warp == 2) && lane == 0) {
            // ================= MMA issuers: one thread per unit =================
            const int x = warp - 1;
            constexpr uint32_t idesc_s = make_idesc_bf16(128, C::BKV, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_bf16(128, D, 0, 1);
            const uint32_t sq = sbase + C::OFF_Q + x * C::Q_BYTES;
            auto issue_s = [&](int t) {
                const int s = t % C::kStages, u = t & 1;
                mbar_wait(b_kfull(s), (t / C::kStages) & 1, 17);
                tc_fence_after();
                const uint32_t sk = sbase + C::OFF_K + s * C::KV_BYTES;
#pragma unroll
                for (int kb = 0; kb < C::NB; ++kb) {
                    const uint64_t adesc = make_smem_desc_sw128(sq + kb * (C::BQ * 128), 0, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(sk + kb * (C::BKV * 128), 0, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16<1>(tmem + C::TM_S + (x * 2 + u) * C::BKV, adesc + ((k * 32) >> 4), bdesc + ((k * 32) >> 4),
                                     idesc_s, (kb | k) != 0);
                }
                umma_commit(b_sfull(x, u));
                umma_commit(b_kempty(s));
            };
            mbar_wait(b_qfull, 0, 13);
            tc_fence_after();
            issue_s(0);
            if (n_kv > 1) issue_s(1);
            for (int t = 0; t < n_kv; ++t) {
                const int s = t % C::kStages;
                mbar_wait(b_vfull(s), (t / C::kStages) & 1, 15);
                mbar_wait(b_pfull(x), t & 1, 14);
                tc_fence_after();
                const uint32_t sp = sbase + C::OFF_P + x * C::P_BYTES;
                const uint32_t sv = sbase + C::OFF_V + s * C::KV_BYTES;
                const uint64_t adesc = make_smem_desc_sw128(sp, 0, 1024);
                const uint64_t bdesc = make_smem_desc_sw128(sv, C::BKV * 128, 1024);
#pragma unroll
                for (int k = 0; k < C::BKV / 16; ++k)
                    umma_bf16<1>(tmem + C::TM_O + x * D, adesc + ((k * 32) >> 4), bdesc + ((k * 2048) >> 4), idesc_o, (t | k) != 0);
                umma_commit(b_pvdone(x));
                umma_commit(b_vempty(s));
                if (t + 2 < n_kv) issue_s(t + 2);
            }
        }
    } else {
        // ================= softmax warpgroups (one per unit) =================
        reg_inc<224>();
        const int x = (warp - 4) >> 2;
        const int wq = warp & 3;
        const int r = wq * 32 + lane;           // row inside the unit == TMEM lane
        const int blk = x * BPU + r / R;        // block index inside the CTA
        const int n = n0 + blk;
        const int hg = (r % R) / p.bs, o = r % p.bs;
        const bool valid = n < p.N;
        const bool kept = valid && p.keep[b * p.N + n] != 0;
        const int a_r = kept ? p.anchors[b * p.N + n] : 0;       // context keys [lo_r, a_r)
        const int W = p.window, lo_r = (W > 0 && kept) ? a_r + o - (W - 1) : 0;
        const uint32_t t_lane = tmem + ((uint32_t)(wq * 32) << 16);
        const uint32_t t_o = t_lane + C::TM_O + x * D;
        float m_ref = -INFINITY, l = 0.f;
        const float c = p.scale_log2;
        for (int t = 0; t < n_kv; ++t) {
            const int u = t & 1;
            const int kv0 = (t_lo + t) * C::BKV;
            const bool own = t == n_ctx;
            mbar_wait(b_sfull(x, u), (t >> 1) & 1, 22 + x);
            tc_fence_after();
            uint32_t sv[C::BKV];
            tmem_ld_32x32b_x32(t_lane + C::TM_S + (x * 2 + u) * C::BKV, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
            tmem_ld_32x32b_x32(t_lane + C::TM_S + (x * 2 + u) * C::BKV + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
            tmem_ld_wait();
            if (own) {                          // block-diagonal: key column cc belongs to CTA block cc / bs
#pragma unroll
                for (int cc = 0; cc < C::BKV; ++cc)
                    if (!kept || cc / p.bs != blk || (W > 0 && cc % p.bs > o)) sv[cc] = 0xff800000u;
            } else if (!kept || kv0 + C::BKV > a_r || kv0 < lo_r) {
#pragma unroll
                for (int cc = 0; cc < C::BKV; ++cc)
                    if (!kept || kv0 + cc >= a_r || kv0 + cc < lo_r) sv[cc] = 0xff800000u;
            }
            float mx = -INFINITY;
#pragma unroll
            for (int cc = 0; cc < C::BKV; ++cc) mx = fmaxf(mx, __uint_as_float(sv[cc]));
            mx *= c;
            const bool jump = mx > m_ref + C::kRescaleThreshold;
            if (__any_sync(0xffffffffu, jump)) {
                const float m_new = jump ? mx : m_ref;
                const float f = (m_ref == -INFINITY) ? 0.f : exp2f(m_ref - m_new);
                l *= f;
                m_ref = m_new;
                if (t > 0) {
                    mbar_wait(b_pvdone(x), (t - 1) & 1, 26 + x);
                    tc_fence_after();
#pragma unroll
                    for (int cc = 0; cc < D / 32; ++cc) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(t_o + cc * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * f);
                        tmem_st_32x32b_x32(t_o + cc * 32, v);
                    }
                    tmem_st_wait();
                }
            }
            const float base = (m_ref == -INFINITY) ? 0.f : m_ref;
            if (t > 0) mbar_wait(b_pvdone(x), (t - 1) & 1, 24 + x);
            uint8_t* sp = sgen + C::OFF_P + x * C::P_BYTES;
            float rs = 0.f;
#pragma unroll
            for (int j = 0; j < C::BKV / 8; ++j) {
                float e[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { e[k] = ex2_approx(fmaf(__uint_as_float(sv[j * 8 + k]), c, -base)); rs += e[k]; }
                uint4 pk;
                pk.x = pack_bf16x2(e[0], e[1]); pk.y = pack_bf16x2(e[2], e[3]);
                pk.z = pack_bf16x2(e[4], e[5]); pk.w = pack_bf16x2(e[6], e[7]);
                *reinterpret_cast<uint4*>(sp + sw128(r, j)) = pk;
            }
            l += rs;
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_pfull(x));
        }
        // ---- epilogue: O / l -> out[(b, n, o), head]; a dropped block yields zeros (l == 0)
        mbar_wait(b_pvdone(x), (n_kv - 1) & 1, 28 + x);
        tc_fence_after();
        const float inv = (l > 0.f) ? 1.f / l : 0.f;
        const int head = kvh * p.g + hg;
        const int64_t orow = ((int64_t)b * p.N + min(n, p.N - 1)) * p.bs + o;
        __nv_bfloat16* optr = p.out + orow * p.ldo + head * D;
#pragma unroll
        for (int cc = 0; cc < D / 32; ++cc) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_o + cc * 32, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 uu;
                    uu.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]) * inv, __uint_as_float(v[q * 8 + 1]) * inv);
                    uu.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]) * inv, __uint_as_float(v[q * 8 + 3]) * inv);
                    uu.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]) * inv, __uint_as_float(v[q * 8 + 5]) * inv);
                    uu.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]) * inv, __uint_as_float(v[q * 8 + 7]) * inv);
                    reinterpret_cast<uint4*>(optr + cc * 32)[q] = uu;
                }
            }
        }
        if (valid) p.lse[orow * p.nh + head] = (l > 0.f) ? (m_ref + log2f(l)) * 0.6931471805599453f : 0.f;
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc<1>(tmem, 512); }
}

template <int D>
static int fwd_tc_t(const AttnArgs& a, cudaStream_t st) {
    using C = TcCfg<D>;
    const int g = a.nh / a.nkv, R = g * a.bs;
    const int64_t Q = (int64_t)a.N * a.bs;
    CUtensorMap tq, tkc, tvc, tkn, tvn;
    SF_TRY_RC(make_tmap_3d_bf16(&tq, a.q, (int64_t)a.nh * D, Q, a.B, a.ldq, a.bs));
    SF_TRY_RC(make_tmap_3d_bf16(&tkc, a.kc, (int64_t)a.nkv * D, a.S, a.B, a.ldkc, C::BKV));
    // V operands may be views into fused projection rows (row stride ld > nkv*D): the map covers the nkv*D columns of the view
    const int64_t KVc = (int64_t)a.nkv * D;
    SF_TRY_RC(make_tmap_3d_bf16(&tvc, a.vc, KVc, a.S, a.B, a.ldvc, C::BKV));
    SF_TRY_RC(make_tmap_3d_bf16(&tkn, a.kn, KVc, Q, a.B, a.ldkn, C::BKV));
    SF_TRY_RC(make_tmap_3d_bf16(&tvn, a.vn, KVc, Q, a.B, a.ldvn, C::BKV));
    TcParams p{};
    p.out = a.out; p.ldo = a.ldo; p.lse = a.lse; p.anchors = a.anchors; p.keep = a.keep;
    p.B = a.B; p.S = a.S; p.N = a.N; p.bs = a.bs; p.nh = a.nh; p.nkv = a.nkv; p.g = g; p.window = a.window;
    p.scale_log2 = a.scale * 1.4426950408889634f;
    static bool set = false;
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return set_error(-22, "dflash attn_fwd_tc smem attr: %s", cudaGetErrorString(e));
        set = true;
    }
    const int BPC = 2 * (C::BQ / R);
    dim3 grid((a.N + BPC - 1) / BPC, a.nkv, a.B);
    attn_fwd_tc_kernel<D><<<grid, 384, C::SMEM, st>>>(tq, tkc, tvc, tkn, tvn, p);
    SF_CUDA_CHECK_LAUNCH("dflash attn_fwd_tc");
    return 0;
}

// shapes the tensor-core path covers; everything else stays on the CUDA-core kernels
bool attn_tc_supported(const AttnArgs& a) {
    const int g = a.nh / a.nkv, R = g * a.bs;
    return (a.d == 64 || a.d == 128) && a.bs % 8 == 0 && (R == 64 || R == 128) && (g == 4 || g == 8) && 256 / g <= 64;
}
int attn_fwd_tc(const AttnArgs& a, cudaStream_t st) { return a.d == 128 ? fwd_tc_t<128>(a, st) : fwd_tc_t<64>(a, st); }

}  // namespace dflash
}  // namespace sf
