// sf_dflash_attn_tc_bwd.cu — DFlash block attention backward on tcgen05 / TMEM / TMA.
//
// STATUS: default together with sf_dflash_attn_tc.cu (same evidence: profiles/r02_dflash_attn_tc_check.txt and
// tests/test_dflash_gpu.py::test_dflash_tc_shapes).
//
// Derived from attn_bwd_dq_tc_kernel / attn_bwd_dkv_tc_kernel (sf_attention_tc_bwd.cu): same warp roles, rings, TMEM plans.
//   df_bwd_dq_tc   CTA = one unit (128 query rows = 128/R anchor blocks x g heads x bs slots), kv head, sequence.
//                  Keys: context tiles [0, max anchor of the unit) with the per-row limit, then the unit's own noise keys
//                  (block-diagonal).  dQ accumulates in TMEM.
//   df_bwd_ctx_tc  CTA = 128 context keys, kv head, sequence.  One iteration per kept anchor block whose anchor lies beyond
//                  the tile start (R = 64 query rows = the block's g heads x bs slots; anchors are sorted, kept blocks first,
//                  so those blocks form one contiguous range); all 64 query columns of an iteration share the key limit.
//   df_bwd_own     the dK/dV of a block's own bs noise keys (R x bs x d work) stay on CUDA cores.
//   df_delta       delta = rowsum(dO * O) per (draft row, head).
// lse / delta are [Mq, nh] (natural-log lse, as the forward kernels write it).
#include "sf_gemm.cuh"
#include "sf_host.h"
#include "sf_dflash.h"

namespace sf {

int make_tmap_3d_bf16(CUtensorMap* tm, const void* base, int64_t cols, int64_t rows, int64_t batches, int64_t ld, int box_rows);

namespace dflash {

__device__ __forceinline__ uint32_t sw128c(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

struct TcBwdParams {
    const float* lse; const float* delta;          // [Mq, nh]
    const int32_t* anchors; const uint8_t* keep;   // [B, N]
    __nv_bfloat16* dq; int64_t lddq;               // [Mq, nh*D]
    __nv_bfloat16* dkc; int64_t lddkc;             // [Mc, nkv*D]
    __nv_bfloat16* dvc; int64_t lddvc;
    int B, S, N, bs, nh, nkv, g;
    int window;                                    // sliding-window layer (0 = full attention), see sf_dflash_attn_tc.cu
    float scale_log2;
};

// ============================================================================================ dQ
template <int D>
struct DqCfg {
    static constexpr int BQ = 128, BKV = 64, NB = D / 64, kKStages = 5, kVStages = 4;
    static constexpr int QT_BYTES = BQ * D * 2;
    static constexpr int KT_BYTES = BKV * D * 2;
    static constexpr int DS_BYTES = BQ * BKV * 2;
    static constexpr int OFF_Q = 0, OFF_DO = QT_BYTES;
    static constexpr int OFF_K = 2 * QT_BYTES;
    static constexpr int OFF_V = OFF_K + kKStages * KT_BYTES;
    static constexpr int OFF_DS = OFF_V + kVStages * KT_BYTES;
    static constexpr int OFF_BAR = OFF_DS + DS_BYTES;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static constexpr int TM_S = 0, TM_DP = 128, TM_DQ = 256;
};

template <int D>
__global__ void __launch_bounds__(384, 1)
df_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                    const __grid_constant__ CUtensorMap tm_kc, const __grid_constant__ CUtensorMap tm_vc,
                    const __grid_constant__ CUtensorMap tm_kn, const __grid_constant__ CUtensorMap tm_vn, const TcBwdParams p) {
    using C = DqCfg<D>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t bar0 = sbase + C::OFF_BAR;
    const uint32_t b_qfull = bar0;
    auto b_kfull = [&](int s) { return bar0 + 8u * (1 + s); };
    auto b_kempty = [&](int s) { return bar0 + 8u * (6 + s); };
    auto b_vfull = [&](int s) { return bar0 + 8u * (11 + s); };
    auto b_vempty = [&](int s) { return bar0 + 8u * (15 + s); };
    auto b_sdpfull = [&](int u) { return bar0 + 8u * (19 + u); };
    const uint32_t b_dsfull = bar0 + 8u * 21;
    const uint32_t b_mmadone = bar0 + 8u * 22;
    const uint32_t tmem_slot = bar0 + 8u * 23;
    auto b_sdpfree = [&](int u) { return bar0 + 8u * (24 + u); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kvh = blockIdx.y, b = blockIdx.z;
    const int R = p.g * p.bs, BPU = C::BQ / R;
    const int n0 = blockIdx.x * BPU;              // first anchor block of this unit
    int amax = 0, amin = 0x7fffffff;
    for (int i = 0; i < BPU; ++i) {
        const int n = n0 + i;
        if (n < p.N && p.keep[b * p.N + n]) { amax = max(amax, p.anchors[b * p.N + n]); amin = min(amin, p.anchors[b * p.N + n]); }
    }
    const int t_lo = (p.window > 0 && amax > 0) ? max(0, amin - (p.window - 1)) / C::BKV : 0;   // first context tile inside a window
    const int n_ctx = (amax + C::BKV - 1) / C::BKV - t_lo;
    const int n_kv = n_ctx + 1;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_kc); tma_prefetch_desc(&tm_vc); tma_prefetch_desc(&tm_kn); tma_prefetch_desc(&tm_vn); }
    if (warp == 1 && lane == 0) {
        mbar_init(b_qfull, 1);
        for (int s = 0; s < C::kKStages; ++s) { mbar_init(b_kfull(s), 1); mbar_init(b_kempty(s), 1); }
        for (int s = 0; s < C::kVStages; ++s) { mbar_init(b_vfull(s), 1); mbar_init(b_vempty(s), 1); }
        for (int u = 0; u < 2; ++u) { mbar_init(b_sdpfull(u), 1); mbar_init(b_sdpfree(u), 8); }
        mbar_init(b_dsfull, 8);
        mbar_init(b_mmadone, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t*>(sgen + C::OFF_BAR + 8 * 23);

    if (warp < 4) {
        if (warp == 0 && lane == 0) {
            // Q and dO, gathered by (block, head) boxes of bs rows; then the K ring (context tiles, then the own tile)
            mbar_expect_tx(b_qfull, 2 * C::QT_BYTES);
            for (int blk = 0; blk < BPU; ++blk)
                for (int hg = 0; hg < p.g; ++hg)
                    for (int kbk = 0; kbk < C::NB; ++kbk) {
                        const uint32_t off = kbk * (C::BQ * 128) + (blk * R + hg * p.bs) * 128;
                        tma_load_3d(sbase + C::OFF_Q + off, &tm_q, b_qfull, (kvh * p.g + hg) * D + kbk * 64, (n0 + blk) * p.bs, b);
                        tma_load_3d(sbase + C::OFF_DO + off, &tm_do, b_qfull, (kvh * p.g + hg) * D + kbk * 64, (n0 + blk) * p.bs, b);
                    }
            for (int t = 0; t < n_kv; ++t) {
                const int sk = t % C::kKStages;
                mbar_wait(b_kempty(sk), ((t / C::kKStages) & 1) ^ 1u, 41);
                mbar_expect_tx(b_kfull(sk), C::KT_BYTES);
                for (int kbk = 0; kbk < C::NB; ++kbk) {
                    const uint32_t dst = sbase + C::OFF_K + sk * C::KT_BYTES + kbk * (C::BKV * 128);
                    if (t < n_ctx) tma_load_3d(dst, &tm_kc, b_kfull(sk), kvh * D + kbk * 64, (t_lo + t) * C::BKV, b);
                    else           tma_load_3d(dst, &tm_kn, b_kfull(sk), kvh * D + kbk * 64, n0 * p.bs, b);
                }
            }
        } else if (warp == 3 && lane == 0) {
            for (int t = 0; t < n_kv; ++t) {
                const int sv = t % C::kVStages;
                mbar_wait(b_vempty(sv), ((t / C::kVStages) & 1) ^ 1u, 42);
                mbar_expect_tx(b_vfull(sv), C::KT_BYTES);
                for (int kbk = 0; kbk < C::NB; ++kbk) {
                    const uint32_t dst = sbase + C::OFF_V + sv * C::KT_BYTES + kbk * (C::BKV * 128);
                    if (t < n_ctx) tma_load_3d(dst, &tm_vc, b_vfull(sv), kvh * D + kbk * 64, (t_lo + t) * C::BKV, b);
                    else           tma_load_3d(dst, &tm_vn, b_vfull(sv), kvh * D + kbk * 64, n0 * p.bs, b);
                }
            }
        } else if (warp == 1 && lane == 0) {
            // ================= MMA issuer A: S = Q K^T, dP = dO V^T =================
            constexpr uint32_t idesc_s = make_idesc_bf16(128, C::BKV, 0, 0);
            mbar_wait(b_qfull, 0, 45);
            tc_fence_after();
            for (int t = 0; t < n_kv; ++t) {
                const int sk = t % C::kKStages, sv = t % C::kVStages, u = t & 1;
                mbar_wait(b_sdpfree(u), ((t >> 1) & 1) ^ 1u, 50);
                mbar_wait(b_kfull(sk), (t / C::kKStages) & 1, 43);
                mbar_wait(b_vfull(sv), (t / C::kVStages) & 1, 44);
                tc_fence_after();
#pragma unroll
                for (int kbk = 0; kbk < C::NB; ++kbk) {
                    const uint64_t a = make_smem_desc_sw128(sbase + C::OFF_Q + kbk * (C::BQ * 128), 0, 1024);
                    const uint64_t bk = make_smem_desc_sw128(sbase + C::OFF_K + sk * C::KT_BYTES + kbk * (C::BKV * 128), 0, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16<1>(tmem + C::TM_S + u * C::BKV, a + ((k * 32) >> 4), bk + ((k * 32) >> 4), idesc_s, (kbk | k) != 0);
                }
#pragma unroll
                for (int kbk = 0; kbk < C::NB; ++kbk) {
                    const uint64_t a = make_smem_desc_sw128(sbase + C::OFF_DO + kbk * (C::BQ * 128), 0, 1024);
                    const uint64_t bv = make_smem_desc_sw128(sbase + C::OFF_V + sv * C::KT_BYTES + kbk * (C::BKV * 128), 0, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16<1>(tmem + C::TM_DP + u * C::BKV, a + ((k * 32) >> 4), bv + ((k * 32) >> 4), idesc_s, (kbk | k) != 0);
                }
                umma_commit(b_sdpfull(u));
                umma_commit(b_vempty(sv));
            }
        } else if (warp == 2 && lane == 0) {
            // ================= MMA issuer B: dQ += dS K =================
            constexpr uint32_t idesc_dq = make_idesc_bf16(128, D, 0, 1);
            for (int t = 0; t < n_kv; ++t) {
                const int sk = t % C::kKStages;
                mbar_wait(b_kfull(sk), (t / C::kKStages) & 1, 43);
                mbar_wait(b_dsfull, t & 1, 46);
                tc_fence_after();
                const uint64_t ads = make_smem_desc_sw128(sbase + C::OFF_DS, 0, 1024);
                const uint64_t bk = make_smem_desc_sw128(sbase + C::OFF_K + sk * C::KT_BYTES, C::BKV * 128, 1024);   // MN-major
#pragma unroll
                for (int k = 0; k < C::BKV / 16; ++k)
                    umma_bf16<1>(tmem + C::TM_DQ, ads + ((k * 32) >> 4), bk + ((k * 2048) >> 4), idesc_dq, (t | k) != 0);
                umma_commit(b_mmadone);
                umma_commit(b_kempty(sk));
            }
        }
    } else {
        // ================= dS warpgroups: one thread per query row, WG x handles kv columns [32x, 32x+32) =================
        const int x = (warp - 4) >> 2;
        const int wq = warp & 3;
        const int r = wq * 32 + lane;
        const int blk = r / R;
        const int n = n0 + blk;
        const int hg = (r % R) / p.bs, o = r % p.bs;
        const bool valid = n < p.N;
        const bool kept = valid && p.keep[b * p.N + n] != 0;
        const int a_r = kept ? p.anchors[b * p.N + n] : 0;
        const int W = p.window, lo_r = (W > 0 && kept) ? a_r + o - (W - 1) : 0;
        const int head = kvh * p.g + hg;
        const int64_t orow = ((int64_t)b * p.N + min(n, p.N - 1)) * p.bs + o;
        const uint32_t t_lane = tmem + ((uint32_t)(wq * 32) << 16);
        const float c = p.scale_log2;
        const float scale = c * kLn2;
        const float L = kept ? p.lse[orow * p.nh + head] * kLog2e : INFINITY;
        const float Dl = kept ? p.delta[orow * p.nh + head] : 0.f;
        for (int t = 0; t < n_kv; ++t) {
            const int u = t & 1;
            const int kv0 = (t_lo + t) * C::BKV;
            const bool own = t == n_ctx;
            mbar_wait(b_sdpfull(u), (t >> 1) & 1, 47);
            tc_fence_after();
            uint32_t sv[32], dv[32];
            tmem_ld_32x32b_x32(t_lane + C::TM_S + u * C::BKV + x * 32, sv);
            tmem_ld_32x32b_x32(t_lane + C::TM_DP + u * C::BKV + x * 32, dv);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_sdpfree(u));
            if (own) {
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    if (!kept || (x * 32 + e) / p.bs != blk || (W > 0 && (x * 32 + e) % p.bs > o)) sv[e] = 0xff800000u;
            } else if (!kept || kv0 + C::BKV > a_r || kv0 < lo_r) {
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    if (!kept || kv0 + x * 32 + e >= a_r || kv0 + x * 32 + e < lo_r) sv[e] = 0xff800000u;
            }
            uint32_t dsk[16];
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(sv[e]), c, -L));
                const float p1 = ex2_approx(fmaf(__uint_as_float(sv[e + 1]), c, -L));
                dsk[e / 2] = pack_bf16x2(p0 * (__uint_as_float(dv[e]) - Dl) * scale, p1 * (__uint_as_float(dv[e + 1]) - Dl) * scale);
            }
            if (t > 0) mbar_wait(b_mmadone, (t - 1) & 1, 48);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(sgen + C::OFF_DS + sw128c(r, x * 4 + j)) = make_uint4(dsk[j * 4], dsk[j * 4 + 1], dsk[j * 4 + 2], dsk[j * 4 + 3]);
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_dsfull);
        }
        mbar_wait(b_mmadone, (n_kv - 1) & 1, 49);
        tc_fence_after();
        __nv_bfloat16* dqp = p.dq + orow * p.lddq + head * D;
#pragma unroll
        for (int cc = x * (D / 64); cc < (x + 1) * (D / 64); ++cc) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_lane + C::TM_DQ + cc * 32, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 ov;
                    ov.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
                    ov.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
                    ov.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
                    ov.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
                    reinterpret_cast<uint4*>(dqp + cc * 32)[q] = ov;
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc<1>(tmem, 512); }
}

// ============================================================================================ context dK / dV
template <int D>
struct CtxCfg {
    static constexpr int BK = 128, BQ = 64, NB = D / 64, kStages = 4;
    static constexpr int KT_BYTES = BK * D * 2;
    static constexpr int QT_BYTES = BQ * D * 2;
    static constexpr int PT_BYTES = BK * BQ * 2;
    static constexpr int OFF_K = 0, OFF_V = KT_BYTES;
    static constexpr int OFF_Q = 2 * KT_BYTES;
    static constexpr int OFF_P = OFF_Q + kStages * 2 * QT_BYTES;
    static constexpr int OFF_DS = OFF_P + PT_BYTES;
    static constexpr int OFF_LD = OFF_DS + PT_BYTES;
    static constexpr int OFF_BAR = OFF_LD + 2 * 2 * BQ * 4;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static constexpr int TM_ST = 0, TM_DP = 128, TM_DV = 256, TM_DK = 256 + D;
};

template <int D>
__global__ void __launch_bounds__(384, 1)
df_bwd_ctx_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                     const __grid_constant__ CUtensorMap tm_kc, const __grid_constant__ CUtensorMap tm_vc, const TcBwdParams p) {
    using C = CtxCfg<D>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t bar0 = sbase + C::OFF_BAR;
    const uint32_t b_kvfull = bar0;
    auto b_qfull = [&](int s) { return bar0 + 8u * (1 + s); };
    auto b_qempty = [&](int s) { return bar0 + 8u * (5 + s); };
    auto b_sdpfull = [&](int u) { return bar0 + 8u * (9 + u); };
    const uint32_t b_pdsfull = bar0 + 8u * 11;
    const uint32_t b_mmadone = bar0 + 8u * 12;
    const uint32_t tmem_slot = bar0 + 8u * 13;
    auto b_ldfull = [&](int u) { return bar0 + 8u * (14 + u); };
    auto b_ldempty = [&](int u) { return bar0 + 8u * (16 + u); };
    auto b_sdpfree = [&](int u) { return bar0 + 8u * (18 + u); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
    const int k0 = kb * C::BK;
    // kept blocks come first and their anchors are sorted (sample_anchor_positions): the blocks that see key k0 are
    // the contiguous range [n_lo, n_hi)
    // (sliding-window layer: additionally only the blocks whose lowest window bound a - (W - 1) does not lie above the tile)
    int n_hi = 0, n_lo = 0;
    for (int n = 0; n < p.N; ++n) {
        if (!p.keep[b * p.N + n]) break;
        const int an = p.anchors[b * p.N + n];
        if (p.window > 0 && an - (p.window - 1) > k0 + C::BK - 1) break;
        n_hi = n + 1;
        if (an <= k0) n_lo = n + 1;
    }
    const int n_it = n_hi - n_lo;
    if (n_it <= 0) {      // nobody attends to this tile: its gradients are zero (uniform early exit, before any barrier)
        const int rows = min(C::BK, p.S - k0);
        for (int i = threadIdx.x; i < rows * (D / 8); i += blockDim.x) {
            const int rr = i / (D / 8), c8 = i % (D / 8);
            reinterpret_cast<uint4*>(p.dkc + ((int64_t)b * p.S + k0 + rr) * p.lddkc + kvh * D)[c8] = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4*>(p.dvc + ((int64_t)b * p.S + k0 + rr) * p.lddvc + kvh * D)[c8] = make_uint4(0, 0, 0, 0);
        }
        return;
    }

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_kc); tma_prefetch_desc(&tm_vc); }
    if (warp == 1 && lane == 0) {
        mbar_init(b_kvfull, 1);
        for (int s = 0; s < C::kStages; ++s) { mbar_init(b_qfull(s), 1); mbar_init(b_qempty(s), 1); }
        for (int u = 0; u < 2; ++u) { mbar_init(b_sdpfull(u), 1); mbar_init(b_ldfull(u), 1); mbar_init(b_ldempty(u), 8); mbar_init(b_sdpfree(u), 8); }
        mbar_init(b_pdsfull, 8);
        mbar_init(b_mmadone, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t*>(sgen + C::OFF_BAR + 8 * 13);

    if (warp < 4) {
        if (warp == 3) {
            // ================= L / delta stager: the block's 64 rows (hg, o) =================
            for (int it = 0; it < n_it; ++it) {
                const int u = it & 1;
                const int n = n_lo + it;
                mbar_wait(b_ldempty(u), ((it >> 1) & 1) ^ 1u, 30);
                float* dst = reinterpret_cast<float*>(sgen + C::OFF_LD) + u * 2 * C::BQ;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = e * 32 + lane;                     // row inside the block: hg * bs + o
                    const int64_t idx = (((int64_t)b * p.N + n) * p.bs + j % p.bs) * p.nh + kvh * p.g + j / p.bs;
                    dst[j] = p.lse[idx] * kLog2e;
                    dst[C::BQ + j] = p.delta[idx];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(b_ldfull(u));
            }
        } else if (warp == 0 && lane == 0) {
            // ================= TMA producer =================
            mbar_expect_tx(b_kvfull, 2 * C::KT_BYTES);
            for (int kbk = 0; kbk < C::NB; ++kbk) {
                tma_load_3d(sbase + C::OFF_K + kbk * (C::BK * 128), &tm_kc, b_kvfull, kvh * D + kbk * 64, k0, b);
                tma_load_3d(sbase + C::OFF_V + kbk * (C::BK * 128), &tm_vc, b_kvfull, kvh * D + kbk * 64, k0, b);
            }
            for (int it = 0; it < n_it; ++it) {
                const int s = it % C::kStages;
                const uint32_t ph = ((it / C::kStages) & 1) ^ 1u;
                const int n = n_lo + it;
                mbar_wait(b_qempty(s), ph, 31);
                mbar_expect_tx(b_qfull(s), 2 * C::QT_BYTES);
                const uint32_t sq = sbase + C::OFF_Q + s * 2 * C::QT_BYTES;
                for (int hg = 0; hg < p.g; ++hg)
                    for (int kbk = 0; kbk < C::NB; ++kbk) {
                        const uint32_t off = kbk * (C::BQ * 128) + hg * p.bs * 128;
                        tma_load_3d(sq + off, &tm_q, b_qfull(s), (kvh * p.g + hg) * D + kbk * 64, n * p.bs, b);
                        tma_load_3d(sq + C::QT_BYTES + off, &tm_do, b_qfull(s), (kvh * p.g + hg) * D + kbk * 64, n * p.bs, b);
                    }
            }
        } else if (warp == 1 && lane == 0) {
            // ================= MMA issuer A: S^T = K Q^T, dP^T = V dO^T =================
            constexpr uint32_t idesc_st = make_idesc_bf16(128, C::BQ, 0, 0);
            mbar_wait(b_kvfull, 0, 32);
            tc_fence_after();
            for (int it = 0; it < n_it; ++it) {
                const int s = it % C::kStages, u = it & 1;
                mbar_wait(b_sdpfree(u), ((it >> 1) & 1) ^ 1u, 39);
                mbar_wait(b_qfull(s), (it / C::kStages) & 1, 33);
                tc_fence_after();
                const uint32_t sq = sbase + C::OFF_Q + s * 2 * C::QT_BYTES;
                const uint32_t sdo = sq + C::QT_BYTES;
#pragma unroll
                for (int kbk = 0; kbk < C::NB; ++kbk) {
                    const uint64_t a = make_smem_desc_sw128(sbase + C::OFF_K + kbk * (C::BK * 128), 0, 1024);
                    const uint64_t bq = make_smem_desc_sw128(sq + kbk * (C::BQ * 128), 0, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16<1>(tmem + C::TM_ST + u * C::BQ, a + ((k * 32) >> 4), bq + ((k * 32) >> 4), idesc_st, (kbk | k) != 0);
                }
#pragma unroll
                for (int kbk = 0; kbk < C::NB; ++kbk) {
                    const uint64_t a = make_smem_desc_sw128(sbase + C::OFF_V + kbk * (C::BK * 128), 0, 1024);
                    const uint64_t bd = make_smem_desc_sw128(sdo + kbk * (C::BQ * 128), 0, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16<1>(tmem + C::TM_DP + u * C::BQ, a + ((k * 32) >> 4), bd + ((k * 32) >> 4), idesc_st, (kbk | k) != 0);
                }
                umma_commit(b_sdpfull(u));
            }
        } else if (warp == 2 && lane == 0) {
            // ================= MMA issuer B: dV += P^T dO, dK += dS^T Q =================
            constexpr uint32_t idesc_acc = make_idesc_bf16(128, D, 0, 1);
            for (int it = 0; it < n_it; ++it) {
                const int s = it % C::kStages;
                mbar_wait(b_qfull(s), (it / C::kStages) & 1, 33);
                mbar_wait(b_pdsfull, it & 1, 34);
                tc_fence_after();
                const uint32_t sq = sbase + C::OFF_Q + s * 2 * C::QT_BYTES;
                const uint32_t sdo = sq + C::QT_BYTES;
                const uint64_t ap = make_smem_desc_sw128(sbase + C::OFF_P, 0, 1024);
                const uint64_t ads = make_smem_desc_sw128(sbase + C::OFF_DS, 0, 1024);
                const uint64_t bdo = make_smem_desc_sw128(sdo, C::BQ * 128, 1024);
                const uint64_t bq = make_smem_desc_sw128(sq, C::BQ * 128, 1024);
#pragma unroll
                for (int k = 0; k < C::BQ / 16; ++k)
                    umma_bf16<1>(tmem + C::TM_DV, ap + ((k * 32) >> 4), bdo + ((k * 2048) >> 4), idesc_acc, (it | k) != 0);
#pragma unroll
                for (int k = 0; k < C::BQ / 16; ++k)
                    umma_bf16<1>(tmem + C::TM_DK, ads + ((k * 32) >> 4), bq + ((k * 2048) >> 4), idesc_acc, (it | k) != 0);
                umma_commit(b_mmadone);
                umma_commit(b_qempty(s));
            }
        }
    } else {
        // ================= P^T / dS^T warpgroups: one thread per key row, WG x handles q columns [32x, 32x+32) ===========
        const int x = (warp - 4) >> 2;
        const int wq = warp & 3;
        const int r = wq * 32 + lane;
        const int key = k0 + r;
        const uint32_t t_lane = tmem + ((uint32_t)(wq * 32) << 16);
        const float c = p.scale_log2;
        const float scale = c * kLn2;
        for (int it = 0; it < n_it; ++it) {
            const int u = it & 1;
            const int a_n = p.anchors[b * p.N + n_lo + it];          // every query column of this iteration shares the limit
            // sliding-window layer: query slot o sees this key iff key >= a_n + o - (W - 1), i.e. o <= slack
            const int slack = p.window > 0 ? key - (a_n - (p.window - 1)) : 0x3fffffff;
            const bool key_ok = key < a_n && slack >= 0;             // (a_n <= S - 2, so key < S as well)
            mbar_wait(b_sdpfull(u), (it >> 1) & 1, 36);
            mbar_wait(b_ldfull(u), (it >> 1) & 1, 35);
            tc_fence_after();
            uint32_t sv[32], dv[32];
            tmem_ld_32x32b_x32(t_lane + C::TM_ST + u * C::BQ + x * 32, sv);
            tmem_ld_32x32b_x32(t_lane + C::TM_DP + u * C::BQ + x * 32, dv);
            const float4* L4 = reinterpret_cast<const float4*>(sgen + C::OFF_LD) + (u * 2 * C::BQ + x * 32) / 4;
            const float4* D4 = L4 + C::BQ / 4;
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_sdpfree(u));
            uint32_t pk[16], dk[16];
            if (!key_ok) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { pk[j] = 0u; dk[j] = 0u; }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 l4 = L4[j], d4 = D4[j];
                    const float Lv[4] = {l4.x, l4.y, l4.z, l4.w}, Dv[4] = {d4.x, d4.y, d4.z, d4.w};
                    float pv[4], dsv[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float pr = ex2_approx(fmaf(__uint_as_float(sv[j * 4 + e]), c, -Lv[e]));
                        if ((x * 32 + j * 4 + e) % p.bs > slack) pr = 0.f;       // query column = hg * bs + o (never true for full layers)
                        pv[e] = pr;
                        dsv[e] = pr * (__uint_as_float(dv[j * 4 + e]) - Dv[e]) * scale;
                    }
                    pk[j * 2] = pack_bf16x2(pv[0], pv[1]); pk[j * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
                    dk[j * 2] = pack_bf16x2(dsv[0], dsv[1]); dk[j * 2 + 1] = pack_bf16x2(dsv[2], dsv[3]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(b_ldempty(u));
            if (it > 0) mbar_wait(b_mmadone, (it - 1) & 1, 37);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                *reinterpret_cast<uint4*>(sgen + C::OFF_P + sw128c(r, x * 4 + j)) = make_uint4(pk[j * 4], pk[j * 4 + 1], pk[j * 4 + 2], pk[j * 4 + 3]);
                *reinterpret_cast<uint4*>(sgen + C::OFF_DS + sw128c(r, x * 4 + j)) = make_uint4(dk[j * 4], dk[j * 4 + 1], dk[j * 4 + 2], dk[j * 4 + 3]);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_pdsfull);
        }
        // ---- epilogue: this CTA owns the tile's rows for this kv head: plain bf16 stores
        mbar_wait(b_mmadone, (n_it - 1) & 1, 38);
        tc_fence_after();
        const bool row_ok = key < p.S;
        __nv_bfloat16* dkp = p.dkc + ((int64_t)b * p.S + min(key, p.S - 1)) * p.lddkc + kvh * D;
        __nv_bfloat16* dvp = p.dvc + ((int64_t)b * p.S + min(key, p.S - 1)) * p.lddvc + kvh * D;
#pragma unroll
        for (int cc = x * (D / 64); cc < (x + 1) * (D / 64); ++cc) {
            uint32_t a[32], v[32];
            tmem_ld_32x32b_x32(t_lane + C::TM_DK + cc * 32, a);
            tmem_ld_32x32b_x32(t_lane + C::TM_DV + cc * 32, v);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 ok, ov;
                    ok.x = pack_bf16x2(__uint_as_float(a[q * 8 + 0]), __uint_as_float(a[q * 8 + 1])); ok.y = pack_bf16x2(__uint_as_float(a[q * 8 + 2]), __uint_as_float(a[q * 8 + 3]));
                    ok.z = pack_bf16x2(__uint_as_float(a[q * 8 + 4]), __uint_as_float(a[q * 8 + 5])); ok.w = pack_bf16x2(__uint_as_float(a[q * 8 + 6]), __uint_as_float(a[q * 8 + 7]));
                    ov.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1])); ov.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
                    ov.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5])); ov.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
                    reinterpret_cast<uint4*>(dkp + cc * 32)[q] = ok;
                    reinterpret_cast<uint4*>(dvp + cc * 32)[q] = ov;
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc<1>(tmem, 512); }
}

// ============================================================================================ delta and own-block dK / dV
// delta[(row, head)] = sum_c dO * O : one warp per (draft row, head)
__global__ void __launch_bounds__(256) df_delta_kernel(const __nv_bfloat16* __restrict__ o, int64_t ldo, const __nv_bfloat16* __restrict__ g,
                                                       int64_t ldg, int nh, int d, int64_t Mq, float* __restrict__ delta) {
    const int64_t item = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= Mq * nh) return;
    const int h = (int)(item % nh);
    const int64_t r = item / nh;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += __bfloat162float(o[r * ldo + h * d + c]) * __bfloat162float(g[r * ldg + h * d + c]);
    s = warp_sum(s);
    if (lane == 0) delta[r * nh + h] = s;
}

// dK / dV of a block's own bs noise keys: CTA = (block, kv head, sequence), R = g*bs query rows; P and dS of the own tile are
// recomputed from lse / delta (fp32, CUDA cores: R x bs x d work).  16-byte global loads, fp32 shared-memory tiles with a row
// stride of d + 4 floats (float4-aligned, conflict-free for the access patterns below), float4 inner products: the first version
// (element-wise 2-byte loads, scalar shared-memory reads, two LDS per FMA) took 3.1 ms per launch at the BASELINE shape — 5 % of
// the DFlash step for 16 GFLOP of work.
__global__ void __launch_bounds__(256) df_bwd_own_kernel(AttnArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int n = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
    const int d = a.d, bs = a.bs, g = a.nh / a.nkv, R = g * bs, ds = d + 4, ts = bs + 1, d8 = d / 8, d4 = d / 4;
    float* Qs = sm;                 // [R][ds]
    float* Gs = Qs + R * ds;        // [R][ds]
    float* Ks = Gs + R * ds;        // [bs][ds]
    float* Vs = Ks + bs * ds;       // [bs][ds]
    float* Ps = Vs + bs * ds;       // [R][ts]
    float* Ds = Ps + R * ts;        // [R][ts]
    const int t = threadIdx.x;
    const int64_t qrow0 = ((int64_t)b * a.N + n) * bs;
    if (!a.keep[b * a.N + n]) {
        for (int i = t; i < bs * d8; i += 256) {
            const int k = i / d8, c8 = i % d8;
            reinterpret_cast<uint4*>(a.dkn + (qrow0 + k) * a.lddkn + kvh * d)[c8] = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4*>(a.dvn + (qrow0 + k) * a.lddvn + kvh * d)[c8] = make_uint4(0, 0, 0, 0);
        }
        return;
    }
    auto put8 = [](float* dst, uint4 raw) {
        const __nv_bfloat162* e = reinterpret_cast<const __nv_bfloat162*>(&raw);
        const float2 f0 = __bfloat1622float2(e[0]), f1 = __bfloat1622float2(e[1]), f2 = __bfloat1622float2(e[2]), f3 = __bfloat1622float2(e[3]);
        reinterpret_cast<float4*>(dst)[0] = make_float4(f0.x, f0.y, f1.x, f1.y);
        reinterpret_cast<float4*>(dst)[1] = make_float4(f2.x, f2.y, f3.x, f3.y);
    };
    for (int i = t; i < R * d8; i += 256) {
        const int rr = i / d8, c8 = i % d8;
        const int64_t row = qrow0 + rr % bs;
        const int hh = kvh * g + rr / bs;
        put8(Qs + rr * ds + c8 * 8, __ldg(reinterpret_cast<const uint4*>(a.q + row * a.ldq + hh * d) + c8));
        put8(Gs + rr * ds + c8 * 8, __ldg(reinterpret_cast<const uint4*>(a.dout + row * a.lddo + hh * d) + c8));
    }
    for (int i = t; i < bs * d8; i += 256) {
        const int k = i / d8, c8 = i % d8;
        put8(Ks + k * ds + c8 * 8, __ldg(reinterpret_cast<const uint4*>(a.kn + (qrow0 + k) * a.ldkn + kvh * d) + c8));
        put8(Vs + k * ds + c8 * 8, __ldg(reinterpret_cast<const uint4*>(a.vn + (qrow0 + k) * a.ldvn + kvh * d) + c8));
    }
    __syncthreads();
    // P and dS of the own tile.  The kernel is shared-memory-bandwidth bound (2 MB of tile reads per CTA with one pair per thread),
    // so every thread computes a 2 x 2 block of (query row, key) pairs from 8 float4 loads per 4 columns.  Tile order: the 4 lanes
    // of a quad take 4 key pairs of the same row pair (rows broadcast, keys 8 banks apart), the quads walk the row pairs.
    const int nrt = R / 2;
    for (int tile = t; tile < nrt * (bs / 2); tile += 256) {
        const int kt = (tile & 3) + 4 * (tile / (4 * nrt)), rt = (tile >> 2) % nrt;
        const int rr0 = rt * 2, k0 = kt * 2;
        float s[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, dp[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        for (int c = 0; c < d4; ++c) {
            float4 qa[2], ga[2], ka[2], va[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                qa[u] = reinterpret_cast<const float4*>(Qs + (rr0 + u) * ds)[c];
                ga[u] = reinterpret_cast<const float4*>(Gs + (rr0 + u) * ds)[c];
                ka[u] = reinterpret_cast<const float4*>(Ks + (k0 + u) * ds)[c];
                va[u] = reinterpret_cast<const float4*>(Vs + (k0 + u) * ds)[c];
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    s[u][v] += qa[u].x * ka[v].x + qa[u].y * ka[v].y + qa[u].z * ka[v].z + qa[u].w * ka[v].w;
                    dp[u][v] += ga[u].x * va[v].x + ga[u].y * va[v].y + ga[u].z * va[v].z + ga[u].w * va[v].w;
                }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int rr = rr0 + u;
            const int64_t orow = qrow0 + rr % bs;
            const int head = kvh * g + rr / bs;
            const float lse = a.lse[orow * a.nh + head], del = a.delta[orow * a.nh + head];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const int k = k0 + v;
                const bool allowed = a.window == 0 || k <= rr % bs;      // sliding-window layer: own slots <= the query's slot
                const float pr = allowed ? __expf(s[u][v] * a.scale - lse) : 0.f;
                Ps[rr * ts + k] = pr;
                Ds[rr * ts + k] = pr * (dp[u][v] - del) * a.scale;
            }
        }
    }
    __syncthreads();
    // dK[k, c..c+3] = sum_rr dS[rr, k] Q[rr, c..c+3], dV likewise from P and dO: two keys and four columns per thread and pass
    for (int i = t; i < (bs / 2) * d4; i += 256) {
        const int k0 = (i / d4) * 2, c4 = i % d4;
        float4 dk[2], dv[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) { dk[v] = make_float4(0.f, 0.f, 0.f, 0.f); dv[v] = dk[v]; }
        for (int rr = 0; rr < R; ++rr) {
            const float4 qa = reinterpret_cast<const float4*>(Qs + rr * ds)[c4], ga = reinterpret_cast<const float4*>(Gs + rr * ds)[c4];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const float dsv = Ds[rr * ts + k0 + v], pv = Ps[rr * ts + k0 + v];
                dk[v].x += dsv * qa.x; dk[v].y += dsv * qa.y; dk[v].z += dsv * qa.z; dk[v].w += dsv * qa.w;
                dv[v].x += pv * ga.x; dv[v].y += pv * ga.y; dv[v].z += pv * ga.z; dv[v].w += pv * ga.w;
            }
        }
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            uint2 ok, ov;
            ok.x = pack_bf16x2(dk[v].x, dk[v].y); ok.y = pack_bf16x2(dk[v].z, dk[v].w);
            ov.x = pack_bf16x2(dv[v].x, dv[v].y); ov.y = pack_bf16x2(dv[v].z, dv[v].w);
            reinterpret_cast<uint2*>(a.dkn + (qrow0 + k0 + v) * a.lddkn + kvh * d)[c4] = ok;
            reinterpret_cast<uint2*>(a.dvn + (qrow0 + k0 + v) * a.lddvn + kvh * d)[c4] = ov;
        }
    }
}

// ============================================================================================ host
static int g_own_smem_set = 0;
template <int D>
static int bwd_tc_t(const AttnArgs& a, cudaStream_t st) {
    const int g = a.nh / a.nkv, R = g * a.bs;
    const int64_t Q = (int64_t)a.N * a.bs, Mq = (int64_t)a.B * Q, A = (int64_t)a.nh * D;
    CUtensorMap tq, tdo, tkc64, tkc128, tvc64, tvc128, tkn, tvn;
    const int64_t KVc = (int64_t)a.nkv * D;      // maps cover the nkv*D columns of each (possibly strided) view
    SF_TRY_RC(make_tmap_3d_bf16(&tq, a.q, A, Q, a.B, a.ldq, a.bs));
    SF_TRY_RC(make_tmap_3d_bf16(&tdo, a.dout, A, Q, a.B, a.lddo, a.bs));
    SF_TRY_RC(make_tmap_3d_bf16(&tkc64, a.kc, KVc, a.S, a.B, a.ldkc, 64));
    SF_TRY_RC(make_tmap_3d_bf16(&tkc128, a.kc, KVc, a.S, a.B, a.ldkc, 128));
    SF_TRY_RC(make_tmap_3d_bf16(&tvc64, a.vc, KVc, a.S, a.B, a.ldvc, 64));
    SF_TRY_RC(make_tmap_3d_bf16(&tvc128, a.vc, KVc, a.S, a.B, a.ldvc, 128));
    SF_TRY_RC(make_tmap_3d_bf16(&tkn, a.kn, KVc, Q, a.B, a.ldkn, 64));
    SF_TRY_RC(make_tmap_3d_bf16(&tvn, a.vn, KVc, Q, a.B, a.ldvn, 64));
    TcBwdParams p{};
    p.lse = a.lse; p.delta = a.delta; p.anchors = a.anchors; p.keep = a.keep;
    p.dq = a.dq; p.lddq = a.lddq; p.dkc = a.dkc; p.lddkc = a.lddkc; p.dvc = a.dvc; p.lddvc = a.lddvc;
    p.B = a.B; p.S = a.S; p.N = a.N; p.bs = a.bs; p.nh = a.nh; p.nkv = a.nkv; p.g = g; p.window = a.window;
    p.scale_log2 = a.scale * kLog2e;
    df_delta_kernel<<<(unsigned)((Mq * a.nh * 32 + 255) / 256), 256, 0, st>>>(a.out, a.ldo, a.dout, a.lddo, a.nh, D, Mq, a.delta);
    SF_CUDA_CHECK_LAUNCH("dflash delta");
    {
        using C = DqCfg<D>;
        static bool set = false;
        if (!set) {
            cudaError_t e = cudaFuncSetAttribute(df_bwd_dq_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
            if (e != cudaSuccess) return set_error(-22, "dflash bwd_dq_tc smem attr: %s", cudaGetErrorString(e));
            set = true;
        }
        const int BPU = C::BQ / R;
        df_bwd_dq_tc_kernel<D><<<dim3((a.N + BPU - 1) / BPU, a.nkv, a.B), 384, C::SMEM, st>>>(tq, tdo, tkc64, tvc64, tkn, tvn, p);
        SF_CUDA_CHECK_LAUNCH("dflash bwd_dq_tc");
    }
    {
        const int smem = (int)((2 * R * (D + 4) + 2 * a.bs * (D + 4) + 2 * R * (a.bs + 1)) * 4);
        // one high-water mark for both head dims: the kernel is not a template, so a per-instantiation mark let the d = 64 call
        // lower the limit a d = 128 call had raised ("invalid argument" on the next d = 128 launch)
        if (smem > g_own_smem_set) {
            cudaError_t e = cudaFuncSetAttribute(df_bwd_own_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return set_error(-22, "dflash bwd_own smem attr: %s", cudaGetErrorString(e));
            g_own_smem_set = smem;
        }
        df_bwd_own_kernel<<<dim3(a.N, a.nkv, a.B), 256, smem, st>>>(a);
        SF_CUDA_CHECK_LAUNCH("dflash bwd_own");
    }
    {
        using C = CtxCfg<D>;
        static bool set = false;
        if (!set) {
            cudaError_t e = cudaFuncSetAttribute(df_bwd_ctx_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
            if (e != cudaSuccess) return set_error(-22, "dflash bwd_ctx_tc smem attr: %s", cudaGetErrorString(e));
            set = true;
        }
        df_bwd_ctx_tc_kernel<D><<<dim3((a.S + C::BK - 1) / C::BK, a.nkv, a.B), 384, C::SMEM, st>>>(tq, tdo, tkc128, tvc128, p);
        SF_CUDA_CHECK_LAUNCH("dflash bwd_ctx_tc");
    }
    return 0;
}

// the backward additionally needs one anchor block = 64 query rows (the context kernel iterates per block)
bool attn_tc_bwd_supported(const AttnArgs& a) { return attn_tc_supported(a) && (a.nh / a.nkv) * a.bs == 64; }
int attn_bwd_tc(const AttnArgs& a, cudaStream_t st) { return a.d == 128 ? bwd_tc_t<128>(a, st) : bwd_tc_t<64>(a, st); }

}  // namespace dflash
}  // namespace sf
