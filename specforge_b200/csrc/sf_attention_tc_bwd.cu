// sf_attention_tc_bwd.cu — TTT attention backward over block 0 on tcgen05 (TMEM accumulators, TMA-fed smem).
//
// Two kernels (no atomics, deterministic), mirroring sf_attention.cu's split:
//   attn_bwd_dkv_tc : CTA = (128 keys, kv head, batch).  Loops the g query heads of the group x 64-row query tiles at or
//                     below the diagonal.  S^T = K Q^T and dP^T = V dO^T land in TMEM with KEYS as rows, so one thread
//                     per key row turns them into P^T / dS^T (bf16, SWIZZLE_128B K-major smem tiles = A operands) and
//                     dV += P^T dO, dK += dS^T Q accumulate in TMEM across the whole loop.
//   attn_bwd_dq_tc  : CTA = (128 queries, head, batch).  S = Q K^T, dP = dO V^T per 64-key tile, one thread per query
//                     row writes dS, dQ += dS K accumulates in TMEM.
// A [rows x 64] bf16 TMA box is simultaneously a K-major operand over its columns and an MN-major operand over its
// rows, so each Q / dO / K tile is loaded once and consumed through two descriptors.
#include "sf_gemm.cuh"
#include "sf_host.h"

namespace sf {

int make_tmap_3d_bf16(CUtensorMap* tm, const void* base, int64_t cols, int64_t rows, int64_t batches, int64_t ld, int box_rows);

template <int kRegs> __device__ __forceinline__ void reg_inc_b() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void reg_dec_b() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
__device__ __forceinline__ uint32_t sw128b(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

struct AttnBwdTcParams {
    const float* lse; const float* delta;     // [B, nh, S]
    const int* kvlen; const uint8_t* key_mask; const int* nonprefix;
    float* dk_acc; float* dv_acc; int64_t ldacc;      // fp32 block-0 accumulators [B*S, nkv*D]
    __nv_bfloat16* dq; int64_t lddq; const float* dq_diag;   // dq out (bf16) + optional fp32 diagonal-term dq [B*S, nh*D]
    int B, S, nh, nkv;
    int q_col0, k_col0, v_col0;
    float scale_log2;
};

// ============================================================================================ dK / dV
template <int D>
struct DkvCfg {
    static constexpr int BK = 128, BQ = 64, NB = D / 64, kStages = 4;   // Q/dO ring: loads run 3 tiles ahead (TMA latency)
    static constexpr int KT_BYTES = BK * D * 2;            // K (or V) tile
    static constexpr int QT_BYTES = BQ * D * 2;            // Q (or dO) stage
    static constexpr int PT_BYTES = BK * BQ * 2;           // P^T (or dS^T) tile
    static constexpr int OFF_K = 0, OFF_V = KT_BYTES;
    static constexpr int OFF_Q = 2 * KT_BYTES;             // [stage] Q then dO
    static constexpr int OFF_P = OFF_Q + kStages * 2 * QT_BYTES;
    static constexpr int OFF_DS = OFF_P + PT_BYTES;
    static constexpr int OFF_LD = OFF_DS + PT_BYTES;       // [2 buffers][L(64) | delta(64)] floats
    static constexpr int OFF_BAR = OFF_LD + 2 * 2 * BQ * 4;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static constexpr int TM_ST = 0, TM_DP = 128, TM_DV = 256, TM_DK = 256 + D;   // S^T[2], dP^T[2] double-buffered
};

template <int D>
__global__ void __launch_bounds__(384, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                       const __grid_constant__ CUtensorMap tm_kv, const AttnBwdTcParams p) {
    using C = DkvCfg<D>;
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
    const int g = p.nh / p.nkv;
    const int k0 = kb * C::BK;
    int kvlen = p.S;
    bool general_mask = false;
    if (p.kvlen) { kvlen = p.kvlen[b]; general_mask = p.nonprefix && p.nonprefix[b]; if (general_mask) kvlen = p.S; }
    if (k0 >= kvlen) return;                               // every key of this tile is masked: contributes nothing
    const int first_qt = k0 / C::BQ;
    const int n_qt = (p.S + C::BQ - 1) / C::BQ;
    const int per_head = n_qt - first_qt;
    const int n_it = per_head * g;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_kv); }
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
            // ================= L / delta stager: 64 + 64 floats per iteration into a 2-deep smem ring =================
            for (int it = 0; it < n_it; ++it) {
                const int u = it & 1;
                const int h = kvh * g + it / per_head;
                const int q0 = (first_qt + it % per_head) * C::BQ;
                mbar_wait(b_ldempty(u), ((it >> 1) & 1) ^ 1u, 30);
                float* dst = reinterpret_cast<float*>(sgen + C::OFF_LD) + u * 2 * C::BQ;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int qi = q0 + e * 32 + lane;
                    const int64_t idx = ((int64_t)b * p.nh + h) * p.S + qi;
                    dst[e * 32 + lane] = qi < p.S ? p.lse[idx] : INFINITY;
                    dst[C::BQ + e * 32 + lane] = qi < p.S ? p.delta[idx] : 0.f;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(b_ldfull(u));
            }
        } else if (warp == 0 && lane == 0) {
            // ================= TMA producer =================
            mbar_expect_tx(b_kvfull, 2 * C::KT_BYTES);
            for (int kbk = 0; kbk < C::NB; ++kbk) {
                tma_load_3d(sbase + C::OFF_K + kbk * (C::BK * 128), &tm_kv, b_kvfull, p.k_col0 + kvh * D + kbk * 64, k0, b);
                tma_load_3d(sbase + C::OFF_V + kbk * (C::BK * 128), &tm_kv, b_kvfull, p.v_col0 + kvh * D + kbk * 64, k0, b);
            }
            for (int it = 0; it < n_it; ++it) {
                const int s = it % C::kStages;
                const uint32_t ph = ((it / C::kStages) & 1) ^ 1u;
                const int h = kvh * g + it / per_head;
                const int q0 = (first_qt + it % per_head) * C::BQ;
                mbar_wait(b_qempty(s), ph, 31);
                mbar_expect_tx(b_qfull(s), 2 * C::QT_BYTES);
                const uint32_t sq = sbase + C::OFF_Q + s * 2 * C::QT_BYTES;
                for (int kbk = 0; kbk < C::NB; ++kbk) {
                    tma_load_3d(sq + kbk * (C::BQ * 128), &tm_q, b_qfull(s), p.q_col0 + h * D + kbk * 64, q0, b);
                    tma_load_3d(sq + C::QT_BYTES + kbk * (C::BQ * 128), &tm_do, b_qfull(s), h * D + kbk * 64, q0, b);
                }
            }
        } else if (warp == 1 && lane == 0) {
            // ================= MMA issuer A: S^T = K Q^T, dP^T = V dO^T (runs ahead, double-buffered in TMEM) =========
            // Two issuer threads because tcgen05.mma issue blocks while the in-order tensor queue is full: the thread
            // that feeds S^T/dP^T batches must not delay the accumulate MMAs of a finished P^T/dS^T tile (issuer B).
            constexpr uint32_t idesc_st = make_idesc_bf16(128, C::BQ, 0, 0);   // [128 keys x 64 q] = K(K-major) Q^T(K-major)
            mbar_wait(b_kvfull, 0, 32);
            tc_fence_after();
            for (int it = 0; it < n_it; ++it) {
                const int s = it % C::kStages, u = it & 1;
                mbar_wait(b_sdpfree(u), ((it >> 1) & 1) ^ 1u, 39);     // warpgroups finished reading S^T/dP^T(it-2)
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
            constexpr uint32_t idesc_acc = make_idesc_bf16(128, D, 0, 1);      // [128 x D] += A(K-major) B(MN-major)
            for (int it = 0; it < n_it; ++it) {
                const int s = it % C::kStages;
                mbar_wait(b_qfull(s), (it / C::kStages) & 1, 33);
                mbar_wait(b_pdsfull, it & 1, 34);
                tc_fence_after();
                const uint32_t sq = sbase + C::OFF_Q + s * 2 * C::QT_BYTES;
                const uint32_t sdo = sq + C::QT_BYTES;
                const uint64_t ap = make_smem_desc_sw128(sbase + C::OFF_P, 0, 1024);
                const uint64_t ads = make_smem_desc_sw128(sbase + C::OFF_DS, 0, 1024);
                const uint64_t bdo = make_smem_desc_sw128(sdo, C::BQ * 128, 1024);    // MN-major over the 64 q rows
                const uint64_t bq = make_smem_desc_sw128(sq, C::BQ * 128, 1024);
#pragma unroll
                for (int k = 0; k < C::BQ / 16; ++k)
                    umma_bf16<1>(tmem + C::TM_DV, ap + ((k * 32) >> 4), bdo + ((k * 2048) >> 4), idesc_acc, (it | k) != 0);
#pragma unroll
                for (int k = 0; k < C::BQ / 16; ++k)
                    umma_bf16<1>(tmem + C::TM_DK, ads + ((k * 32) >> 4), bq + ((k * 2048) >> 4), idesc_acc, (it | k) != 0);
                umma_commit(b_mmadone);
                umma_commit(b_qempty(s));     // S^T/dP^T of this tile completed before its P^T/dS^T existed
            }
        }
    } else {
        // ================= P^T / dS^T warpgroups: one thread per key row, WG x handles q columns [32x, 32x+32) ===========
        const int x = (warp - 4) >> 2;
        const int wq = warp & 3;
        const int r = wq * 32 + lane;
        const int key = k0 + r;
        bool key_ok = key < kvlen;
        if (key_ok && general_mask) key_ok = p.key_mask[(int64_t)b * p.S + key] != 0;
        const uint32_t t_lane = tmem + ((uint32_t)(wq * 32) << 16);
        const float c = p.scale_log2;
        const float scale = c * 0.6931471805599453f;
        for (int it = 0; it < n_it; ++it) {
            const int u = it & 1;
            const int q0 = (first_qt + it % per_head) * C::BQ;
            const bool causal_tile = q0 < k0 + C::BK - 1;         // some (key, q) pairs of this tile have key > q
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
                        float sval = __uint_as_float(sv[j * 4 + e]);
                        if (causal_tile && key > q0 + x * 32 + j * 4 + e) sval = -INFINITY;
                        const float pr = ex2_approx(fmaf(sval, c, -Lv[e]));
                        pv[e] = pr;
                        dsv[e] = pr * (__uint_as_float(dv[j * 4 + e]) - Dv[e]) * scale;
                    }
                    pk[j * 2] = pack_bf16x2(pv[0], pv[1]); pk[j * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
                    dk[j * 2] = pack_bf16x2(dsv[0], dsv[1]); dk[j * 2 + 1] = pack_bf16x2(dsv[2], dsv[3]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(b_ldempty(u));
            // the single P^T / dS^T buffers are free once the accumulate MMAs of iteration it-1 completed
            if (it > 0) mbar_wait(b_mmadone, (it - 1) & 1, 37);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                *reinterpret_cast<uint4*>(sgen + C::OFF_P + sw128b(r, x * 4 + j)) = make_uint4(pk[j * 4], pk[j * 4 + 1], pk[j * 4 + 2], pk[j * 4 + 3]);
                *reinterpret_cast<uint4*>(sgen + C::OFF_DS + sw128b(r, x * 4 + j)) = make_uint4(dk[j * 4], dk[j * 4 + 1], dk[j * 4 + 2], dk[j * 4 + 3]);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_pdsfull);
        }
        // ---- epilogue: accumulate dK / dV of this tile into the fp32 block-0 buffers (this CTA owns the rows)
        mbar_wait(b_mmadone, (n_it - 1) & 1, 38);
        tc_fence_after();
        const bool row_ok = key < p.S;
        float* dkp = p.dk_acc + ((int64_t)b * p.S + min(key, p.S - 1)) * p.ldacc + kvh * D;
        float* dvp = p.dv_acc + ((int64_t)b * p.S + min(key, p.S - 1)) * p.ldacc + kvh * D;
#pragma unroll
        for (int cc = x * (D / 64); cc < (x + 1) * (D / 64); ++cc) {
            uint32_t a[32], v[32];
            tmem_ld_32x32b_x32(t_lane + C::TM_DK + cc * 32, a);
            tmem_ld_32x32b_x32(t_lane + C::TM_DV + cc * 32, v);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 x = reinterpret_cast<float4*>(dkp + cc * 32)[q];
                    x.x += __uint_as_float(a[q * 4]); x.y += __uint_as_float(a[q * 4 + 1]); x.z += __uint_as_float(a[q * 4 + 2]); x.w += __uint_as_float(a[q * 4 + 3]);
                    reinterpret_cast<float4*>(dkp + cc * 32)[q] = x;
                    float4 y = reinterpret_cast<float4*>(dvp + cc * 32)[q];
                    y.x += __uint_as_float(v[q * 4]); y.y += __uint_as_float(v[q * 4 + 1]); y.z += __uint_as_float(v[q * 4 + 2]); y.w += __uint_as_float(v[q * 4 + 3]);
                    reinterpret_cast<float4*>(dvp + cc * 32)[q] = y;
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc<1>(tmem, 512); }
}

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
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                      const __grid_constant__ CUtensorMap tm_kv, const AttnBwdTcParams p) {
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
    const int qb = gridDim.x - 1 - blockIdx.x;
    const int h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / (p.nh / p.nkv);
    const int q0 = qb * C::BQ;
    int kvlen = p.S;
    bool general_mask = false;
    if (p.kvlen) { kvlen = p.kvlen[b]; general_mask = p.nonprefix && p.nonprefix[b]; if (general_mask) kvlen = p.S; }
    int n_kv = (min(q0 + C::BQ, p.S) + C::BKV - 1) / C::BKV;
    { const int lim = max(1, (kvlen + C::BKV - 1) / C::BKV); if (n_kv > lim) n_kv = lim; }

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_kv); }
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
            mbar_expect_tx(b_qfull, 2 * C::QT_BYTES);
            for (int kbk = 0; kbk < C::NB; ++kbk) {
                tma_load_3d(sbase + C::OFF_Q + kbk * (C::BQ * 128), &tm_q, b_qfull, p.q_col0 + h * D + kbk * 64, q0, b);
                tma_load_3d(sbase + C::OFF_DO + kbk * (C::BQ * 128), &tm_do, b_qfull, h * D + kbk * 64, q0, b);
            }
            for (int t = 0; t < n_kv; ++t) {
                const int sk = t % C::kKStages;
                mbar_wait(b_kempty(sk), ((t / C::kKStages) & 1) ^ 1u, 41);
                mbar_expect_tx(b_kfull(sk), C::KT_BYTES);
                for (int kbk = 0; kbk < C::NB; ++kbk)
                    tma_load_3d(sbase + C::OFF_K + sk * C::KT_BYTES + kbk * (C::BKV * 128), &tm_kv, b_kfull(sk),
                                p.k_col0 + kvh * D + kbk * 64, t * C::BKV, b);
            }
        } else if (warp == 3 && lane == 0) {
            for (int t = 0; t < n_kv; ++t) {
                const int sv = t % C::kVStages;
                mbar_wait(b_vempty(sv), ((t / C::kVStages) & 1) ^ 1u, 42);
                mbar_expect_tx(b_vfull(sv), C::KT_BYTES);
                for (int kbk = 0; kbk < C::NB; ++kbk)
                    tma_load_3d(sbase + C::OFF_V + sv * C::KT_BYTES + kbk * (C::BKV * 128), &tm_kv, b_vfull(sv),
                                p.v_col0 + kvh * D + kbk * 64, t * C::BKV, b);
            }
        } else if (warp == 1 && lane == 0) {
            // ================= MMA issuer A: S = Q K^T, dP = dO V^T (runs ahead) =================
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
        const int row = q0 + r;
        const bool row_ok = row < p.S;
        const uint32_t t_lane = tmem + ((uint32_t)(wq * 32) << 16);
        const float c = p.scale_log2;
        const float scale = c * 0.6931471805599453f;
        const float L = row_ok ? p.lse[((int64_t)b * p.nh + h) * p.S + row] : INFINITY;
        const float Dl = row_ok ? p.delta[((int64_t)b * p.nh + h) * p.S + row] : 0.f;
        for (int t = 0; t < n_kv; ++t) {
            const int u = t & 1;
            const int kv0 = t * C::BKV;
            const bool need_mask = (kv0 + C::BKV - 1 > q0 + wq * 32) || (kv0 + C::BKV > kvlen) || general_mask;
            mbar_wait(b_sdpfull(u), (t >> 1) & 1, 47);
            tc_fence_after();
            uint32_t sv[32], dv[32];
            tmem_ld_32x32b_x32(t_lane + C::TM_S + u * C::BKV + x * 32, sv);
            tmem_ld_32x32b_x32(t_lane + C::TM_DP + u * C::BKV + x * 32, dv);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_sdpfree(u));
            if (need_mask) {
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int key = kv0 + x * 32 + e;
                    bool ok = key <= row && key < kvlen;
                    if (ok && general_mask) ok = p.key_mask[(int64_t)b * p.S + key] != 0;
                    if (!ok) sv[e] = 0xff800000u;   // -inf -> p = 0
                }
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
                *reinterpret_cast<uint4*>(sgen + C::OFF_DS + sw128b(r, x * 4 + j)) = make_uint4(dsk[j * 4], dsk[j * 4 + 1], dsk[j * 4 + 2], dsk[j * 4 + 3]);
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_dsfull);
        }
        mbar_wait(b_mmadone, (n_kv - 1) & 1, 49);
        tc_fence_after();
        __nv_bfloat16* dqp = p.dq + ((int64_t)b * p.S + min(row, p.S - 1)) * p.lddq + h * D;
        const float* ddp = p.dq_diag ? p.dq_diag + ((int64_t)b * p.S + min(row, p.S - 1)) * (int64_t)(p.nh * D) + h * D : nullptr;
#pragma unroll
        for (int cc = x * (D / 64); cc < (x + 1) * (D / 64); ++cc) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_lane + C::TM_DQ + cc * 32, v);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]);
                    if (ddp) {
                        const float4 d0 = *reinterpret_cast<const float4*>(ddp + cc * 32 + q * 8);
                        const float4 d1 = *reinterpret_cast<const float4*>(ddp + cc * 32 + q * 8 + 4);
                        f[0] += d0.x; f[1] += d0.y; f[2] += d0.z; f[3] += d0.w; f[4] += d1.x; f[5] += d1.y; f[6] += d1.z; f[7] += d1.w;
                    }
                    uint4 o;
                    o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
                    reinterpret_cast<uint4*>(dqp + cc * 32)[q] = o;
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc<1>(tmem, 512); }
}

// ============================================================================================ host
template <int D>
static int bwd_tc_t(const AttnDesc& a, cudaStream_t st) {
    CUtensorMap tq64, tq128, tdo64, tdo128, tkv64, tkv128;
    const int64_t A = (int64_t)a.nh * D;
    SF_TRY_RC(make_tmap_3d_bf16(&tq64, a.q_row_base, a.ldq, a.S, a.B, a.ldq, 64));
    SF_TRY_RC(make_tmap_3d_bf16(&tq128, a.q_row_base, a.ldq, a.S, a.B, a.ldq, 128));
    SF_TRY_RC(make_tmap_3d_bf16(&tdo64, a.dout, A, a.S, a.B, a.lddo, 64));
    SF_TRY_RC(make_tmap_3d_bf16(&tdo128, a.dout, A, a.S, a.B, a.lddo, 128));
    SF_TRY_RC(make_tmap_3d_bf16(&tkv64, a.kv_row_base, a.ldkv, a.S, a.B, a.ldkv, 64));
    SF_TRY_RC(make_tmap_3d_bf16(&tkv128, a.kv_row_base, a.ldkv, a.S, a.B, a.ldkv, 128));
    AttnBwdTcParams p{};
    p.lse = a.lse; p.delta = a.delta_ws; p.kvlen = a.kvlen; p.key_mask = a.key_mask; p.nonprefix = a.nonprefix;
    p.dk_acc = a.dk_acc[0]; p.dv_acc = a.dv_acc[0]; p.ldacc = a.ldacc;
    p.dq = (__nv_bfloat16*)a.dq; p.lddq = a.lddq; p.dq_diag = (a.J > 0) ? a.dq_diag_ws : nullptr;
    p.B = a.B; p.S = a.S; p.nh = a.nh; p.nkv = a.nkv; p.q_col0 = a.q_col0; p.k_col0 = a.k_col0; p.v_col0 = a.v_col0;
    p.scale_log2 = (1.0f / sqrtf((float)D)) * 1.4426950408889634f;
    {
        using C = DkvCfg<D>;
        static bool set = false;
        if (!set) {
            cudaError_t e = cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
            if (e != cudaSuccess) return set_error(-22, "attn_bwd_dkv_tc smem attr: %s", cudaGetErrorString(e));
            set = true;
        }
        dim3 grid((a.S + C::BK - 1) / C::BK, a.nkv, a.B);
        attn_bwd_dkv_tc_kernel<D><<<grid, 384, C::SMEM, st>>>(tq64, tdo64, tkv128, p);
        SF_CUDA_CHECK_LAUNCH("attn_bwd_dkv_tc");
    }
    {
        using C = DqCfg<D>;
        static bool set = false;
        if (!set) {
            cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
            if (e != cudaSuccess) return set_error(-22, "attn_bwd_dq_tc smem attr: %s", cudaGetErrorString(e));
            set = true;
        }
        dim3 grid((a.S + C::BQ - 1) / C::BQ, a.nh, a.B);
        attn_bwd_dq_tc_kernel<D><<<grid, 384, C::SMEM, st>>>(tq128, tdo128, tkv64, p);
        SF_CUDA_CHECK_LAUNCH("attn_bwd_dq_tc");
    }
    return 0;
}

int attn_bwd_tc(const AttnDesc& a, cudaStream_t st) {
    return a.head_dim == 128 ? bwd_tc_t<128>(a, st) : bwd_tc_t<64>(a, st);
}

}  // namespace sf
