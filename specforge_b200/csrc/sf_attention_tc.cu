// sf_attention_tc.cu — TTT attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), forward.
//
// Same semantics as sf_attention.cu (block-0 causal flash attention whose running softmax state is seeded with
// the T-1 diagonal scores; reference specforge/modeling/draft/llama3_eagle.py:717-785).
//
// CTA = (128 query rows) x (one or two query heads of the same GQA group) x batch element.  Both heads share every K/V tile
// (K/V smem traffic halved); each head has its own MMA issuer thread and softmax warpgroup, so one head's exponentials run
// under the other head's MMAs.
//   warp 0     TMA producer   Q tiles once + the K ring (4 stages; 3-D tensor maps [cols, S, B]: rows past the end of a
//                             sequence are zero-filled, never another sequence's data)
//   warp 3     TMA producer   V ring (own thread: a late PV never delays the K loads)
//   warps 1,2  MMA issuers    one thread per head: S = Q K^T -> TMEM (double-buffered), O += P V -> TMEM (accumulated
//                             across all kv tiles); warp 2 also allocates the 512 TMEM columns
//   warps 4-7  softmax warpgroup of head a: one thread per query row (= TMEM lane), running max / sum in registers, P written
//   warps 8-11 same for head b               as bf16 into a SWIZZLE_128B K-major smem tile (the A operand of the PV MMA); O
//                                            stays in TMEM and is rescaled lazily (only when the running max grew by > 2^8).
#include "sf_gemm.cuh"   // CUtensorMap, make_* helpers
#include "sf_host.h"
#include <cudaTypedefs.h>
#include <cstdlib>

namespace sf {

int make_tmap_3d_bf16(CUtensorMap* tm, const void* base, int64_t cols, int64_t rows, int64_t batches, int64_t ld,
                      int box_rows);

template <int kRegs> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }

// byte offset of the 16-byte chunk `chunk` of row `row` inside a [rows x 64] bf16 SWIZZLE_128B block
__device__ __forceinline__ uint32_t sw128(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

struct AttnTcParams {
    __nv_bfloat16* out; int64_t ldo;
    float* lse;                    // [B, nh, S] (log2 domain)
    const float* sd;               // [B, nh, S, J] diagonal scores or null
    const __nv_bfloat16* vdiag[8]; int64_t ldkv;
    const int* kvlen;              // [B] number of leading valid keys (prefix masks), or null
    const uint8_t* key_mask;       // [B, S] general mask (used only for batch rows flagged in `nonprefix`)
    const int* nonprefix;          // [B] 1 if the mask of this row is not a prefix, or null
    int B, S, nh, nkv, J;
    int heads_per_cta;             // 1 or 2
    int q_col0, k_col0, v_col0;    // column of head 0 inside the fused qkv row
    float scale_log2;
};

template <int D>
struct FwdCfg {
    static constexpr int BQ = 128, BKV = 64, NB = D / 64;     // NB = 64-column blocks per head row
    static constexpr int Q_BYTES = BQ * D * 2;                  // per head
    static constexpr int KV_BYTES = BKV * D * 2;
    static constexpr int P_BYTES = BQ * BKV * 2;
    static constexpr int kStages = 4;         // K and V rings: loads run 3 tiles ahead of the MMAs (TMA latency >> one tile)
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = 2 * Q_BYTES;
    static constexpr int OFF_V = OFF_K + kStages * KV_BYTES;
    static constexpr int OFF_P = OFF_V + kStages * KV_BYTES;   // [head], single-buffered
    static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
    static constexpr int SMEM = OFF_BAR + 512 + 1024;
    static constexpr int TM_S = 0;            // S[head][buf] at col (head*2+buf)*64
    static constexpr int TM_O = 256;          // O[head] at 256 + head*D  (accumulated across kv tiles)
    static constexpr float kRescaleThreshold = 8.0f;   // log2 units: P stays <= 2^8 between lazy rescales
};

// Forward.  O accumulates in TMEM across the kv tiles (PV MMAs with accumulate); the softmax threads keep a
// *reference* max m_ref and only rescale O / l when the running max exceeds it by > 2^8 (exact: softmax is
// shift invariant, the final division by l uses the same reference) — the per-tile TMEM round trip of O is gone.
// S and P are double-buffered per head so the MMA issuer computes S(t+1) while the warpgroup exponentiates S(t).
template <int D>
__global__ void __launch_bounds__(384, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const AttnTcParams p) {
    using C = FwdCfg<D>;
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
    const int qb = gridDim.x - 1 - blockIdx.x;
    const int b = blockIdx.z;
    const int nx = p.heads_per_cta;
    const int h0 = blockIdx.y * nx;
    const int kvh = h0 / (p.nh / p.nkv);
    const int q0 = qb * C::BQ;
    int kvlen = p.S;
    bool general_mask = false;
    if (p.kvlen) { kvlen = p.kvlen[b]; general_mask = p.nonprefix && p.nonprefix[b]; if (general_mask) kvlen = p.S; }
    int n_kv = (min(q0 + C::BQ, p.S) + C::BKV - 1) / C::BKV;
    { const int lim = max(1, (kvlen + C::BKV - 1) / C::BKV); if (n_kv > lim) n_kv = lim; }

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv); }
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
            // ================= TMA producer =================
            mbar_expect_tx(b_qfull, C::Q_BYTES * nx);
            for (int x = 0; x < nx; ++x)
                for (int kb = 0; kb < C::NB; ++kb)
                    tma_load_3d(sbase + C::OFF_Q + x * C::Q_BYTES + kb * (C::BQ * 128), &tm_q, b_qfull,
                                p.q_col0 + (h0 + x) * D + kb * 64, q0, b);
            for (int t = 0; t < n_kv; ++t) {
                const int s = t % C::kStages;
                mbar_wait(b_kempty(s), ((t / C::kStages) & 1) ^ 1u, 11);
                mbar_expect_tx(b_kfull(s), C::KV_BYTES);
                for (int kb = 0; kb < C::NB; ++kb)
                    tma_load_3d(sbase + C::OFF_K + s * C::KV_BYTES + kb * (C::BKV * 128), &tm_kv, b_kfull(s),
                                p.k_col0 + kvh * D + kb * 64, t * C::BKV, b);
            }
        } else if (warp == 3 && lane == 0) {
            // ================= V producer (own thread: a late PV never delays the K loads) =================
            for (int t = 0; t < n_kv; ++t) {
                const int s = t % C::kStages;
                mbar_wait(b_vempty(s), ((t / C::kStages) & 1) ^ 1u, 12);
                mbar_expect_tx(b_vfull(s), C::KV_BYTES);
                for (int kb = 0; kb < C::NB; ++kb)
                    tma_load_3d(sbase + C::OFF_V + s * C::KV_BYTES + kb * (C::BKV * 128), &tm_kv, b_vfull(s),
                                p.v_col0 + kvh * D + kb * 64, t * C::BKV, b);
            }
        } else if ((warp == 1 || warp == 2) && lane == 0 && (warp - 1) < nx) {
            // ================= MMA issuers: one thread per head =================
            // tcgen05.mma issue blocks while the in-order tensor queue is full, so one thread serving two heads would
            // sit inside one head's S batch while the other head's finished P tile waits.  Each head gets its own
            // issuer (ordering is only needed within a head): PV_x(t) the moment P_x(t) lands, then S_x(t+2) into the
            // S buffer head x just finished reading.  K/V ring stages are released by both issuers (barrier count nx).
            const int x = warp - 1;
            constexpr uint32_t idesc_s = make_idesc_bf16(128, C::BKV, 0, 0);   // S[128 x 64]  = Q(K-major) K^T(K-major)
            constexpr uint32_t idesc_o = make_idesc_bf16(128, D, 0, 1);        // O[128 x D]  += P(K-major) V(MN-major)
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
                const uint64_t bdesc = make_smem_desc_sw128(sv, C::BKV * 128, 1024);   // MN-major: LBO = 64-col block stride
#pragma unroll
                for (int k = 0; k < C::BKV / 16; ++k)
                    umma_bf16<1>(tmem + C::TM_O + x * D, adesc + ((k * 32) >> 4), bdesc + ((k * 2048) >> 4), idesc_o, (t | k) != 0);
                umma_commit(b_pvdone(x));
                umma_commit(b_vempty(s));
                if (t + 2 < n_kv) issue_s(t + 2);
            }
        }
    } else {
        // ================= softmax warpgroups =================
        reg_inc<224>();
        const int x = (warp - 4) >> 2;          // 0: head a, 1: head b
        const int wq = warp & 3;
        const int r = wq * 32 + lane;           // query row inside the tile == TMEM lane
        if (x < nx) {
            const int h = h0 + x;
            const int row = q0 + r;
            const bool row_ok = row < p.S;
            const uint32_t t_lane = tmem + ((uint32_t)(wq * 32) << 16);
            const uint32_t t_o = t_lane + C::TM_O + x * D;
            float m_ref = -INFINITY, l = 0.f;
            const float* sdr = nullptr;
            if (p.J > 0) {
                sdr = p.sd + (((int64_t)b * p.nh + h) * p.S + min(row, p.S - 1)) * p.J;
                for (int i = 0; i < p.J; ++i) m_ref = fmaxf(m_ref, sdr[i]);
                for (int i = 0; i < p.J; ++i) l += exp2f(sdr[i] - m_ref);
            }
            const float c = p.scale_log2;
            for (int t = 0; t < n_kv; ++t) {
                const int u = t & 1;
                const int kv0 = t * C::BKV;
                const bool need_mask = (kv0 + C::BKV - 1 > q0 + wq * 32) || (kv0 + C::BKV > kvlen) || general_mask;
                mbar_wait(b_sfull(x, u), (t >> 1) & 1, 22 + x);
                tc_fence_after();
                uint32_t sv[C::BKV];
                tmem_ld_32x32b_x32(t_lane + C::TM_S + (x * 2 + u) * C::BKV, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
                tmem_ld_32x32b_x32(t_lane + C::TM_S + (x * 2 + u) * C::BKV + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
                tmem_ld_wait();
                float mx = -INFINITY;
                if (need_mask) {
#pragma unroll
                    for (int cc = 0; cc < C::BKV; ++cc) {
                        const int key = kv0 + cc;
                        bool ok = key <= row && key < kvlen;
                        if (ok && general_mask) ok = p.key_mask[(int64_t)b * p.S + key] != 0;
                        if (!ok) sv[cc] = 0xff800000u;   // -inf
                    }
                }
#pragma unroll
                for (int cc = 0; cc < C::BKV; ++cc) mx = fmaxf(mx, __uint_as_float(sv[cc]));
                mx *= c;                                   // c > 0
                // ---- lazy rescale (warp-collective TMEM round trip only when some row's max jumped by > 2^8)
                const bool jump = mx > m_ref + C::kRescaleThreshold;   // also true when m_ref == -inf and mx finite
                if (__any_sync(0xffffffffu, jump)) {
                    const float m_new = jump ? mx : m_ref;
                    const float f = (m_ref == -INFINITY) ? 0.f : exp2f(m_ref - m_new);   // jump==false -> 1
                    l *= f;
                    m_ref = m_new;
                    if (t > 0) {
                        // O holds tiles < t: PV(t-1) must have landed before it is rescaled
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
                // the (single) P buffer must have been consumed by PV(t-1)
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
            // ---- epilogue: O from TMEM, diagonal P*V terms (never masked), normalise, store
            mbar_wait(b_pvdone(x), (n_kv - 1) & 1, 28 + x);
            tc_fence_after();
            const float inv = (l > 0.f) ? 1.f / l : 0.f;
            __nv_bfloat16* orow = p.out + ((int64_t)b * p.S + min(row, p.S - 1)) * p.ldo + h * D;
#pragma unroll
            for (int cc = 0; cc < D / 32; ++cc) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(t_o + cc * 32, v);
                tmem_ld_wait();
                float o[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) o[e] = __uint_as_float(v[e]);
                if (row_ok) {
                    for (int i = 0; i < p.J; ++i) {
                        const float w = exp2f(sdr[i] - m_ref);
                        const uint4* vp = reinterpret_cast<const uint4*>(p.vdiag[i] + ((int64_t)b * p.S + row) * p.ldkv + kvh * D + cc * 32);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 uu = __ldg(vp + q);
                            const uint32_t wds[4] = {uu.x, uu.y, uu.z, uu.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const __nv_bfloat162 bb = *reinterpret_cast<const __nv_bfloat162*>(&wds[e]);
                                o[q * 8 + 2 * e] += w * __bfloat162float(bb.x);
                                o[q * 8 + 2 * e + 1] += w * __bfloat162float(bb.y);
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 uu;
                        uu.x = pack_bf16x2(o[q * 8 + 0] * inv, o[q * 8 + 1] * inv);
                        uu.y = pack_bf16x2(o[q * 8 + 2] * inv, o[q * 8 + 3] * inv);
                        uu.z = pack_bf16x2(o[q * 8 + 4] * inv, o[q * 8 + 5] * inv);
                        uu.w = pack_bf16x2(o[q * 8 + 6] * inv, o[q * 8 + 7] * inv);
                        reinterpret_cast<uint4*>(orow + cc * 32)[q] = uu;
                    }
                }
            }
            if (row_ok) p.lse[((int64_t)b * p.nh + h) * p.S + row] = m_ref + log2f(l);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc<1>(tmem, 512); }
}

// ---------------------------------------------------------------------------------------------------- host
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}
// 3-D map over [batches, rows, cols] bf16 (row stride ld, batch stride rows*ld); box = [1, box_rows, 64], SWIZZLE_128B.
int make_tmap_3d_bf16(CUtensorMap* tm, const void* base, int64_t cols, int64_t rows, int64_t batches, int64_t ld, int box_rows) {
    auto fn = encode_fn();
    if (!fn) return set_error(-38, "cuTensorMapEncodeTiled entry point not found");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 8)) return set_error(-22, "attention: operand must be 16-byte aligned, ld %% 8 == 0");
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batches};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows * (cuuint64_t)ld * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(-22, "cuTensorMapEncodeTiled(3d) failed (%d)", (int)r);
    return 0;
}

template <int D>
static int fwd_tc_t(const AttnDesc& a, cudaStream_t st) {
    using C = FwdCfg<D>;
    CUtensorMap tq, tkv;
    // the q / k / v views all live in fused [B*S, ld] rows: map the whole row, pick heads by column coordinate
    const __nv_bfloat16* qrow = reinterpret_cast<const __nv_bfloat16*>(a.q_row_base);
    const __nv_bfloat16* kvrow = reinterpret_cast<const __nv_bfloat16*>(a.kv_row_base);
    SF_TRY_RC(make_tmap_3d_bf16(&tq, qrow, a.ldq, a.S, a.B, a.ldq, C::BQ));
    SF_TRY_RC(make_tmap_3d_bf16(&tkv, kvrow, a.ldkv, a.S, a.B, a.ldkv, C::BKV));
    AttnTcParams p{};
    p.out = (__nv_bfloat16*)a.out; p.ldo = a.ldo; p.lse = a.lse; p.sd = a.sd_ws; p.ldkv = a.ldkv;
    for (int i = 0; i < 8; ++i) p.vdiag[i] = (i < a.J) ? (const __nv_bfloat16*)a.v[i + 1] : nullptr;
    p.kvlen = a.kvlen; p.key_mask = a.key_mask; p.nonprefix = a.nonprefix;
    p.B = a.B; p.S = a.S; p.nh = a.nh; p.nkv = a.nkv; p.J = a.J;
    const int g = a.nh / a.nkv;
    p.heads_per_cta = (g % 2 == 0) ? 2 : 1;
    p.q_col0 = a.q_col0; p.k_col0 = a.k_col0; p.v_col0 = a.v_col0;
    p.scale_log2 = (1.0f / sqrtf((float)D)) * 1.4426950408889634f;
    static bool set = false;
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return set_error(-22, "attn_fwd_tc smem attr: %s", cudaGetErrorString(e));
        set = true;
    }
    dim3 grid((a.S + C::BQ - 1) / C::BQ, a.nh / p.heads_per_cta, a.B);
    attn_fwd_tc_kernel<D><<<grid, 384, C::SMEM, st>>>(tq, tkv, p);
    SF_CUDA_CHECK_LAUNCH("attn_fwd_tc");
    return 0;
}

int attn_fwd_tc(const AttnDesc& a, cudaStream_t st) {
    return a.head_dim == 128 ? fwd_tc_t<128>(a, st) : fwd_tc_t<64>(a, st);
}

}  // namespace sf
