// sf_attention_tc.cu — TTT attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), forward.
//
// Same semantics as sf_attention.cu (block-0 causal flash attention whose running softmax state is seeded with
// the T-1 diagonal scores; reference specforge/modeling/draft/llama3_eagle.py:717-785).
//
// CTA = (128 query rows) x (one or two query heads of the same GQA group) x batch element.  Both heads share
// every K/V tile, so K/V smem traffic is halved, and the two softmax warpgroups ping-pong against the single MMA
// issuer: while warpgroup A exponentiates S_a(t), the tensor core runs P_b(t-1)·V and Q_b·K(t)^T.
//   warp 0     TMA producer   (Q tiles once, K/V tiles through 2-stage rings; 3-D tensor maps [cols, S, B] so rows
//                              past the end of a sequence are zero-filled, never another sequence's data)
//   warp 1     MMA issuer     (S = Q K^T  -> TMEM;  O_tile = P V -> TMEM; one thread)
//   warp 2     TMEM allocator
//   warps 4-7  softmax warpgroup for head a: one thread per query row (TMEM lane), running max/sum in registers,
//   warps 8-11 same for head b               P written as bf16 into a SWIZZLE_128B K-major smem tile (the A operand
//                                            of the PV MMA), O accumulated in registers from the per-tile TMEM result
//                                            with the deferred rescale alpha_{t-1} (off the critical path).
#include "sf_gemm.cuh"   // CUtensorMap, make_* helpers
#include "sf_host.h"
#include <cudaTypedefs.h>

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
    static constexpr int kStages = 2;
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = 2 * Q_BYTES;
    static constexpr int OFF_V = OFF_K + kStages * KV_BYTES;
    static constexpr int OFF_P = OFF_V + kStages * KV_BYTES;
    static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static constexpr int TM_S = 0;            // S_a at col 0, S_b at col 64
    static constexpr int TM_O = 128;          // O_a at 128, O_b at 128 + D
};

template <int D>
__global__ void __launch_bounds__(384, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const AttnTcParams p) {
    using C = FwdCfg<D>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t bar0 = sbase + C::OFF_BAR;
    // barrier map
    const uint32_t b_qfull = bar0;
    auto b_kfull = [&](int s) { return bar0 + 8u * (1 + s); };
    auto b_kempty = [&](int s) { return bar0 + 8u * (3 + s); };
    auto b_vfull = [&](int s) { return bar0 + 8u * (5 + s); };
    auto b_vempty = [&](int s) { return bar0 + 8u * (7 + s); };
    auto b_sfull = [&](int x) { return bar0 + 8u * (9 + x); };
    auto b_pfull = [&](int x) { return bar0 + 8u * (11 + x); };
    auto b_pempty = [&](int x) { return bar0 + 8u * (13 + x); };
    auto b_ofull = [&](int x) { return bar0 + 8u * (15 + x); };
    auto b_oempty = [&](int x) { return bar0 + 8u * (17 + x); };
    const uint32_t tmem_slot = bar0 + 8u * 19;

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
        for (int s = 0; s < 2; ++s) { mbar_init(b_kfull(s), 1); mbar_init(b_kempty(s), 1); mbar_init(b_vfull(s), 1); mbar_init(b_vempty(s), 1); }
        for (int x = 0; x < 2; ++x) {
            mbar_init(b_sfull(x), 1); mbar_init(b_pfull(x), 4); mbar_init(b_pempty(x), 1);
            mbar_init(b_ofull(x), 1); mbar_init(b_oempty(x), 4);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t*>(sgen + C::OFF_BAR + 8 * 19);

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
                const int s = t & 1;
                const uint32_t ph = ((t >> 1) & 1) ^ 1u;
                mbar_wait(b_kempty(s), ph, 11);
                mbar_expect_tx(b_kfull(s), C::KV_BYTES);
                for (int kb = 0; kb < C::NB; ++kb)
                    tma_load_3d(sbase + C::OFF_K + s * C::KV_BYTES + kb * (C::BKV * 128), &tm_kv, b_kfull(s),
                                p.k_col0 + kvh * D + kb * 64, t * C::BKV, b);
                mbar_wait(b_vempty(s), ph, 12);
                mbar_expect_tx(b_vfull(s), C::KV_BYTES);
                for (int kb = 0; kb < C::NB; ++kb)
                    tma_load_3d(sbase + C::OFF_V + s * C::KV_BYTES + kb * (C::BKV * 128), &tm_kv, b_vfull(s),
                                p.v_col0 + kvh * D + kb * 64, t * C::BKV, b);
            }
        } else if (warp == 1 && lane == 0) {
            // ================= MMA issuer =================
            constexpr uint32_t idesc_s = make_idesc_bf16(128, C::BKV, 0, 0);   // S[128 x 64]  = Q(K-major) K^T(K-major)
            constexpr uint32_t idesc_o = make_idesc_bf16(128, D, 0, 1);        // O[128 x D]   = P(K-major) V(MN-major)
            mbar_wait(b_qfull, 0, 13);
            tc_fence_after();
            for (int t = 0; t <= n_kv; ++t) {
                for (int x = 0; x < nx; ++x) {
                    if (t >= 1) {
                        // O_tile_x = P_x(t-1) V_{t-1}
                        const int u = t - 1, s = u & 1;
                        mbar_wait(b_pfull(x), u & 1, 14);
                        if (x == 0) mbar_wait(b_vfull(s), (u >> 1) & 1, 15);
                        mbar_wait(b_oempty(x), (u & 1) ^ 1u, 16);
                        tc_fence_after();
                        const uint32_t sp = sbase + C::OFF_P + x * C::P_BYTES;
                        const uint32_t sv = sbase + C::OFF_V + s * C::KV_BYTES;
                        const uint64_t adesc = make_smem_desc_sw128(sp, 0, 1024);
                        const uint64_t bdesc = make_smem_desc_sw128(sv, C::BKV * 128, 1024);   // MN-major: LBO = 64-col block stride
#pragma unroll
                        for (int k = 0; k < C::BKV / 16; ++k)
                            umma_bf16<1>(tmem + C::TM_O + x * D, adesc + ((k * 32) >> 4), bdesc + ((k * 2048) >> 4), idesc_o, k != 0);
                        umma_commit(b_ofull(x));
                        umma_commit(b_pempty(x));
                        if (x == nx - 1) umma_commit(b_vempty(s));
                    }
                    if (t < n_kv) {
                        // S_x = Q_x K_t^T   (S_x(t-1) has been fully read: p_full[x](t-1) was awaited above)
                        const int s = t & 1;
                        if (x == 0) mbar_wait(b_kfull(s), (t >> 1) & 1, 17);
                        tc_fence_after();
                        const uint32_t sq = sbase + C::OFF_Q + x * C::Q_BYTES;
                        const uint32_t sk = sbase + C::OFF_K + s * C::KV_BYTES;
#pragma unroll
                        for (int kb = 0; kb < C::NB; ++kb) {
                            const uint64_t adesc = make_smem_desc_sw128(sq + kb * (C::BQ * 128), 0, 1024);
                            const uint64_t bdesc = make_smem_desc_sw128(sk + kb * (C::BKV * 128), 0, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16<1>(tmem + C::TM_S + x * C::BKV, adesc + ((k * 32) >> 4), bdesc + ((k * 32) >> 4), idesc_s,
                                             (kb | k) != 0);
                        }
                        umma_commit(b_sfull(x));
                        if (x == nx - 1) umma_commit(b_kempty(s));
                    }
                }
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
            float m = -INFINITY, l = 0.f;
            const float* sdr = nullptr;
            if (p.J > 0) {
                sdr = p.sd + (((int64_t)b * p.nh + h) * p.S + min(row, p.S - 1)) * p.J;
                for (int i = 0; i < p.J; ++i) m = fmaxf(m, sdr[i]);
                for (int i = 0; i < p.J; ++i) l += exp2f(sdr[i] - m);
            }
            float o[D];
#pragma unroll
            for (int i = 0; i < D; ++i) o[i] = 0.f;
            float alpha_prev = 1.f;
            uint8_t* sp = sgen + C::OFF_P + x * C::P_BYTES;

            auto accumulate_o = [&](int u) {
                mbar_wait(b_ofull(x), u & 1, 20 + x);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < D / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_lane + C::TM_O + x * D + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 32; ++e) o[c * 32 + e] = o[c * 32 + e] * alpha_prev + __uint_as_float(v[e]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_oempty(x));
            };

            for (int t = 0; t < n_kv; ++t) {
                const int kv0 = t * C::BKV;
                const bool need_mask = (kv0 + C::BKV - 1 > q0 + wq * 32) || (kv0 + C::BKV > kvlen) || general_mask;
                mbar_wait(b_sfull(x), t & 1, 22 + x);
                tc_fence_after();
                uint32_t sv[C::BKV];
                tmem_ld_32x32b_x32(t_lane + C::TM_S + x * C::BKV, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
                tmem_ld_32x32b_x32(t_lane + C::TM_S + x * C::BKV + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
                tmem_ld_wait();
                float mx = m;
#pragma unroll
                for (int c = 0; c < C::BKV; ++c) {
                    float v = __uint_as_float(sv[c]) * p.scale_log2;
                    if (need_mask) {
                        const int key = kv0 + c;
                        bool ok = key <= row && key < kvlen;
                        if (ok && general_mask) ok = p.key_mask[(int64_t)b * p.S + key] != 0;
                        if (!ok) v = -INFINITY;
                    }
                    sv[c] = __float_as_uint(v);
                    mx = fmaxf(mx, v);
                }
                const float base = (mx == -INFINITY) ? 0.f : mx;
                const float alpha = exp2f(m - base);
                m = mx;
                // P tile buffer must have been consumed by PV(t-1)
                mbar_wait(b_pempty(x), (t & 1) ^ 1u, 24 + x);
                float rs = 0.f;
#pragma unroll
                for (int j = 0; j < C::BKV / 8; ++j) {
                    float e[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) { e[k] = exp2f(__uint_as_float(sv[j * 8 + k]) - base); rs += e[k]; }
                    uint4 pk;
                    pk.x = pack_bf16x2(e[0], e[1]); pk.y = pack_bf16x2(e[2], e[3]);
                    pk.z = pack_bf16x2(e[4], e[5]); pk.w = pack_bf16x2(e[6], e[7]);
                    *reinterpret_cast<uint4*>(sp + sw128(r, j)) = pk;
                }
                l = l * alpha + rs;
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_pfull(x));
                if (t >= 1) accumulate_o(t - 1);
                alpha_prev = alpha;
            }
            accumulate_o(n_kv - 1);
            // ---- epilogue: diagonal P*V terms (never masked), normalise, store
            if (row_ok) {
                for (int i = 0; i < p.J; ++i) {
                    const float w = exp2f(sdr[i] - m);
                    const uint4* vp = reinterpret_cast<const uint4*>(p.vdiag[i] + ((int64_t)b * p.S + row) * p.ldkv + kvh * D);
#pragma unroll
                    for (int c = 0; c < D / 8; ++c) {
                        const uint4 u = __ldg(vp + c);
                        const uint32_t wds[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __nv_bfloat162 bb = *reinterpret_cast<const __nv_bfloat162*>(&wds[e]);
                            o[c * 8 + 2 * e] += w * __bfloat162float(bb.x);
                            o[c * 8 + 2 * e + 1] += w * __bfloat162float(bb.y);
                        }
                    }
                }
                const float inv = (l > 0.f) ? 1.f / l : 0.f;
                uint4* op = reinterpret_cast<uint4*>(p.out + ((int64_t)b * p.S + row) * p.ldo + h * D);
#pragma unroll
                for (int c = 0; c < D / 8; ++c) {
                    uint4 u;
                    u.x = pack_bf16x2(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv);
                    u.y = pack_bf16x2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
                    u.z = pack_bf16x2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv);
                    u.w = pack_bf16x2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
                    op[c] = u;
                }
                p.lse[((int64_t)b * p.nh + h) * p.S + row] = m + log2f(l);
            }
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
