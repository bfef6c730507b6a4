// sf_gemm.cuh — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[m, n] = sum_k A(m, k) * B(n, k)        (bf16 inputs, fp32 accumulation in TMEM)
//
// Each operand may be stored K-major (row-major [rows, K], the contraction dim contiguous) or
// MN-major (row-major [K, rows], the contraction dim is the slow one).  That covers, on plain
// row-major tensors and without any transposed copies:
//   forward  Y = X W^T          A = X  (K-major)    B = W  (K-major)
//   dgrad    dX = dY W          A = dY (K-major)    B = W  (MN-major)
//   wgrad    dW = dY^T X        A = dY (MN-major)   B = X  (MN-major)
// (reference ops replaced: every nn.Linear on the EAGLE3 draft path —
//  specforge/modeling/draft/llama3_eagle.py:555-563,1513-1515,1668-1693 — and their autograd.)
//
// Structure: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator,
// warps 4-7 = epilogue (TMEM -> registers -> global).  kStages-deep smem ring between TMA and
// MMA, two TMEM accumulator stages between MMA and epilogue, static persistent tile schedule.
// kCtaGroup == 2 pairs two CTAs (one cluster) on a 256 x BLOCK_N tile with cta_group::2 MMAs.
#pragma once
#include "sf_ptx.cuh"
#include <cuda.h>

namespace sf {

enum : int { MAJOR_K = 0, MAJOR_MN = 1 };
enum : int {
    EPI_BF16 = 0,        // D(bf16) = acc
    EPI_BF16_RESID = 1,  // D(bf16) = bf16(acc) + R   (R bf16; sum rounded once more to bf16)
    EPI_F32 = 2,         // D(f32)  = acc
    EPI_F32_ACCUM = 3,   // D(f32) += acc
    // Fused SwiGLU (llama3_eagle.py:1547).  Forward: B = [gate ; up] weight ([2I, K]); each 256-wide tile holds 128
    // gate and the matching 128 up columns (CTA rank 0 stages gate rows, rank 1 up rows; 2-CTA only).
    // D = gu [M, 2I] (bf16 gate | up, saved for backward), D2 = act [M, I] = bf16(bf16(silu(g)) * u).
    EPI_SWIGLU = 4,
    // Backward: acc = d(act) tile [128 x 256] over columns j of I; R = gu [M, 2I]; D = d(gu) [M, 2I]:
    // d(gate) = d(act) * u * silu'(g), d(up) = d(act) * silu(g).
    EPI_SWIGLU_BWD = 5,
    // Row statistics fused into the epilogue (the reductions of _compute_target_p / LogSoftmaxLoss pass A move into the GEMM
    // that produces the row).  Per (row, n-block) the epilogue emits partial online-softmax state over the bf16-ROUNDED outputs
    // x = bf16(acc):  stats[k][row][n_blk], k = 0 max, 1 sum exp(x - max), 2 first index of the max (int bits); a small merge
    // kernel combines the n-blocks in ascending order (first index wins ties = torch.argmax).
    //   EPI_BF16_STATS: D(bf16) = acc as EPI_BF16, plus the three partials          (draft lm_head -> loss pass A)
    //   EPI_TEACHER:    no D at all.  Columns flagged in t2d_bits are appended, in vocabulary order, to xg[orow, :] (the gathered
    //                   draft-vocab teacher logits, orow = row + (row / S) * T = the [B, S+T, DV] padded layout), with two more
    //                   partials k = 3 max, 4 sum exp over the draft-vocab columns only.  The [M, V] teacher logits never exist
    //                   (eagle3/model.py:487-501; this is also core/compact_teacher.py:57-150's streaming logsumexp/argmax).
    EPI_BF16_STATS = 6,
    EPI_TEACHER = 7,
    // D(bf16) = RoPE(bf16(acc)) on the columns < rope_cols (the q and k heads of a fused [q;k;v] projection), plain bf16 beyond:
    // x' = x cos + rotate_half(x) sin at position (row % S) + rope_pos0, cos / sin the reference's bf16 tables [rows, head_dim]
    // (llama3_eagle.py:133-142; fp32 math on the bf16-rounded linear output, one rounding — what the standalone rope kernel
    // computes).  Needs head_dim in {64, 128} and rope_cols % head_dim == 0; a head never straddles a 128-column part.
    EPI_BF16_ROPE = 8,
};

struct GemmParams {
    void* D;
    const __nv_bfloat16* R;
    int M, N, K;
    int ldd, ldr;
    int epi;
    int num_m_blocks, num_n_blocks;
    void* D2; int ldd2;   // EPI_SWIGLU: act output
    int n_half;           // EPI_SWIGLU / EPI_SWIGLU_BWD: I (columns of gate == columns of up)
    int group_m;          // M-blocks per raster group (L2 reuse window of the A operand)
    int stages;           // depth of the TMA -> MMA smem ring (6 or 7 x 32 KB)
    int staged;           // 1: the epilogue warps own a 2 KB staging block each (after the barriers) for coalesced bf16 transfers
    // EPI_BF16_STATS / EPI_TEACHER
    float* stats;                  // [3 or 5][M][partial blocks per row]  (a row's partials are contiguous for the merge)
    const uint32_t* t2d_bits;      // [ceil(N / 32)] bit e of word w: column 32 w + e is in the draft vocabulary
    const int* t2d_prefix;         // [ceil(N / 32)] draft-vocab columns before column 32 w
    __nv_bfloat16* xg; int S, T, DV;
    // EPI_BF16_ROPE (S as above = sequence length)
    const __nv_bfloat16* rope_cos; const __nv_bfloat16* rope_sin; int rope_cols, rope_pos0, head_dim;
    unsigned long long* trace;     // diagnostic (sf_debug_gemm_trace): cluster 0 / CTA 0 writes clock64() stamps, 16 per tile
};

template <int kCtaGroup, int kAMajor, int kBMajor, int kBlockN>
struct GemmCfg {
    static constexpr int BLOCK_M = 128;  // rows per CTA
    static constexpr int TILE_M = 128 * kCtaGroup;
    static constexpr int BLOCK_N = kBlockN;
    static constexpr int BLOCK_K = 64;
    static constexpr int UMMA_K = 16;
    static constexpr int B_ROWS = kBlockN / kCtaGroup;  // B rows staged by each CTA
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int kStages = (224 * 1024) / STAGE_BYTES;      // default ring depth: 7 x 32 KB (in-step A/B on one box: -1.8 % step time vs 6)
    static constexpr int kMaxStages = (224 * 1024) / STAGE_BYTES;   // deepest ring that fits the 227 KB of a CTA
    static constexpr int kAccStages = 2;
    static constexpr int TMEM_COLS = 512;
    static constexpr int STAGING_BYTES = 8 * 2048;                  // optional: 2 KB per epilogue warp, up to 8 (GemmParams::staged)
    static constexpr int smem_bytes(int stages, int staged = 0) { return stages * STAGE_BYTES + 1024 /*align*/ + 512 /*barriers*/ + (staged ? STAGING_BYTES : 0); }
    static constexpr int SMEM_BYTES = smem_bytes(kStages);
    static constexpr int kThreads = 256;
    static_assert(kBlockN * kAccStages <= 512, "TMEM overflow");
    static_assert(kBlockN % 64 == 0 && kBlockN <= 256, "bad BLOCK_N");
};

// ---- warp-staged row-block transfers --------------------------------------------------------------------------------------
// In the epilogue a thread owns one accumulator ROW (TMEM lane), so writing its 32 bf16 columns straight to global memory makes
// every 16-byte store instruction of the warp touch 32 different 128-byte lines.  Staged through a 2 KB per-warp shared-memory
// block ([32 rows x 64 B], 16-byte units XOR-swizzled so the row-wise and the 8-rows-per-instruction accesses are both
// bank-conflict-free) the same data leaves as 4 instructions that each write 8 rows x 64 contiguous bytes (whole sectors).
__device__ __forceinline__ uint32_t stg_off(int r, int u) { return (uint32_t)(r * 64 + ((u ^ ((r >> 1) & 3)) << 4)); }
// lane r holds row r's 32 bf16 (o[16]); dst -> (row 0 of the block, first column); rows >= rows_valid are not written
__device__ __forceinline__ void warp_store_32x32(uint8_t* wbuf, int lane, const uint32_t (&o)[16], __nv_bfloat16* dst, int64_t ld,
                                                 int rows_valid) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(wbuf + stg_off(lane, q)) = make_uint4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
    __syncwarp();
    const int u = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + (lane >> 2);
        const uint4 v = *reinterpret_cast<const uint4*>(wbuf + stg_off(r, u));
        if (r < rows_valid) *reinterpret_cast<uint4*>(dst + (size_t)r * ld + u * 8) = v;
    }
    __syncwarp();
}
__device__ __forceinline__ void warp_load_32x32(uint8_t* wbuf, int lane, uint32_t (&o)[16], const __nv_bfloat16* src, int64_t ld,
                                                int rows_valid) {
    const int u = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + (lane >> 2);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < rows_valid) v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * ld + u * 8));
        *reinterpret_cast<uint4*>(wbuf + stg_off(r, u)) = v;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 v = *reinterpret_cast<const uint4*>(wbuf + stg_off(lane, q));
        o[q * 4] = v.x; o[q * 4 + 1] = v.y; o[q * 4 + 2] = v.z; o[q * 4 + 3] = v.w;
    }
    __syncwarp();
}

// The same load split in two so that the global reads of chunk c+1 can be in flight while chunk c is processed (the epilogue of a
// tile is a chain of dependent HBM/L2 round trips otherwise: 8 chunks x ~1.5 k cycles, enough to make a K = 4096 GEMM
// epilogue-paced — the SwiGLU-backward dgrad ran at 800 TFLOP/s inside the step before this).
__device__ __forceinline__ void warp_load_issue(int lane, uint4 (&t)[4], const __nv_bfloat16* src, int64_t ld, int rows_valid) {
    const int u = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + (lane >> 2);
        t[i] = make_uint4(0, 0, 0, 0);
        if (r < rows_valid) t[i] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * ld + u * 8));
    }
}
__device__ __forceinline__ void warp_load_commit(uint8_t* wbuf, int lane, const uint4 (&t)[4], uint32_t (&o)[16]) {
    const int u = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(wbuf + stg_off(i * 8 + (lane >> 2), u)) = t[i];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 v = *reinterpret_cast<const uint4*>(wbuf + stg_off(lane, q));
        o[q * 4] = v.x; o[q * 4 + 1] = v.y; o[q * 4 + 2] = v.z; o[q * 4 + 3] = v.w;
    }
    __syncwarp();
}
// direct (per-thread row) variant: this thread's 32 bf16 of one row as 4 x 16 bytes
__device__ __forceinline__ void row_load_issue(uint4 (&t)[4], const __nv_bfloat16* src, bool ok) {
#pragma unroll
    for (int q = 0; q < 4; ++q) t[q] = ok ? __ldg(reinterpret_cast<const uint4*>(src) + q) : make_uint4(0, 0, 0, 0);
}

// L2 prefetch of the epilogue INPUTS of this thread's row in the NEXT tile of its CTA (residual row, or the saved gate / up rows of
// the SwiGLU backward): issued at the start of the current tile's epilogue, a whole main loop before they are read, so that the
// register prefetch above finds them in L2 instead of paying an HBM round trip per 32-column chunk.
__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
template <int kBlockN>
__device__ __forceinline__ void gemm_epilogue_prefetch(const GemmParams& p, const int row, const int n0) {
    if (row >= p.M) return;
    if (p.epi == EPI_BF16_RESID) {
        const __nv_bfloat16* r = p.R + (size_t)row * p.ldr + n0;
#pragma unroll
        for (int i = 0; i < kBlockN / 64; ++i)
            if (n0 + i * 64 < p.N) prefetch_l2(r + i * 64);
    } else if (p.epi == EPI_SWIGLU_BWD) {
        const __nv_bfloat16* g = p.R + (size_t)row * p.ldr + n0;
#pragma unroll
        for (int i = 0; i < kBlockN / 64; ++i)
            if (n0 + i * 64 < p.n_half) { prefetch_l2(g + i * 64); prefetch_l2(g + p.n_half + i * 64); }
    }
}

// The epilogue of one thread: its row of the [128 x kBlockN] accumulator block at TMEM address t_row (lane = row), processed in
// 32-column chunks; shared by the 256 x 256 (gemm_kernel) and 512 x 256 (gemm_wide_kernel) tilings.  wbuf: this warp's 2 KB
// staging block (bf16 outputs / inputs then move as described above) or nullptr (direct per-thread accesses).
template <int kBlockN, int kEpiSet = 0>
__device__ __forceinline__ void gemm_epilogue_rows(const GemmParams& p, const int row, const bool row_ok, const int n_blk,
                                                   const int n0, const uint32_t t_row, uint8_t* wbuf, const int lane,
                                                   const int part = 0, const int nparts = 1) {
            // part / nparts: this call covers the part-th of nparts equal column ranges of the tile (the 512 x 256 tiling drains
            // each accumulator half with two warpgroups side by side); the row-statistics partials are then per (n-block, part)
            const int row0 = row - lane;                                   // first row of this warp's 32-row block
            const int rows_valid = min(32, max(0, p.M - row0));
            // kEpiSet: 0 = every epilogue, 1 = the plain ones only (bf16 / residual / fp32 / accumulate), 2 = the fused ones only
            // (heavy_epilogue() in sf_gemm.cu) — the 384-thread kernels are capped at 168 registers, and one function body for all
            // variants made both groups spill.
            constexpr bool kLight = kEpiSet == 1;
            if (!kLight && p.epi == EPI_SWIGLU) {
                // columns [0,128) of the accumulator = gate(j0 + .), [128,256) = up(j0 + .)
                const int j0 = n_blk * (kBlockN / 2);
                __nv_bfloat16* gu = reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row * p.ldd;
                __nv_bfloat16* act = reinterpret_cast<__nv_bfloat16*>(p.D2) + (size_t)row * p.ldd2;
#pragma unroll 1
                for (int c = part * (kBlockN / 64 / nparts); c < (part + 1) * (kBlockN / 64 / nparts); ++c) {
                    uint32_t g[32], u[32];
                    tmem_ld_32x32b_x32(t_row + c * 32, g);
                    tmem_ld_32x32b_x32(t_row + kBlockN / 2 + c * 32, u);
                    tmem_ld_wait();
                    const int col = j0 + c * 32;
                    if (col >= p.n_half || (!wbuf && !row_ok)) continue;
                    uint32_t og[16], ou[16], oa[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        og[e] = pack_bf16x2(__uint_as_float(g[2 * e]), __uint_as_float(g[2 * e + 1]));
                        ou[e] = pack_bf16x2(__uint_as_float(u[2 * e]), __uint_as_float(u[2 * e + 1]));
                        const __nv_bfloat162 gb = *reinterpret_cast<const __nv_bfloat162*>(&og[e]);
                        const __nv_bfloat162 ub = *reinterpret_cast<const __nv_bfloat162*>(&ou[e]);
                        const float g0 = __bfloat162float(gb.x), g1 = __bfloat162float(gb.y);
                        const float s0 = __bfloat162float(__float2bfloat16_rn(g0 / (1.f + __expf(-g0))));
                        const float s1 = __bfloat162float(__float2bfloat16_rn(g1 / (1.f + __expf(-g1))));
                        oa[e] = pack_bf16x2(s0 * __bfloat162float(ub.x), s1 * __bfloat162float(ub.y));
                    }
                    // host guarantees n_half % 128 == 0: every 32-column chunk is whole
                    if (wbuf) {
                        __nv_bfloat16* gu0 = reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row0 * p.ldd;
                        warp_store_32x32(wbuf, lane, og, gu0 + col, p.ldd, rows_valid);
                        warp_store_32x32(wbuf, lane, ou, gu0 + p.n_half + col, p.ldd, rows_valid);
                        warp_store_32x32(wbuf, lane, oa, reinterpret_cast<__nv_bfloat16*>(p.D2) + (size_t)row0 * p.ldd2 + col, p.ldd2, rows_valid);
                        continue;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        reinterpret_cast<uint4*>(gu + col)[q] = make_uint4(og[q * 4], og[q * 4 + 1], og[q * 4 + 2], og[q * 4 + 3]);
                        reinterpret_cast<uint4*>(gu + p.n_half + col)[q] = make_uint4(ou[q * 4], ou[q * 4 + 1], ou[q * 4 + 2], ou[q * 4 + 3]);
                        reinterpret_cast<uint4*>(act + col)[q] = make_uint4(oa[q * 4], oa[q * 4 + 1], oa[q * 4 + 2], oa[q * 4 + 3]);
                    }
                }
            } else if (!kLight && p.epi == EPI_SWIGLU_BWD) {
                const __nv_bfloat16* gu = p.R + (size_t)row * p.ldr;
                const __nv_bfloat16* gu0 = p.R + (size_t)row0 * p.ldr;
                __nv_bfloat16* dgu = reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row * p.ldd;
                const int c_begin = part * (kBlockN / 32 / nparts), c_end = (part + 1) * (kBlockN / 32 / nparts);
                uint4 gn[4], un[4];                                  // the NEXT chunk's gate / up inputs, already in flight
                auto issue = [&](int c) {
                    const int col = n0 + c * 32;
                    if (c >= c_end || col >= p.n_half) return;
                    if (wbuf) { warp_load_issue(lane, gn, gu0 + col, p.ldr, rows_valid); warp_load_issue(lane, un, gu0 + p.n_half + col, p.ldr, rows_valid); }
                    else { row_load_issue(gn, gu + col, row_ok); row_load_issue(un, gu + p.n_half + col, row_ok); }
                };
                issue(c_begin);
#pragma unroll 1
                for (int c = c_begin; c < c_end; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + c * 32, v);
                    uint4 gc[4], uc[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { gc[q] = gn[q]; uc[q] = un[q]; }
                    issue(c + 1);
                    tmem_ld_wait();
                    const int col = n0 + c * 32;
                    if (col >= p.n_half || (!wbuf && !row_ok)) continue;
                    // host guarantees n_half % 32 == 0
                    uint32_t gsw[16], usw[16], dgw[16], duw[16];
                    if (wbuf) {
                        warp_load_commit(wbuf, lane, gc, gsw);
                        warp_load_commit(wbuf, lane, uc, usw);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t gw[4], uw[4];
                        if (wbuf) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) { gw[e] = gsw[q * 4 + e]; uw[e] = usw[q * 4 + e]; }
                        } else {
                            const uint4 g4 = gc[q], u4 = uc[q];
                            gw[0] = g4.x; gw[1] = g4.y; gw[2] = g4.z; gw[3] = g4.w;
                            uw[0] = u4.x; uw[1] = u4.y; uw[2] = u4.z; uw[3] = u4.w;
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __nv_bfloat162 gb = *reinterpret_cast<const __nv_bfloat162*>(&gw[e]);
                            const __nv_bfloat162 ub = *reinterpret_cast<const __nv_bfloat162*>(&uw[e]);
                            float dg[2], du[2];
#pragma unroll
                            for (int w = 0; w < 2; ++w) {
                                const float gg = __bfloat162float(w ? gb.y : gb.x), uu = __bfloat162float(w ? ub.y : ub.x);
                                const float da = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[q * 8 + e * 2 + w])));
                                const float sg = 1.f / (1.f + __expf(-gg));
                                dg[w] = da * uu * (sg * (1.f + gg * (1.f - sg)));
                                du[w] = da * gg * sg;
                            }
                            dgw[q * 4 + e] = pack_bf16x2(dg[0], dg[1]);
                            duw[q * 4 + e] = pack_bf16x2(du[0], du[1]);
                        }
                        if (!wbuf) {
                            reinterpret_cast<uint4*>(dgu + col)[q] = make_uint4(dgw[q * 4], dgw[q * 4 + 1], dgw[q * 4 + 2], dgw[q * 4 + 3]);
                            reinterpret_cast<uint4*>(dgu + p.n_half + col)[q] = make_uint4(duw[q * 4], duw[q * 4 + 1], duw[q * 4 + 2], duw[q * 4 + 3]);
                        }
                    }
                    if (wbuf) {
                        __nv_bfloat16* dgu0 = reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row0 * p.ldd;
                        warp_store_32x32(wbuf, lane, dgw, dgu0 + col, p.ldd, rows_valid);
                        warp_store_32x32(wbuf, lane, duw, dgu0 + p.n_half + col, p.ldd, rows_valid);
                    }
                }
            } else if (!kLight && (p.epi == EPI_BF16_STATS || p.epi == EPI_TEACHER)) {
                const bool gather = p.epi == EPI_TEACHER;
                float m = -INFINITY, d = 0.f, md = -INFINITY, dd = 0.f;
                int idx = 0x7fffffff;
                __nv_bfloat16* xo = nullptr;
                if (gather && row_ok) xo = p.xg + ((size_t)row + (size_t)(row / p.S) * p.T) * p.DV;
#pragma unroll 1
                for (int c = part * (kBlockN / 32 / nparts); c < (part + 1) * (kBlockN / 32 / nparts); ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + c * 32, v);
                    tmem_ld_wait();
                    const int col = n0 + c * 32;
                    if (col >= p.N) continue;
                    const bool full = (col + 32 <= p.N);
                    const bool staged = wbuf && full && !gather;
                    if (!row_ok && !staged) continue;
                    float x[32];
                    uint32_t o[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        o[e] = pack_bf16x2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
                        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&o[e]);
                        x[2 * e] = __bfloat162float(b2.x);
                        x[2 * e + 1] = __bfloat162float(b2.y);
                    }
                    if (!full) {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (col + e >= p.N) x[e] = -INFINITY;
                    }
                    if (staged) {
                        warp_store_32x32(wbuf, lane, o, reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row0 * p.ldd + col, p.ldd, rows_valid);
                        if (!row_ok) continue;
                    } else if (!gather) {
                        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row * p.ldd + col;
                        if (full) {
                            uint4* dp = reinterpret_cast<uint4*>(dst);
#pragma unroll
                            for (int q = 0; q < 4; ++q) dp[q] = make_uint4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                if (col + e < p.N) dst[e] = __float2bfloat16_rn(x[e]);
                        }
                    }
                    float cm = x[0];
                    int ci = 0;
#pragma unroll
                    for (int e = 1; e < 32; ++e)
                        if (x[e] > cm) { cm = x[e]; ci = e; }
                    if (cm > m) { d *= __expf(m - cm); m = cm; idx = col + ci; }
#pragma unroll
                    for (int e = 0; e < 32; ++e) d += __expf(x[e] - m);
                    if (gather) {
                        const uint32_t mask = __ldg(p.t2d_bits + (col >> 5));
                        if (mask) {
                            __nv_bfloat16* dst = xo + __ldg(p.t2d_prefix + (col >> 5));
                            float dm = -INFINITY;
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                if ((mask >> e) & 1u) dm = fmaxf(dm, x[e]);
                            if (dm > md) { dd *= __expf(md - dm); md = dm; }
                            int cnt = 0;
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                if ((mask >> e) & 1u) {       // warp-uniform: every lane holds the same columns
                                    dd += __expf(x[e] - md);
                                    dst[cnt++] = __float2bfloat16_rn(x[e]);
                                }
                        }
                    }
                }
                if (row_ok) {   // a part with no valid column writes the neutral state (max -inf, sum 0), which the merges skip
                    const size_t nbt = (size_t)p.num_n_blocks * nparts;       // partial blocks per row; layout [k][row][block]
                    const size_t plane = nbt * p.M;
                    float* sp = p.stats + (size_t)row * nbt + ((size_t)n_blk * nparts + part);
                    sp[0] = m; sp[plane] = d; sp[2 * plane] = __int_as_float(idx);
                    if (gather) { sp[3 * plane] = md; sp[4 * plane] = dd; }
                }
            } else if (!kLight && p.epi == EPI_BF16_ROPE) {
                // chunk pairs (c, c + half/32) of one head: x1 = columns [hc, hc+32), x2 = columns [hc + half, hc + half + 32)
                const int hchunks = p.head_dim / 32, pairs = hchunks / 2;              // 4 / 2 (d = 128) or 2 / 1 (d = 64)
                const int pos = (row_ok ? row % p.S : 0) + p.rope_pos0;
                const __nv_bfloat16* cosr = p.rope_cos + (size_t)pos * p.head_dim;
                const __nv_bfloat16* sinr = p.rope_sin + (size_t)pos * p.head_dim;
                __nv_bfloat16* drow0 = reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row0 * p.ldd;
                __nv_bfloat16* drow = reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row * p.ldd;
                const int c_begin = part * (kBlockN / 32 / nparts), c_end = (part + 1) * (kBlockN / 32 / nparts);
#pragma unroll 1
                for (int cb = c_begin; cb < c_end; cb += hchunks) {                      // one head per iteration
#pragma unroll 1
                    for (int cc = 0; cc < pairs; ++cc) {
                        const int c1 = cb + cc, c2 = c1 + pairs;
                        uint32_t v1[32], v2[32];
                        tmem_ld_32x32b_x32(t_row + c1 * 32, v1);
                        tmem_ld_32x32b_x32(t_row + c2 * 32, v2);
                        tmem_ld_wait();
                        const int col1 = n0 + c1 * 32, col2 = n0 + c2 * 32;
                        if (col1 >= p.N) continue;                                       // host: N % head_dim == 0 -> whole pairs
                        uint32_t o1[16], o2[16];
                        if (col1 < p.rope_cols) {
                            const int hc = (col1 % p.head_dim);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 c4 = __ldg(reinterpret_cast<const uint4*>(cosr + hc) + q);
                                const uint4 s4 = __ldg(reinterpret_cast<const uint4*>(sinr + hc) + q);
                                const uint32_t cw[4] = {c4.x, c4.y, c4.z, c4.w}, sw[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const __nv_bfloat162 cb2 = *reinterpret_cast<const __nv_bfloat162*>(&cw[e]);
                                    const __nv_bfloat162 sb2 = *reinterpret_cast<const __nv_bfloat162*>(&sw[e]);
                                    const int i = q * 8 + e * 2;
                                    float a[2], b[2];
#pragma unroll
                                    for (int w = 0; w < 2; ++w) {
                                        const float x1 = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v1[i + w])));
                                        const float x2 = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v2[i + w])));
                                        const float cf = __bfloat162float(w ? cb2.y : cb2.x), sf_ = __bfloat162float(w ? sb2.y : sb2.x);
                                        a[w] = x1 * cf - x2 * sf_;
                                        b[w] = x2 * cf + x1 * sf_;
                                    }
                                    o1[q * 4 + e] = pack_bf16x2(a[0], a[1]);
                                    o2[q * 4 + e] = pack_bf16x2(b[0], b[1]);
                                }
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                o1[e] = pack_bf16x2(__uint_as_float(v1[2 * e]), __uint_as_float(v1[2 * e + 1]));
                                o2[e] = pack_bf16x2(__uint_as_float(v2[2 * e]), __uint_as_float(v2[2 * e + 1]));
                            }
                        }
                        if (wbuf) {
                            warp_store_32x32(wbuf, lane, o1, drow0 + col1, p.ldd, rows_valid);
                            warp_store_32x32(wbuf, lane, o2, drow0 + col2, p.ldd, rows_valid);
                        } else if (row_ok) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                reinterpret_cast<uint4*>(drow + col1)[q] = make_uint4(o1[q * 4], o1[q * 4 + 1], o1[q * 4 + 2], o1[q * 4 + 3]);
                                reinterpret_cast<uint4*>(drow + col2)[q] = make_uint4(o2[q * 4], o2[q * 4 + 1], o2[q * 4 + 2], o2[q * 4 + 3]);
                            }
                        }
                    }
                }
            } else if (kEpiSet != 2) {
            const int c_begin = part * (kBlockN / 32 / nparts), c_end = (part + 1) * (kBlockN / 32 / nparts);
            const bool resid = p.epi == EPI_BF16_RESID;
            uint4 rn[4];                                             // the NEXT chunk's residual, already in flight
            auto issue_resid = [&](int c) {
                const int col = n0 + c * 32;
                if (!resid || c >= c_end || col + 32 > p.N) return;   // ragged last chunk: loaded element-wise below
                if (wbuf) warp_load_issue(lane, rn, p.R + (size_t)row0 * p.ldr + col, p.ldr, rows_valid);
                else row_load_issue(rn, p.R + (size_t)row * p.ldr + col, row_ok);
            };
            issue_resid(c_begin);
            // One 32-column chunk: v holds this thread's row of the accumulator (already waited for).
            auto chunk = [&](const uint32_t (&v)[32], const int c) {
                uint4 rc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) rc[q] = rn[q];
                issue_resid(c + 1);
                const int col = n0 + c * 32;
                if (col >= p.N) return;
                const bool full = (col + 32 <= p.N);
                if (wbuf && full && (p.epi == EPI_BF16 || p.epi == EPI_BF16_RESID)) {   // warp-staged: every lane takes part
                    uint32_t o[16];
                    if (p.epi == EPI_BF16_RESID) {
                        uint32_t rw[16];
                        warp_load_commit(wbuf, lane, rc, rw);
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const __nv_bfloat162 rb = *reinterpret_cast<const __nv_bfloat162*>(&rw[e]);
                            // match the reference's two roundings: linear output -> bf16, then add
                            const float a0 = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[2 * e])));
                            const float a1 = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[2 * e + 1])));
                            o[e] = pack_bf16x2(a0 + __bfloat162float(rb.x), a1 + __bfloat162float(rb.y));
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) o[e] = pack_bf16x2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
                    }
                    warp_store_32x32(wbuf, lane, o, reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row0 * p.ldd + col, p.ldd, rows_valid);
                    return;
                }
                if (!row_ok) return;
                if (p.epi == EPI_BF16 || p.epi == EPI_BF16_RESID) {
                    __nv_bfloat16* dst =
                        reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row * p.ldd + col;
                    if (full) {
                        uint32_t o[16];
                        if (p.epi == EPI_BF16_RESID) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 r4 = rc[q];
                                const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    __nv_bfloat162 rb = *reinterpret_cast<const __nv_bfloat162*>(&rr[e]);
                                    // match the reference's two roundings: linear output -> bf16, then add
                                    float a0 = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[q * 8 + e * 2])));
                                    float a1 = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[q * 8 + e * 2 + 1])));
                                    o[q * 4 + e] = pack_bf16x2(a0 + __bfloat162float(rb.x),
                                                               a1 + __bfloat162float(rb.y));
                                }
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e)
                                o[e] = pack_bf16x2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
                        }
                        uint4* dp = reinterpret_cast<uint4*>(dst);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            dp[q] = make_uint4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
                    } else {
                        for (int e = 0; e < 32 && col + e < p.N; ++e) {
                            float a = __uint_as_float(v[e]);
                            if (p.epi == EPI_BF16_RESID)
                                a = __bfloat162float(__float2bfloat16_rn(a)) +
                                    __bfloat162float(p.R[(size_t)row * p.ldr + col + e]);
                            dst[e] = __float2bfloat16_rn(a);
                        }
                    }
                } else {
                    float* dst = reinterpret_cast<float*>(p.D) + (size_t)row * p.ldd + col;
                    if (full) {
                        float4* dp = reinterpret_cast<float4*>(dst);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 o = make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                                                   __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
                            if (p.epi == EPI_F32_ACCUM) {
                                float4 old = dp[q];
                                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                            }
                            dp[q] = o;
                        }
                    } else {
                        for (int e = 0; e < 32 && col + e < p.N; ++e) {
                            float a = __uint_as_float(v[e]);
                            if (p.epi == EPI_F32_ACCUM) a += dst[e];
                            dst[e] = a;
                        }
                    }
                }
            };
            // The TMEM read of chunk c+1 is in flight while chunk c is converted, staged and stored: a chunk is otherwise a serial
            // chain (tcgen05.ld ~0.5 k cycles with 8 warps on the 64 B/clk port, then ~0.7 k of conversion / staging / stores),
            // which made the 512 x 256 tiling's two accumulator halves take ~5 k cycles each to drain (profiles/r02_gemm_wide_trace_*).
            uint32_t va[32], vb[32];
            tmem_ld_32x32b_x32(t_row + c_begin * 32, va);
            tmem_ld_wait();
#pragma unroll 1
            for (int c = c_begin; c < c_end; c += 2) {
                if (c + 1 < c_end) tmem_ld_32x32b_x32(t_row + (c + 1) * 32, vb);
                chunk(va, c);
                if (c + 1 < c_end) {
                    tmem_ld_wait();
                    if (c + 2 < c_end) tmem_ld_32x32b_x32(t_row + (c + 2) * 32, va);
                    chunk(vb, c + 1);
                    if (c + 2 < c_end) tmem_ld_wait();
                }
            }
            }
}

// kEpiWarps: 4 (one per TMEM lane quadrant) or 8 — two per quadrant, each taking half of the tile's columns.  With four, every
// SM sub-partition hosts ONE epilogue warp, so nothing hides the latency of its dependent ALU / MUFU chain: the exp-heavy epilogues
// (SwiGLU backward ~35 instructions per element) then take as long as the K = 4096 main loop and pace the GEMM (1.02 PFLOP/s in the
// step).  Eight warps double the epilogue's issue rate at the price of a 168-register cap.
template <int kCtaGroup, int kAMajor, int kBMajor, int kBlockN, int kEpiWarps = 4>   // 8 warps: used for the fused epilogues only
__global__ void __launch_bounds__(128 + 32 * kEpiWarps, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const GemmParams p) {
    using Cfg = GemmCfg<kCtaGroup, kAMajor, kBMajor, kBlockN>;
    const int kStages = p.stages;
    constexpr int BLOCK_K = Cfg::BLOCK_K;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + kStages * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    uint32_t* tmem_slot_ptr =
        reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    // Let a programmatically-dependent successor (GemmDesc::overlap_prev) be scheduled as soon as this grid's CTAs exit.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (kCtaGroup == 2) ? cluster_ctarank() : 0u;
    const bool is_leader = (cta_rank == 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), kEpiWarps * kCtaGroup);  // one arrive per epilogue warp per CTA
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<kCtaGroup>(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int num_k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int num_tiles = p.num_m_blocks * p.num_n_blocks;
    const int cluster_id = blockIdx.x / kCtaGroup;
    const int num_clusters = gridDim.x / kCtaGroup;
    const int kGroupM = p.group_m;

    auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
        const int per_group = kGroupM * p.num_n_blocks;
        const int group = tile / per_group;
        const int first_m = group * kGroupM;
        const int gsz = min(kGroupM, p.num_m_blocks - first_m);
        const int in_group = tile - group * per_group;
        m_blk = first_m + in_group % gsz;
        n_blk = in_group / gsz;
    };

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            int m_blk, n_blk;
            tile_coords(tile, m_blk, n_blk);
            const int m0 = m_blk * Cfg::TILE_M + (int)cta_rank * Cfg::BLOCK_M;
            int n0 = n_blk * Cfg::BLOCK_N + (int)cta_rank * Cfg::B_ROWS;
            if (p.epi == EPI_SWIGLU) n0 = n_blk * (Cfg::BLOCK_N / 2) + (int)cta_rank * p.n_half;   // gate rows | up rows
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u, 1);
                const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                const uint32_t sb = sa + Cfg::A_BYTES;
                const int k0 = kb * BLOCK_K;
                uint32_t bar;
                if constexpr (kCtaGroup == 2) {
                    if (is_leader) mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES * 2);
                    bar = mapa(full_bar(stage), 0);
                } else {
                    mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                    bar = full_bar(stage);
                }
                auto load = [&](uint32_t dst, const CUtensorMap* tm, int c0, int c1) {
                    if constexpr (kCtaGroup == 2) tma_load_2d_pair(dst, tm, bar, c0, c1);
                    else tma_load_2d(dst, tm, bar, c0, c1);
                };
                if constexpr (kAMajor == MAJOR_K) {
                    load(sa, &tmap_a, k0, m0);
                } else {
#pragma unroll
                    for (int j = 0; j < Cfg::BLOCK_M / 64; ++j)
                        load(sa + j * (BLOCK_K * 128), &tmap_a, m0 + 64 * j, k0);
                }
                if constexpr (kBMajor == MAJOR_K) {
                    load(sb, &tmap_b, k0, n0);
                } else {
#pragma unroll
                    for (int j = 0; j < Cfg::B_ROWS / 64; ++j)
                        load(sb + j * (BLOCK_K * 128), &tmap_b, n0 + 64 * j, k0);
                }
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1 && lane == 0 && is_leader) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc =
            make_idesc_bf16(Cfg::TILE_M, Cfg::BLOCK_N, kAMajor == MAJOR_MN, kBMajor == MAJOR_MN);
        // K-major SW128: 8-row groups are 1024 B apart (SBO); one UMMA_K step = 32 B inside the atom.
        // MN-major SW128: 64-element MN chunks are BLOCK_K*128 B apart (LBO), 8-k-row groups 1024 B
        // apart (SBO); one UMMA_K step = 16 k-rows = 2048 B.
        constexpr uint32_t a_lbo = (kAMajor == MAJOR_K) ? 0u : (uint32_t)(BLOCK_K * 128);
        constexpr uint32_t b_lbo = (kBMajor == MAJOR_K) ? 0u : (uint32_t)(BLOCK_K * 128);
        constexpr uint32_t a_kstep = (kAMajor == MAJOR_K) ? 32u : 2048u;
        constexpr uint32_t b_kstep = (kBMajor == MAJOR_K) ? 32u : 2048u;
        int stage = 0;
        uint32_t phase = 0;
        int astage = 0;
        uint32_t aphase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            mbar_wait(tempty_bar(astage), aphase ^ 1u, 2);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + astage * Cfg::BLOCK_N;
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(full_bar(stage), phase, 3);
                tc_fence_after();
                const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                const uint32_t sb = sa + Cfg::A_BYTES;
                const uint64_t adesc = make_smem_desc_sw128(sa, a_lbo, 1024);
                const uint64_t bdesc = make_smem_desc_sw128(sb, b_lbo, 1024);
#pragma unroll
                for (int k = 0; k < BLOCK_K / Cfg::UMMA_K; ++k) {
                    umma_bf16<kCtaGroup>(d_tmem, adesc + ((k * a_kstep) >> 4),
                                         bdesc + ((k * b_kstep) >> 4), idesc,
                                         (kb | k) != 0 ? 1u : 0u);
                }
                if constexpr (kCtaGroup == 2) umma_commit_pair(empty_bar(stage), 0b11);
                else umma_commit(empty_bar(stage));
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            if constexpr (kCtaGroup == 2) umma_commit_pair(tfull_bar(astage), 0b11);
            else umma_commit(tfull_bar(astage));
            if (++astage == 2) { astage = 0; aphase ^= 1u; }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int wq = warp & 3;  // TMEM lane quadrant this warp may read
        int astage = 0;
        uint32_t aphase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            int m_blk, n_blk;
            tile_coords(tile, m_blk, n_blk);
            const int row = m_blk * Cfg::TILE_M + (int)cta_rank * Cfg::BLOCK_M + wq * 32 + lane;
            const int n0 = n_blk * Cfg::BLOCK_N;
            mbar_wait(tfull_bar(astage), aphase, 4);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + astage * Cfg::BLOCK_N;
            const bool row_ok = row < p.M;
            uint8_t* wbuf = p.staged ? smem_raw + (bar_base + 512u - smem_u32(smem_raw)) + (warp - 4) * 2048 : nullptr;
            if (tile + num_clusters < num_tiles && warp < 8 && (p.epi == EPI_BF16_RESID || p.epi == EPI_SWIGLU_BWD)) {
                int nm, nn;
                tile_coords(tile + num_clusters, nm, nn);
                gemm_epilogue_prefetch<Cfg::BLOCK_N>(p, nm * Cfg::TILE_M + (int)cta_rank * Cfg::BLOCK_M + wq * 32 + lane, nn * Cfg::BLOCK_N);
            }
            gemm_epilogue_rows<Cfg::BLOCK_N, kEpiWarps == 8 ? 2 : 0>(p, row, row_ok, n_blk, n0, t_row, wbuf, lane, (warp - 4) >> 2, kEpiWarps / 4);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (kCtaGroup == 2) mbar_arrive_cluster(tempty_bar(astage), 0);
                else mbar_arrive(tempty_bar(astage));
            }
            if (++astage == 2) { astage = 0; aphase ^= 1u; }
        }
    }

    // ===================== teardown =====================
    __syncwarp();  // single-lane roles rejoin their warp before the aligned block/cluster barrier
    tc_fence_before();
    if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<kCtaGroup>(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace sf
