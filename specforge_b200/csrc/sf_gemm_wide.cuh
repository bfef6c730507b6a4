// sf_gemm_wide.cuh — the 512 x 256 tiling of the tcgen05 GEMM (same contract and epilogues as sf_gemm.cuh).
//
// Why a second tiling: what bounds the 256 x 256 kernel on a power-capped B200 is the operand traffic L2 -> shared memory
// (64 B per SM clock at full tensor rate; measured next to cuBLAS in profiles/r02_gemm_ncu_vs_cublas.csv: 4.29 GB vs 3.22 GB
// for the o_proj GEMM, 86 % vs 92 % tensor-pipe activity, 1.39 vs 1.48 GHz under the same 1 kW cap).  A CTA pair that owns
// 512 rows x 256 columns reads (512 + 256) x 64 operand elements per k-block for 512 x 256 x 64 MACs — 48 B per SM clock,
// 25 % less — at the price of filling all 512 TMEM columns with ONE tile's accumulators:
//   CTA rank r owns rows [256 r, 256 r + 256) of the pair tile as two 128-row halves h; half h accumulates in TMEM columns
//   [256 h, 256 h + 256) through its own cta_group::2 UMMA (M = 256 across the pair, N = 256), both halves reading the same B stage.
// Without a second accumulator stage the epilogue cannot hide behind the next tile's whole main loop; instead each half has its
// own full/empty barrier pair and its own epilogue warpgroup (warps 4-7: half 0, warps 8-11: half 1), the MMA issuer finishes
// half 0 of the last k-block first, and the next tile's half-0 MMAs start as soon as half 0 is drained — the exposed time per
// tile is about one half-drain minus one k-block of MMAs.
#pragma once
#include "sf_gemm.cuh"

namespace sf {

template <int kAMajor, int kBMajor>
struct GemmWideCfg {
    static constexpr int BLOCK_M = 256;          // rows per CTA: two 128-row halves
    static constexpr int TILE_M = 512;           // rows per CTA pair
    static constexpr int BLOCK_N = 256;
    static constexpr int BLOCK_K = 64;
    static constexpr int UMMA_K = 16;
    static constexpr int B_ROWS = 128;           // B rows staged by each CTA (the pair shares the 256)
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int A_HALF_BYTES = A_BYTES / 2;
    static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 48 KB
    static constexpr int kStages = 4;
    static constexpr int TMEM_COLS = 512;
    static constexpr int STAGING_BYTES = 8 * 2048;          // 2 KB per epilogue warp (coalesced bf16 transfers, sf_gemm.cuh)
    static constexpr int SMEM_BYTES = kStages * STAGE_BYTES + 1024 /*align*/ + 512 /*barriers*/ + STAGING_BYTES;
    static constexpr int kThreads = 384;
};

template <int kAMajor, int kBMajor, int kEpiSet>
__global__ void __launch_bounds__(384, 1)
gemm_wide_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
    using Cfg = GemmWideCfg<kAMajor, kBMajor>;
    constexpr int kStages = Cfg::kStages;
    constexpr int BLOCK_K = Cfg::BLOCK_K;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + kStages * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int h) { return bar_base + 8u * (2 * kStages + h); };
    auto tempty_bar = [&](int h) { return bar_base + 8u * (2 * kStages + 2 + h); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
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
        for (int h = 0; h < 2; ++h) {
            mbar_init(tfull_bar(h), 1);
            mbar_init(tempty_bar(h), 8 * 2);   // all eight epilogue warps of each CTA of the pair drain every half together
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<2>(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int num_k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int num_tiles = p.num_m_blocks * p.num_n_blocks;
    const int cluster_id = blockIdx.x / 2;
    const int num_clusters = gridDim.x / 2;
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
                if (is_leader) mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES * 2);
                const uint32_t bar = mapa(full_bar(stage), 0);
                if constexpr (kAMajor == MAJOR_K) {
                    tma_load_2d_pair(sa, &tmap_a, bar, k0, m0);                       // one [256 rows x 64 k] box
                } else {
#pragma unroll
                    for (int j = 0; j < Cfg::BLOCK_M / 64; ++j)
                        tma_load_2d_pair(sa + j * (BLOCK_K * 128), &tmap_a, bar, m0 + 64 * j, k0);
                }
                if constexpr (kBMajor == MAJOR_K) {
                    tma_load_2d_pair(sb, &tmap_b, bar, k0, n0);
                } else {
#pragma unroll
                    for (int j = 0; j < Cfg::B_ROWS / 64; ++j)
                        tma_load_2d_pair(sb + j * (BLOCK_K * 128), &tmap_b, bar, n0 + 64 * j, k0);
                }
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1 && lane == 0 && is_leader) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_bf16(256, Cfg::BLOCK_N, kAMajor == MAJOR_MN, kBMajor == MAJOR_MN);
        constexpr uint32_t a_lbo = (kAMajor == MAJOR_K) ? 0u : (uint32_t)(BLOCK_K * 128);
        constexpr uint32_t b_lbo = (kBMajor == MAJOR_K) ? 0u : (uint32_t)(BLOCK_K * 128);
        constexpr uint32_t a_kstep = (kAMajor == MAJOR_K) ? 32u : 2048u;
        constexpr uint32_t b_kstep = (kBMajor == MAJOR_K) ? 32u : 2048u;
        int stage = 0;
        uint32_t phase = 0, tphase = 0;
        // Half 0 runs ONE k-block ahead of half 1 ( h0(0) | h0(1) h1(0) | h0(2) h1(1) | ... | h1(n-1) ): half 0's accumulator is
        // complete two half-steps before half 1's, its drain starts earlier, and at the start of the next tile the tensor pipe has
        // two half-steps of half-0 work to do while half 1 is still being drained.
        auto issue_half = [&](int h, int st, bool first) {
            const uint32_t sa = smem_base + st * Cfg::STAGE_BYTES;
            const uint64_t adesc = make_smem_desc_sw128(sa + h * Cfg::A_HALF_BYTES, a_lbo, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(sa + Cfg::A_BYTES, b_lbo, 1024);
            const uint32_t d_tmem = tmem_base + h * Cfg::BLOCK_N;
#pragma unroll
            for (int k = 0; k < BLOCK_K / Cfg::UMMA_K; ++k)
                umma_bf16<2>(d_tmem, adesc + ((k * a_kstep) >> 4), bdesc + ((k * b_kstep) >> 4), idesc, (first && k == 0) ? 0u : 1u);
        };
        int tcount = 0;
        unsigned long long* tr = (p.trace && cluster_id == 0) ? p.trace : nullptr;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            if (tr && tcount < 32) tr[tcount * 16 + 0] = clock64();
            mbar_wait(full_bar(stage), phase, 3);
            if (tr && tcount < 32) tr[tcount * 16 + 1] = clock64();
            mbar_wait(tempty_bar(0), tphase ^ 1u, 2);     // half 0's accumulator drained by the epilogue warps
            if (tr && tcount < 32) tr[tcount * 16 + 2] = clock64();
            tc_fence_after();
            issue_half(0, stage, true);
            if (num_k_blocks == 1) umma_commit_pair(tfull_bar(0), 0b11);
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                int nstage = stage + 1;
                uint32_t nphase = phase;
                if (nstage == kStages) { nstage = 0; nphase ^= 1u; }
                if (kb + 1 < num_k_blocks) {
                    mbar_wait(full_bar(nstage), nphase, 3);
                    tc_fence_after();
                    issue_half(0, nstage, false);
                    if (kb + 2 == num_k_blocks) umma_commit_pair(tfull_bar(0), 0b11);
                }
                if (kb == 0) {
                    mbar_wait(tempty_bar(1), tphase ^ 1u, 2);
                    if (tr && tcount < 32) tr[tcount * 16 + 3] = clock64();
                    tc_fence_after();
                }
                issue_half(1, stage, kb == 0);
                umma_commit_pair(empty_bar(stage), 0b11);
                if (kb + 1 == num_k_blocks) umma_commit_pair(tfull_bar(1), 0b11);
                stage = nstage; phase = nphase;
                if (tr && tcount < 32 && kb == num_k_blocks / 2) tr[tcount * 16 + 4] = clock64();
            }
            if (tr && tcount < 32) tr[tcount * 16 + 5] = clock64();
            ++tcount;
            tphase ^= 1u;
        }
    } else if (warp >= 4) {
        // ===================== epilogue: both warpgroups drain half 0, then half 1 (128 columns each) =====================
        const int part = (warp - 4) >> 2;
        const int wq = warp & 3;  // TMEM lane quadrant this warp may read
        uint32_t tphase = 0;
        uint8_t* wbuf = p.staged ? smem_raw + (bar_base + 512u - smem_u32(smem_raw)) + (warp - 4) * 2048 : nullptr;
        int tcount = 0;
        unsigned long long* tr = (p.trace && cluster_id == 0 && is_leader && warp == 4 && lane == 0) ? p.trace : nullptr;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tcount) {
            int m_blk, n_blk;
            tile_coords(tile, m_blk, n_blk);
            const int n0 = n_blk * Cfg::BLOCK_N;
            if (tile + num_clusters < num_tiles && part == 0 && (p.epi == EPI_BF16_RESID || p.epi == EPI_SWIGLU_BWD)) {
                int nm, nn;
                tile_coords(tile + num_clusters, nm, nn);
                for (int h = 0; h < 2; ++h)
                    gemm_epilogue_prefetch<Cfg::BLOCK_N>(p, nm * Cfg::TILE_M + (int)cta_rank * Cfg::BLOCK_M + h * 128 + wq * 32 + lane, nn * Cfg::BLOCK_N);
            }
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int row = m_blk * Cfg::TILE_M + (int)cta_rank * Cfg::BLOCK_M + h * 128 + wq * 32 + lane;
                if (tr && tcount < 32) tr[tcount * 16 + 8 + h * 3] = clock64();
                mbar_wait(tfull_bar(h), tphase, 4);
                if (tr && tcount < 32) tr[tcount * 16 + 9 + h * 3] = clock64();
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + h * Cfg::BLOCK_N;
                gemm_epilogue_rows<Cfg::BLOCK_N, kEpiSet>(p, row, row < p.M, n_blk, n0, t_row, wbuf, lane, part, 2);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_bar(h), 0);
                if (tr && tcount < 32) tr[tcount * 16 + 10 + h * 3] = clock64();
            }
            tphase ^= 1u;
        }
    }

    // ===================== teardown =====================
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<2>(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace sf
