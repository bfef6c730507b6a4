// sf_gemm.cu — host launcher + C-ABI for the tcgen05 GEMM (see sf_gemm.cuh).
#include "sf_gemm.cuh"
#include "sf_gemm_wide.cuh"
#include "sf_host.h"

#include <cudaTypedefs.h>
#include <mutex>
#include <vector>

namespace sf {

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

// 2-D bf16 tensor map over a row-major [rows, cols] matrix with row stride `ld` elements,
// box = [box_rows, box_cols], 128-byte swizzle (box_cols * 2 must be 128).
int make_tmap_2d_bf16(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld,
                      int box_rows, int box_cols) {
    auto fn = get_encode_fn();
    if (!fn) return set_error(-38, "cuTensorMapEncodeTiled entry point not found (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 8) != 0)
        return set_error(-22, "TMA operand must be 16-byte aligned with a leading dim multiple of 8 (got ld=%lld)",
                         (long long)ld);
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(-22, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

// ---- optional live timing of every GEMM launch (bench.py's roofline): CUDA events on the launching stream
struct GemmProf {
    bool on = false;
    std::vector<cudaEvent_t> ev;   // pairs
    std::vector<double> flops;
    std::vector<long long> shape;  // 4 per launch: M, N, K, tiling (256 or 512 rows per CTA pair)
    size_t used = 0;
};
static GemmProf g_prof;
static unsigned long long* g_trace = nullptr;   // sf_debug_gemm_trace
static std::mutex g_prof_mu;

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

// the epilogues whose per-element work is long enough to pace a K <= 8192 GEMM with four epilogue warps (sf_gemm.cuh)
static bool heavy_epilogue(int epi) {
    return epi == EPI_TEACHER || epi == EPI_BF16_STATS || epi == EPI_SWIGLU || epi == EPI_SWIGLU_BWD || epi == EPI_BF16_ROPE;
}
static bool use_epi8(const GemmDesc& g) { return opt(OPT_GEMM_EPI8) >= 0 && heavy_epilogue(g.epi); }   // the 8-warp instance holds the fused epilogues only

template <int G, int AM, int BM, int BN, int EW = 4>
static int launch_cfg(const GemmDesc& g, cudaStream_t stream) {
    using Cfg = GemmCfg<G, AM, BM, BN>;
    CUtensorMap ta, tb;
    int rc;
    if (AM == MAJOR_K) rc = make_tmap_2d_bf16(&ta, g.A, g.M, g.K, g.lda, Cfg::BLOCK_M, 64);
    else               rc = make_tmap_2d_bf16(&ta, g.A, g.K, g.M, g.lda, 64, 64);
    if (rc) return rc;
    if (BM == MAJOR_K) rc = make_tmap_2d_bf16(&tb, g.B, g.N, g.K, g.ldb, Cfg::B_ROWS, 64);
    else               rc = make_tmap_2d_bf16(&tb, g.B, g.K, g.N, g.ldb, 64, 64);
    if (rc) return rc;

    GemmParams p;
    p.D = g.D; p.R = reinterpret_cast<const __nv_bfloat16*>(g.R);
    p.M = g.M; p.N = g.N; p.K = g.K; p.ldd = (int)g.ldd; p.ldr = (int)g.ldr; p.epi = g.epi;
    p.D2 = g.D2; p.ldd2 = (int)g.ldd2; p.n_half = g.n_half;
    p.stats = g.stats; p.t2d_bits = g.t2d_bits; p.t2d_prefix = g.t2d_prefix; p.xg = reinterpret_cast<__nv_bfloat16*>(g.xg);
    p.S = g.S; p.T = g.T; p.DV = g.DV;
    p.rope_cos = reinterpret_cast<const __nv_bfloat16*>(g.rope_cos); p.rope_sin = reinterpret_cast<const __nv_bfloat16*>(g.rope_sin);
    p.rope_cols = g.rope_cols; p.rope_pos0 = g.rope_pos0; p.head_dim = g.head_dim;
    p.trace = g_trace;
    p.num_m_blocks = (g.M + Cfg::TILE_M - 1) / Cfg::TILE_M;
    p.num_n_blocks = (g.N + Cfg::BLOCK_N - 1) / Cfg::BLOCK_N;
    // Raster: tiles walk N inside groups of `group_m` M-blocks, so a group's A rows stay L2-resident while B streams.
    // Forward / dgrad GEMMs (K <= 16K) measured fastest with 16 (half the B re-reads of 8: -4 % step time); the
    // weight-gradient GEMMs (K = T*M) cannot keep an A panel in L2 at all and want the squarest wave (8 x 9 tiles).
    const int gm = opt(OPT_GEMM_GROUP_M), gmk = opt(OPT_GEMM_GROUP_M_MIDK), gmw = opt(OPT_GEMM_GROUP_M_WGRAD);
    p.group_m = g.K > 32768 ? (gmw > 0 ? gmw : 8) : g.K > 4608 ? (gmk > 0 ? gmk : 16) : (gm > 0 ? gm : 16);
    p.stages = opt(OPT_GEMM_STAGES) > 0 ? opt(OPT_GEMM_STAGES) : Cfg::kStages;
    if (p.stages < 2) p.stages = 2;
    if (p.stages > Cfg::kMaxStages) p.stages = Cfg::kMaxStages;
    // warp-staged (coalesced) bf16 epilogue transfers: needs 8 KB, i.e. one ring stage less than the deepest ring
    // gemm_epi_staged: 0 (default) / 2 = only the SwiGLU-backward epilogue (2 x 64 B read + 2 x 64 B written per chunk) — in the step
    // 238.9 vs 242.7 ms (none, value 3) vs 240.3 (every launch, value 1), profiles/r02_ab_gemm_tiling.jsonl; -1 = no staging in either tiling
    const int so = opt(OPT_GEMM_EPI_STAGED);
    p.staged = (so == 1 || ((so == 0 || so == 2) && g.epi == EPI_SWIGLU_BWD)) ? 1 : 0;
    while (p.staged && Cfg::smem_bytes(p.stages, 1) > 227 * 1024) --p.stages;
    const int smem_bytes = Cfg::smem_bytes(p.stages, p.staged);
    const int tiles = p.num_m_blocks * p.num_n_blocks;
    int clusters = num_sms() / G;
    if (tiles < clusters) clusters = tiles;

    auto kern = gemm_kernel<G, AM, BM, BN, EW>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes(Cfg::kMaxStages, 0));
        if (e != cudaSuccess) return set_error(-22, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * G);
    cfg.blockDim = dim3(128 + 32 * EW);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (g.overlap_prev && !opt(OPT_NO_PDL) && !g_prof.on) {   // per-launch profiling events would sit between the two kernels
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 2;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (g_prof.on) {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        if (g_prof.used + 2 > g_prof.ev.size()) {
            for (int i = 0; i < 2; ++i) { cudaEvent_t ev; cudaEventCreate(&ev); g_prof.ev.push_back(ev); }
        }
        e0 = g_prof.ev[g_prof.used]; e1 = g_prof.ev[g_prof.used + 1];
        g_prof.used += 2;
        g_prof.flops.push_back(2.0 * (double)g.M * (double)g.N * (double)g.K);
        g_prof.shape.insert(g_prof.shape.end(), {(long long)g.M, (long long)g.N, (long long)g.K, (long long)Cfg::TILE_M});
        cudaEventRecord(e0, stream);
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
    if (e1) cudaEventRecord(e1, stream);
    if (e != cudaSuccess) return set_error(-5, "gemm launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return 0;
}

// ---- 512 x 256 tiling (sf_gemm_wide.cuh)
template <int AM, int BM, int EPISET>
static int launch_wide_impl(const GemmDesc& g, cudaStream_t stream) {
    using Cfg = GemmWideCfg<AM, BM>;
    CUtensorMap ta, tb;
    int rc;
    if (AM == MAJOR_K) rc = make_tmap_2d_bf16(&ta, g.A, g.M, g.K, g.lda, Cfg::BLOCK_M, 64);
    else               rc = make_tmap_2d_bf16(&ta, g.A, g.K, g.M, g.lda, 64, 64);
    if (rc) return rc;
    if (BM == MAJOR_K) rc = make_tmap_2d_bf16(&tb, g.B, g.N, g.K, g.ldb, Cfg::B_ROWS, 64);
    else               rc = make_tmap_2d_bf16(&tb, g.B, g.K, g.N, g.ldb, 64, 64);
    if (rc) return rc;
    GemmParams p;
    p.D = g.D; p.R = reinterpret_cast<const __nv_bfloat16*>(g.R);
    p.M = g.M; p.N = g.N; p.K = g.K; p.ldd = (int)g.ldd; p.ldr = (int)g.ldr; p.epi = g.epi;
    p.D2 = g.D2; p.ldd2 = (int)g.ldd2; p.n_half = g.n_half;
    p.stats = g.stats; p.t2d_bits = g.t2d_bits; p.t2d_prefix = g.t2d_prefix; p.xg = reinterpret_cast<__nv_bfloat16*>(g.xg);
    p.S = g.S; p.T = g.T; p.DV = g.DV;
    p.rope_cos = reinterpret_cast<const __nv_bfloat16*>(g.rope_cos); p.rope_sin = reinterpret_cast<const __nv_bfloat16*>(g.rope_sin);
    p.rope_cols = g.rope_cols; p.rope_pos0 = g.rope_pos0; p.head_dim = g.head_dim;
    p.trace = g_trace;
    p.num_m_blocks = (g.M + Cfg::TILE_M - 1) / Cfg::TILE_M;
    p.num_n_blocks = (g.N + Cfg::BLOCK_N - 1) / Cfg::BLOCK_N;
    const int gm = opt(OPT_GEMM_GROUP_M), gmk = opt(OPT_GEMM_GROUP_M_MIDK), gmw = opt(OPT_GEMM_GROUP_M_WGRAD);
    // same raster windows as the 256-row tiling, in 512-row blocks
    p.group_m = (g.K > 32768 ? (gmw > 0 ? gmw : 8) : g.K > 4608 ? (gmk > 0 ? gmk : 16) : (gm > 0 ? gm : 16)) / 2;
    if (p.group_m < 1) p.group_m = 1;
    p.stages = Cfg::kStages;
    p.staged = opt(OPT_GEMM_EPI_STAGED) >= 0 ? 1 : 0;
    const int tiles = p.num_m_blocks * p.num_n_blocks;
    int clusters = num_sms() / 2;
    if (tiles < clusters) clusters = tiles;
    auto kern = gemm_wide_kernel<AM, BM, EPISET>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return set_error(-22, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * 2);
    cfg.blockDim = dim3(Cfg::kThreads);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (g.overlap_prev && !opt(OPT_NO_PDL) && !g_prof.on) {
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 2;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (g_prof.on) {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        if (g_prof.used + 2 > g_prof.ev.size()) {
            for (int i = 0; i < 2; ++i) { cudaEvent_t ev; cudaEventCreate(&ev); g_prof.ev.push_back(ev); }
        }
        e0 = g_prof.ev[g_prof.used]; e1 = g_prof.ev[g_prof.used + 1];
        g_prof.used += 2;
        g_prof.flops.push_back(2.0 * (double)g.M * (double)g.N * (double)g.K);
        g_prof.shape.insert(g_prof.shape.end(), {(long long)g.M, (long long)g.N, (long long)g.K, (long long)Cfg::TILE_M});
        cudaEventRecord(e0, stream);
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
    if (e1) cudaEventRecord(e1, stream);
    if (e != cudaSuccess) return set_error(-5, "gemm (512x256 tiling) launch failed: %s", cudaGetErrorString(e));
    count_launch();
    return 0;
}
template <int AM, int BM>
static int launch_wide(const GemmDesc& g, cudaStream_t stream) {
    return heavy_epilogue(g.epi) ? launch_wide_impl<AM, BM, 2>(g, stream) : launch_wide_impl<AM, BM, 1>(g, stream);
}
// Which tiling.  The wide one moves 25 % fewer operand bytes per FLOP but cannot hide its epilogue behind the next tile's main loop
// (all 512 TMEM columns hold one tile), so the accumulator drain is paid once per tile — and TMEM reads run at 64 B/clk per SM
// (tools/gemm_trace.py: ~5 k cycles per 128 x 256 half next to the other half's MMAs, ~11 k exposed per tile of 65 k at K = 4096).
// Measured on B200 next to the 256 x 256 tiling and cuBLAS (profiles/r02_gemm_vs_cublas_d.txt, isolated kernels, random operands):
// K = 4096 tiles lose 3 %, K = 8192 gains 4 %, K = 12 288 gains 9 %, K >= 24 576 gains 13-20 % (0.95-0.97 of cuBLAS).  Inside the step
// (bench.py --ab gemm_wide=0,2,-1, alternated step by step on one box, profiles/r02_ab_gemm_tiling.jsonl): wide for K > 4096 only
// 230.2 ms, wide everywhere but the statistics epilogues 243.3 ms, never 240.4 ms.
// In-step rates by shape (bench.py roofline.gemm_by_shape) show why: with its drain exposed the wide tiling runs the SwiGLU
// epilogues at 780-1065 TFLOP/s at K = 4096 (1350 in the 256 x 256 tiling, where the epilogue hides behind the next tile).
// gemm_wide: -1 never; 0 (default) M >= 512 and K > 4096, K > 8192 for the heavy epilogues (SwiGLU fwd/bwd, RoPE, row statistics);
// 1 every GEMM except the row-statistics epilogues at K <= 4096; 2 always; 3 only K > 8192; 4 every plain epilogue, fused ones for K > 8192.
static bool use_wide(const GemmDesc& g) {
    const int o = opt(OPT_GEMM_WIDE);
    if (o < 0 || g.M < 512) return false;
    if (o == 2) return true;
    if (o == 3) return g.K > 8192;
    const bool stats = g.epi == EPI_TEACHER || g.epi == EPI_BF16_STATS;
    if (o == 1) return g.K > 4096 || !stats;
    const bool heavy = heavy_epilogue(g.epi);
    if (o == 4) return !heavy || g.K > 8192;
    return g.K > (heavy ? 8192 : 4096);
}

// Number of per-row partial blocks the row-statistics epilogues (EPI_BF16_STATS / EPI_TEACHER) of this GEMM will write: one per
// 256-column n-block in the 256 x 256 tiling, two (column halves, one per epilogue warpgroup) in the 512 x 256 tiling.
int gemm_stats_blocks(const GemmDesc& g) {
    const int nb = (g.N + 255) / 256;
    int G = g.cta_group;
    if (G == 0) G = (g.M > 128) ? 2 : 1;
    if (G == 2 && g.cta_group == 0 && use_wide(g)) return 2 * nb;
    return (G == 2 && use_epi8(g) && (g.a_major == MAJOR_K)) ? 2 * nb : nb;    // two epilogue warps per lane quadrant: two column halves
}

int gemm(const GemmDesc& g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return set_error(-22, "gemm: empty problem %dx%dx%d", g.M, g.N, g.K);
    if (g.epi < 0 || g.epi > 8) return set_error(-22, "gemm: bad epilogue %d", g.epi);
    if (g.epi == EPI_BF16_ROPE && (!g.rope_cos || !g.rope_sin || g.S <= 0 || (g.head_dim != 64 && g.head_dim != 128) || g.rope_cols % g.head_dim ||
                                   g.N % g.head_dim || g.rope_cols > g.N))
        return set_error(-22, "gemm: RoPE epilogue needs cos/sin tables, S, head_dim 64|128, rope_cols and N multiples of head_dim");
    if ((g.epi == EPI_BF16_STATS || g.epi == EPI_TEACHER) && !g.stats) return set_error(-22, "gemm: statistics epilogue without a stats buffer");
    if (g.epi == EPI_TEACHER && (!g.t2d_bits || !g.t2d_prefix || !g.xg || g.S <= 0 || g.DV <= 0)) return set_error(-22, "gemm: teacher epilogue needs t2d_bits, t2d_prefix, xg, S, DV");
    if (g.epi == EPI_SWIGLU) {
        if (g.cta_group == 1 || g.M <= 128) return set_error(-22, "gemm: fused SwiGLU needs the 2-CTA configuration (M > 128)");
        if (!g.D2 || g.n_half <= 0 || g.N != 2 * g.n_half || g.n_half % 128) return set_error(-22, "gemm: fused SwiGLU needs D2 and N == 2*n_half, n_half %% 128 == 0");
        if (g.a_major || g.b_major) return set_error(-22, "gemm: fused SwiGLU forward is K-major x K-major");
    }
    if (g.epi == EPI_SWIGLU_BWD && (!g.R || g.n_half <= 0 || g.N != g.n_half || g.n_half % 32)) return set_error(-22, "gemm: fused SwiGLU backward needs R = gu and N == n_half");
    if (g.epi == EPI_BF16_RESID && !g.R) return set_error(-22, "gemm: residual epilogue without R");
    const int elt = (g.epi == EPI_F32 || g.epi == EPI_F32_ACCUM) ? 4 : 2;
    if (g.epi != EPI_TEACHER && ((reinterpret_cast<uintptr_t>(g.D) & 15) || (g.ldd * elt) % 16))
        return set_error(-22, "gemm: D must be 16-byte aligned with 16-byte aligned rows");
    if (g.epi == EPI_BF16_RESID && ((reinterpret_cast<uintptr_t>(g.R) & 15) || (g.ldr % 8)))
        return set_error(-22, "gemm: R must be 16-byte aligned with ld multiple of 8");
    int G = g.cta_group;
    if (G == 0) G = (g.M > 128) ? 2 : 1;
    const int key = (G == 2 ? 4 : 0) | (g.a_major ? 2 : 0) | (g.b_major ? 1 : 0);
    if (G == 2 && g.cta_group == 0 && use_wide(g)) {
        switch (key) {
            case 4: return launch_wide<MAJOR_K, MAJOR_K>(g, stream);
            case 5: return launch_wide<MAJOR_K, MAJOR_MN>(g, stream);
            case 7: return launch_wide<MAJOR_MN, MAJOR_MN>(g, stream);
            default: break;
        }
    }
    switch (key) {
        case 0: return launch_cfg<1, MAJOR_K, MAJOR_K, 256>(g, stream);
        case 1: return launch_cfg<1, MAJOR_K, MAJOR_MN, 256>(g, stream);
        case 3: return launch_cfg<1, MAJOR_MN, MAJOR_MN, 256>(g, stream);
        case 4: return use_epi8(g) ? launch_cfg<2, MAJOR_K, MAJOR_K, 256, 8>(g, stream) : launch_cfg<2, MAJOR_K, MAJOR_K, 256>(g, stream);
        case 5: return use_epi8(g) ? launch_cfg<2, MAJOR_K, MAJOR_MN, 256, 8>(g, stream) : launch_cfg<2, MAJOR_K, MAJOR_MN, 256>(g, stream);
        case 7: return launch_cfg<2, MAJOR_MN, MAJOR_MN, 256>(g, stream);
        default: return set_error(-22, "gemm: unsupported operand majors a=%d b=%d", g.a_major, g.b_major);
    }
}

}  // namespace sf

extern "C" void sf_profile_gemm(int enable) {
    std::lock_guard<std::mutex> lk(sf::g_prof_mu);
    sf::g_prof.on = enable != 0;
    sf::g_prof.used = 0;
    sf::g_prof.flops.clear();
    sf::g_prof.shape.clear();
}
// Per-launch detail of the recorded GEMMs (after a stream sync): mnkt[4 i ..] = M, N, K, rows per pair tile; ms[i] = device time.
extern "C" long long sf_profile_gemm_detail(long long* mnkt, double* ms, long long max_n) {
    std::lock_guard<std::mutex> lk(sf::g_prof_mu);
    const size_t n = sf::g_prof.used / 2;
    long long out = 0;
    for (size_t i = 0; i < n && out < max_n; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, sf::g_prof.ev[2 * i], sf::g_prof.ev[2 * i + 1]) != cudaSuccess) continue;
        for (int k = 0; k < 4; ++k) mnkt[4 * out + k] = sf::g_prof.shape[4 * i + k];
        ms[out++] = t;
    }
    return out;
}
// Sums the recorded launches (call after synchronising the stream).  Returns the number of launches.
extern "C" long long sf_profile_gemm_collect(double* total_ms, double* total_flops) {
    std::lock_guard<std::mutex> lk(sf::g_prof_mu);
    double ms = 0, fl = 0;
    const size_t n = sf::g_prof.used / 2;
    for (size_t i = 0; i < n; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, sf::g_prof.ev[2 * i], sf::g_prof.ev[2 * i + 1]) == cudaSuccess) { ms += t; fl += sf::g_prof.flops[i]; }
    }
    if (total_ms) *total_ms = ms;
    if (total_flops) *total_flops = fl;
    return (long long)n;
}

extern "C" int sf_gemm_bf16(const void* A, int64_t lda, int a_major, const void* B, int64_t ldb, int b_major,
                            void* D, int64_t ldd, const void* R, int64_t ldr, int M, int N, int K, int epi,
                            int cta_group, void* stream) {
    sf::GemmDesc g;
    g.A = A; g.lda = lda; g.a_major = a_major; g.B = B; g.ldb = ldb; g.b_major = b_major;
    g.D = D; g.ldd = ldd; g.R = R; g.ldr = ldr; g.M = M; g.N = N; g.K = K; g.epi = epi; g.cta_group = cta_group;
    return sf::gemm(g, reinterpret_cast<cudaStream_t>(stream));
}

// Every epilogue, including the fused SwiGLU ones (epi 4: D = gu [M, 2*n_half], D2 = act [M, n_half]; epi 5: R = gu, D = d(gu)).
extern "C" int sf_gemm_bf16_ex(const void* A, int64_t lda, int a_major, const void* B, int64_t ldb, int b_major,
                               void* D, int64_t ldd, const void* R, int64_t ldr, void* D2, int64_t ldd2, int n_half,
                               int M, int N, int K, int epi, int cta_group, void* stream) {
    sf::GemmDesc g;
    g.A = A; g.lda = lda; g.a_major = a_major; g.B = B; g.ldb = ldb; g.b_major = b_major;
    g.D = D; g.ldd = ldd; g.R = R; g.ldr = ldr; g.M = M; g.N = N; g.K = K; g.epi = epi; g.cta_group = cta_group;
    g.D2 = D2; g.ldd2 = ldd2; g.n_half = n_half;
    return sf::gemm(g, reinterpret_cast<cudaStream_t>(stream));
}

// Diagnostic: device buffer (>= 32 tiles x 16 stamps of clock64()) that cluster 0 / CTA 0 of the 512 x 256 GEMM kernel fills with the
// times of its MMA-issuer and epilogue hand-offs; nullptr switches the tracing off.  See tools/gemm_trace.py.
extern "C" void sf_debug_gemm_trace(unsigned long long* dev_buf) { sf::g_trace = dev_buf; }
