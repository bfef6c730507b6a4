// sf_host.h — host-side shared declarations for libspecforge_b200 (errors, launch counter, op descs).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace sf {

// Records a message retrievable through sf_last_error() and returns `code` (negative errno style).
int set_error(int code, const char* fmt, ...);
// Every kernel launch made by this library goes through here; bench.py reports the count.
void count_launch(int n = 1);

// Tuning / A-B switches.  Each starts from its environment variable (SF_<NAME>, upper case) and can be changed at run time
// through sf_debug_option() so that two settings can be alternated inside ONE process (boxes of the pool differ by +-8 %).
enum Opt { OPT_NO_PDL, OPT_LOSS_SIDE, OPT_NO_OVERLAP, OPT_NO_SWIGLU_FUSION, OPT_GEMM_GROUP_M, OPT_GEMM_GROUP_M_MIDK, OPT_GEMM_GROUP_M_WGRAD, OPT_DFLASH_ATTN_TC, OPT_GEMM_STAGES, OPT_NO_TEACHER_FUSION, OPT_NO_LOSS_STATS_FUSION, OPT_GEMM_WIDE, OPT_GEMM_EPI_STAGED, OPT_NO_ROPE_FUSION, OPT_GEMM_EPI8, OPT_DFLASH_ATTN_WINDOW, OPT_COUNT };
int opt(Opt o);

struct GemmDesc {
    const void* A; int64_t lda; int a_major;   // a_major/b_major: 0 = K-major, 1 = MN-major
    const void* B; int64_t ldb; int b_major;
    void* D; int64_t ldd;
    const void* R; int64_t ldr;                // residual (EPI_BF16_RESID) or null
    int M, N, K;
    int epi;                                   // sf::EPI_*
    int cta_group;                             // 0 = auto, 1, 2
    void* D2 = nullptr; int64_t ldd2 = 0;      // EPI_SWIGLU: act output [M, I]
    int n_half = 0;                            // EPI_SWIGLU(_BWD): I
    // EPI_BF16_STATS / EPI_TEACHER (see sf_gemm.cuh): partial row statistics [3 or 5][ceil(N/256)][M], draft-vocab gather
    float* stats = nullptr; const uint32_t* t2d_bits = nullptr; const int* t2d_prefix = nullptr;
    void* xg = nullptr; int S = 0, T = 0, DV = 0;
    // EPI_BF16_ROPE: RoPE of the columns < rope_cols at position (row % S) + rope_pos0 (cos / sin: bf16 [rows, head_dim] tables)
    const void* rope_cos = nullptr; const void* rope_sin = nullptr; int rope_cols = 0, rope_pos0 = 0, head_dim = 0;
    int overlap_prev = 0;                      // 1: independent of the previous kernel in the stream — programmatic
                                               // dependent launch lets its CTAs start on the SMs the previous GEMM's tail frees
};
int gemm(const GemmDesc& g, cudaStream_t stream);
int gemm_stats_blocks(const GemmDesc& g);   // partial blocks per row the statistics epilogues of g write (tiling-dependent)

// TTT attention at step j = J (J diagonal blocks).  k[i]/v[i]: [B*S, nkv*D] views of block i (row stride ldkv).
struct AttnDesc {
    const void* q; int64_t ldq;
    const void* k[9]; const void* v[9]; int64_t ldkv;
    void* out; int64_t ldo;
    float* lse;                 // [B, nh, S]
    float* sd_ws;               // [B, nh, S, J] workspace (diagonal scores)
    const uint8_t* key_mask;    // [B, S] or null
    const int* kvlen;           // [B] leading valid keys (required with key_mask), see mask_prefix()
    const int* nonprefix;       // [B] 1 if that row's mask is not of prefix form
    const void* q_row_base;     // base of the fused rows holding q (column q_col0 + h*d); tcgen05 path uses TMA on it
    const void* kv_row_base;    // base of the fused rows holding block-0 k / v
    int q_col0, k_col0, v_col0;
    int B, S, nh, nkv, head_dim, J;
    // backward only
    const void* dout; int64_t lddo;
    float* delta_ws;            // [B, nh, S]
    float* dk_acc[9]; float* dv_acc[9]; int64_t ldacc;   // fp32 accumulators per block, [B*S, nkv*D]
    void* dq; int64_t lddq;     // bf16 out
    float* dq_diag_ws;          // [B*S, nh*D] fp32 workspace
};
int attn_fwd(const AttnDesc& a, cudaStream_t st);
int attn_fwd_tc(const AttnDesc& a, cudaStream_t st);
int attn_bwd_tc(const AttnDesc& a, cudaStream_t st);
int mask_prefix(const uint8_t* key_mask, int B, int S, int* kvlen, int* nonprefix, cudaStream_t st);
#define SF_TRY_RC(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)
int attn_bwd(const AttnDesc& a, cudaStream_t st);

int rmsnorm_fwd(const void* x, int64_t ldx, const int64_t* ids, int S, int shift, const void* w, void* out, int64_t ldo,
                int64_t M, int H, float eps, float* rstd, cudaStream_t st);
int64_t rmsnorm_bwd_ws_bytes(int H);
int rmsnorm_bwd(const void* x, int64_t ldx, const int64_t* ids, int S, int shift, const void* w, const void* dy,
                int64_t lddy, const void* add1, const void* add2, void* dx, float* dw, float* partial_ws, int64_t M, int H,
                float eps, cudaStream_t st);
int rope(void* x, const float* src32, int64_t ld, int64_t ld32, int n_heads, int head_dim, const void* cos_t,
         const void* sin_t, int S, int pos_offset, int64_t M, int inverse, cudaStream_t st);
int cvt_f32_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t M, int cols, float scale, cudaStream_t st);
int swiglu_fwd(const void* gu, void* act, int64_t M, int I, cudaStream_t st);
int swiglu_bwd(const void* gu, const void* dact, void* dgu, int64_t M, int I, cudaStream_t st);
int shift_left(const void* src, void* dst, int64_t B, int S, int H, cudaStream_t st);
int add_bf16(const void* a, const void* b, void* out, int64_t n, cudaStream_t st);
int embedding_gather(const void* table, int64_t V, int H, const int64_t* ids, int64_t n, void* out, cudaStream_t st);

int teacher(const void* tl, int64_t ld, const int* d2t_idx, const uint8_t* t2d, const int* loss_mask, void* xg, float* tstats,
            int64_t* ids, int* position_mask, int B, int S, int T, int V, int DV, int64_t row0, int64_t nrows, int pad,
            int lean, cudaStream_t st);
int t2d_index(const uint8_t* t2d, int V, uint32_t* bits, int* prefix, cudaStream_t st);
int teacher_merge(const float* stats, int nb, int64_t M, const uint8_t* t2d, const int* loss_mask, float* tstats, int64_t* ids,
                  int* position_mask, void* xg, int B, int S, int T, int DV, cudaStream_t st);
// stats / stats_nb: optional pass-A partials [3][stats_nb][M] left by the lm_head GEMM's EPI_BF16_STATS epilogue
int loss_step(void* logits, int64_t ld, const void* xg, const float* tstats, const int64_t* tgt_ids,
              const int* position_mask, const int* loss_mask, const int64_t* d2t, int B, int S, int T, int DV, int step,
              float step_weight, int write_grad, int lk_type, float kl_scale, float kl_decay, float* row_ws, float* metrics,
              const float* stats, int stats_nb, cudaStream_t st);
int grad_norm(const void* g, int64_t n, float gscale, float* partials_ws, float* out, cudaStream_t st);
int adamw(const void* g, float* master, float* m1, float* m2, void* param, int64_t n, const float* gnorm, float max_norm,
          float gscale, float lr, float beta1, float beta2, float eps, float wd, int step, cudaStream_t st);
int cvt_flat_f32_bf16(const float* src, void* dst, int64_t n, const float* scale_dev, cudaStream_t st);

#define SF_CUDA_CHECK_LAUNCH(what)                                                            \
    do {                                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess) return ::sf::set_error(-5, "%s: %s", what, cudaGetErrorString(e__)); \
        ::sf::count_launch();                                                                 \
    } while (0)

}  // namespace sf
