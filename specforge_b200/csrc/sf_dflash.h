// sf_dflash.h — internal declarations shared by sf_dflash_kernels.cu (kernels) and sf_dflash.cu (step + C ABI).
#pragma once
#include "sf_host.h"

namespace sf {
namespace dflash {

constexpr int kNormBlocks = 296;   // persistent grid of the per-head norm backward (partials reduced in block order)

struct AttnArgs {
    const __nv_bfloat16* q; int64_t ldq;          // roped q   [Mq, nh*d]
    const __nv_bfloat16* kn; int64_t ldkn;        // roped noise k [Mq, nkv*d]
    const __nv_bfloat16* vn; int64_t ldvn;        // noise v (view into the fused qkv rows)
    const __nv_bfloat16* kc; int64_t ldkc;        // roped context k [Mc, nkv*d]
    const __nv_bfloat16* vc; int64_t ldvc;        // context v (view into the fused kv rows)
    __nv_bfloat16* out; int64_t ldo;              // [Mq, nh*d]
    float* lse;                                   // [Mq, nh]  natural-log LSE of the scaled scores
    const int32_t* anchors; const uint8_t* keep;  // [B, N]
    int B, S, N, bs, nh, nkv, d;
    int window;                                   // sliding-window layers (dflash_family_model.py:73-84): slot o of a block anchored at a sees the
                                                  // context keys [a + o - (window - 1), a) and its block's slots <= o; 0 = full attention
    float scale;
    // backward
    const __nv_bfloat16* dout; int64_t lddo;
    float* delta;                                 // [Mq, nh]
    __nv_bfloat16* dq; int64_t lddq;              // [Mq, nh*d]
    __nv_bfloat16* dkn; int64_t lddkn;            // [Mq, nkv*d]
    __nv_bfloat16* dvn; int64_t lddvn;
    __nv_bfloat16* dkc; int64_t lddkc;            // [Mc, nkv*d]
    __nv_bfloat16* dvc; int64_t lddvc;
};

// sums[0] = loss_den, sums[1] = acc_den (rows), sums[2] = loss_num, sums[3] = correct (ce)
int rows(const int64_t* input_ids, const int64_t* loss_mask, const int32_t* anchors, const uint8_t* keep, int B, int S, int N, int bs,
         float gamma, int mask_id, int32_t* pos, int32_t* tgt, int32_t* noise_id, float* w, float* lw, float* sums, cudaStream_t st);
int gather_rows(const void* table, const int32_t* ids, void* out, int64_t M, int H, cudaStream_t st);
int headnorm_rope_fwd(const void* x, int64_t ldx, int n_heads, int d, const void* w, const void* cos_t, const void* sin_t,
                      const int32_t* pos, int S, float eps, void* out, int64_t ldo, int64_t M, cudaStream_t st);
int headnorm_rope_bwd(const void* x, int64_t ldx, int n_heads, int d, const void* w, const void* cos_t, const void* sin_t,
                      const int32_t* pos, int S, float eps, const void* g, int64_t ldg, void* dx, int64_t lddx, float* dw,
                      int accumulate, float* partial_ws, int64_t M, cudaStream_t st);
int attn_fwd(const AttnArgs& a, cudaStream_t st);       // dispatch: tcgen05 where supported, else CUDA cores
int attn_fwd_cc(const AttnArgs& a, cudaStream_t st);    // CUDA-core tiles (sf_dflash_kernels.cu)
int attn_bwd_cc(const AttnArgs& a, cudaStream_t st);
bool attn_tc_supported(const AttnArgs& a);          // sf_dflash_attn_tc.cu (default where the shape is covered)
int attn_fwd_tc(const AttnArgs& a, cudaStream_t st);
bool attn_tc_bwd_supported(const AttnArgs& a);      // sf_dflash_attn_tc_bwd.cu (default where the shape is covered)
int attn_bwd_tc(const AttnArgs& a, cudaStream_t st);
int attn_bwd(const AttnArgs& a, cudaStream_t st);
int ce(void* logits, int64_t ld, int V, const int32_t* tgt, const float* w, float* lw, float* sums, int write_grad,
       float* row_loss, float* row_correct, int64_t M, int grad_of_numerator, int loss_type, float dpace_alpha, int bs, int batch,
       float* row_state, cudaStream_t st);
int finalize_loss(const float* sums, float* metrics, float* loss, cudaStream_t st);

}  // namespace dflash
}  // namespace sf
