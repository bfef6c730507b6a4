/* specforge_b200.h — C ABI of libspecforge_b200.so: the B200 (sm_100a) EAGLE3 draft-head training step.
 *
 * The reference (sgl-project/SpecForge @ a6b6b02) has no FFI for this path — its seams are Python protocols
 * (SURVEY.md §8b).  This header is the boundary a binding sits on: plain pointers and sizes, no torch types.
 * All data pointers are DEVICE pointers owned by the caller (PyTorch), `stream` is a cudaStream_t passed as
 * void*.  Every function returns 0 on success or a negative errno-style code; sf_last_error() holds the
 * message (thread-local).  Nothing allocates: the caller provides a workspace sized by
 * sf_eagle3_workspace_bytes().  Re-entrant per device; a workspace must not be shared by concurrent steps.
 *
 * Reference interfaces replaced (file:line in the reference tree):
 *   sf_eagle3_forward   Eagle3TrainStrategy.forward_loss      specforge/training/strategies/base.py:237-304
 *                       TargetHead.preprocess/forward         specforge/modeling/target/target_head.py:100-108
 *                       _compute_target_p(_padded)            specforge/algorithms/eagle3/model.py:445-501
 *                       OnlineEagle3Model.forward (TTT loop)  specforge/algorithms/eagle3/model.py:244-442
 *                       LlamaForCausalLMEagle3 fc/backbone/compute_logits
 *                                                             specforge/modeling/draft/llama3_eagle.py:1625-1798
 *                       LogSoftmaxLoss / acceptance / top-1   specforge/core/loss.py:15-228, core/lk_loss.py:43-80
 *   sf_eagle3_backward  torch.autograd of the above (backend.backward)   specforge/training/backend.py:310-320
 *   sf_optimizer_step   BF16Optimizer.step                    specforge/optimizer.py:95-168
 *   sf_gemm_bf16 ...    the individual ops, exported for kernel-level parity tests
 */
#ifndef SPECFORGE_B200_H
#define SPECFORGE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ---- */
const char* sf_version(void);
const char* sf_last_error(void);
/* Tuning / A-B switch by name ("no_pdl", "loss_side", "no_overlap", "no_swiglu_fusion", "gemm_group_m",
 * "gemm_group_m_midk", "gemm_group_m_wgrad", "dflash_attn_tc" (-1: CUDA-core DFlash attention instead of tcgen05), "gemm_stages",
 * "gemm_wide" (-1: 256x256 tiling only, 0: 512x256 for M >= 512 by the K / epilogue rule in sf_gemm.cu), "gemm_epi_staged" (warp-staged
 * coalesced epilogue transfers; 0: 512x256 tiling + the SwiGLU-backward epilogue, 1: everywhere, 3: 512x256 only, -1: nowhere),
 * "gemm_epi8" (-1: four instead of eight epilogue warps for the fused epilogues), "dflash_attn_window" (window the op-level
 * sf_dflash_attention_fwd / _bwd calls apply; the step takes it from sf_dflash_config), "no_teacher_fusion", "no_loss_stats_fusion",
 * "no_rope_fusion"); each defaults to its SF_<NAME> environment variable.  Diagnostic only. */
int sf_debug_option(const char* name, int value);
long long sf_launch_count(void);      /* kernels launched by this library since the last reset */
void sf_launch_count_reset(void);

/* Optional live timing of every tcgen05 GEMM launch with CUDA events on the launching stream (bench.py roofline).
 * enable(1) resets the record; collect() (after a stream sync) returns the launch count and the summed device time
 * and algorithmic FLOPs (2*M*N*K) of the recorded launches. */
void sf_profile_gemm(int enable);
long long sf_profile_gemm_collect(double* total_ms, double* total_flops);
/* per launch: mnkt[4 i ..] = {M, N, K, rows per CTA-pair tile (256 | 512)}, ms[i] = device time; returns the number written */
long long sf_profile_gemm_detail(long long* mnkt, double* ms, long long max_n);
/* Diagnostic: device buffer of >= 512 uint64 that cluster 0 of the 512 x 256 GEMM tiling fills with clock64() stamps of its
 * MMA-issuer / epilogue hand-offs (16 per tile; tools/gemm_trace.py); NULL switches it off. */
void sf_debug_gemm_trace(unsigned long long* dev_buf);

/* ---- model / step description ---- */
typedef struct sf_eagle3_config {
    int32_t batch;            /* B  sequences per micro-batch on this GPU */
    int32_t seq_len;          /* S  */
    int32_t ttt_length;       /* T  (<= 9) */
    int32_t hidden_size;      /* H  draft hidden */
    int32_t target_hidden;    /* H_t target hidden (fc input is 3*H_t) */
    int32_t intermediate;     /* I  */
    int32_t num_heads;        /* nh */
    int32_t num_kv_heads;     /* nkv */
    int32_t head_dim;         /* d  (64 or 128) */
    int32_t vocab;            /* V  target vocab */
    int32_t draft_vocab;      /* DV */
    int32_t fc_norm;          /* EAGLE3.1 per-third RMSNorm before fc (llama3_eagle.py:1679-1687) */
    int32_t norm_output;      /* final norm before lm_head (default 1) */
    int32_t rope_rows;        /* rows in the cos/sin tables (>= S + T) */
    float rms_eps;
    float ploss_decay;        /* 0.8 */
    int32_t lk_loss_type;     /* 0 = KL only (default), 1 = "lambda", 2 = "alpha"  (core/lk_loss.py:83-99) */
    float kl_scale;           /* lambda: w = kl_scale * exp(-kl_decay * acceptance) */
    float kl_decay;
} sf_eagle3_config;

/* Offsets (in elements) of each trainable parameter inside the flat bf16 parameter buffer; the fp32 gradient
 * buffer uses the same layout.  q/k/v and gate/up are adjacent so that they form the fused [q;k;v] and
 * [gate;up] matrices the kernels read.  Parameter shapes are the reference state-dict shapes. */
enum {
    SF_P_FC = 0, SF_P_Q, SF_P_K, SF_P_V, SF_P_O, SF_P_GATE, SF_P_UP, SF_P_DOWN,
    SF_P_HIDDEN_NORM, SF_P_INPUT_NORM, SF_P_POST_NORM, SF_P_NORM, SF_P_LM_HEAD,
    SF_P_FC_NORM0, SF_P_FC_NORM1, SF_P_FC_NORM2, SF_P_COUNT
};
int sf_eagle3_param_layout(const sf_eagle3_config* cfg, int64_t offsets[SF_P_COUNT], int64_t sizes[SF_P_COUNT],
                           int64_t* total_elems);

typedef struct sf_eagle3_frozen {     /* not trained */
    const void* embed_tokens;         /* bf16 [V, H]   (frozen target embedding) */
    const void* target_head;          /* bf16 [V, H_t] (TargetHead.fc.weight)    */
    const void* rope_cos;             /* bf16 [rope_rows, d] */
    const void* rope_sin;             /* bf16 [rope_rows, d] */
    const uint8_t* t2d;               /* [V]  1 if the target token is in the draft vocab */
    const int64_t* d2t;               /* [DV] target id = draft id + d2t[draft id] */
} sf_eagle3_frozen;

typedef struct sf_eagle3_batch {      /* one collated micro-batch, already on the device */
    const int64_t* input_ids;         /* [B, S] */
    const int64_t* attention_mask;    /* [B, S] 1 = real token (key padding mask) */
    const int64_t* loss_mask;         /* [B, S] */
    const void* hidden_state;         /* bf16 [B, S, 3*H_t] aux hidden states */
    const void* target;               /* bf16 [B, S, H_t]  target last hidden state (unshifted) */
} sf_eagle3_batch;

size_t sf_eagle3_workspace_bytes(const sf_eagle3_config* cfg);

/* Teacher + TTT-unrolled forward + loss/metrics.  metrics: device float [T][8] =
 * {ploss, acc_correct, acc_denom, acceptance_rate, accept_num, accept_den, loss_denom, kl_weight};
 * loss: device float scalar = sum_j decay^j * ploss_j.  need_grad != 0 also leaves d(loss)/d(logits) and the
 * activations in the workspace for sf_eagle3_backward. */
int sf_eagle3_forward(const sf_eagle3_config* cfg, const void* params_flat, const sf_eagle3_frozen* frozen,
                      const sf_eagle3_batch* batch, void* workspace, size_t workspace_bytes, float* metrics,
                      float* loss, int need_grad, void* stream);

/* Location of a named tensor of the last step inside the workspace: "h" bf16 [T+1, M, H], "qkv" bf16 [T, M, (nh+2nkv)d] (after
 * RoPE), "attn" bf16 [T, M, nh*d], "hf" bf16 [T, M, H], "logits" bf16 [T, M, DV] (the logits after need_grad=0, d(loss)/d(logits)
 * after need_grad=1), "teacher_xg" bf16 [B, S+T, DV], "teacher_stats" f32 [B, S+T, 4], "teacher_ids" i64 [B, S+T],
 * "position_mask" i32 [B, S].  For parity tests (logits vs the reference's lm_head output) and forward-only inspection. */
int sf_eagle3_workspace_view(const sf_eagle3_config* cfg, const char* name, int64_t* offset_bytes, int64_t* size_bytes);

/* Full backward of the step just run by sf_eagle3_forward(need_grad=1) on the same workspace.
 * grads_flat_f32 (+)= loss_scale * dLoss/dParams  (accumulate != 0 adds into the buffer).  loss_scale != 1 is applied as one pass
 * over the finished buffer and is therefore refused (-EINVAL) together with `accumulate` or the gradient-ready callback: scale an
 * accumulation window once, through sf_grads_to_bf16's scale_dev or sf_optimizer_step's grad_scale. */
int sf_eagle3_backward(const sf_eagle3_config* cfg, const void* params_flat, const sf_eagle3_frozen* frozen,
                       const sf_eagle3_batch* batch, void* workspace, size_t workspace_bytes, float loss_scale,
                       float* grads_flat_f32, int accumulate, void* stream);

/* Same, with a host callback invoked right after the LAST kernel producing a group of adjacent parameters'
 * gradients has been enqueued (norm weights first, then lm_head, down, gate+up, o, q+k+v, fc[+fc_norm]): the caller
 * can record an event and start the data-parallel all-reduce of that slice on another stream while the remaining
 * weight-gradient GEMMs still run.  first_param indexes the SF_P_* enum; the slice is n_params adjacent entries. */
typedef void (*sf_grad_ready_fn)(int32_t first_param, int32_t n_params, void* user);
int sf_eagle3_backward_ex(const sf_eagle3_config* cfg, const void* params_flat, const sf_eagle3_frozen* frozen,
                          const sf_eagle3_batch* batch, void* workspace, size_t workspace_bytes, float loss_scale,
                          float* grads_flat_f32, int accumulate, sf_grad_ready_fn on_ready, void* user, void* stream);

/* grads_bf16[i] = bf16(grads_f32[i] * (scale_dev ? *scale_dev : 1))  — the bf16 gradient buffer DDP all-reduces
 * (backend.py:233-253).  scale_dev is an optional DEVICE float (e.g. the 1/accumulation_steps that autograd hands
 * to backward), so applying it needs no host synchronisation. */
int sf_grads_to_bf16(const float* grads_f32, void* grads_bf16, int64_t n, const float* scale_dev, void* stream);

/* BF16Optimizer.step (optimizer.py:140-168) fused: ||g|| over bf16(g * grad_scale), clip coefficient
 * min(1, max_norm/(||g||+1e-6)), AdamW on fp32 masters, bf16 write-back.  step is 1-based.
 * grad_norm_out: device float (the pre-clip norm); scratch: device float[1024]. */
int sf_optimizer_step(const void* grads_bf16, float* master, float* exp_avg, float* exp_avg_sq, void* params_bf16,
                      int64_t n, float grad_scale, float max_grad_norm, float lr, float beta1, float beta2, float eps,
                      float weight_decay, int32_t step, float* grad_norm_out, float* scratch, void* stream);

/* ---- individual ops (kernel-level parity tests) ---- */
/* D[m,n] = sum_k A(m,k) B(n,k); *_major: 0 = K-major ([rows,K] row-major), 1 = MN-major ([K,rows] row-major);
 * epi: 0 bf16, 1 bf16 + residual R, 2 fp32, 3 fp32 accumulate; cta_group: 0 auto, 1, 2. */
int sf_gemm_bf16(const void* A, int64_t lda, int a_major, const void* B, int64_t ldb, int b_major, void* D, int64_t ldd,
                 const void* R, int64_t ldr, int M, int N, int K, int epi, int cta_group, void* stream);
/* Same with the fused-SwiGLU epilogues (llama3_eagle.py:1547): epi 4 — B = [gate ; up] weight [2*n_half, K], D = gu [M, 2*n_half]
 * (bf16 gate | up, kept for backward), D2 = act [M, n_half] = bf16(bf16(silu(gate)) * up), N = 2*n_half, n_half % 128 == 0, M > 128;
 * epi 5 — acc = d(act) [M, n_half], R = gu [M, 2*n_half], D = d(gu) [M, 2*n_half], N = n_half. */
int sf_gemm_bf16_ex(const void* A, int64_t lda, int a_major, const void* B, int64_t ldb, int b_major, void* D, int64_t ldd,
                    const void* R, int64_t ldr, void* D2, int64_t ldd2, int n_half, int M, int N, int K, int epi, int cta_group,
                    void* stream);
/* out[i, :] = table[ids[i], :] (bf16 [V, H] table, int64 ids; embed_input_ids, llama3_eagle.py:1759-1760) */
int sf_embedding_gather(const void* table, int64_t V, int H, const int64_t* ids, int64_t n, void* out, void* stream);
/* D(bf16) [M, N] = RoPE(A W^T): the fused [q;k;v] projection with the rotary embedding applied in the GEMM epilogue to the columns
 * < rope_cols (heads of head_dim 64 | 128) at position (row % S) + pos_offset; cos / sin: bf16 [rows, head_dim] tables
 * (llama3_eagle.py:133-142,673-675,730-734).  Identical to sf_gemm_bf16 followed by sf_rope. */
int sf_gemm_bf16_rope(const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd, int M, int N, int K,
                      const void* cos_t, const void* sin_t, int S, int pos_offset, int head_dim, int rope_cols, void* stream);
int sf_rmsnorm_fwd(const void* x, int64_t ldx, const void* w, void* out, int64_t ldo, int64_t M, int H, float eps,
                   void* stream);
/* dw (+)= column sums via per-block partials in `scratch` (sf_rmsnorm_bwd_scratch_bytes(H) bytes): deterministic. */
int64_t sf_rmsnorm_bwd_scratch_bytes(int H);
int sf_rmsnorm_bwd(const void* x, int64_t ldx, const void* w, const void* dy, int64_t lddy, const void* add, void* dx,
                   float* dw, float* scratch, int64_t M, int H, float eps, void* stream);
/* TTT attention at step J over the fused per-step qkv buffers qkv[i] = [B*S, (nh+2nkv)*d] (RoPE already applied). */
/* key_mask: optional [B,S] bytes (1 = attend); kvlen_ws: 2*B ints of scratch, required with key_mask. */
int sf_ttt_attention_fwd(const void* const* qkv, int J, void* out, float* lse, float* sd_ws, const uint8_t* key_mask,
                         int* kvlen_ws, int B, int S, int nh, int nkv, int head_dim, void* stream);
int sf_ttt_attention_bwd(const void* const* qkv, int J, const void* out, const void* dout, const float* lse,
                         float* sd_ws, const uint8_t* key_mask, float* const* dk_acc, float* const* dv_acc, void* dq,
                         float* delta_ws, float* dq_diag_ws, int* kvlen_ws, int B, int S, int nh, int nkv, int head_dim,
                         void* stream);
int sf_swiglu_fwd(const void* gu, void* act, int64_t M, int I, void* stream);
int sf_swiglu_bwd(const void* gu, const void* dact, void* dgu, int64_t M, int I, void* stream);
int sf_rope(void* x, int64_t ld, int n_heads, int head_dim, const void* cos_t, const void* sin_t, int S, int pos_offset,
            int64_t M, int inverse, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Packed shards of offline feature records (host side of the input feed).
 * Replaces, for reading, the reference's one-`torch.save`-file-per-sample layout (scripts/prepare_hidden_states.py:446-480,
 * :578-595; reader runtime/data_plane/feature_store.py:235-240) followed by per-sample truncation
 * (algorithms/eagle3/data.py:10-27) and host-side pad + concatenate (data/utils.py:106-200).  File layout: see
 * specforge_b200/csrc/sf_shard.cpp.  No CUDA involved; destination buffers are plain host pointers (pin them).
 * ------------------------------------------------------------------------------------------------------------------ */
enum { SF_DT_U8 = 0, SF_DT_I32 = 1, SF_DT_I64 = 2, SF_DT_BF16 = 3, SF_DT_F16 = 4, SF_DT_F32 = 5, SF_DT_BOOL = 6 };
int sf_shard_open(const char* path, void** handle);
void sf_shard_close(void* handle);
int64_t sf_shard_num_records(void* handle);
int sf_shard_num_features(void* handle);
/* name40: caller buffer of 40 bytes (NUL-padded raw key, e.g. "aux_hidden_state"); width = elements per token */
int sf_shard_feature_info(void* handle, int feature, char* name40, int* dtype, int* elem_bytes, int64_t* width);
int64_t sf_shard_record_tokens(void* handle, int64_t record);
/* recompute the payload CRC-32 from disk: 0 = matches the index, 1 = mismatch, <0 = error */
int sf_shard_verify_record(void* handle, int64_t record);
/* Gather n_rec records into batch-major buffers: dst[f] is [n_rec, pad_tokens, width_f] (row-major) for feature f, or NULL
 * to skip it.  Each record contributes its first min(num_tokens, max_tokens) tokens; the rest of its pad_tokens rows are
 * zero-filled.  Fails if a truncated record is longer than pad_tokens.  n_threads pread() workers (<= 0: 4). */
int sf_shard_read_batch(void* handle, const int64_t* records, int n_rec, int64_t max_tokens, int64_t pad_tokens,
                        void* const* dst, int n_threads);

/* ------------------------------------------------------------------------------------------------------------------
 * DFlash block-parallel draft training step (first correct CUDA version).
 * Replaces OnlineDFlashModel.forward + autograd (algorithms/common/dflash_family_model.py:385-461) around
 * DFlashDraftModel.forward (modeling/draft/dflash.py:431-460).  Anchors are sampled by the caller (the reference draws them
 * with torch's RNG, dflash_family_model.py:179-210).  Same conventions as the EAGLE3 entry points: device pointers, one flat
 * bf16 parameter buffer, one flat fp32 gradient buffer, caller-owned workspace, sf_optimizer_step for the update.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct sf_dflash_config {
    int32_t batch, seq_len;           /* context: B x S target tokens                                                     */
    int32_t num_blocks, block_size;   /* N anchor blocks per sequence x bs draft slots  (config.block_size, num_anchors)  */
    int32_t hidden_size;              /* H (the target's hidden size: the context feature is num_target_feats * H wide)   */
    int32_t num_target_feats;         /* len(dflash_config.target_layer_ids)                                              */
    int32_t intermediate, num_heads, num_kv_heads, head_dim, num_layers, vocab;
    int32_t mask_token_id, rope_rows;
    float rms_eps;
    float loss_decay_gamma;           /* <= 0: no decay  (dflash_family_model.py:349-358)                                 */
    int32_t grad_of_numerator;        /* 0: backward yields d(loss_num / loss_den) (a self-contained step); 1: d(loss_num) —    */
                                      /* the reference controller's loss_terms contract, which backprops the numerator and then  */
                                      /* divides the synchronised gradients by the GLOBAL denominator (controller.py:334-398)     */
    int32_t loss_type;                /* 0 "dflash"; D-PACE (dflash_family_model.py:245-279,360-369): 1 "dpace",                   */
                                      /* 2 "dpace-cumulative-confidence-only", 3 "dpace-continuation-value-only"; loss_den = batch */
    float dpace_alpha;                /* smoothing of the draft's confidence on the target token (default 0.5)                    */
    int32_t sliding_window;           /* window W of the "sliding_attention" layers (config.sliding_window; dflash.py:38-68):      */
                                      /* slot o of a block anchored at a sees context keys [a + o - (W - 1), a) and own slots <= o */
                                      /* (dflash_family_model.py:73-84); 0 when no layer slides                                    */
    uint32_t sliding_layers;          /* bit l set: layer l is "sliding_attention" (config.layer_types)                            */
} sf_dflash_config;
/* parameter order inside the flat buffer: for each layer the SF_DF_* slices below (q, k, v contiguous = one fused GEMM
 * operand, likewise gate, up), then fc [H, F*H], hidden_norm [H], norm [H].  Names/shapes: dflash.py:336-375. */
enum { SF_DF_Q = 0, SF_DF_K, SF_DF_V, SF_DF_O, SF_DF_GATE, SF_DF_UP, SF_DF_DOWN, SF_DF_Q_NORM, SF_DF_K_NORM, SF_DF_INPUT_LN,
       SF_DF_POST_LN, SF_DF_PER_LAYER };
typedef struct sf_dflash_frozen {
    const void* embed_tokens;         /* [V, H] bf16  target embedding (noise tokens)        */
    const void* lm_head;              /* [V, H] bf16  target lm_head                         */
    const void* rope_cos;             /* [rope_rows, head_dim] bf16 (cos of cat(freqs, freqs)) */
    const void* rope_sin;
} sf_dflash_frozen;
typedef struct sf_dflash_batch {
    const int64_t* input_ids;         /* [B, S]                                              */
    const void* hidden_states;        /* [B, S, F*H] bf16 concatenated target layers         */
    const int64_t* loss_mask;         /* [B, S] 0/1                                          */
    const int32_t* anchors;           /* [B, N] sorted anchor positions (0 for dropped)      */
    const uint8_t* keep;              /* [B, N] block_keep_mask                              */
} sf_dflash_batch;
int sf_dflash_num_params(const sf_dflash_config* cfg);      /* 11 * num_layers + 3 */
int sf_dflash_param_layout(const sf_dflash_config* cfg, int64_t* offsets, int64_t* sizes, int64_t* total);
size_t sf_dflash_workspace_bytes(const sf_dflash_config* cfg);
/* metrics[4] = {loss_num, loss_den, correct, accuracy_den} (device, dflash_family_model.py:441-459); loss = num / den */
int sf_dflash_forward(const sf_dflash_config* cfg, const void* params_flat, const sf_dflash_frozen* frozen,
                      const sf_dflash_batch* batch, void* workspace, size_t workspace_bytes, float* metrics, float* loss,
                      int need_grad, void* stream);
int sf_dflash_backward(const sf_dflash_config* cfg, const void* params_flat, const sf_dflash_frozen* frozen,
                       const sf_dflash_batch* batch, void* workspace, size_t workspace_bytes, float* grads_flat_f32,
                       int accumulate, void* stream);

/* The DFlash block attention alone (contiguous [rows, heads*d] bf16 tensors; q/kn/vn/out/dq/dkn/dvn have B*N*bs rows, kc/vc/dkc/dvc
 * B*S rows; lse / delta_ws are [B*N*bs, nh] fp32).  impl: 0 = CUDA-core tiles, 1 = tcgen05, -1 = the step's choice (tcgen05 where the shape is covered).
 * Reference semantics: dflash_family_model.py:47-89 (mask) + dflash.py:185-213 (attention, dropped blocks give zeros). */
int sf_dflash_attention_fwd(const void* q, const void* kn, const void* vn, const void* kc, const void* vc, void* out, float* lse,
                            const int32_t* anchors, const uint8_t* keep, int B, int S, int N, int bs, int nh, int nkv, int d,
                            int impl, void* stream);
int sf_dflash_attention_bwd(const void* q, const void* kn, const void* vn, const void* kc, const void* vc, const void* out,
                            const float* lse, const void* dout, const int32_t* anchors, const uint8_t* keep, void* dq, void* dkn,
                            void* dvn, void* dkc, void* dvc, float* delta_ws, int B, int S, int N, int bs, int nh, int nkv, int d,
                            int impl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPECFORGE_B200_H */
