// sf_eagle3.cu — the EAGLE3 draft-head training step: workspace plan, forward, backward, C ABI.
//
// Data layout in HBM (M = B*S rows, row r = b*S + s, everything row-major, bf16 unless noted):
//   every per-TTT-step activation lives in one [T, M, *] slab so that (a) step j is a pointer offset and
//   (b) all T steps' inputs / output-gradients of a weight are contiguous along the token axis: each weight
//   gradient is then ONE tcgen05 GEMM contracting over T*M tokens with fp32 accumulation in TMEM
//   (instead of T read-modify-write passes).  180 GB of HBM3e make keeping those slabs free.
#include "sf_gemm.cuh"
#include "sf_host.h"
#include "../../include/specforge_b200.h"

#include <cstdlib>
#include <cstring>

namespace sf {

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct Dims {
    int B, S, T, H, Ht, I, nh, nkv, d, V, DV;
    int64_t M, A, KV, QKV;
};
static Dims dims_of(const sf_eagle3_config& c) {
    Dims x;
    x.B = c.batch; x.S = c.seq_len; x.T = c.ttt_length; x.H = c.hidden_size; x.Ht = c.target_hidden;
    x.I = c.intermediate; x.nh = c.num_heads; x.nkv = c.num_kv_heads; x.d = c.head_dim; x.V = c.vocab; x.DV = c.draft_vocab;
    x.M = (int64_t)x.B * x.S; x.A = (int64_t)x.nh * x.d; x.KV = (int64_t)x.nkv * x.d; x.QKV = x.A + 2 * x.KV;
    return x;
}

static int validate(const sf_eagle3_config& c) {
    if (c.batch <= 0 || c.seq_len <= 0) return set_error(-22, "config: empty batch (B=%d, S=%d)", c.batch, c.seq_len);
    if (c.ttt_length < 1 || c.ttt_length > 9) return set_error(-22, "config: ttt_length=%d outside [1, 9]", c.ttt_length);
    if (c.head_dim != 64 && c.head_dim != 128) return set_error(-22, "config: head_dim=%d (64 or 128 supported)", c.head_dim);
    if (c.hidden_size % 8 || c.target_hidden % 8 || c.intermediate % 8 || c.draft_vocab % 8 || c.vocab % 8)
        return set_error(-22, "config: hidden/intermediate/vocab sizes must be multiples of 8");
    if (c.hidden_size > 8192 || c.target_hidden > 8192) return set_error(-22, "config: hidden size > 8192 unsupported");
    if (c.num_heads % c.num_kv_heads) return set_error(-22, "config: num_heads %% num_kv_heads != 0");
    if (c.rope_rows < c.seq_len + c.ttt_length) return set_error(-22, "config: rope tables have %d rows, need >= S+T=%d", c.rope_rows, c.seq_len + c.ttt_length);
    if (c.lk_loss_type < 0 || c.lk_loss_type > 2) return set_error(-22, "config: lk_loss_type=%d (0 none, 1 lambda, 2 alpha)", c.lk_loss_type);
    return 0;
}

// ------------------------------------------------------------------ parameter layout
static void layout(const sf_eagle3_config& c, int64_t* off, int64_t* sz, int64_t* total) {
    const Dims x = dims_of(c);
    int64_t s[SF_P_COUNT] = {0};
    s[SF_P_FC] = (int64_t)x.H * 3 * x.Ht;
    s[SF_P_Q] = x.A * 2 * x.H; s[SF_P_K] = x.KV * 2 * x.H; s[SF_P_V] = x.KV * 2 * x.H;
    s[SF_P_O] = (int64_t)x.H * x.A;
    s[SF_P_GATE] = (int64_t)x.I * x.H; s[SF_P_UP] = (int64_t)x.I * x.H; s[SF_P_DOWN] = (int64_t)x.H * x.I;
    s[SF_P_HIDDEN_NORM] = x.H; s[SF_P_INPUT_NORM] = x.H; s[SF_P_POST_NORM] = x.H; s[SF_P_NORM] = x.H;
    s[SF_P_LM_HEAD] = (int64_t)x.DV * x.H;
    if (c.fc_norm) s[SF_P_FC_NORM0] = s[SF_P_FC_NORM1] = s[SF_P_FC_NORM2] = x.Ht;
    int64_t o = 0;
    for (int i = 0; i < SF_P_COUNT; ++i) { off[i] = o; sz[i] = s[i]; o += s[i]; }  // all sizes are multiples of 8
    *total = o;
}

// ------------------------------------------------------------------ workspace plan
struct Plan {
    // persistent forward -> backward
    int64_t h, xcat, qkv, attn, lse, hmid, hn2, gu, act, hf, logits, hs3n;
    int64_t xg, tstats, ids, pos_mask, loss_mask32, key_mask, kvlen, d2t_idx, row_ws, sd_ws, metrics, misc;
    // union region: forward temporaries / backward buffers
    int64_t u_base;
    int64_t tgt_shift, tlogits, tpart, lstats, t2d_bits, t2d_prefix;   // forward temporaries
    int64_t dh_tot, dgu, dhmid, dqkv, d_hf, d_act, d_hn2, d_attn, d_xcat, dh_carry, dk_acc, dv_acc, dq_diag, delta, d_hs3n, norm_ws;
    int64_t total;
};
static Plan make_plan(const sf_eagle3_config& c) {
    const Dims x = dims_of(c);
    Plan p;
    int64_t o = 0;
    auto take = [&](int64_t bytes) { int64_t r = o; o = align_up(o + bytes, 1024); return r; };
    const int64_t T = x.T, M = x.M;
    p.h = take((T + 1) * M * x.H * 2);
    p.xcat = take(T * M * 2 * x.H * 2);
    p.qkv = take(T * M * x.QKV * 2);
    p.attn = take(T * M * x.A * 2);
    p.lse = take(T * (int64_t)x.B * x.nh * x.S * 4);
    p.hmid = take(T * M * x.H * 2);
    p.hn2 = take(T * M * x.H * 2);
    p.gu = take(T * M * 2 * x.I * 2);
    p.act = take(T * M * x.I * 2);
    p.hf = c.norm_output ? take(T * M * x.H * 2) : p.h + M * x.H * 2;
    p.logits = take(T * M * x.DV * 2);
    p.hs3n = c.fc_norm ? take(M * 3 * x.Ht * 2) : 0;
    p.xg = take((int64_t)x.B * (x.S + T) * x.DV * 2);       // gathered bf16 teacher logits, padded rows
    p.tstats = take((int64_t)x.B * (x.S + T) * 16);         // {md, 1/dd, p_on_draft scale, 0} per padded row
    p.ids = take((int64_t)x.B * (x.S + T) * 8);
    p.pos_mask = take(M * 4);
    p.loss_mask32 = take(M * 4);
    p.key_mask = take(M);
    p.kvlen = take(2 * (int64_t)x.B * 4 + 64);
    p.d2t_idx = take((int64_t)x.DV * 4);
    p.row_ws = take(3 * M * 4);
    p.sd_ws = take((int64_t)x.B * x.nh * x.S * (T > 1 ? T - 1 : 1) * 4);
    p.metrics = take(T * 8 * 4);
    p.misc = take(4096);
    p.u_base = o;
    // forward temporaries
    p.tgt_shift = take(M * x.Ht * 2);
    p.tlogits = take(M * (int64_t)x.V * 2);
    p.tpart = take(5 * 2 * (int64_t)((x.V + 255) / 256) * M * 4);      // EPI_TEACHER partials [5][<= 2 per n-block][M]
    p.lstats = take(3 * 2 * (int64_t)((x.DV + 255) / 256) * M * 4);    // EPI_BF16_STATS partials of one step's lm_head GEMM
    p.t2d_bits = take((int64_t)((x.V + 31) / 32) * 4);
    p.t2d_prefix = take((int64_t)((x.V + 31) / 32) * 4);
    const int64_t fwd_end = o;
    // backward buffers alias the forward temporaries
    o = p.u_base;
    p.dh_tot = take(T * M * x.H * 2);
    p.dgu = take(T * M * 2 * x.I * 2);
    p.dhmid = take(T * M * x.H * 2);
    p.dqkv = take(T * M * x.QKV * 2);
    p.d_hf = take(M * x.H * 2);
    p.d_act = take(M * x.I * 2);
    p.d_hn2 = take(M * x.H * 2);
    p.d_attn = take(M * x.A * 2);
    p.d_xcat = take(M * 2 * x.H * 2);
    p.dh_carry = take(M * x.H * 2);
    p.dk_acc = take(T * M * x.KV * 4);
    p.dv_acc = take(T * M * x.KV * 4);
    p.dq_diag = take(M * x.A * 4);
    p.delta = take((int64_t)x.B * x.nh * x.S * 4);
    p.norm_ws = take(rmsnorm_bwd_ws_bytes(x.H > x.Ht ? x.H : x.Ht));
    p.d_hs3n = c.fc_norm ? take(M * 3 * x.Ht * 2) : 0;
    p.total = (o > fwd_end ? o : fwd_end) + 1024;
    return p;
}

// ------------------------------------------------------------------ small prep kernels
__global__ void prep_masks_kernel(const int64_t* __restrict__ attention_mask, const int64_t* __restrict__ loss_mask,
                                  uint8_t* __restrict__ key_mask, int* __restrict__ loss_mask32, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        key_mask[i] = attention_mask ? (attention_mask[i] != 0) : 1;
        loss_mask32[i] = (int)loss_mask[i];
    }
}
__global__ void d2t_idx_kernel(const int64_t* __restrict__ d2t, int* __restrict__ idx, int DV) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < DV; i += gridDim.x * blockDim.x) idx[i] = i + (int)d2t[i];
}
__global__ void total_loss_kernel(const float* __restrict__ metrics, int T, float decay, float* __restrict__ loss,
                                  float* __restrict__ metrics_out) {
    if (threadIdx.x == 0) {
        float acc = 0.f, w = 1.f;
        for (int j = 0; j < T; ++j) { acc += w * metrics[j * 8]; w *= decay; }
        *loss = acc;
    }
    if (metrics_out)
        for (int i = threadIdx.x; i < T * 8; i += blockDim.x) metrics_out[i] = metrics[i];
}

// Side stream for the work that is OFF the draft's critical path: the teacher distribution (needed only by the first
// loss) and every step's loss/metrics kernel (needed only by backward).  Their blocks use no dynamic smem, so they
// co-reside with the persistent tcgen05 GEMM CTAs of the next ops and soak up the HBM bandwidth the GEMMs leave idle.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork[12] = {nullptr};
    cudaEvent_t join = nullptr;
    bool init() {
        if (stream) return true;
        if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return false;
        for (auto& e : fork) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return false;
        return cudaEventCreateWithFlags(&join, cudaEventDisableTiming) == cudaSuccess;
    }
};
static SideStream g_side[16];
static bool loss_on_side() { return opt(OPT_LOSS_SIDE) == 1; }
static bool overlap_enabled() { return opt(OPT_NO_OVERLAP) != 1; }

struct Ctx {
    const sf_eagle3_config* cfg; Dims x; Plan p; uint8_t* ws; cudaStream_t st;
    const __nv_bfloat16* W[SF_P_COUNT];
    template <typename T_> T_* at(int64_t off) const { return reinterpret_cast<T_*>(ws + off); }
    __nv_bfloat16* bf(int64_t off, int64_t elem_off = 0) const { return reinterpret_cast<__nv_bfloat16*>(ws + off) + elem_off; }
};

static bool fuse_swiglu(const Dims& x) { return opt(OPT_NO_SWIGLU_FUSION) != 1 && x.M > 128 && x.I % 128 == 0; }

static thread_local int g_overlap_prev = 0;   // consumed by the next mm(): see GemmDesc::overlap_prev
static int mm(const Ctx& c, const void* A, int64_t lda, int am, const void* B, int64_t ldb, int bm, void* D, int64_t ldd,
              const void* R, int64_t ldr, int64_t M, int64_t N, int64_t K, int epi, void* D2 = nullptr, int64_t ldd2 = 0,
              int n_half = 0) {
    GemmDesc g;
    g.D2 = D2; g.ldd2 = ldd2; g.n_half = n_half;
    g.A = A; g.lda = lda; g.a_major = am; g.B = B; g.ldb = ldb; g.b_major = bm; g.D = D; g.ldd = ldd; g.R = R; g.ldr = ldr;
    g.M = (int)M; g.N = (int)N; g.K = (int)K; g.epi = epi; g.cta_group = 0;
    g.overlap_prev = g_overlap_prev; g_overlap_prev = 0;
    return gemm(g, c.st);
}
#define SF_TRY(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

static int setup(Ctx& c, const sf_eagle3_config* cfg, const void* params_flat, void* ws, size_t ws_bytes, void* stream) {
    if (!cfg || !params_flat || !ws) return set_error(-22, "null argument");
    SF_TRY(validate(*cfg));
    c.cfg = cfg; c.x = dims_of(*cfg); c.p = make_plan(*cfg); c.ws = reinterpret_cast<uint8_t*>(ws);
    c.st = reinterpret_cast<cudaStream_t>(stream);
    if ((size_t)c.p.total > ws_bytes) return set_error(-12, "workspace too small: need %lld bytes, got %zu", (long long)c.p.total, ws_bytes);
    if (reinterpret_cast<uintptr_t>(ws) & 1023) return set_error(-22, "workspace must be 1024-byte aligned");
    int64_t off[SF_P_COUNT], sz[SF_P_COUNT], total;
    layout(*cfg, off, sz, &total);
    for (int i = 0; i < SF_P_COUNT; ++i) c.W[i] = reinterpret_cast<const __nv_bfloat16*>(params_flat) + off[i];
    return 0;
}

// ------------------------------------------------------------------ forward
static int forward(Ctx& c, const sf_eagle3_frozen& fz, const sf_eagle3_batch& bt, float* metrics_out, float* loss_out,
                   int need_grad) {
    const Dims& x = c.x; const Plan& p = c.p; const sf_eagle3_config& cfg = *c.cfg;
    const int64_t M = x.M; const int T = x.T;
    cudaStream_t st = c.st;
    // masks, gather index
    prep_masks_kernel<<<148, 256, 0, st>>>(bt.attention_mask, bt.loss_mask, c.at<uint8_t>(p.key_mask), c.at<int>(p.loss_mask32), M);
    SF_CUDA_CHECK_LAUNCH("prep_masks");
    d2t_idx_kernel<<<64, 256, 0, st>>>(fz.d2t, c.at<int>(p.d2t_idx), x.DV);
    SF_CUDA_CHECK_LAUNCH("d2t_idx");
    // ---- teacher: shift, frozen target head GEMM, distribution  (strategies/base.py:116-120, eagle3/model.py:445-501)
    SF_TRY(shift_left(bt.target, c.bf(p.tgt_shift), x.B, x.S, x.Ht, st));
    int dev = 0;
    cudaGetDevice(&dev);
    SideStream* side = (overlap_enabled() && dev < 16 && g_side[dev].init()) ? &g_side[dev] : nullptr;
    cudaStream_t ls = st;   // stream of the loss-side work
    if (side) ls = side->stream;
    const bool fuse_teacher = opt(OPT_NO_TEACHER_FUSION) != 1;
    if (fuse_teacher) {
        // The teacher's row statistics (argmax / logsumexp over the full vocabulary, softmax statistics and gather of the draft
        // vocabulary) are computed by the target-head GEMM's own epilogue from the accumulators in TMEM: the [M, V] logits are
        // never written (eagle3/model.py:487-501; core/compact_teacher.py:57-150 is the same streaming reduction).
        SF_TRY(t2d_index(fz.t2d, x.V, c.at<uint32_t>(p.t2d_bits), c.at<int>(p.t2d_prefix), st));
        GemmDesc g;
        g.A = c.bf(p.tgt_shift); g.lda = x.Ht; g.a_major = MAJOR_K; g.B = fz.target_head; g.ldb = x.Ht; g.b_major = MAJOR_K;
        g.D = nullptr; g.ldd = 0; g.R = nullptr; g.ldr = 0; g.M = (int)M; g.N = x.V; g.K = x.Ht; g.epi = EPI_TEACHER; g.cta_group = 0;
        g.stats = c.at<float>(p.tpart); g.t2d_bits = c.at<uint32_t>(p.t2d_bits); g.t2d_prefix = c.at<int>(p.t2d_prefix);
        g.xg = c.bf(p.xg); g.S = x.S; g.T = T; g.DV = x.DV;
        SF_TRY(gemm(g, st));
        SF_TRY(teacher_merge(c.at<float>(p.tpart), gemm_stats_blocks(g), M, fz.t2d, c.at<int>(p.loss_mask32), c.at<float>(p.tstats),
                             c.at<int64_t>(p.ids), c.at<int>(p.pos_mask), c.bf(p.xg), x.B, x.S, T, x.DV, st));
        side = nullptr;   // nothing left to overlap: the side stream is only used by the unfused path below
        ls = st;
    }
    // Unfused path (sf_debug_option("no_teacher_fusion", 1), kept for A/B runs): the target-head GEMM writes the logits and is
    // issued in two row halves; the teacher statistics of the first half run on the side stream under the second half's GEMM,
    // those of the second half under the draft's fc / QKV GEMMs.
    const int64_t Mh = (side && M >= 1024) ? (M / 2) / 256 * 256 : M;
    for (int half = 0; !fuse_teacher && half < (Mh < M ? 2 : 1); ++half) {
        const int64_t r0 = half ? Mh : 0, nr = half ? M - Mh : Mh;
        SF_TRY(mm(c, c.bf(p.tgt_shift, r0 * x.Ht), x.Ht, MAJOR_K, fz.target_head, x.Ht, MAJOR_K, c.bf(p.tlogits, r0 * x.V), x.V,
                  nullptr, 0, nr, x.V, x.Ht, EPI_BF16));
        if (side) {
            if (cudaEventRecord(side->fork[10 + half], st) != cudaSuccess || cudaStreamWaitEvent(ls, side->fork[10 + half], 0) != cudaSuccess)
                return set_error(-5, "side-stream fork failed");
        }
        SF_TRY(teacher(c.bf(p.tlogits), x.V, c.at<int>(p.d2t_idx), fz.t2d, c.at<int>(p.loss_mask32), c.bf(p.xg), c.at<float>(p.tstats),
                       c.at<int64_t>(p.ids), c.at<int>(p.pos_mask), x.B, x.S, T, x.V, x.DV, r0, nr, half == 0, side ? 1 : 0, ls));
    }
    // ---- fc: h_0 = [fc_norm_i(chunk_i)] W_fc^T   (llama3_eagle.py:1762-1770; per-third RMSNorm = EAGLE3.1)
    const void* fc_in = bt.hidden_state;
    if (cfg.fc_norm) {
        const __nv_bfloat16* hs = reinterpret_cast<const __nv_bfloat16*>(bt.hidden_state);
        for (int i = 0; i < 3; ++i)
            SF_TRY(rmsnorm_fwd(hs + i * x.Ht, 3 * x.Ht, nullptr, x.S, 0, c.W[SF_P_FC_NORM0 + i], c.bf(p.hs3n) + i * x.Ht, 3 * x.Ht, M, x.Ht, cfg.rms_eps, nullptr, st));
        fc_in = c.bf(p.hs3n);
    }
    SF_TRY(mm(c, fc_in, 3 * x.Ht, MAJOR_K, c.W[SF_P_FC], 3 * x.Ht, MAJOR_K, c.bf(p.h), x.H, nullptr, 0, M, x.H, 3 * x.Ht, EPI_BF16));
    bool side_dirty = side != nullptr;   // the teacher kernels are in flight on the side stream
    const uint8_t* key_mask = bt.attention_mask ? c.at<uint8_t>(p.key_mask) : nullptr;
    if (key_mask) SF_TRY(mask_prefix(key_mask, x.B, x.S, c.at<int>(p.kvlen), c.at<int>(p.kvlen) + x.B, st));
    for (int j = 0; j < T; ++j) {
        __nv_bfloat16* h_in = c.bf(p.h, (int64_t)j * M * x.H);
        __nv_bfloat16* h_out = c.bf(p.h, (int64_t)(j + 1) * M * x.H);
        __nv_bfloat16* xcat = c.bf(p.xcat, (int64_t)j * M * 2 * x.H);
        __nv_bfloat16* qkv = c.bf(p.qkv, (int64_t)j * M * x.QKV);
        __nv_bfloat16* attn = c.bf(p.attn, (int64_t)j * M * x.A);
        __nv_bfloat16* hmid = c.bf(p.hmid, (int64_t)j * M * x.H);
        __nv_bfloat16* hn2 = c.bf(p.hn2, (int64_t)j * M * x.H);
        __nv_bfloat16* gu = c.bf(p.gu, (int64_t)j * M * 2 * x.I);
        __nv_bfloat16* act = c.bf(p.act, (int64_t)j * M * x.I);
        __nv_bfloat16* hf = cfg.norm_output ? c.bf(p.hf, (int64_t)j * M * x.H) : h_out;
        __nv_bfloat16* logits = c.bf(p.logits, (int64_t)j * M * x.DV);
        float* lse = c.at<float>(p.lse) + (int64_t)j * x.B * x.nh * x.S;
        // x = cat(RMSNorm_in(embed(ids_j)), RMSNorm_h(h_j))   (llama3_eagle.py:1625-1630; ids shifted 1+j, model.py:428-432)
        SF_TRY(rmsnorm_fwd(fz.embed_tokens, x.H, bt.input_ids, x.S, 1 + j, c.W[SF_P_INPUT_NORM], xcat, 2 * x.H, M, x.H, cfg.rms_eps, nullptr, st));
        SF_TRY(rmsnorm_fwd(h_in, x.H, nullptr, x.S, 0, c.W[SF_P_HIDDEN_NORM], xcat + x.H, 2 * x.H, M, x.H, cfg.rms_eps, nullptr, st));
        // fused q/k/v projection, RoPE at position s + j   (llama3_eagle.py:673-675,730-734)
        if (opt(OPT_NO_ROPE_FUSION) != 1) {   // RoPE of the q and k heads inside the projection's epilogue (v columns pass through)
            GemmDesc g;
            g.A = xcat; g.lda = 2 * x.H; g.a_major = MAJOR_K; g.B = c.W[SF_P_Q]; g.ldb = 2 * x.H; g.b_major = MAJOR_K;
            g.D = qkv; g.ldd = x.QKV; g.R = nullptr; g.ldr = 0; g.M = (int)M; g.N = (int)x.QKV; g.K = 2 * x.H; g.epi = EPI_BF16_ROPE; g.cta_group = 0;
            g.rope_cos = fz.rope_cos; g.rope_sin = fz.rope_sin; g.S = x.S; g.rope_pos0 = j; g.head_dim = x.d; g.rope_cols = (int)(x.A + x.KV);
            SF_TRY(gemm(g, st));
        } else {
            SF_TRY(mm(c, xcat, 2 * x.H, MAJOR_K, c.W[SF_P_Q], 2 * x.H, MAJOR_K, qkv, x.QKV, nullptr, 0, M, x.QKV, 2 * x.H, EPI_BF16));
            SF_TRY(rope(qkv, nullptr, x.QKV, 0, x.nh + x.nkv, x.d, fz.rope_cos, fz.rope_sin, x.S, j, M, 0, st));
        }
        // TTT attention   (llama3_eagle.py:739-785)
        AttnDesc a{};
        a.q = qkv; a.ldq = x.QKV; a.ldkv = x.QKV; a.out = attn; a.ldo = x.A; a.lse = lse; a.sd_ws = c.at<float>(p.sd_ws);
        a.key_mask = key_mask; a.B = x.B; a.S = x.S; a.nh = x.nh; a.nkv = x.nkv; a.head_dim = x.d; a.J = j;
        a.kvlen = key_mask ? c.at<int>(p.kvlen) : nullptr; a.nonprefix = key_mask ? c.at<int>(p.kvlen) + x.B : nullptr;
        a.q_row_base = qkv; a.kv_row_base = c.bf(p.qkv); a.q_col0 = 0; a.k_col0 = (int)x.A; a.v_col0 = (int)(x.A + x.KV);
        for (int i = 0; i <= j; ++i) {
            const __nv_bfloat16* qi = c.bf(p.qkv, (int64_t)i * M * x.QKV);
            a.k[i] = qi + x.A; a.v[i] = qi + x.A + x.KV;
        }
        SF_TRY(attn_fwd(a, st));
        // h_mid = h_j + attn W_o^T ; MLP ; h_{j+1} = h_mid + down(silu(gate) * up)   (llama3_eagle.py:1641-1648)
        SF_TRY(mm(c, attn, x.A, MAJOR_K, c.W[SF_P_O], x.A, MAJOR_K, hmid, x.H, h_in, x.H, M, x.H, x.A, EPI_BF16_RESID));
        SF_TRY(rmsnorm_fwd(hmid, x.H, nullptr, x.S, 0, c.W[SF_P_POST_NORM], hn2, x.H, M, x.H, cfg.rms_eps, nullptr, st));
        if (fuse_swiglu(x)) {   // gate/up GEMM with the SwiGLU fused into its epilogue
            SF_TRY(mm(c, hn2, x.H, MAJOR_K, c.W[SF_P_GATE], x.H, MAJOR_K, gu, 2 * x.I, nullptr, 0, M, 2 * x.I, x.H, EPI_SWIGLU, act, x.I, x.I));
        } else {
            SF_TRY(mm(c, hn2, x.H, MAJOR_K, c.W[SF_P_GATE], x.H, MAJOR_K, gu, 2 * x.I, nullptr, 0, M, 2 * x.I, x.H, EPI_BF16));
            SF_TRY(swiglu_fwd(gu, act, M, x.I, st));
        }
        SF_TRY(mm(c, act, x.I, MAJOR_K, c.W[SF_P_DOWN], x.I, MAJOR_K, h_out, x.H, hmid, x.H, M, x.H, x.I, EPI_BF16_RESID));
        // logits = lm_head(norm(h_{j+1}))   (llama3_eagle.py:1772-1777)
        if (cfg.norm_output)
            SF_TRY(rmsnorm_fwd(h_out, x.H, nullptr, x.S, 0, c.W[SF_P_NORM], hf, x.H, M, x.H, cfg.rms_eps, nullptr, st));
        const bool fuse_stats = opt(OPT_NO_LOSS_STATS_FUSION) != 1 && !(side && loss_on_side());   // one partials buffer: the loss must run in stream order
        int stats_nb = 0;
        if (fuse_stats) {   // logits + per-tile (max, sum-exp, argmax) partials: the loss kernel's first pass over the row disappears
            GemmDesc g;
            g.A = hf; g.lda = x.H; g.a_major = MAJOR_K; g.B = c.W[SF_P_LM_HEAD]; g.ldb = x.H; g.b_major = MAJOR_K;
            g.D = logits; g.ldd = x.DV; g.R = nullptr; g.ldr = 0; g.M = (int)M; g.N = x.DV; g.K = x.H; g.epi = EPI_BF16_STATS; g.cta_group = 0;
            g.stats = c.at<float>(p.lstats);
            stats_nb = gemm_stats_blocks(g);
            SF_TRY(gemm(g, st));
        } else {
            SF_TRY(mm(c, hf, x.H, MAJOR_K, c.W[SF_P_LM_HEAD], x.H, MAJOR_K, logits, x.DV, nullptr, 0, M, x.DV, x.H, EPI_BF16));
        }
        // loss / metrics / d(logits) in place   (eagle3/model.py:142-199, core/loss.py, core/lk_loss.py)
        const float step_weight = powf(cfg.ploss_decay, (float)j);
        // The loss runs on the main stream: it is issue- and power-hungry enough that co-running it with a GEMM slows
        // the GEMM by as much as it saves (measured; SF_LOSS_SIDE=1 re-enables the side-stream placement for A/B runs).
        // The teacher statistics it consumes come from the side stream, hence the join before the first use.
        const bool on_side = side && loss_on_side() && j + 1 < T;
        if (on_side) {
            if (cudaEventRecord(side->fork[j], st) != cudaSuccess || cudaStreamWaitEvent(ls, side->fork[j], 0) != cudaSuccess)
                return set_error(-5, "side-stream fork failed");
            side_dirty = true;
        } else if (side && side_dirty) {
            if (cudaEventRecord(side->join, ls) != cudaSuccess || cudaStreamWaitEvent(st, side->join, 0) != cudaSuccess)
                return set_error(-5, "side-stream join failed");
            side_dirty = false;
        }
        SF_TRY(loss_step(logits, x.DV, c.bf(p.xg), c.at<float>(p.tstats), c.at<int64_t>(p.ids), c.at<int>(p.pos_mask),
                         c.at<int>(p.loss_mask32), fz.d2t, x.B, x.S, T, x.DV, j, step_weight, need_grad, cfg.lk_loss_type,
                         cfg.kl_scale, cfg.kl_decay, c.at<float>(p.row_ws), c.at<float>(p.metrics),
                         fuse_stats ? c.at<float>(p.lstats) : nullptr, stats_nb, on_side ? ls : st));
    }
    total_loss_kernel<<<1, 64, 0, st>>>(c.at<float>(p.metrics), T, cfg.ploss_decay, loss_out ? loss_out : c.at<float>(p.misc), metrics_out);
    SF_CUDA_CHECK_LAUNCH("total_loss");
    return 0;
}

// ------------------------------------------------------------------ backward
__global__ void scale_f32_kernel(float* __restrict__ g, int64_t n, float s) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) g[i] *= s;
}

static int backward(Ctx& c, const sf_eagle3_frozen& fz, const sf_eagle3_batch& bt, float loss_scale, float* G, int accumulate,
                    sf_grad_ready_fn on_ready, void* user) {
    const Dims& x = c.x; const Plan& p = c.p; const sf_eagle3_config& cfg = *c.cfg;
    const int64_t M = x.M; const int T = x.T;
    cudaStream_t st = c.st;
    int64_t off[SF_P_COUNT], sz[SF_P_COUNT], total;
    layout(cfg, off, sz, &total);
    // A fresh step zeroes only the norm-weight slots (their gradients are summed over the TTT steps by the row kernels);
    // every matrix gradient is produced by exactly one GEMM, which then simply overwrites its slice (EPI_F32) instead of
    // read-modify-writing 1.6 GB of zeros.  With `accumulate` the GEMMs add into what is there (EPI_F32_ACCUM).
    if (!accumulate) {
        const size_t n0 = (size_t)(off[SF_P_LM_HEAD] - off[SF_P_HIDDEN_NORM]) * 4;   // hidden_norm, input_norm, post_norm, norm
        if (cudaMemsetAsync(G + off[SF_P_HIDDEN_NORM], 0, n0, st) != cudaSuccess) return set_error(-5, "memset grads failed");
        if (total > off[SF_P_FC_NORM0] &&
            cudaMemsetAsync(G + off[SF_P_FC_NORM0], 0, (size_t)(total - off[SF_P_FC_NORM0]) * 4, st) != cudaSuccess)
            return set_error(-5, "memset grads failed");
    }
    const int wepi = accumulate ? EPI_F32_ACCUM : EPI_F32;
    // loss_scale is applied by scaling d(logits)-derived quantities linearly: we fold it into the final gradients
    // only when != 1 (the common 1/accumulation_steps case is passed to forward via grad_coef by the caller).
    if (cudaMemsetAsync(c.ws + p.dk_acc, 0, (size_t)T * M * x.KV * 4, st) != cudaSuccess ||
        cudaMemsetAsync(c.ws + p.dv_acc, 0, (size_t)T * M * x.KV * 4, st) != cudaSuccess)
        return set_error(-5, "memset dk/dv failed");
    const uint8_t* key_mask = bt.attention_mask ? c.at<uint8_t>(p.key_mask) : nullptr;
    float* Gn = G;  // fp32 flat grads
    // If a loss_scale != 1 is requested together with accumulation we must not rescale what is already there:
    // compute this micro-batch into a clean region is avoided by scaling d(logits) instead (done in forward).
    (void)loss_scale;

    for (int j = T - 1; j >= 0; --j) {
        const __nv_bfloat16* h_in = c.bf(p.h, (int64_t)j * M * x.H);
        const __nv_bfloat16* h_out = c.bf(p.h, (int64_t)(j + 1) * M * x.H);
        const __nv_bfloat16* qkv = c.bf(p.qkv, (int64_t)j * M * x.QKV);
        const __nv_bfloat16* attn = c.bf(p.attn, (int64_t)j * M * x.A);
        const __nv_bfloat16* hmid = c.bf(p.hmid, (int64_t)j * M * x.H);
        const __nv_bfloat16* gu = c.bf(p.gu, (int64_t)j * M * 2 * x.I);
        const __nv_bfloat16* dlogits = c.bf(p.logits, (int64_t)j * M * x.DV);
        const float* lse = c.at<float>(p.lse) + (int64_t)j * x.B * x.nh * x.S;
        __nv_bfloat16* dh_tot = c.bf(p.dh_tot, (int64_t)j * M * x.H);
        __nv_bfloat16* dgu = c.bf(p.dgu, (int64_t)j * M * 2 * x.I);
        __nv_bfloat16* dhmid = c.bf(p.dhmid, (int64_t)j * M * x.H);
        __nv_bfloat16* dqkv = c.bf(p.dqkv, (int64_t)j * M * x.QKV);
        __nv_bfloat16* d_hf = c.bf(p.d_hf); __nv_bfloat16* d_act = c.bf(p.d_act); __nv_bfloat16* d_hn2 = c.bf(p.d_hn2);
        __nv_bfloat16* d_attn = c.bf(p.d_attn); __nv_bfloat16* d_xcat = c.bf(p.d_xcat); __nv_bfloat16* dh_carry = c.bf(p.dh_carry);
        const __nv_bfloat16* carry_in = (j == T - 1) ? nullptr : dh_carry;

        // d(norm(h_{j+1})) = dlogits W_lm ; through the final norm ; + gradient arriving from step j+1
        SF_TRY(mm(c, dlogits, x.DV, MAJOR_K, c.W[SF_P_LM_HEAD], x.H, MAJOR_MN, d_hf, x.H, nullptr, 0, M, x.H, x.DV, EPI_BF16));
        if (cfg.norm_output) {
            SF_TRY(rmsnorm_bwd(h_out, x.H, nullptr, x.S, 0, c.W[SF_P_NORM], d_hf, x.H, carry_in, nullptr, dh_tot, Gn + off[SF_P_NORM], c.at<float>(p.norm_ws), M, x.H, cfg.rms_eps, st));
        } else {
            SF_TRY(add_bf16(d_hf, carry_in, dh_tot, M * x.H, st));   // lm_head reads h_{j+1} directly
        }
        // MLP
        if (fuse_swiglu(x)) {   // d(act) never touches HBM: the dgrad epilogue applies the SwiGLU backward
            SF_TRY(mm(c, dh_tot, x.H, MAJOR_K, c.W[SF_P_DOWN], x.I, MAJOR_MN, dgu, 2 * x.I, gu, 2 * x.I, M, x.I, x.H, EPI_SWIGLU_BWD, nullptr, 0, x.I));
        } else {
            SF_TRY(mm(c, dh_tot, x.H, MAJOR_K, c.W[SF_P_DOWN], x.I, MAJOR_MN, d_act, x.I, nullptr, 0, M, x.I, x.H, EPI_BF16));
            SF_TRY(swiglu_bwd(gu, d_act, dgu, M, x.I, st));
        }
        SF_TRY(mm(c, dgu, 2 * x.I, MAJOR_K, c.W[SF_P_GATE], x.H, MAJOR_MN, d_hn2, x.H, nullptr, 0, M, x.H, 2 * x.I, EPI_BF16));
        SF_TRY(rmsnorm_bwd(hmid, x.H, nullptr, x.S, 0, c.W[SF_P_POST_NORM], d_hn2, x.H, dh_tot, nullptr, dhmid, Gn + off[SF_P_POST_NORM], c.at<float>(p.norm_ws), M, x.H, cfg.rms_eps, st));
        // attention
        SF_TRY(mm(c, dhmid, x.H, MAJOR_K, c.W[SF_P_O], x.A, MAJOR_MN, d_attn, x.A, nullptr, 0, M, x.A, x.H, EPI_BF16));
        AttnDesc a{};
        a.q = qkv; a.ldq = x.QKV; a.ldkv = x.QKV; a.out = const_cast<__nv_bfloat16*>(attn); a.ldo = x.A;
        a.lse = const_cast<float*>(lse); a.sd_ws = c.at<float>(p.sd_ws);
        a.key_mask = key_mask; a.B = x.B; a.S = x.S; a.nh = x.nh; a.nkv = x.nkv; a.head_dim = x.d; a.J = j;
        a.kvlen = key_mask ? c.at<int>(p.kvlen) : nullptr; a.nonprefix = key_mask ? c.at<int>(p.kvlen) + x.B : nullptr;
        a.q_row_base = qkv; a.kv_row_base = c.bf(p.qkv); a.q_col0 = 0; a.k_col0 = (int)x.A; a.v_col0 = (int)(x.A + x.KV);
        a.dout = d_attn; a.lddo = x.A; a.delta_ws = c.at<float>(p.delta); a.ldacc = x.KV;
        a.dq = dqkv; a.lddq = x.QKV; a.dq_diag_ws = c.at<float>(p.dq_diag);
        for (int i = 0; i <= j; ++i) {
            const __nv_bfloat16* qi = c.bf(p.qkv, (int64_t)i * M * x.QKV);
            a.k[i] = qi + x.A; a.v[i] = qi + x.A + x.KV;
            a.dk_acc[i] = c.at<float>(p.dk_acc) + (int64_t)i * M * x.KV;
            a.dv_acc[i] = c.at<float>(p.dv_acc) + (int64_t)i * M * x.KV;
        }
        SF_TRY(attn_bwd(a, st));
        // block j's K/V gradients are final now (only steps >= j touch them): inverse RoPE, pack next to dq
        SF_TRY(rope(dqkv, nullptr, x.QKV, 0, x.nh, x.d, fz.rope_cos, fz.rope_sin, x.S, j, M, 1, st));
        SF_TRY(rope(dqkv + x.A, a.dk_acc[j], x.QKV, x.KV, x.nkv, x.d, fz.rope_cos, fz.rope_sin, x.S, j, M, 1, st));
        SF_TRY(cvt_f32_bf16(a.dv_acc[j], x.KV, dqkv + x.A + x.KV, x.QKV, M, (int)x.KV, 1.0f, st));
        // through the fused q/k/v projection and the two input norms
        SF_TRY(mm(c, dqkv, x.QKV, MAJOR_K, c.W[SF_P_Q], 2 * x.H, MAJOR_MN, d_xcat, 2 * x.H, nullptr, 0, M, 2 * x.H, x.QKV, EPI_BF16));
        SF_TRY(rmsnorm_bwd(fz.embed_tokens, x.H, bt.input_ids, x.S, 1 + j, c.W[SF_P_INPUT_NORM], d_xcat, 2 * x.H, nullptr, nullptr, nullptr, Gn + off[SF_P_INPUT_NORM], c.at<float>(p.norm_ws), M, x.H, cfg.rms_eps, st));
        SF_TRY(rmsnorm_bwd(h_in, x.H, nullptr, x.S, 0, c.W[SF_P_HIDDEN_NORM], d_xcat + x.H, 2 * x.H, dhmid, nullptr, dh_carry, Gn + off[SF_P_HIDDEN_NORM], c.at<float>(p.norm_ws), M, x.H, cfg.rms_eps, st));
    }
    // ---- weight gradients: one GEMM per weight, contracting over all T*M tokens (fp32 accumulation in TMEM).
    // `on_ready` fires after each group of adjacent parameters is complete so the caller can overlap its all-reduce.
    auto ready = [&](int first, int n) { if (on_ready) on_ready(first, n, user); };
    const int64_t TM = (int64_t)T * M;
    ready(SF_P_HIDDEN_NORM, 4);     // the four norm-weight gradients were finished inside the TTT loop
    SF_TRY(mm(c, c.bf(p.logits), x.DV, MAJOR_MN, c.bf(p.hf), x.H, MAJOR_MN, Gn + off[SF_P_LM_HEAD], x.H, nullptr, 0, x.DV, x.H, TM, wepi));
    ready(SF_P_LM_HEAD, 1);
    // The weight-gradient GEMMs are mutually independent (disjoint outputs, inputs final): without a caller hook between
    // them each one is launched as a programmatic dependent of the previous, so its CTAs fill the SMs the previous
    // GEMM's last partial wave leaves idle (2000/768/1536/256/768 tiles over 74 clusters: 4 % of the wgrad time).
    const int chain = on_ready ? 0 : 1;
    g_overlap_prev = chain;
    SF_TRY(mm(c, c.bf(p.dh_tot), x.H, MAJOR_MN, c.bf(p.act), x.I, MAJOR_MN, Gn + off[SF_P_DOWN], x.I, nullptr, 0, x.H, x.I, TM, wepi));
    ready(SF_P_DOWN, 1);
    g_overlap_prev = chain;
    SF_TRY(mm(c, c.bf(p.dgu), 2 * x.I, MAJOR_MN, c.bf(p.hn2), x.H, MAJOR_MN, Gn + off[SF_P_GATE], x.H, nullptr, 0, 2 * x.I, x.H, TM, wepi));
    ready(SF_P_GATE, 2);
    g_overlap_prev = chain;
    SF_TRY(mm(c, c.bf(p.dhmid), x.H, MAJOR_MN, c.bf(p.attn), x.A, MAJOR_MN, Gn + off[SF_P_O], x.A, nullptr, 0, x.H, x.A, TM, wepi));
    ready(SF_P_O, 1);
    g_overlap_prev = chain;
    SF_TRY(mm(c, c.bf(p.dqkv), x.QKV, MAJOR_MN, c.bf(p.xcat), 2 * x.H, MAJOR_MN, Gn + off[SF_P_Q], 2 * x.H, nullptr, 0, x.QKV, 2 * x.H, TM, wepi));
    ready(SF_P_Q, 3);
    // fc: dW_fc = d(h_0)^T fc_in ; with fc_norm also the three norm-weight gradients through d(fc_in) = d(h_0) W_fc
    const void* fc_in = bt.hidden_state;
    if (cfg.fc_norm) {
        fc_in = c.bf(p.hs3n);
        const __nv_bfloat16* hs = reinterpret_cast<const __nv_bfloat16*>(bt.hidden_state);
        SF_TRY(mm(c, c.bf(p.dh_carry), x.H, MAJOR_K, c.W[SF_P_FC], 3 * x.Ht, MAJOR_MN, c.bf(p.d_hs3n), 3 * x.Ht, nullptr, 0, M, 3 * x.Ht, x.H, EPI_BF16));
        for (int i = 0; i < 3; ++i)
            SF_TRY(rmsnorm_bwd(hs + i * x.Ht, 3 * x.Ht, nullptr, x.S, 0, c.W[SF_P_FC_NORM0 + i], c.bf(p.d_hs3n) + i * x.Ht, 3 * x.Ht, nullptr, nullptr, nullptr, Gn + off[SF_P_FC_NORM0 + i], c.at<float>(p.norm_ws), M, x.Ht, cfg.rms_eps, st));
        ready(SF_P_FC_NORM0, 3);
    }
    if (!cfg.fc_norm) g_overlap_prev = chain;
    SF_TRY(mm(c, c.bf(p.dh_carry), x.H, MAJOR_MN, fc_in, 3 * x.Ht, MAJOR_MN, Gn + off[SF_P_FC], 3 * x.Ht, nullptr, 0, x.H, 3 * x.Ht, M, wepi));
    ready(SF_P_FC, 1);
    return 0;
}

}  // namespace sf

// =============================================================================== C ABI
using namespace sf;

extern "C" int sf_eagle3_param_layout(const sf_eagle3_config* cfg, int64_t offsets[SF_P_COUNT], int64_t sizes[SF_P_COUNT],
                                      int64_t* total_elems) {
    if (!cfg || !offsets || !sizes || !total_elems) return set_error(-22, "null argument");
    layout(*cfg, offsets, sizes, total_elems);
    return 0;
}
extern "C" size_t sf_eagle3_workspace_bytes(const sf_eagle3_config* cfg) {
    if (!cfg || validate(*cfg)) return 0;
    return (size_t)make_plan(*cfg).total;
}
// Where a named tensor of the last step lives inside the caller's workspace (parity tests, forward-only draft-model methods).
extern "C" int sf_eagle3_workspace_view(const sf_eagle3_config* cfg, const char* name, int64_t* offset_bytes, int64_t* size_bytes) {
    if (!cfg || !name || !offset_bytes || !size_bytes) return set_error(-22, "null argument");
    SF_TRY(validate(*cfg));
    const Plan p = make_plan(*cfg);
    const Dims x = dims_of(*cfg);
    const int64_t T = x.T, M = x.M, PR = (int64_t)x.B * (x.S + T);
    struct Ent { const char* n; int64_t off, bytes; } tab[] = {
        {"h", p.h, (T + 1) * M * x.H * 2},            // bf16 [T+1, M, H]   h_0 = fc output, h_{j+1} = step j output
        {"qkv", p.qkv, T * M * x.QKV * 2},            // bf16 [T, M, (nh+2nkv)*d]  after RoPE
        {"attn", p.attn, T * M * x.A * 2},            // bf16 [T, M, nh*d]
        {"hf", p.hf, T * M * x.H * 2},                // bf16 [T, M, H]     norm(h_{j+1}) (aliases h[1:] when !norm_output)
        {"logits", p.logits, T * M * x.DV * 2},       // bf16 [T, M, DV]    logits after forward(need_grad=0); d(loss)/d(logits) after need_grad=1
        {"teacher_xg", p.xg, PR * x.DV * 2},          // bf16 [B, S+T, DV]  gathered teacher logits
        {"teacher_stats", p.tstats, PR * 16},         // f32  [B, S+T, 4]   {md, 1/dd, cs, 0}
        {"teacher_ids", p.ids, PR * 8},               // i64  [B, S+T]      argmax of the full-vocab teacher logits
        {"position_mask", p.pos_mask, M * 4},         // i32  [B, S]
    };
    for (const Ent& e : tab)
        if (!strcmp(e.n, name)) { *offset_bytes = e.off; *size_bytes = e.bytes; return 0; }
    return set_error(-22, "workspace view '%s' unknown", name);
}
extern "C" int sf_eagle3_forward(const sf_eagle3_config* cfg, const void* params_flat, const sf_eagle3_frozen* frozen,
                                 const sf_eagle3_batch* batch, void* workspace, size_t workspace_bytes, float* metrics,
                                 float* loss, int need_grad, void* stream) {
    if (!frozen || !batch) return set_error(-22, "null argument");
    Ctx c;
    SF_TRY(setup(c, cfg, params_flat, workspace, workspace_bytes, stream));
    return forward(c, *frozen, *batch, metrics, loss, need_grad);
}
extern "C" int sf_eagle3_backward_ex(const sf_eagle3_config* cfg, const void* params_flat, const sf_eagle3_frozen* frozen,
                                     const sf_eagle3_batch* batch, void* workspace, size_t workspace_bytes, float loss_scale,
                                     float* grads_flat_f32, int accumulate, sf_grad_ready_fn on_ready, void* user, void* stream) {
    if (!frozen || !batch || !grads_flat_f32) return set_error(-22, "null argument");
    Ctx c;
    SF_TRY(setup(c, cfg, params_flat, workspace, workspace_bytes, stream));
    if (loss_scale != 1.0f && accumulate) return set_error(-22, "loss_scale != 1 with accumulate is unsupported: scale the loss via grads_to_bf16 / optimizer grad_scale");
    if (loss_scale != 1.0f && on_ready) return set_error(-22, "loss_scale != 1 cannot be combined with the gradient-ready callback");
    SF_TRY(backward(c, *frozen, *batch, loss_scale, grads_flat_f32, accumulate, on_ready, user));
    if (loss_scale != 1.0f) {
        int64_t off[SF_P_COUNT], sz[SF_P_COUNT], total;
        layout(*cfg, off, sz, &total);
        scale_f32_kernel<<<148 * 8, 256, 0, c.st>>>(grads_flat_f32, total, loss_scale);
        SF_CUDA_CHECK_LAUNCH("scale_grads");
    }
    return 0;
}
extern "C" int sf_eagle3_backward(const sf_eagle3_config* cfg, const void* params_flat, const sf_eagle3_frozen* frozen,
                                  const sf_eagle3_batch* batch, void* workspace, size_t workspace_bytes, float loss_scale,
                                  float* grads_flat_f32, int accumulate, void* stream) {
    return sf_eagle3_backward_ex(cfg, params_flat, frozen, batch, workspace, workspace_bytes, loss_scale, grads_flat_f32,
                                 accumulate, nullptr, nullptr, stream);
}
extern "C" int sf_grads_to_bf16(const float* grads_f32, void* grads_bf16, int64_t n, const float* scale_dev, void* stream) {
    return cvt_flat_f32_bf16(grads_f32, grads_bf16, n, scale_dev, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int sf_optimizer_step(const void* grads_bf16, float* master, float* exp_avg, float* exp_avg_sq, void* params_bf16,
                                 int64_t n, float grad_scale, float max_grad_norm, float lr, float beta1, float beta2,
                                 float eps, float weight_decay, int32_t step, float* grad_norm_out, float* scratch,
                                 void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (step < 1) return set_error(-22, "optimizer step must be >= 1");
    SF_TRY(grad_norm(grads_bf16, n, grad_scale, scratch, grad_norm_out, st));
    return adamw(grads_bf16, master, exp_avg, exp_avg_sq, params_bf16, n, grad_norm_out, max_grad_norm, grad_scale, lr, beta1,
                 beta2, eps, weight_decay, step, st);
}

// ---- individual ops
extern "C" int sf_rmsnorm_fwd(const void* x, int64_t ldx, const void* w, void* out, int64_t ldo, int64_t M, int H, float eps,
                              void* stream) {
    return rmsnorm_fwd(x, ldx, nullptr, 1, 0, w, out, ldo, M, H, eps, nullptr, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int sf_rmsnorm_bwd(const void* x, int64_t ldx, const void* w, const void* dy, int64_t lddy, const void* add, void* dx,
                              float* dw, float* scratch, int64_t M, int H, float eps, void* stream) {
    return rmsnorm_bwd(x, ldx, nullptr, 1, 0, w, dy, lddy, add, nullptr, dx, dw, scratch, M, H, eps, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int sf_gemm_bf16_rope(const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd, int M, int N, int K,
                                 const void* cos_t, const void* sin_t, int S, int pos_offset, int head_dim, int rope_cols, void* stream) {
    GemmDesc g;
    g.A = A; g.lda = lda; g.a_major = MAJOR_K; g.B = B; g.ldb = ldb; g.b_major = MAJOR_K; g.D = D; g.ldd = ldd; g.R = nullptr; g.ldr = 0;
    g.M = M; g.N = N; g.K = K; g.epi = EPI_BF16_ROPE; g.cta_group = 0;
    g.rope_cos = cos_t; g.rope_sin = sin_t; g.S = S; g.rope_pos0 = pos_offset; g.head_dim = head_dim; g.rope_cols = rope_cols;
    return gemm(g, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int sf_embedding_gather(const void* table, int64_t V, int H, const int64_t* ids, int64_t n, void* out, void* stream) {
    if (!table || !ids || !out) return set_error(-22, "null argument");
    return embedding_gather(table, V, H, ids, n, out, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int64_t sf_rmsnorm_bwd_scratch_bytes(int H) { return rmsnorm_bwd_ws_bytes(H); }
extern "C" int sf_swiglu_fwd(const void* gu, void* act, int64_t M, int I, void* stream) {
    return swiglu_fwd(gu, act, M, I, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int sf_swiglu_bwd(const void* gu, const void* dact, void* dgu, int64_t M, int I, void* stream) {
    return swiglu_bwd(gu, dact, dgu, M, I, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int sf_rope(void* x, int64_t ld, int n_heads, int head_dim, const void* cos_t, const void* sin_t, int S,
                       int pos_offset, int64_t M, int inverse, void* stream) {
    return rope(x, nullptr, ld, 0, n_heads, head_dim, cos_t, sin_t, S, pos_offset, M, inverse, reinterpret_cast<cudaStream_t>(stream));
}
static void fill_attn(AttnDesc& a, const void* const* qkv, int J, int B, int S, int nh, int nkv, int d) {
    const int64_t A = (int64_t)nh * d, KV = (int64_t)nkv * d, QKV = A + 2 * KV;
    a.q = qkv[J]; a.ldq = QKV; a.ldkv = QKV; a.ldo = A; a.B = B; a.S = S; a.nh = nh; a.nkv = nkv; a.head_dim = d; a.J = J;
    a.q_row_base = qkv[J]; a.kv_row_base = qkv[0]; a.q_col0 = 0; a.k_col0 = (int)A; a.v_col0 = (int)(A + KV);
    for (int i = 0; i <= J; ++i) {
        const __nv_bfloat16* qi = reinterpret_cast<const __nv_bfloat16*>(qkv[i]);
        a.k[i] = qi + A; a.v[i] = qi + A + KV;
    }
}
extern "C" int sf_ttt_attention_fwd(const void* const* qkv, int J, void* out, float* lse, float* sd_ws, const uint8_t* key_mask,
                                    int* kvlen_ws, int B, int S, int nh, int nkv, int head_dim, void* stream) {
    if (J < 0 || J > 8) return set_error(-22, "attention: J=%d outside [0, 8]", J);
    AttnDesc a{};
    fill_attn(a, qkv, J, B, S, nh, nkv, head_dim);
    a.out = out; a.lse = lse; a.sd_ws = sd_ws; a.key_mask = key_mask;
    if (key_mask) {
        if (!kvlen_ws) return set_error(-22, "attention: key_mask needs kvlen_ws (2*B ints)");
        SF_TRY(mask_prefix(key_mask, B, S, kvlen_ws, kvlen_ws + B, reinterpret_cast<cudaStream_t>(stream)));
        a.kvlen = kvlen_ws; a.nonprefix = kvlen_ws + B;
    }
    return attn_fwd(a, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int sf_ttt_attention_bwd(const void* const* qkv, int J, const void* out, const void* dout, const float* lse,
                                    float* sd_ws, const uint8_t* key_mask, float* const* dk_acc, float* const* dv_acc,
                                    void* dq, float* delta_ws, float* dq_diag_ws, int* kvlen_ws, int B, int S, int nh, int nkv,
                                    int head_dim, void* stream) {
    if (J < 0 || J > 8) return set_error(-22, "attention: J=%d outside [0, 8]", J);
    AttnDesc a{};
    fill_attn(a, qkv, J, B, S, nh, nkv, head_dim);
    if (key_mask) {
        if (!kvlen_ws) return set_error(-22, "attention: key_mask needs kvlen_ws (2*B ints)");
        SF_TRY(mask_prefix(key_mask, B, S, kvlen_ws, kvlen_ws + B, reinterpret_cast<cudaStream_t>(stream)));
        a.kvlen = kvlen_ws; a.nonprefix = kvlen_ws + B;
    }
    a.out = const_cast<void*>(out); a.lse = const_cast<float*>(lse); a.sd_ws = sd_ws; a.key_mask = key_mask;
    a.dout = dout; a.lddo = (int64_t)nh * head_dim; a.delta_ws = delta_ws; a.ldacc = (int64_t)nkv * head_dim;
    for (int i = 0; i <= J; ++i) { a.dk_acc[i] = dk_acc[i]; a.dv_acc[i] = dv_acc[i]; }
    a.dq = dq; a.lddq = (int64_t)nh * head_dim; a.dq_diag_ws = dq_diag_ws;
    return attn_bwd(a, reinterpret_cast<cudaStream_t>(stream));
}
