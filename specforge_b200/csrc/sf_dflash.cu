// sf_dflash.cu — the DFlash block-parallel draft training step: layout, workspace, forward, backward, C ABI.
//
// Reference: `OnlineDFlashModel.forward` (algorithms/common/dflash_family_model.py:385-461) around `DFlashDraftModel.forward`
// (modeling/draft/dflash.py:431-460).  Rows: context rows r = b*S + s (Mc = B*S), draft rows r = (b*N + n)*bs + o
// (Mq = B*N*bs).  Anchors are sampled on the host (dflash_family_model.py:179-210 needs torch's RNG stream) and passed in.
//
//   noise = embed(noise_ids)                                   [Mq, H]
//   ctx   = RMSNorm_hidden(hidden_states W_fc^T)                [Mc, H]
//   per layer:  hn = RMSNorm_in(x);  [q|k|v]_n = hn W_qkv^T;  [k|v]_c = ctx W_kv^T   (W_kv = rows A.. of the fused W_qkv)
//               q,k <- per-head RMSNorm + RoPE(positions);  attn = block attention(q; [k_c;k_n], [v_c;v_n])
//               x1 = x + attn W_o^T;  x = x1 + down(silu(gate(hn2)) * up(hn2)),  hn2 = RMSNorm_post(x1)
//   logits = RMSNorm_norm(x_L) W_head^T (frozen, full vocabulary);  loss = sum(w nll) / sum(w)
//
// Every projection is the tcgen05 GEMM of sf_gemm.cuh (SwiGLU fused into the gate/up and down-dgrad epilogues as in the
// EAGLE3 step); norms reuse sf_elementwise.cu; the DFlash-specific kernels are in sf_dflash_kernels.cu.
#include "sf_gemm.cuh"
#include "sf_host.h"
#include "sf_dflash.h"
#include "../../include/specforge_b200.h"

namespace sf {
namespace dflash {

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct DD {
    int B, S, N, bs, Q, H, F, I, nh, nkv, d, L, V;
    int64_t Mc, Mq, A, KV, QKV;
};
static DD dims_of(const sf_dflash_config& c) {
    DD x;
    x.B = c.batch; x.S = c.seq_len; x.N = c.num_blocks; x.bs = c.block_size; x.Q = x.N * x.bs; x.H = c.hidden_size;
    x.F = c.num_target_feats; x.I = c.intermediate; x.nh = c.num_heads; x.nkv = c.num_kv_heads; x.d = c.head_dim;
    x.L = c.num_layers; x.V = c.vocab;
    x.Mc = (int64_t)x.B * x.S; x.Mq = (int64_t)x.B * x.Q; x.A = (int64_t)x.nh * x.d; x.KV = (int64_t)x.nkv * x.d; x.QKV = x.A + 2 * x.KV;
    return x;
}
static int validate(const sf_dflash_config& c) {
    if (c.batch <= 0 || c.seq_len <= 0 || c.num_blocks <= 0 || c.block_size <= 1) return set_error(-22, "dflash config: empty batch / blocks");
    if (c.num_layers < 1 || c.num_layers > 16) return set_error(-22, "dflash config: num_layers=%d outside [1, 16]", c.num_layers);
    if (c.hidden_size % 8 || c.intermediate % 8 || c.head_dim % 8 || c.vocab % 8) return set_error(-22, "dflash config: sizes must be multiples of 8");
    if (c.hidden_size > 8192) return set_error(-22, "dflash config: hidden size > 8192 unsupported");
    if (c.num_heads % c.num_kv_heads) return set_error(-22, "dflash config: num_heads %% num_kv_heads != 0");
    if (c.rope_rows < c.seq_len + c.block_size) return set_error(-22, "dflash config: rope tables have %d rows, need >= S+block=%d", c.rope_rows, c.seq_len + c.block_size);
    if (c.mask_token_id < 0 || c.mask_token_id >= c.vocab) return set_error(-22, "dflash config: mask_token_id outside the vocabulary");
    if (c.loss_type < 0 || c.loss_type > 3) return set_error(-22, "dflash config: loss_type=%d (0 dflash, 1 dpace, 2 cumulative-confidence-only, 3 continuation-value-only)", c.loss_type);
    if (c.loss_type != 0 && (c.dpace_alpha < 0.f || c.dpace_alpha > 1.f)) return set_error(-22, "dflash config: dpace_alpha must be in [0, 1]");
    if (c.sliding_layers >> c.num_layers) return set_error(-22, "dflash config: sliding_layers names a layer >= num_layers=%d", c.num_layers);
    if (c.sliding_layers && c.sliding_window <= 0)      // dflash.py:62-67
        return set_error(-22, "dflash config: sliding_attention layers require a positive sliding_window");
    return 0;
}

// parameter order: per layer {q, k, v, o, gate, up, down, q_norm, k_norm, input_ln, post_ln}, then fc, hidden_norm, norm
static int n_params(const sf_dflash_config& c) { return SF_DF_PER_LAYER * c.num_layers + 3; }
static void layout(const sf_dflash_config& c, int64_t* off, int64_t* sz, int64_t* total) {
    const DD x = dims_of(c);
    int64_t o = 0;
    int i = 0;
    auto put = [&](int64_t n) { off[i] = o; sz[i] = n; o += n; ++i; };
    for (int l = 0; l < x.L; ++l) {
        put(x.A * x.H); put(x.KV * x.H); put(x.KV * x.H); put((int64_t)x.H * x.A);
        put((int64_t)x.I * x.H); put((int64_t)x.I * x.H); put((int64_t)x.H * x.I);
        put(x.d); put(x.d); put(x.H); put(x.H);
    }
    put((int64_t)x.H * x.F * x.H); put(x.H); put(x.H);
    *total = o;
}

struct Plan {
    int64_t pos, tgt, nid, w, lw, sums, row_loss, row_correct, row_state;
    int64_t ctx_raw, ctx, x, hn, qkv, kvc, qr, krn, krc, attn, lse, x1, hn2, gu, act, hf, logits;
    // backward
    int64_t dxa, dxb, dtmp, dgu, dact, dattn, dqr, dkrn, dkrc, dqkv, dkvc, delta, dctx32, dctx, norm_ws, head_ws;
    int64_t total;
};
static Plan make_plan(const sf_dflash_config& c) {
    const DD x = dims_of(c);
    Plan p{};
    int64_t o = 0;
    auto take = [&](int64_t bytes) { int64_t r = o; o = align_up(o + bytes, 1024); return r; };
    const int64_t Mq = x.Mq, Mc = x.Mc, L = x.L;
    p.pos = take(Mq * 4); p.tgt = take(Mq * 4); p.nid = take(Mq * 4); p.w = take(Mq * 4); p.lw = take(Mq * 4);
    p.sums = take(64); p.row_loss = take(Mq * 4); p.row_correct = take(Mq * 4); p.row_state = take(Mq * 8);
    p.ctx_raw = take(Mc * x.H * 2); p.ctx = take(Mc * x.H * 2);
    p.x = take((L + 1) * Mq * x.H * 2);
    p.hn = take(L * Mq * x.H * 2);
    p.qkv = take(L * Mq * x.QKV * 2);
    p.kvc = take(L * Mc * 2 * x.KV * 2);
    p.qr = take(L * Mq * x.A * 2); p.krn = take(L * Mq * x.KV * 2); p.krc = take(L * Mc * x.KV * 2);
    p.attn = take(L * Mq * x.A * 2); p.lse = take(L * Mq * x.nh * 4);
    p.x1 = take(L * Mq * x.H * 2); p.hn2 = take(L * Mq * x.H * 2);
    p.gu = take(L * Mq * 2 * x.I * 2); p.act = take(L * Mq * x.I * 2);
    p.hf = take(Mq * x.H * 2); p.logits = take(Mq * (int64_t)x.V * 2);
    p.dxa = take(Mq * x.H * 2); p.dxb = take(Mq * x.H * 2); p.dtmp = take(Mq * x.H * 2);
    p.dgu = take(Mq * 2 * x.I * 2); p.dact = take(Mq * x.I * 2); p.dattn = take(Mq * x.A * 2);
    p.dqr = take(Mq * x.A * 2); p.dkrn = take(Mq * x.KV * 2);      // dV goes straight into d(qkv) / d(kv_c)
    p.dkrc = take(Mc * x.KV * 2);
    p.dqkv = take(Mq * x.QKV * 2); p.dkvc = take(Mc * 2 * x.KV * 2);
    p.delta = take(Mq * x.nh * 4);
    p.dctx32 = take(Mc * x.H * 4); p.dctx = take(Mc * x.H * 2);
    p.norm_ws = take(rmsnorm_bwd_ws_bytes(x.H));
    p.head_ws = take((int64_t)kNormBlocks * x.d * 4);
    p.total = o;
    return p;
}

struct Ctx {
    const sf_dflash_config* cfg; DD x; Plan p; uint8_t* ws; cudaStream_t st;
    int64_t off[SF_DF_PER_LAYER * 16 + 3], sz[SF_DF_PER_LAYER * 16 + 3], total;
    const __nv_bfloat16* params;
    const __nv_bfloat16* W(int l, int which) const { return params + off[l * SF_DF_PER_LAYER + which]; }
    const __nv_bfloat16* Wg(int which) const { return params + off[x.L * SF_DF_PER_LAYER + which]; }   // 0 fc, 1 hidden_norm, 2 norm
    int64_t goff(int l, int which) const { return off[l * SF_DF_PER_LAYER + which]; }
    int64_t ggoff(int which) const { return off[x.L * SF_DF_PER_LAYER + which]; }
    template <typename T> T* at(int64_t o) const { return reinterpret_cast<T*>(ws + o); }
    __nv_bfloat16* bf(int64_t o, int64_t elem_off = 0) const { return reinterpret_cast<__nv_bfloat16*>(ws + o) + elem_off; }
};

static int mm(const Ctx& c, const void* A, int64_t lda, int am, const void* B, int64_t ldb, int bm, void* D, int64_t ldd,
              const void* R, int64_t ldr, int64_t M, int64_t N, int64_t K, int epi, void* D2 = nullptr, int64_t ldd2 = 0, int n_half = 0) {
    GemmDesc g;
    g.D2 = D2; g.ldd2 = ldd2; g.n_half = n_half;
    g.A = A; g.lda = lda; g.a_major = am; g.B = B; g.ldb = ldb; g.b_major = bm; g.D = D; g.ldd = ldd; g.R = R; g.ldr = ldr;
    g.M = (int)M; g.N = (int)N; g.K = (int)K; g.epi = epi; g.cta_group = 0;
    return gemm(g, c.st);
}
#define SF_TRY(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

static bool fuse_swiglu(const DD& x) { return opt(OPT_NO_SWIGLU_FUSION) != 1 && x.Mq > 128 && x.I % 128 == 0; }

static int setup(Ctx& c, const sf_dflash_config* cfg, const void* params_flat, void* ws, size_t ws_bytes, void* stream) {
    if (!cfg || !params_flat || !ws) return set_error(-22, "null argument");
    SF_TRY(validate(*cfg));
    c.cfg = cfg; c.x = dims_of(*cfg); c.p = make_plan(*cfg); c.ws = reinterpret_cast<uint8_t*>(ws);
    c.st = reinterpret_cast<cudaStream_t>(stream);
    if ((size_t)c.p.total > ws_bytes) return set_error(-12, "dflash workspace too small: need %lld bytes, got %zu", (long long)c.p.total, ws_bytes);
    if (reinterpret_cast<uintptr_t>(ws) & 1023) return set_error(-22, "workspace must be 1024-byte aligned");
    layout(*cfg, c.off, c.sz, &c.total);
    c.params = reinterpret_cast<const __nv_bfloat16*>(params_flat);
    return 0;
}

static AttnArgs attn_args(const Ctx& c, const sf_dflash_batch& bt, int l) {
    const DD& x = c.x; const Plan& p = c.p;
    AttnArgs a{};
    a.q = c.bf(p.qr, (int64_t)l * x.Mq * x.A); a.ldq = x.A;
    a.kn = c.bf(p.krn, (int64_t)l * x.Mq * x.KV); a.ldkn = x.KV;
    a.vn = c.bf(p.qkv, (int64_t)l * x.Mq * x.QKV) + x.A + x.KV; a.ldvn = x.QKV;
    a.kc = c.bf(p.krc, (int64_t)l * x.Mc * x.KV); a.ldkc = x.KV;
    a.vc = c.bf(p.kvc, (int64_t)l * x.Mc * 2 * x.KV) + x.KV; a.ldvc = 2 * x.KV;
    a.out = c.bf(p.attn, (int64_t)l * x.Mq * x.A); a.ldo = x.A;
    a.lse = c.at<float>(p.lse) + (int64_t)l * x.Mq * x.nh;
    a.anchors = bt.anchors; a.keep = bt.keep;
    a.B = x.B; a.S = x.S; a.N = x.N; a.bs = x.bs; a.nh = x.nh; a.nkv = x.nkv; a.d = x.d;
    a.scale = 1.0f / sqrtf((float)x.d);
    a.window = ((c.cfg->sliding_layers >> l) & 1u) ? c.cfg->sliding_window : 0;
    return a;
}

// ------------------------------------------------------------------ forward
static int forward(Ctx& c, const sf_dflash_frozen& fz, const sf_dflash_batch& bt, float* metrics_out, float* loss_out, int need_grad) {
    const DD& x = c.x; const Plan& p = c.p; const sf_dflash_config& cfg = *c.cfg;
    cudaStream_t st = c.st;
    const int64_t Mq = x.Mq, Mc = x.Mc;
    SF_TRY(rows(bt.input_ids, bt.loss_mask, bt.anchors, bt.keep, x.B, x.S, x.N, x.bs, cfg.loss_decay_gamma, cfg.mask_token_id,
                c.at<int32_t>(p.pos), c.at<int32_t>(p.tgt), c.at<int32_t>(p.nid), c.at<float>(p.w), c.at<float>(p.lw), c.at<float>(p.sums), st));
    SF_TRY(gather_rows(fz.embed_tokens, c.at<int32_t>(p.nid), c.bf(p.x), Mq, x.H, st));
    // ctx = hidden_norm(fc(target hidden))   (dflash.py:442)
    SF_TRY(mm(c, bt.hidden_states, (int64_t)x.F * x.H, MAJOR_K, c.Wg(0), (int64_t)x.F * x.H, MAJOR_K, c.bf(p.ctx_raw), x.H, nullptr, 0, Mc, x.H,
              (int64_t)x.F * x.H, EPI_BF16));
    SF_TRY(rmsnorm_fwd(c.bf(p.ctx_raw), x.H, nullptr, x.S, 0, c.Wg(1), c.bf(p.ctx), x.H, Mc, x.H, cfg.rms_eps, nullptr, st));
    for (int l = 0; l < x.L; ++l) {
        __nv_bfloat16* xin = c.bf(p.x, (int64_t)l * Mq * x.H);
        __nv_bfloat16* xout = c.bf(p.x, (int64_t)(l + 1) * Mq * x.H);
        __nv_bfloat16* hn = c.bf(p.hn, (int64_t)l * Mq * x.H);
        __nv_bfloat16* qkv = c.bf(p.qkv, (int64_t)l * Mq * x.QKV);
        __nv_bfloat16* kvc = c.bf(p.kvc, (int64_t)l * Mc * 2 * x.KV);
        __nv_bfloat16* x1 = c.bf(p.x1, (int64_t)l * Mq * x.H);
        __nv_bfloat16* hn2 = c.bf(p.hn2, (int64_t)l * Mq * x.H);
        __nv_bfloat16* gu = c.bf(p.gu, (int64_t)l * Mq * 2 * x.I);
        __nv_bfloat16* act = c.bf(p.act, (int64_t)l * Mq * x.I);
        SF_TRY(rmsnorm_fwd(xin, x.H, nullptr, x.S, 0, c.W(l, SF_DF_INPUT_LN), hn, x.H, Mq, x.H, cfg.rms_eps, nullptr, st));
        SF_TRY(mm(c, hn, x.H, MAJOR_K, c.W(l, SF_DF_Q), x.H, MAJOR_K, qkv, x.QKV, nullptr, 0, Mq, x.QKV, x.H, EPI_BF16));
        SF_TRY(mm(c, c.bf(p.ctx), x.H, MAJOR_K, c.W(l, SF_DF_K), x.H, MAJOR_K, kvc, 2 * x.KV, nullptr, 0, Mc, 2 * x.KV, x.H, EPI_BF16));
        // per-head RMSNorm + RoPE: q and noise k at position anchor + o, context k at position s   (dflash.py:160-177)
        SF_TRY(headnorm_rope_fwd(qkv, x.QKV, x.nh, x.d, c.W(l, SF_DF_Q_NORM), fz.rope_cos, fz.rope_sin, c.at<int32_t>(p.pos), x.S, cfg.rms_eps,
                                 c.bf(p.qr, (int64_t)l * Mq * x.A), x.A, Mq, st));
        SF_TRY(headnorm_rope_fwd(qkv + x.A, x.QKV, x.nkv, x.d, c.W(l, SF_DF_K_NORM), fz.rope_cos, fz.rope_sin, c.at<int32_t>(p.pos), x.S,
                                 cfg.rms_eps, c.bf(p.krn, (int64_t)l * Mq * x.KV), x.KV, Mq, st));
        SF_TRY(headnorm_rope_fwd(kvc, 2 * x.KV, x.nkv, x.d, c.W(l, SF_DF_K_NORM), fz.rope_cos, fz.rope_sin, nullptr, x.S, cfg.rms_eps,
                                 c.bf(p.krc, (int64_t)l * Mc * x.KV), x.KV, Mc, st));
        AttnArgs a = attn_args(c, bt, l);
        SF_TRY(attn_fwd(a, st));
        SF_TRY(mm(c, a.out, x.A, MAJOR_K, c.W(l, SF_DF_O), x.A, MAJOR_K, x1, x.H, xin, x.H, Mq, x.H, x.A, EPI_BF16_RESID));
        SF_TRY(rmsnorm_fwd(x1, x.H, nullptr, x.S, 0, c.W(l, SF_DF_POST_LN), hn2, x.H, Mq, x.H, cfg.rms_eps, nullptr, st));
        if (fuse_swiglu(x)) {
            SF_TRY(mm(c, hn2, x.H, MAJOR_K, c.W(l, SF_DF_GATE), x.H, MAJOR_K, gu, 2 * x.I, nullptr, 0, Mq, 2 * x.I, x.H, EPI_SWIGLU, act, x.I, x.I));
        } else {
            SF_TRY(mm(c, hn2, x.H, MAJOR_K, c.W(l, SF_DF_GATE), x.H, MAJOR_K, gu, 2 * x.I, nullptr, 0, Mq, 2 * x.I, x.H, EPI_BF16));
            SF_TRY(swiglu_fwd(gu, act, Mq, x.I, st));
        }
        SF_TRY(mm(c, act, x.I, MAJOR_K, c.W(l, SF_DF_DOWN), x.I, MAJOR_K, xout, x.H, x1, x.H, Mq, x.H, x.I, EPI_BF16_RESID));
    }
    SF_TRY(rmsnorm_fwd(c.bf(p.x, (int64_t)x.L * Mq * x.H), x.H, nullptr, x.S, 0, c.Wg(2), c.bf(p.hf), x.H, Mq, x.H, cfg.rms_eps, nullptr, st));
    SF_TRY(mm(c, c.bf(p.hf), x.H, MAJOR_K, fz.lm_head, x.H, MAJOR_K, c.bf(p.logits), x.V, nullptr, 0, Mq, x.V, x.H, EPI_BF16));
    SF_TRY(ce(c.bf(p.logits), x.V, x.V, c.at<int32_t>(p.tgt), c.at<float>(p.w), c.at<float>(p.lw), c.at<float>(p.sums), need_grad,
              c.at<float>(p.row_loss), c.at<float>(p.row_correct), Mq, cfg.grad_of_numerator, cfg.loss_type, cfg.dpace_alpha, x.bs, x.B,
              c.at<float>(p.row_state), st));
    // metrics = {loss_num, loss_den, correct, acc_den}; loss = loss_num / loss_den
    return finalize_loss(c.at<float>(p.sums), metrics_out, loss_out, st);
}

// ------------------------------------------------------------------ backward
static int backward(Ctx& c, const sf_dflash_frozen& fz, const sf_dflash_batch& bt, float* G, int accumulate) {
    const DD& x = c.x; const Plan& p = c.p; const sf_dflash_config& cfg = *c.cfg;
    cudaStream_t st = c.st;
    const int64_t Mq = x.Mq, Mc = x.Mc;
    if (!accumulate && cudaMemsetAsync(G, 0, (size_t)c.total * 4, st) != cudaSuccess) return set_error(-5, "memset grads failed");
    float* nws = c.at<float>(p.norm_ws);
    float* hws = c.at<float>(p.head_ws);
    // d(hf) = d(logits) W_head ; through the final norm
    SF_TRY(mm(c, c.bf(p.logits), x.V, MAJOR_K, fz.lm_head, x.H, MAJOR_MN, c.bf(p.dtmp), x.H, nullptr, 0, Mq, x.H, x.V, EPI_BF16));
    __nv_bfloat16* dx = c.bf(p.dxa);        // gradient w.r.t. the layer output, ping-pongs with dxb
    __nv_bfloat16* dx_next = c.bf(p.dxb);
    SF_TRY(rmsnorm_bwd(c.bf(p.x, (int64_t)x.L * Mq * x.H), x.H, nullptr, x.S, 0, c.Wg(2), c.bf(p.dtmp), x.H, nullptr, nullptr, dx,
                       G + c.ggoff(2), nws, Mq, x.H, cfg.rms_eps, st));
    for (int l = x.L - 1; l >= 0; --l) {
        const __nv_bfloat16* xin = c.bf(p.x, (int64_t)l * Mq * x.H);
        const __nv_bfloat16* hn = c.bf(p.hn, (int64_t)l * Mq * x.H);
        const __nv_bfloat16* qkv = c.bf(p.qkv, (int64_t)l * Mq * x.QKV);
        const __nv_bfloat16* kvc = c.bf(p.kvc, (int64_t)l * Mc * 2 * x.KV);
        const __nv_bfloat16* x1 = c.bf(p.x1, (int64_t)l * Mq * x.H);
        const __nv_bfloat16* hn2 = c.bf(p.hn2, (int64_t)l * Mq * x.H);
        const __nv_bfloat16* gu = c.bf(p.gu, (int64_t)l * Mq * 2 * x.I);
        const __nv_bfloat16* act = c.bf(p.act, (int64_t)l * Mq * x.I);
        // ---- MLP: x_out = x1 + down(act)
        SF_TRY(mm(c, dx, x.H, MAJOR_MN, act, x.I, MAJOR_MN, G + c.goff(l, SF_DF_DOWN), x.I, nullptr, 0, x.H, x.I, Mq, EPI_F32_ACCUM));
        if (fuse_swiglu(x) && x.I % 32 == 0) {
            SF_TRY(mm(c, dx, x.H, MAJOR_K, c.W(l, SF_DF_DOWN), x.I, MAJOR_MN, c.bf(p.dgu), 2 * x.I, gu, 2 * x.I, Mq, x.I, x.H, EPI_SWIGLU_BWD, nullptr, 0, x.I));
        } else {
            SF_TRY(mm(c, dx, x.H, MAJOR_K, c.W(l, SF_DF_DOWN), x.I, MAJOR_MN, c.bf(p.dact), x.I, nullptr, 0, Mq, x.I, x.H, EPI_BF16));
            SF_TRY(swiglu_bwd(gu, c.bf(p.dact), c.bf(p.dgu), Mq, x.I, st));
        }
        SF_TRY(mm(c, c.bf(p.dgu), 2 * x.I, MAJOR_MN, hn2, x.H, MAJOR_MN, G + c.goff(l, SF_DF_GATE), x.H, nullptr, 0, 2 * x.I, x.H, Mq, EPI_F32_ACCUM));
        SF_TRY(mm(c, c.bf(p.dgu), 2 * x.I, MAJOR_K, c.W(l, SF_DF_GATE), x.H, MAJOR_MN, c.bf(p.dtmp), x.H, nullptr, 0, Mq, x.H, 2 * x.I, EPI_BF16));
        // d(x1) = d(x_out) + RMSNorm_post^T(d hn2)
        SF_TRY(rmsnorm_bwd(x1, x.H, nullptr, x.S, 0, c.W(l, SF_DF_POST_LN), c.bf(p.dtmp), x.H, dx, nullptr, dx_next, G + c.goff(l, SF_DF_POST_LN),
                           nws, Mq, x.H, cfg.rms_eps, st));
        // ---- attention: x1 = x_in + attn W_o^T
        AttnArgs a = attn_args(c, bt, l);
        SF_TRY(mm(c, dx_next, x.H, MAJOR_MN, a.out, x.A, MAJOR_MN, G + c.goff(l, SF_DF_O), x.A, nullptr, 0, x.H, x.A, Mq, EPI_F32_ACCUM));
        SF_TRY(mm(c, dx_next, x.H, MAJOR_K, c.W(l, SF_DF_O), x.A, MAJOR_MN, c.bf(p.dattn), x.A, nullptr, 0, Mq, x.A, x.H, EPI_BF16));
        a.dout = c.bf(p.dattn); a.lddo = x.A; a.delta = c.at<float>(p.delta);
        a.dq = c.bf(p.dqr); a.lddq = x.A;
        a.dkn = c.bf(p.dkrn); a.lddkn = x.KV;
        a.dvn = c.bf(p.dqkv) + x.A + x.KV; a.lddvn = x.QKV;          // dV of the noise rows goes straight into d(qkv)
        a.dkc = c.bf(p.dkrc); a.lddkc = x.KV;
        a.dvc = c.bf(p.dkvc) + x.KV; a.lddvc = 2 * x.KV;             // dV of the context rows straight into d(kv_c)
        SF_TRY(attn_bwd(a, st));
        // back through RoPE + per-head norms into the raw projections
        SF_TRY(headnorm_rope_bwd(qkv, x.QKV, x.nh, x.d, c.W(l, SF_DF_Q_NORM), fz.rope_cos, fz.rope_sin, c.at<int32_t>(p.pos), x.S, cfg.rms_eps,
                                 c.bf(p.dqr), x.A, c.bf(p.dqkv), x.QKV, G + c.goff(l, SF_DF_Q_NORM), 1, hws, Mq, st));
        SF_TRY(headnorm_rope_bwd(qkv + x.A, x.QKV, x.nkv, x.d, c.W(l, SF_DF_K_NORM), fz.rope_cos, fz.rope_sin, c.at<int32_t>(p.pos), x.S,
                                 cfg.rms_eps, c.bf(p.dkrn), x.KV, c.bf(p.dqkv) + x.A, x.QKV, G + c.goff(l, SF_DF_K_NORM), 1, hws, Mq, st));
        SF_TRY(headnorm_rope_bwd(kvc, 2 * x.KV, x.nkv, x.d, c.W(l, SF_DF_K_NORM), fz.rope_cos, fz.rope_sin, nullptr, x.S, cfg.rms_eps,
                                 c.bf(p.dkrc), x.KV, c.bf(p.dkvc), 2 * x.KV, G + c.goff(l, SF_DF_K_NORM), 1, hws, Mc, st));
        // weight gradients of the fused [q;k;v] (noise rows) and of its [k;v] rows again from the context rows
        SF_TRY(mm(c, c.bf(p.dqkv), x.QKV, MAJOR_MN, hn, x.H, MAJOR_MN, G + c.goff(l, SF_DF_Q), x.H, nullptr, 0, x.QKV, x.H, Mq, EPI_F32_ACCUM));
        SF_TRY(mm(c, c.bf(p.dkvc), 2 * x.KV, MAJOR_MN, c.bf(p.ctx), x.H, MAJOR_MN, G + c.goff(l, SF_DF_K), x.H, nullptr, 0, 2 * x.KV, x.H, Mc, EPI_F32_ACCUM));
        // d(ctx) accumulates over the layers (fp32), d(hn) goes through the input norm
        SF_TRY(mm(c, c.bf(p.dkvc), 2 * x.KV, MAJOR_K, c.W(l, SF_DF_K), x.H, MAJOR_MN, c.at<float>(p.dctx32), x.H, nullptr, 0, Mc, x.H, 2 * x.KV,
                  l == x.L - 1 ? EPI_F32 : EPI_F32_ACCUM));
        SF_TRY(mm(c, c.bf(p.dqkv), x.QKV, MAJOR_K, c.W(l, SF_DF_Q), x.H, MAJOR_MN, c.bf(p.dtmp), x.H, nullptr, 0, Mq, x.H, x.QKV, EPI_BF16));
        SF_TRY(rmsnorm_bwd(xin, x.H, nullptr, x.S, 0, c.W(l, SF_DF_INPUT_LN), c.bf(p.dtmp), x.H, dx_next, nullptr, dx, G + c.goff(l, SF_DF_INPUT_LN),
                           nws, Mq, x.H, cfg.rms_eps, st));
        // dx now holds d(x_in) = d(x1) + RMSNorm_in^T(d hn): it is the next (lower) layer's output gradient
    }
    // context branch: ctx = hidden_norm(ctx_raw), ctx_raw = hidden_states W_fc^T (the noise embedding is frozen: dx is dropped)
    SF_TRY(cvt_f32_bf16(c.at<float>(p.dctx32), x.H, c.bf(p.dctx), x.H, Mc, x.H, 1.0f, st));
    SF_TRY(rmsnorm_bwd(c.bf(p.ctx_raw), x.H, nullptr, x.S, 0, c.Wg(1), c.bf(p.dctx), x.H, nullptr, nullptr, c.bf(p.ctx) /* reuse as d(ctx_raw) */,
                       G + c.ggoff(1), nws, Mc, x.H, cfg.rms_eps, st));
    SF_TRY(mm(c, c.bf(p.ctx), x.H, MAJOR_MN, bt.hidden_states, (int64_t)x.F * x.H, MAJOR_MN, G + c.ggoff(0), (int64_t)x.F * x.H, nullptr, 0, x.H,
              (int64_t)x.F * x.H, Mc, EPI_F32_ACCUM));
    return 0;
}

}  // namespace dflash
}  // namespace sf

// =============================================================================== C ABI
using namespace sf;
using namespace sf::dflash;

extern "C" int sf_dflash_num_params(const sf_dflash_config* cfg) {
    if (!cfg) return set_error(-22, "null argument");
    if (int rc = validate(*cfg)) return rc;
    return n_params(*cfg);
}
extern "C" int sf_dflash_param_layout(const sf_dflash_config* cfg, int64_t* offsets, int64_t* sizes, int64_t* total) {
    if (!cfg || !offsets || !sizes || !total) return set_error(-22, "null argument");
    if (int rc = validate(*cfg)) return rc;
    layout(*cfg, offsets, sizes, total);
    return 0;
}
extern "C" size_t sf_dflash_workspace_bytes(const sf_dflash_config* cfg) {
    if (!cfg || validate(*cfg)) return 0;
    return (size_t)make_plan(*cfg).total;
}
extern "C" int sf_dflash_forward(const sf_dflash_config* cfg, const void* params_flat, const sf_dflash_frozen* frozen,
                                 const sf_dflash_batch* batch, void* workspace, size_t workspace_bytes, float* metrics, float* loss,
                                 int need_grad, void* stream) {
    Ctx c;
    if (!frozen || !batch || !metrics || !loss) return set_error(-22, "null argument");
    SF_TRY(setup(c, cfg, params_flat, workspace, workspace_bytes, stream));
    return forward(c, *frozen, *batch, metrics, loss, need_grad);
}
extern "C" int sf_dflash_backward(const sf_dflash_config* cfg, const void* params_flat, const sf_dflash_frozen* frozen,
                                  const sf_dflash_batch* batch, void* workspace, size_t workspace_bytes, float* grads_flat_f32,
                                  int accumulate, void* stream) {
    Ctx c;
    if (!frozen || !batch || !grads_flat_f32) return set_error(-22, "null argument");
    SF_TRY(setup(c, cfg, params_flat, workspace, workspace_bytes, stream));
    return backward(c, *frozen, *batch, grads_flat_f32, accumulate);
}

// ---- op-level entry points: the DFlash block attention alone, on contiguous tensors (tests / tools; `impl` picks the kernels:
// 0 = CUDA-core tiles, 1 = tcgen05, -1 = whatever the step would use)
static AttnArgs op_args(const void* q, const void* kn, const void* vn, const void* kc, const void* vc, void* out, float* lse,
                        const int32_t* anchors, const uint8_t* keep, int B, int S, int N, int bs, int nh, int nkv, int d) {
    AttnArgs a{};
    const int64_t A = (int64_t)nh * d, KV = (int64_t)nkv * d;
    a.q = (const __nv_bfloat16*)q; a.ldq = A; a.kn = (const __nv_bfloat16*)kn; a.ldkn = KV; a.vn = (const __nv_bfloat16*)vn; a.ldvn = KV;
    a.kc = (const __nv_bfloat16*)kc; a.ldkc = KV; a.vc = (const __nv_bfloat16*)vc; a.ldvc = KV;
    a.out = (__nv_bfloat16*)out; a.ldo = A; a.lse = lse; a.anchors = anchors; a.keep = keep;
    a.B = B; a.S = S; a.N = N; a.bs = bs; a.nh = nh; a.nkv = nkv; a.d = d; a.scale = 1.0f / sqrtf((float)d);
    a.window = opt(OPT_DFLASH_ATTN_WINDOW) > 0 ? opt(OPT_DFLASH_ATTN_WINDOW) : 0;      // diagnostic: the op-level calls have no config
    return a;
}
extern "C" int sf_dflash_attention_fwd(const void* q, const void* kn, const void* vn, const void* kc, const void* vc, void* out, float* lse,
                                       const int32_t* anchors, const uint8_t* keep, int B, int S, int N, int bs, int nh, int nkv, int d,
                                       int impl, void* stream) {
    if (!q || !kn || !vn || !kc || !vc || !out || !lse || !anchors || !keep) return set_error(-22, "null argument");
    const AttnArgs a = op_args(q, kn, vn, kc, vc, out, lse, anchors, keep, B, S, N, bs, nh, nkv, d);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (impl == 1) return attn_tc_supported(a) ? attn_fwd_tc(a, st) : set_error(-22, "dflash attention: shape not covered by the tcgen05 path");
    return impl == 0 ? attn_fwd_cc(a, st) : attn_fwd(a, st);
}
extern "C" int sf_dflash_attention_bwd(const void* q, const void* kn, const void* vn, const void* kc, const void* vc, const void* out,
                                       const float* lse, const void* dout, const int32_t* anchors, const uint8_t* keep, void* dq, void* dkn,
                                       void* dvn, void* dkc, void* dvc, float* delta_ws, int B, int S, int N, int bs, int nh, int nkv, int d,
                                       int impl, void* stream) {
    if (!q || !kn || !vn || !kc || !vc || !out || !lse || !dout || !dq || !dkn || !dvn || !dkc || !dvc || !delta_ws) return set_error(-22, "null argument");
    AttnArgs a = op_args(q, kn, vn, kc, vc, const_cast<void*>(out), const_cast<float*>(lse), anchors, keep, B, S, N, bs, nh, nkv, d);
    const int64_t A = (int64_t)nh * d, KV = (int64_t)nkv * d;
    a.dout = (const __nv_bfloat16*)dout; a.lddo = A; a.delta = delta_ws;
    a.dq = (__nv_bfloat16*)dq; a.lddq = A; a.dkn = (__nv_bfloat16*)dkn; a.lddkn = KV; a.dvn = (__nv_bfloat16*)dvn; a.lddvn = KV;
    a.dkc = (__nv_bfloat16*)dkc; a.lddkc = KV; a.dvc = (__nv_bfloat16*)dvc; a.lddvc = KV;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (impl == 1) return attn_tc_bwd_supported(a) ? attn_bwd_tc(a, st) : set_error(-22, "dflash attention: shape not covered by the tcgen05 path");
    return impl == 0 ? attn_bwd_cc(a, st) : attn_bwd(a, st);
}

