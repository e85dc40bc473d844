// DualARTransformer + generation loops behind the C ABI (include/fsb.h).
// Host-side mirror of fish_speech_core/lib/lm/dual_ar.rs:443-713 and
// lm/generate/single_batch.rs; the device work is in fsb_lm_kernels.cuh and
// fsb_lm_mega.cuh.  No CPU fallback: every entry point needs an sm_100 device.
#include <algorithm>
#include <cmath>
#include <memory>

#include "fsb_lm_kernels.cuh"
#include "fsb_lm_mega_params.cuh"
#include "fsb_tc_gemm.cuh"

namespace fsb {

// mod.rs:67 gates the top-p branch in f64 (`top_p <= 0.0 || top_p >= sum_p as f64`); sum_p is an f32, so the gate is
// `sum_p <= (largest f32 that is <= top_p)`, which the device can test without doubles
static float top_p_gate_of(double top_p) {
    if (top_p <= 0.0) return INFINITY;
    float f = (float)top_p;
    if ((double)f > top_p) f = std::nextafterf(f, -INFINITY);
    return f;
}

struct LayerW {
    DevTensor wqkv, wo, w1, w2, w3;
    TcMap m_wqkv, m_wo, m_w1, m_w2, m_w3;  // TMA maps of the bf16 matrices (tcgen05 prefill)
    DevTensor attn_norm, ffn_norm;  // always f32
};

struct Scratch {
    float *x = nullptr, *xn = nullptr, *qkv = nullptr, *q = nullptr, *att = nullptr, *g1 = nullptr, *g3 = nullptr;
    float *partial = nullptr;
    float *logits = nullptr;       // (B, V)   step API only
    float *slow_logits = nullptr;  // (B, V')
    float *hidden = nullptr;       // (B, D)
    float *fast_x = nullptr;       // (B, D)
    float *fast_logits = nullptr;  // (B, CS)
    uint32_t *toks = nullptr;      // (C+1, Mmax) prompt staging
};

}  // namespace fsb

using namespace fsb;

struct fsb_kv_snapshot {
    float *k = nullptr, *v = nullptr;  // (NL, KV, n, hd) each
    size_t n = 0;
};

struct fsb_lm {
    fsb_model_args cfg;
    fsb_token_config tok;
    fsb_lm_options opt;
    int D, I, H, KV, hd, V, C, CS, QKV, NL, NFL;
    int max_batch, max_len, fast_len;
    int wdt;  // weight dtype
    int prefill_rows;
    size_t toks_cap = 0;  // u32 words of the prompt staging arena
    int4 *d_segs = nullptr;  // segment table of the current prefill pass (<= 256 entries)
    int nsplit;
    int n_slow_logits, slow_row0, slow_rest_base;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bool poisoned = false;
    std::vector<void *> owned;
    DevTensor emb, cb_emb, out_w, fast_emb, fast_out, norm, fast_norm;
    std::vector<LayerW> layers, fast_layers;
    float *cosT = nullptr, *sinT = nullptr;
    float *kc = nullptr, *vc = nullptr;    // (NL, B, KV, max_len, hd)
    float *fkc = nullptr, *fvc = nullptr;  // (NFL, B, KV, fast_len, hd)
    Scratch s;
    // generation state
    GenState h_st;            // host copy of the device struct
    GenState *d_st = nullptr;
    std::vector<int> kv_len;  // host mirror of pos[] between calls
    int *h_pin = nullptr;     // pinned staging (ints)
    std::map<int, cudaGraphExec_t> frame_graphs;  // keyed by bsz (+ 4096 when hidden states are collected)
    std::map<int, cudaGraphExec_t> tail_graphs;
    bool session = false;              // fsb_lm_session_*: rows are slots
    std::vector<int> slot_state;       // 0 free, 1 generating, 2 finished (awaiting collect)
    bool collect_hidden = false;  // generate_blocking_with_hidden: per-op path + store_hidden_kernel per frame
    float *hid = nullptr;         // (max_batch, out_cap, D), allocated on first use
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    fsb_lm_stats stats;
    uint64_t launches = 0;
    // tcgen05 prefill (bf16 weights): split-activation buffers (hi | mid | lo) and their TMA maps per N tile
    bool tc_ok = false;
    float *mega_rep = nullptr;
    unsigned long long *mega_ll = nullptr;
    size_t mega_ll_words = 0;
    size_t smem_optin = 0;  // cudaDevAttrMaxSharedMemoryPerBlockOptin, cached by mega_setup
    __nv_bfloat16 *sp_xn = nullptr, *sp_att = nullptr, *sp_h = nullptr;
    float *tc_ws = nullptr;  // split-K workspace of the decode-sized GEMMs
    size_t tc_ws_floats = 0;
    TcMap mx_xn[3], mx_att[3], mx_h[3];  // index: bn 32 / 64 / 128
    // persistent megakernel (decode_mode 2)
    MegaParams mp;
    bool mega_ok = false;
    int mega_grid = 0;
    MegaLayer *d_mega_slow = nullptr, *d_mega_fast = nullptr;
    float *mega_partial = nullptr, *mega_logits = nullptr;
    unsigned int *mega_bar = nullptr;
    unsigned long long *mega_dbg = nullptr;
    cudaEvent_t prof_m0 = nullptr, prof_m1 = nullptr;
    // wide-batch megakernel (fsb_lm_megab.cuh): 9..32 rows, bf16 weights
    bool megab_ok = false;
    MegaBExtra mbx;
    int megab_rows_cap = 0;  // rows the split-K workspace / attention scratch are sized for
    // profile mode: event pairs around the weight-streaming kernel
    bool profile = false;
    std::vector<cudaEvent_t> prof_ev;
    size_t prof_used = 0;
    uint64_t prof_bytes = 0;
};

namespace fsb {

static size_t esize(int dt) { return dt == FSB_F32 ? 4 : 2; }

template <typename T>
static int dev_alloc(fsb_lm *lm, T **p, size_t n) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? FSB_ERR_OOM : FSB_ERR_CUDA;
    }
    lm->owned.push_back(q);
    *p = reinterpret_cast<T *>(q);
    return FSB_OK;
}

#define LAUNCH_CHECK(lm)                                                          \
    do {                                                                          \
        cudaError_t _e = cudaGetLastError();                                      \
        if (_e != cudaSuccess) {                                                  \
            set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return FSB_ERR_CUDA;                                                  \
        }                                                                         \
        (lm)->launches++;                                                         \
    } while (0)

// ---------------------------------------------------------------- GEMV dispatch
static int pick_gemv_grid(int rows) {
    const int ctas = (rows + kGemvRowsPerCta - 1) / kGemvRowsPerCta;
    return std::max(1, std::min(ctas, 148 * 4));
}

template <typename WT, int EPI>
static int launch_gemv_nb(fsb_lm *lm, GemvArgs a, int nb) {
    const int grid = pick_gemv_grid(a.rows);
    // rows of `x`/`y` are processed in groups of <= 8; a group re-reads the weights (L2-resident)
    for (int b0 = 0; b0 < nb; b0 += 8) {
        GemvArgs g = a;
        const int n = std::min(8, nb - b0);
        g.x = a.x + (size_t)b0 * a.ldx;
        g.y = a.y + (size_t)b0 * a.ldy;
        if (a.resid) g.resid = a.resid + (size_t)b0 * a.ldy;
#define GEMV_CASE(NB)                                                                                        \
    case NB:                                                                                                 \
        gemv_kernel<WT, NB, EPI><<<grid, kGemvThreads, (size_t)NB * a.K * sizeof(float), lm->stream>>>(g); \
        break;
        const bool prof = lm->profile && lm->prof_used + 2 <= lm->prof_ev.size();
        if (prof) cudaEventRecord(lm->prof_ev[lm->prof_used], lm->stream);
        switch (n) {
            GEMV_CASE(1) GEMV_CASE(2) GEMV_CASE(3) GEMV_CASE(4) GEMV_CASE(5) GEMV_CASE(6) GEMV_CASE(7) GEMV_CASE(8)
        }
#undef GEMV_CASE
        if (prof) {
            cudaEventRecord(lm->prof_ev[lm->prof_used + 1], lm->stream);
            lm->prof_used += 2;
            lm->prof_bytes += (uint64_t)a.rows * a.K * sizeof(WT) * (EPI == EPI_SWIGLU ? 2 : 1);
        }
        LAUNCH_CHECK(lm);
    }
    return FSB_OK;
}

// dynamic smem opt-in for every GEMV instantiation (done once, outside graph capture)
template <typename WT, int EPI>
static cudaError_t init_gemv_attrs_epi(int bytes) {
    cudaError_t e = cudaSuccess;
#define GEMV_ATTR(NB)                                                                                          \
    if (e == cudaSuccess && (size_t)bytes / 8 * NB > 48 * 1024)                                                \
        e = cudaFuncSetAttribute(gemv_kernel<WT, NB, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes / 8 * NB);
    GEMV_ATTR(1) GEMV_ATTR(2) GEMV_ATTR(3) GEMV_ATTR(4) GEMV_ATTR(5) GEMV_ATTR(6) GEMV_ATTR(7) GEMV_ATTR(8)
#undef GEMV_ATTR
    return e;
}
static cudaError_t init_gemv_attrs(int max_k) {
    static int max_k_set = 0;  // process-wide, monotonic (see the sampler attributes in lm_create_impl)
    if (max_k <= max_k_set) return cudaSuccess;
    max_k_set = max_k;
    const int bytes = 8 * max_k * (int)sizeof(float);
    cudaError_t e = init_gemv_attrs_epi<float, EPI_STORE>(bytes);
    if (e == cudaSuccess) e = init_gemv_attrs_epi<float, EPI_RESID>(bytes);
    if (e == cudaSuccess) e = init_gemv_attrs_epi<float, EPI_SWIGLU>(bytes);
    if (e == cudaSuccess) e = init_gemv_attrs_epi<__nv_bfloat16, EPI_STORE>(bytes);
    if (e == cudaSuccess) e = init_gemv_attrs_epi<__nv_bfloat16, EPI_RESID>(bytes);
    if (e == cudaSuccess) e = init_gemv_attrs_epi<__nv_bfloat16, EPI_SWIGLU>(bytes);
    return e;
}

template <int EPI>
static int launch_gemv(fsb_lm *lm, const DevTensor &W, const DevTensor *W3, const float *x, int ldx,
                       const float *norm_w, const float *resid, float *y, int ldy, int rows, int K, int nb,
                       const int *n_active, int row0 = 0, int rest_base = 1) {
    GemvArgs a;
    a.W = W.ptr;
    a.W3 = W3 ? W3->ptr : nullptr;
    a.x = x;
    a.norm_w = norm_w;
    a.resid = resid;
    a.y = y;
    a.rows = rows;
    a.K = K;
    a.ldx = ldx;
    a.ldy = ldy;
    a.eps = lm->cfg.norm_eps;
    a.row0 = row0;
    a.rest_base = rest_base;
    a.n_active = n_active;
    if (W.dtype == FSB_F32) return launch_gemv_nb<float, EPI>(lm, a, nb);
    return launch_gemv_nb<__nv_bfloat16, EPI>(lm, a, nb);
}

// ---------------------------------------------------------------- one decode step of a block stack
// x (nb, D) in place.  pos_ptr (device, per row) or pos_imm.
static int decode_layer(fsb_lm *lm, const LayerW &L, float *x, int nb, float *kc, float *vc, int cache_len,
                        const int *pos_ptr, int pos_imm, int rope_delta, const int *n_active, int nsplit,
                        const int *active = nullptr) {
    const int D = lm->D, H = lm->H, KV = lm->KV, hd = lm->hd, I = lm->I, QKV = lm->QKV;
    Scratch &s = lm->s;
    FSB_TRY(launch_gemv<EPI_STORE>(lm, L.wqkv, nullptr, x, D, (const float *)L.attn_norm.ptr, nullptr, s.qkv, QKV,
                                   QKV, D, nb, n_active));
    rope_append_kernel<<<nb, 256, 0, lm->stream>>>(s.qkv, s.q, kc, vc, lm->cosT, lm->sinT, pos_ptr, pos_imm,
                                                   rope_delta, H, KV, hd, cache_len, n_active, active);
    LAUNCH_CHECK(lm);
    const int n_rep = H / KV;
    attn_decode_split_kernel<<<dim3(nsplit, KV, nb), n_rep * 32, n_rep * hd * sizeof(float), lm->stream>>>(
        s.q, kc, vc, pos_ptr, pos_imm, H, KV, hd, cache_len, 1.0f / sqrtf((float)hd), s.partial, n_active, active);
    LAUNCH_CHECK(lm);
    attn_decode_combine_kernel<<<nb * H, hd, 0, lm->stream>>>(s.partial, nsplit, hd, s.att, n_active);
    LAUNCH_CHECK(lm);
    FSB_TRY(launch_gemv<EPI_RESID>(lm, L.wo, nullptr, s.att, H * hd, nullptr, x, x, D, D, H * hd, nb, n_active));
    FSB_TRY(launch_gemv<EPI_SWIGLU>(lm, L.w1, &L.w3, x, D, (const float *)L.ffn_norm.ptr, nullptr, s.g1, I, I, D, nb,
                                    n_active));
    FSB_TRY(launch_gemv<EPI_RESID>(lm, L.w2, nullptr, s.g1, I, nullptr, x, x, D, D, I, nb, n_active));
    return FSB_OK;
}

// Same block step for wide batches (nb > 8, bf16 weights): the five projections run on tcgen05 with the batch
// rows as the MMA N dimension (weights stream once per layer instead of once per 8 rows); attention, RoPE
// and the KV append are the per-op kernels above.
static int decode_layer_tc(fsb_lm *lm, const LayerW &L, float *x, int nb, float *kc, float *vc, int cache_len,
                           const int *pos_ptr, int pos_imm, int rope_delta, const int *n_active, int nsplit,
                           const int *active = nullptr) {
    const int D = lm->D, H = lm->H, KV = lm->KV, hd = lm->hd, I = lm->I, QKV = lm->QKV;
    Scratch &s = lm->s;
    cudaStream_t st = lm->stream;
    const int bn = tc_pick_bn(nb), bi = bn == 32 ? 0 : (bn == 64 ? 1 : 2);
    const int seg = lm->prefill_rows;
    const float eps = lm->cfg.norm_eps;
    FSB_TRY(tc_rmsnorm_split3(x, (const float *)L.attn_norm.ptr, eps, nb, D, lm->sp_xn, (size_t)seg * D, st));
    FSB_TRY(tc_gemm(L.m_wqkv, lm->mx_xn[bi], bn, s.qkv, nullptr, nb, QKV, D, seg, QKV, st, lm->tc_ws, lm->tc_ws_floats));
    rope_append_kernel<<<nb, 256, 0, st>>>(s.qkv, s.q, kc, vc, lm->cosT, lm->sinT, pos_ptr, pos_imm, rope_delta, H, KV, hd,
                                           cache_len, n_active, active);
    LAUNCH_CHECK(lm);
    const int n_rep = H / KV;
    attn_decode_split_kernel<<<dim3(nsplit, KV, nb), n_rep * 32, n_rep * hd * sizeof(float), st>>>(
        s.q, kc, vc, pos_ptr, pos_imm, H, KV, hd, cache_len, 1.0f / sqrtf((float)hd), s.partial, n_active, active);
    LAUNCH_CHECK(lm);
    attn_decode_combine_kernel<<<nb * H, hd, 0, st>>>(s.partial, nsplit, hd, s.att, n_active);
    LAUNCH_CHECK(lm);
    FSB_TRY(tc_split3(s.att, lm->sp_att, (size_t)nb * H * hd, (size_t)seg * H * hd, st));
    FSB_TRY(tc_gemm(L.m_wo, lm->mx_att[bi], bn, x, x, nb, D, H * hd, seg, D, st, lm->tc_ws, lm->tc_ws_floats));
    FSB_TRY(tc_rmsnorm_split3(x, (const float *)L.ffn_norm.ptr, eps, nb, D, lm->sp_xn, (size_t)seg * D, st));
    FSB_TRY(tc_gemm(L.m_w1, lm->mx_xn[bi], bn, s.g1, nullptr, nb, I, D, seg, I, st, lm->tc_ws, lm->tc_ws_floats));
    FSB_TRY(tc_gemm(L.m_w3, lm->mx_xn[bi], bn, s.g3, nullptr, nb, I, D, seg, I, st, lm->tc_ws, lm->tc_ws_floats));
    FSB_TRY(tc_swiglu_split3(s.g1, s.g3, (size_t)nb * I, lm->sp_h, (size_t)seg * I, st));
    FSB_TRY(tc_gemm(L.m_w2, lm->mx_h[bi], bn, x, x, nb, D, I, seg, D, st, lm->tc_ws, lm->tc_ws_floats));
    lm->launches += 9;
    return FSB_OK;
}

static int decode_layer_any(fsb_lm *lm, const LayerW &L, float *x, int nb, float *kc, float *vc, int cache_len,
                            const int *pos_ptr, int pos_imm, int rope_delta, const int *n_active, int nsplit,
                            const int *active) {
    if (lm->tc_ok && nb > 8)
        return decode_layer_tc(lm, L, x, nb, kc, vc, cache_len, pos_ptr, pos_imm, rope_delta, n_active, nsplit, active);
    return decode_layer(lm, L, x, nb, kc, vc, cache_len, pos_ptr, pos_imm, rope_delta, n_active, nsplit, active);
}

template <typename WT>
static void launch_embed(fsb_lm *lm, const uint32_t *toks, int nrows, int S, float *x, const int *n_active) {
    embed_sum_kernel<WT><<<nrows, 256, 0, lm->stream>>>(toks, S, lm->C, lm->D, lm->CS, (const WT *)lm->emb.ptr,
                                                       (const WT *)lm->cb_emb.ptr, lm->tok.semantic_start_id,
                                                       lm->tok.semantic_end_id, lm->tok.has_semantic_end, x,
                                                       n_active);
}
static int embed(fsb_lm *lm, const uint32_t *toks, int nrows, int S, float *x, const int *n_active) {
    if (lm->wdt == FSB_F32) launch_embed<float>(lm, toks, nrows, S, x, n_active);
    else launch_embed<__nv_bfloat16>(lm, toks, nrows, S, x, n_active);
    LAUNCH_CHECK(lm);
    return FSB_OK;
}

static float *slow_k(fsb_lm *lm, int l) { return lm->kc + (size_t)l * lm->max_batch * lm->KV * lm->max_len * lm->hd; }
static float *slow_v(fsb_lm *lm, int l) { return lm->vc + (size_t)l * lm->max_batch * lm->KV * lm->max_len * lm->hd; }
static float *fast_k(fsb_lm *lm, int l) { return lm->fkc + (size_t)l * lm->max_batch * lm->KV * lm->fast_len * lm->hd; }
static float *fast_v(fsb_lm *lm, int l) { return lm->fvc + (size_t)l * lm->max_batch * lm->KV * lm->fast_len * lm->hd; }

// ---------------------------------------------------------------- prefill of one row (S > 1)
template <typename WT, int EPI>
static void launch_gemm(fsb_lm *lm, const float *A, const DevTensor &W, const float *resid, float *Cm, int M, int N,
                        int K) {
    dim3 grid((N + kGemmBN - 1) / kGemmBN, (M + kGemmBM - 1) / kGemmBM);
    gemm_nt_kernel<WT, EPI><<<grid, 256, 0, lm->stream>>>(A, (const WT *)W.ptr, resid, Cm, M, N, K);
}
template <int EPI>
static int gemm(fsb_lm *lm, const float *A, const DevTensor &W, const float *resid, float *Cm, int M, int N, int K) {
    if (W.dtype == FSB_F32) launch_gemm<float, EPI>(lm, A, W, resid, Cm, M, N, K);
    else launch_gemm<__nv_bfloat16, EPI>(lm, A, W, resid, Cm, M, N, K);
    LAUNCH_CHECK(lm);
    return FSB_OK;
}

// One prefill pass over up to `prefill_rows` prompt positions that may belong to SEVERAL batch rows: the dense
// projections (QKV / O / FFN) see all positions as one GEMM (the tcgen05 kernel needs hundreds of tiles to fill 148 SMs;
// one 384-token prompt yields 30-96), while embedding, RoPE + KV append and causal attention run per row segment.
struct PrefillSeg {
    const uint32_t *toks;  // device (C+1, S_total) of the row
    int S_total;           // leading dimension of toks
    int c0, n;             // prompt columns [c0, c0 + n) of the row
    int b, pos0;           // batch row, cache position of column c0
    int rope_delta;        // RoPE row = cache position + rope_delta
    int off;               // first position of the segment inside the pass
    bool last;             // the segment ends the row's prompt: its last position is the row's hidden state
};

static int prefill_pass(fsb_lm *lm, const std::vector<PrefillSeg> &segs) {
    const int D = lm->D, H = lm->H, KV = lm->KV, hd = lm->hd, I = lm->I, QKV = lm->QKV;
    Scratch &s = lm->s;
    cudaStream_t st = lm->stream;
    int S = 0;
    for (const PrefillSeg &g : segs) S = std::max(S, g.off + g.n);
    for (const PrefillSeg &g : segs) {
        // the embed kernel indexes toks[c * S_total + s]; a shifted base selects the segment's columns
        if (lm->wdt == FSB_F32)
            embed_sum_kernel<float><<<g.n, 256, 0, st>>>(g.toks + g.c0, g.S_total, lm->C, D, lm->CS, (const float *)lm->emb.ptr,
                                                        (const float *)lm->cb_emb.ptr, lm->tok.semantic_start_id,
                                                        lm->tok.semantic_end_id, lm->tok.has_semantic_end,
                                                        s.x + (size_t)g.off * D, nullptr);
        else
            embed_sum_kernel<__nv_bfloat16><<<g.n, 256, 0, st>>>(
                g.toks + g.c0, g.S_total, lm->C, D, lm->CS, (const __nv_bfloat16 *)lm->emb.ptr,
                (const __nv_bfloat16 *)lm->cb_emb.ptr, lm->tok.semantic_start_id, lm->tok.semantic_end_id,
                lm->tok.has_semantic_end, s.x + (size_t)g.off * D, nullptr);
        LAUNCH_CHECK(lm);
    }
    const float scale = 1.0f / sqrtf((float)hd);
    const bool tc = lm->tc_ok;
    // Fish shapes (head_dim 64, 8 query heads per KV head): register-blocked attention kernel
    const bool pf8 = hd == 64 && H == 8 * KV && getenv("FSB_PREFILL_ATT_V1") == nullptr;
    if (pf8) {
        static bool attr_done = false;
        if (!attr_done) {
            FSB_CUDA_OK(cudaFuncSetAttribute(attn_prefill8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(kPf2SmemFloats * sizeof(float))));
            attr_done = true;
        }
    }
    const int bn = tc_pick_bn(S), bi = bn == 32 ? 0 : (bn == 64 ? 1 : 2);
    const int seg = lm->prefill_rows;
    // segment table for the batched RoPE / attention launches (rope_delta != 0 only occurs on the single-row step API)
    bool batched = segs.size() > 1 && segs.size() <= 256;
    int max_n = 0;
    for (const PrefillSeg &g : segs) {
        max_n = std::max(max_n, g.n);
        if (g.rope_delta != 0) batched = false;
    }
    if (batched) {
        std::vector<int4> h(segs.size());
        for (size_t i = 0; i < segs.size(); ++i) h[i] = make_int4(segs[i].b, segs[i].pos0, segs[i].n, segs[i].off);
        // (pageable source: the copy is staged by the runtime before the call returns)
        FSB_CUDA_OK(cudaMemcpyAsync(lm->d_segs, h.data(), h.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
    }
    for (int l = 0; l < lm->NL; ++l) {
        const LayerW &L = lm->layers[l];
        if (tc) {
            FSB_TRY(tc_rmsnorm_split3(s.x, (const float *)L.attn_norm.ptr, lm->cfg.norm_eps, S, D, lm->sp_xn, (size_t)seg * D, st));
            FSB_TRY(tc_gemm(L.m_wqkv, lm->mx_xn[bi], bn, s.qkv, nullptr, S, QKV, D, seg, QKV, st));
            lm->launches += 2;
        } else {
            rmsnorm_rows_kernel<<<(S + 3) / 4, 128, 0, st>>>(s.x, (const float *)L.attn_norm.ptr, lm->cfg.norm_eps, S, D, s.xn);
            LAUNCH_CHECK(lm);
            FSB_TRY(gemm<EPI_STORE>(lm, s.xn, L.wqkv, nullptr, s.qkv, S, QKV, D));
        }
        if (batched) {
            // one launch for every row segment of the pass (a 384-token prompt alone yields 96 CTAs of 105 us each)
            rope_append_rows_kernel<<<dim3(max_n, (unsigned)segs.size()), 256, 0, st>>>(
                s.qkv, s.q, slow_k(lm, l), slow_v(lm, l), lm->cosT, lm->sinT, 0, 0, 0, H, KV, hd, lm->max_len, lm->d_segs);
            LAUNCH_CHECK(lm);
            if (pf8)
                attn_prefill8_kernel<<<dim3((max_n + kPf2Q - 1) / kPf2Q, KV, (unsigned)segs.size()), kPf2Threads,
                                       kPf2SmemFloats * sizeof(float), st>>>(s.q, slow_k(lm, l), slow_v(lm, l), 0, 0, 0, H, KV,
                                                                             lm->max_len, s.att, lm->d_segs);
            else
                attn_prefill_kernel<<<dim3((max_n + kPrefQ - 1) / kPrefQ, KV, (unsigned)segs.size()), 512, 0, st>>>(
                    s.q, slow_k(lm, l), slow_v(lm, l), 0, 0, 0, H, KV, hd, lm->max_len, scale, s.att, lm->d_segs);
            LAUNCH_CHECK(lm);
        } else {
            for (const PrefillSeg &g : segs) {
                rope_append_rows_kernel<<<g.n, 256, 0, st>>>(s.qkv + (size_t)g.off * QKV, s.q + (size_t)g.off * H * hd,
                                                             slow_k(lm, l), slow_v(lm, l), lm->cosT, lm->sinT, g.b, g.pos0,
                                                             g.rope_delta, H, KV, hd, lm->max_len);
                LAUNCH_CHECK(lm);
                if (pf8)
                    attn_prefill8_kernel<<<dim3((g.n + kPf2Q - 1) / kPf2Q, KV), kPf2Threads, kPf2SmemFloats * sizeof(float), st>>>(
                        s.q + (size_t)g.off * H * hd, slow_k(lm, l), slow_v(lm, l), g.b, g.pos0, g.n, H, KV, lm->max_len,
                        s.att + (size_t)g.off * H * hd);
                else
                    attn_prefill_kernel<<<dim3((g.n + kPrefQ - 1) / kPrefQ, KV), 512, 0, st>>>(
                        s.q + (size_t)g.off * H * hd, slow_k(lm, l), slow_v(lm, l), g.b, g.pos0, g.n, H, KV, hd, lm->max_len, scale,
                        s.att + (size_t)g.off * H * hd);
                LAUNCH_CHECK(lm);
            }
        }
        if (tc) {
            FSB_TRY(tc_split3(s.att, lm->sp_att, (size_t)S * H * hd, (size_t)seg * H * hd, st));
            FSB_TRY(tc_gemm(L.m_wo, lm->mx_att[bi], bn, s.x, s.x, S, D, H * hd, seg, D, st));
            lm->launches += 2;
        } else {
            FSB_TRY(gemm<EPI_RESID>(lm, s.att, L.wo, s.x, s.x, S, D, H * hd));
        }
        const size_t n = (size_t)S * I;
        if (tc) {
            FSB_TRY(tc_rmsnorm_split3(s.x, (const float *)L.ffn_norm.ptr, lm->cfg.norm_eps, S, D, lm->sp_xn, (size_t)seg * D, st));
            FSB_TRY(tc_gemm(L.m_w1, lm->mx_xn[bi], bn, s.g1, nullptr, S, I, D, seg, I, st));
            FSB_TRY(tc_gemm(L.m_w3, lm->mx_xn[bi], bn, s.g3, nullptr, S, I, D, seg, I, st));
            FSB_TRY(tc_swiglu_split3(s.g1, s.g3, n, lm->sp_h, (size_t)seg * I, st));
            FSB_TRY(tc_gemm(L.m_w2, lm->mx_h[bi], bn, s.x, s.x, S, D, I, seg, D, st));
            lm->launches += 5;
        } else {
            rmsnorm_rows_kernel<<<(S + 3) / 4, 128, 0, st>>>(s.x, (const float *)L.ffn_norm.ptr, lm->cfg.norm_eps, S, D, s.xn);
            LAUNCH_CHECK(lm);
            FSB_TRY(gemm<EPI_STORE>(lm, s.xn, L.w1, nullptr, s.g1, S, I, D));
            FSB_TRY(gemm<EPI_STORE>(lm, s.xn, L.w3, nullptr, s.g3, S, I, D));
            swiglu_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s.g1, s.g3, n, s.g1);
            LAUNCH_CHECK(lm);
            FSB_TRY(gemm<EPI_RESID>(lm, s.g1, L.w2, s.x, s.x, S, D, I));
        }
    }
    for (const PrefillSeg &g : segs)
        if (g.last)
            FSB_CUDA_OK(cudaMemcpyAsync(s.hidden + (size_t)g.b * D, s.x + (size_t)(g.off + g.n - 1) * D, D * sizeof(float),
                                        cudaMemcpyDeviceToDevice, st));
    return FSB_OK;
}

// Prefill `S` prompt columns of row b (chunked to the scratch size); leaves the
// pre-norm last-position state in s.hidden[b].
static int prefill_row(fsb_lm *lm, const uint32_t *toks_dev, int S, int b, int pos0, int rope_delta) {
    for (int c0 = 0; c0 < S; c0 += lm->prefill_rows) {
        const int n = std::min(lm->prefill_rows, S - c0);
        std::vector<PrefillSeg> segs(1);
        segs[0] = PrefillSeg{toks_dev, S, c0, n, b, pos0 + c0, rope_delta, 0, c0 + n == S};
        FSB_TRY(prefill_pass(lm, segs));
    }
    return FSB_OK;
}

// Prefill of a whole batch: the rows' prompts are packed back to back into passes of up to `prefill_rows` positions.
// toks_dev: the rows' (C+1, P_b) arrays back to back.
static int prefill_batch(fsb_lm *lm, const uint32_t *toks_dev, const int32_t *prompt_lens, int bsz) {
    std::vector<PrefillSeg> segs;
    int used = 0;
    size_t tok_off = 0;
    for (int b = 0; b < bsz; ++b) {
        const int P = prompt_lens[b];
        for (int c0 = 0; c0 < P;) {
            if (used == lm->prefill_rows) {
                FSB_TRY(prefill_pass(lm, segs));
                segs.clear();
                used = 0;
            }
            const int n = std::min(lm->prefill_rows - used, P - c0);
            segs.push_back(PrefillSeg{toks_dev + tok_off, P, c0, n, b, lm->kv_len[b] + c0, 0, used, c0 + n == P});
            used += n;
            c0 += n;
        }
        tok_off += (size_t)(lm->C + 1) * P;
    }
    if (!segs.empty()) FSB_TRY(prefill_pass(lm, segs));
    return FSB_OK;
}

// norm + constrained output head on s.hidden -> s.slow_logits (generate/utils.rs:6-33)
static int slow_head(fsb_lm *lm, int nb, const int *n_active) {
    return launch_gemv<EPI_STORE>(lm, lm->out_w, nullptr, lm->s.hidden, lm->D, (const float *)lm->norm.ptr, nullptr,
                                  lm->s.slow_logits, lm->n_slow_logits, lm->n_slow_logits, lm->D, nb, n_active,
                                  lm->slow_row0, lm->slow_rest_base);
}

static size_t sample_smem(int n) {
    int n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    return (size_t)2 * std::max(n_pad, kSampleThreads) * 8 + (size_t)n_pad * 4 + 64 * 4 * 2;
}

// slow sample + C fast steps + frame bookkeeping (single_batch.rs:126-204)
static int frame_tail(fsb_lm *lm, int nb) {
    Scratch &s = lm->s;
    const int *na = lm->h_st.n_active;
    if (lm->collect_hidden) {
        store_hidden_kernel<<<nb, 256, 0, lm->stream>>>(s.hidden, lm->d_st, lm->hid, lm->D);
        LAUNCH_CHECK(lm);
    }
    FSB_TRY(slow_head(lm, nb, na));
    sample_slow_kernel<<<nb, kSampleThreads, sample_smem(lm->n_slow_logits), lm->stream>>>(
        s.slow_logits, lm->n_slow_logits, lm->n_slow_logits, lm->d_st, lm->tok.semantic_start_id, s.hidden, s.fast_x,
        lm->D);
    LAUNCH_CHECK(lm);
    for (int cb = 0; cb < lm->C; ++cb) {
        for (int l = 0; l < lm->NFL; ++l)
            FSB_TRY(decode_layer_any(lm, lm->fast_layers[l], s.fast_x, nb, fast_k(lm, l), fast_v(lm, l), lm->fast_len,
                                     nullptr, cb, 0, na, 1, nullptr));
        FSB_TRY(launch_gemv<EPI_STORE>(lm, lm->fast_out, nullptr, s.fast_x, lm->D, (const float *)lm->fast_norm.ptr,
                                       nullptr, s.fast_logits, lm->CS, lm->CS, lm->D, nb, na));
        if (lm->wdt == FSB_F32)
            sample_fast_kernel<float><<<nb, kSampleThreads, sample_smem(lm->CS), lm->stream>>>(
                s.fast_logits, lm->CS, cb, lm->d_st, (const float *)lm->fast_emb.ptr, s.fast_x, lm->D);
        else
            sample_fast_kernel<__nv_bfloat16><<<nb, kSampleThreads, sample_smem(lm->CS), lm->stream>>>(
                s.fast_logits, lm->CS, cb, lm->d_st, (const __nv_bfloat16 *)lm->fast_emb.ptr, s.fast_x, lm->D);
        LAUNCH_CHECK(lm);
    }
    return FSB_OK;
}

// one decode frame for nb rows: slow step on `prev` codes + tail
static int decode_frame(fsb_lm *lm, int nb) {
    Scratch &s = lm->s;
    const int *na = lm->h_st.n_active;
    FSB_TRY(embed(lm, lm->h_st.prev, nb, 1, s.hidden, na));
    for (int l = 0; l < lm->NL; ++l)
        FSB_TRY(decode_layer_any(lm, lm->layers[l], s.hidden, nb, slow_k(lm, l), slow_v(lm, l), lm->max_len,
                                 lm->h_st.pos, 0, 0, na, lm->nsplit, lm->h_st.active));
    return frame_tail(lm, nb);
}

static int get_graph(fsb_lm *lm, std::map<int, cudaGraphExec_t> &cache, int nb, bool with_slow,
                     cudaGraphExec_t *out) {
    const int key = nb + (lm->collect_hidden ? 4096 : 0);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return FSB_OK;
    }
    FSB_CUDA_OK(cudaStreamBeginCapture(lm->stream, cudaStreamCaptureModeThreadLocal));
    const uint64_t l0 = lm->launches;
    int st = with_slow ? decode_frame(lm, nb) : frame_tail(lm, nb);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(lm->stream, &g);
    const uint64_t per_graph = lm->launches - l0;
    lm->launches = l0;
    if (st != FSB_OK) {
        if (g) cudaGraphDestroy(g);
        return st;
    }
    FSB_CUDA_OK(e);
    cudaGraphExec_t ge = nullptr;
    FSB_CUDA_OK(cudaGraphInstantiate(&ge, g, 0));
    cudaGraphDestroy(g);
    cache[key] = ge;
    // remember how many kernels one replay launches
    cache[-key - 1] = reinterpret_cast<cudaGraphExec_t>((uintptr_t)per_graph);
    *out = ge;
    return FSB_OK;
}
static uint64_t graph_launches(fsb_lm *lm, std::map<int, cudaGraphExec_t> &cache, int nb) {
    return (uint64_t)(uintptr_t)cache[-(nb + (lm->collect_hidden ? 4096 : 0)) - 1];
}

// ---------------------------------------------------------------- megakernel host side
static int mega_nb_template(int nb) { return nb <= 1 ? 1 : nb <= 2 ? 2 : nb <= 4 ? 4 : 8; }

static size_t mega_smem_bytes(const fsb_lm *lm, int NB, int *xs_floats, int *val_floats) {
    const int G = lm->mega_grid;
    const int NE = lm->wdt == FSB_F32 ? 4 : 8;
    const int maxK = std::max(std::max(lm->D, lm->I), lm->H * lm->hd);
    int n_pad = 1;
    while (n_pad < std::max(lm->n_slow_logits, lm->CS)) n_pad <<= 1;
    const int samp_floats = (2 * std::max(n_pad, kMegaThreads) * 8 + n_pad * 4 + 1024) / 4;
    *xs_floats = (std::max(NB * maxK, samp_floats) + 3) & ~3;
    auto tasks = [&](int rows, int K, int mats) {
        const int upr = K / (32 * NE), ksplit = (upr + kMegaPre - 1) / kMegaPre;
        return ((rows + G - 1) / G + 2) * ksplit * mats;
    };
    int nt = tasks(lm->QKV, lm->D, 1);
    nt = std::max(nt, tasks(lm->D, lm->H * lm->hd, 1));
    nt = std::max(nt, tasks(lm->I, lm->D, 2));
    nt = std::max(nt, tasks(lm->D, lm->I, 1));
    nt = std::max(nt, tasks(lm->n_slow_logits, lm->D, 1));
    nt = std::max(nt, tasks(lm->CS, lm->D, 1));
    *val_floats = (nt * NB + 3) & ~3;
    // red | rope rows (slow, fast) | positions | sampler state (flags, cur/prev, rep-pen windows)
    const size_t aux = 64 * NB + NB * 64 + 8 * 64 + 4 * ((NB + 3) / 4) + 4 * NB + 2 * 20 * NB +
                       (sizeof(RepPenState) / 4) * NB * 8 + 8;
    return ((size_t)*xs_floats + (size_t)*val_floats + aux + 2ull * kMegaChunk * kMegaKvStride) * sizeof(float);
}

// ---- single-row kernel (fsb_lm_mega1.cuh): shared-memory carve-up mirrored from Mega1's constructor
static bool mega1_eligible(const fsb_lm *lm) {
    if (getenv("FSB_MEGA_V1")) return false;  // A/B switch: keep the register-prefetch kernel for one row too
    const int Hhd = lm->H * lm->hd;
    const SampleParams &sp = lm->h_st.sp;  // the single-row kernel carries the selection sampler only
    if (!sp.greedy && (sp.top_k > (uint32_t)kSelMaxK || std::max(lm->n_slow_logits, lm->CS) > (1 << kSelIdxBits))) return false;
    return lm->mega_ok && lm->D == kM1Slice && lm->I == 4 * kM1Slice && Hhd == kM1Slice && lm->hd == 64 && lm->KV == 2 && lm->QKV % 2 == 0 &&
           lm->C == 8 && lm->fast_len == 8 && lm->CS == 1024;
}
static size_t mega1_smem_bytes(const fsb_lm *lm, int depth, int *xs_floats, int *kvs_floats) {
    const int n_max = std::max(lm->n_slow_logits, lm->CS);
    const int samp_floats = (sel_scratch_bytes(kM1Threads) + 15) / 16 * 4 + ((n_max + 3) & ~3) + 64;
    *xs_floats = (std::max(lm->I, 2 * lm->H * lm->hd) + 3) & ~3;
    *kvs_floats = (std::max(2 * kM1AttChunk * kM1KvStride, samp_floats) + 3) & ~3;
    const size_t fl = (size_t)*xs_floats + lm->D + kM1ValFloats + 64 + 64 + 8 * 64 + *kvs_floats + 4 + 12 +
                      (sizeof(MegaLayer) / 4) * (lm->NL + lm->NFL) + 4 + 20 + 20 +
                      (sizeof(RepPenState) / 4) * 8 + 4 * kM1MaxDepth + 2;
    return (size_t)depth * kM1ChunkBytes + fl * sizeof(float) + 16;
}

// rows [row0, row0 + nb) of the current batch: every per-row buffer is addressed relative to row0
static int mega_launch(fsb_lm *lm, int row0, int group_index, int nb, int nframes, bool first_is_tail) {
    MegaParams mp = lm->mp;
    mp.nb = nb;
    mp.nframes = nframes;
    mp.first_is_tail = first_is_tail ? 1 : 0;
    mp.row0 = row0;
    const int NB = mega_nb_template(nb);
    int xs_floats = 0, val_floats = 0;
    const size_t smem = mega_smem_bytes(lm, NB, &xs_floats, &val_floats);
    mp.xs_floats = xs_floats;
    mp.val_floats = val_floats;
    mp.st = lm->h_st;
    const int C1 = lm->C + 1;
    const size_t r = (size_t)row0;
    mp.x += r * lm->D; mp.fx += r * lm->D; mp.q += r * lm->H * lm->hd; mp.h += r * lm->I;
    mp.partial += r * lm->H * 2 * mp.n_chunks_max * (lm->hd + 4);
    mp.logits += r * mp.ldl;
    mp.kc += r * lm->KV * lm->max_len * lm->hd; mp.vc += r * lm->KV * lm->max_len * lm->hd;
    mp.fkc += r * lm->KV * lm->fast_len * lm->hd; mp.fvc += r * lm->KV * lm->fast_len * lm->hd;
    mp.st.pos += r; mp.st.active += r; mp.st.eos += r; mp.st.frame += r; mp.st.max_frames += r;
    mp.st.cur += r * C1; mp.st.prev += r * C1; mp.st.out += r * mp.st.out_cap * C1; mp.st.rep += r * lm->C;
    mp.st.n_active += group_index;  // one live-row counter per group
    FSB_CUDA_OK(cudaMemsetAsync(lm->mega_bar, 0, 4 * sizeof(unsigned int), lm->stream));
    if (nb == 1 && mega1_eligible(lm)) {
        int depth = kM1MaxDepth, xf = 0, kf = 0;
        while (depth > 2 && mega1_smem_bytes(lm, depth, &xf, &kf) > lm->smem_optin) --depth;
        const size_t smem1 = mega1_smem_bytes(lm, depth, &xf, &kf);
        if (smem1 <= lm->smem_optin) {
            mp.ring_depth = depth;
            mp.xs_floats = xf;
            mp.kvs_floats = kf;
            mp.sampler_cta = getenv("FSB_MEGA_SAMPLER_CTA") ? lm->mega_grid - 1 : -1;
            // flag-in-data synchronisation instead of grid barriers: opt-in experiment (FSB_MEGA_LL=1).  Measured on
            // B200 it only ties the barrier version: the tag doubles the bytes every CTA pulls through the few L2
            // slices that hold a vector, which is exactly where the post-barrier reload is bound
            const bool ll = getenv("FSB_MEGA_LL") && first_is_tail && mp.sampler_cta < 0 && lm->NL >= 1 && lm->NL < 31 &&
                            lm->NFL >= 1 && lm->NFL < 31 && lm->C + 2 <= 16 && lm->mega_grid <= kM1FlagStride;
            mp.ll = ll ? lm->mega_ll : nullptr;
            if (ll) {
                const LLLayout lay(lm->D, lm->I, lm->H * lm->hd, lm->KV, lm->hd, lm->NFL, lm->fast_len, lm->H, 2 * mp.n_chunks_max, mp.ldl);
                unsigned long long *b = lm->mega_ll;
                mp.ll_xt = b + lay.xt; mp.ll_ht = b + lay.ht; mp.ll_qt = b + lay.qt; mp.ll_nkv = b + lay.nkv; mp.ll_fkv = b + lay.fkv;
                mp.ll_pt = b + lay.pt; mp.ll_lt = b + lay.lt; mp.ll_ct = b + lay.ct; mp.ll_fl = b + lay.fl;
            }
            if (ll) FSB_CUDA_OK(cudaMemsetAsync(lm->mega_ll, 0, lm->mega_ll_words * sizeof(unsigned long long), lm->stream));
            FSB_CUDA_OK(lm->wdt == FSB_F32 ? mega1_launch_f32(mp, lm->mega_grid, smem1, lm->stream)
                                           : mega1_launch_bf16(mp, lm->mega_grid, smem1, lm->stream));
            lm->launches++;
            return FSB_OK;
        }
    }
    FSB_CUDA_OK(lm->wdt == FSB_F32 ? mega_launch_f32(NB, mp, lm->mega_grid, smem, lm->stream)
                                   : mega_launch_bf16(NB, mp, lm->mega_grid, smem, lm->stream));
    lm->launches++;
    return FSB_OK;
}

static int mega_setup(fsb_lm *lm) {
    cudaDeviceProp prop;
    FSB_CUDA_OK(cudaGetDeviceProperties(&prop, lm->opt.device));
    lm->mega_grid = prop.multiProcessorCount;
    lm->smem_optin = (size_t)prop.sharedMemPerBlockOptin;
    int coop = 0;
    FSB_CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, lm->opt.device));
    auto split_ok = [&](int K) {
        const int upr = K / (32 * (lm->wdt == FSB_F32 ? 4 : 8)), ksplit = (upr + kMegaPre - 1) / kMegaPre;
        return upr >= 1 && upr % ksplit == 0;
    };
    const int upr_ok = (lm->D % 256 == 0) && (lm->I % 256 == 0) && (lm->hd == 64) && (lm->QKV % 2 == 0) &&
                       split_ok(lm->D) && split_ok(lm->I) && split_ok(lm->H * lm->hd);
    int xs_floats, val_floats;
    const size_t smem8 = mega_smem_bytes(lm, 1, &xs_floats, &val_floats);  // per-batch check happens at launch
    lm->mega_ok = coop && upr_ok && lm->C <= 8 && std::max(lm->D, lm->I) >= 2 * lm->H * lm->hd && smem8 <= (size_t)prop.sharedMemPerBlockOptin && lm->H / lm->KV <= 8;
    if (!lm->mega_ok) return FSB_OK;
    std::vector<MegaLayer> hs(lm->NL), hf(lm->NFL);
    auto fill = [](const LayerW &L, MegaLayer *m) {
        m->wqkv = L.wqkv.ptr; m->wo = L.wo.ptr; m->w1 = L.w1.ptr; m->w3 = L.w3.ptr; m->w2 = L.w2.ptr;
        m->attn_norm = (const float *)L.attn_norm.ptr; m->ffn_norm = (const float *)L.ffn_norm.ptr;
    };
    for (int l = 0; l < lm->NL; ++l) fill(lm->layers[l], &hs[l]);
    for (int l = 0; l < lm->NFL; ++l) fill(lm->fast_layers[l], &hf[l]);
    FSB_TRY(dev_alloc(lm, &lm->d_mega_slow, lm->NL));
    FSB_TRY(dev_alloc(lm, &lm->d_mega_fast, lm->NFL));
    FSB_CUDA_OK(cudaMemcpy(lm->d_mega_slow, hs.data(), hs.size() * sizeof(MegaLayer), cudaMemcpyHostToDevice));
    FSB_CUDA_OK(cudaMemcpy(lm->d_mega_fast, hf.data(), hf.size() * sizeof(MegaLayer), cudaMemcpyHostToDevice));
    const int B = lm->max_batch;
    const int ldl = std::max(lm->n_slow_logits, lm->CS);
    const int n_chunks_max = (lm->max_len + kMegaChunk - 1) / kMegaChunk;
    FSB_TRY(dev_alloc(lm, &lm->mega_partial, (size_t)B * lm->H * 2 * n_chunks_max * (lm->hd + 4)));
    FSB_TRY(dev_alloc(lm, &lm->mega_logits, (size_t)B * ldl));
    FSB_TRY(dev_alloc(lm, &lm->mega_rep, (size_t)(kM1Rep - 1) * (2 * lm->D + lm->I)));
    {
        const LLLayout lay(lm->D, lm->I, lm->H * lm->hd, lm->KV, lm->hd, lm->NFL, lm->fast_len, lm->H, 2 * n_chunks_max, ldl);
        lm->mega_ll_words = lay.total;
        FSB_TRY(dev_alloc(lm, &lm->mega_ll, lay.total));
    }
    FSB_TRY(dev_alloc(lm, &lm->mega_bar, 4));  // [0] grid barrier, [1] frames confirmed (single-row kernel)
    if (getenv("FSB_MEGA_TIMERS")) {
        FSB_TRY(dev_alloc(lm, &lm->mega_dbg, 256));
        FSB_CUDA_OK(cudaMemset(lm->mega_dbg, 0, 256 * sizeof(unsigned long long)));
    }
    MegaParams &m = lm->mp;
    memset(&m, 0, sizeof(m));
    m.slow = lm->d_mega_slow; m.fast = lm->d_mega_fast;
    m.NL = lm->NL; m.NFL = lm->NFL; m.D = lm->D; m.I = lm->I; m.H = lm->H; m.KV = lm->KV; m.hd = lm->hd;
    m.QKV = lm->QKV; m.C = lm->C; m.CS = lm->CS;
    m.n_slow_logits = lm->n_slow_logits; m.slow_row0 = lm->slow_row0; m.slow_rest_base = lm->slow_rest_base;
    m.emb = lm->emb.ptr; m.cb_emb = lm->cb_emb.ptr; m.out_w = lm->out_w.ptr; m.fast_emb = lm->fast_emb.ptr;
    m.fast_out = lm->fast_out.ptr;
    m.norm = (const float *)lm->norm.ptr; m.fast_norm = (const float *)lm->fast_norm.ptr;
    m.eps = lm->cfg.norm_eps;
    m.kc = lm->kc; m.vc = lm->vc; m.fkc = lm->fkc; m.fvc = lm->fvc;
    m.max_len = lm->max_len; m.fast_len = lm->fast_len; m.max_batch = lm->max_batch;
    m.cosT = lm->cosT; m.sinT = lm->sinT;
    m.x = lm->s.hidden; m.fx = lm->s.fast_x; m.q = lm->s.q; m.partial = lm->mega_partial; m.h = lm->s.g1;
    m.logits = lm->mega_logits; m.ldl = ldl;
    m.n_chunks_max = n_chunks_max;
    m.bar = lm->mega_bar; m.dbg = lm->mega_dbg; m.rep = lm->mega_rep;
    m.slow_kv_stride = (size_t)lm->max_batch * lm->KV * lm->max_len * lm->hd;
    m.fast_kv_stride = (size_t)lm->max_batch * lm->KV * lm->fast_len * lm->hd;
    m.sem_start = lm->tok.semantic_start_id; m.sem_end = lm->tok.semantic_end_id; m.has_end = lm->tok.has_semantic_end;
    FSB_CUDA_OK(cudaEventCreate(&lm->prof_m0));
    FSB_CUDA_OK(cudaEventCreate(&lm->prof_m1));
    return FSB_OK;
}

// tcgen05 prefill: bf16 weights only (fp32 weights keep the CUDA-core GEMM: parity mode)
static int tc_setup(fsb_lm *lm) {
    lm->tc_ok = false;
    if (lm->wdt != FSB_BF16 || getenv("FSB_NO_TCGEN05")) return FSB_OK;
    const int D = lm->D, I = lm->I, Hhd = lm->H * lm->hd, QKV = lm->QKV, M = lm->prefill_rows;
    if (D % 64 || I % 64 || Hhd % 64 || QKV % 128 || D % 128 || I % 128) return FSB_OK;
    FSB_TRY(dev_alloc(lm, &lm->sp_xn, (size_t)3 * M * D));
    FSB_TRY(dev_alloc(lm, &lm->sp_att, (size_t)3 * M * Hhd));
    FSB_TRY(dev_alloc(lm, &lm->sp_h, (size_t)3 * M * I));
    lm->tc_ws_floats = (size_t)16 * 64 * std::max(std::max(I, QKV), D);  // up to 16 splits of a 64-row batch
    FSB_TRY(dev_alloc(lm, &lm->tc_ws, lm->tc_ws_floats));
    FSB_CUDA_OK(cudaMemset(lm->sp_xn, 0, (size_t)3 * M * D * 2));
    FSB_CUDA_OK(cudaMemset(lm->sp_att, 0, (size_t)3 * M * Hhd * 2));
    FSB_CUDA_OK(cudaMemset(lm->sp_h, 0, (size_t)3 * M * I * 2));
    const int bns[3] = {32, 64, 128};
    for (int i = 0; i < 3; ++i) {
        FSB_TRY(tc_make_map_bf16(&lm->mx_xn[i], lm->sp_xn, 3 * M, D, bns[i]));
        FSB_TRY(tc_make_map_bf16(&lm->mx_att[i], lm->sp_att, 3 * M, Hhd, bns[i]));
        FSB_TRY(tc_make_map_bf16(&lm->mx_h[i], lm->sp_h, 3 * M, I, bns[i]));
    }
    for (auto *stack : {&lm->layers, &lm->fast_layers})
        for (auto &L : *stack) {
            FSB_TRY(tc_make_map_bf16(&L.m_wqkv, L.wqkv.ptr, QKV, D, 128));
            FSB_TRY(tc_make_map_bf16(&L.m_wo, L.wo.ptr, D, Hhd, 128));
            FSB_TRY(tc_make_map_bf16(&L.m_w1, L.w1.ptr, I, D, 128));
            FSB_TRY(tc_make_map_bf16(&L.m_w3, L.w3.ptr, I, D, 128));
            FSB_TRY(tc_make_map_bf16(&L.m_w2, L.w2.ptr, D, I, 128));
        }
    FSB_TRY(tc_init());
    lm->tc_ok = true;
    return FSB_OK;
}


// ---- wide-batch kernel (fsb_lm_megab.cuh): tensor maps of every weight matrix + scratch
static bool megab_shapes_ok(const fsb_lm *lm) {
    return lm->mega_ok && lm->tc_ok && lm->wdt == FSB_BF16 && lm->D == 1024 && lm->I == 4096 && lm->H * lm->hd == 1024 &&
           lm->hd == 64 && lm->KV == 2 && lm->C == 8 && lm->fast_len == 8 && lm->CS == 1024 && lm->QKV == 1280 &&
           lm->n_slow_logits <= (1 << kSelIdxBits) && lm->NL >= 1 && lm->NFL >= 1 && lm->mega_grid >= 128;
}
static int megab_setup(fsb_lm *lm) {
    lm->megab_ok = false;
    memset(&lm->mbx, 0, sizeof(lm->mbx));
    if (!megab_shapes_ok(lm) || lm->max_batch < 2 || getenv("FSB_NO_MEGAB")) return FSB_OK;
    const int nl = lm->NL + lm->NFL;
    std::vector<TcMap> maps((size_t)5 * nl + 2);
    for (int i = 0; i < nl; ++i) {
        const LayerW &L = i < lm->NL ? lm->layers[i] : lm->fast_layers[i - lm->NL];
        FSB_TRY(tc_make_map_bf16(&maps[5 * i + 0], L.wqkv.ptr, lm->QKV, lm->D, 128));
        FSB_TRY(tc_make_map_bf16(&maps[5 * i + 1], L.wo.ptr, lm->D, lm->H * lm->hd, 128));
        FSB_TRY(tc_make_map_bf16(&maps[5 * i + 2], L.w1.ptr, lm->I, lm->D, 64));  // a W13 tile = 64 rows of w1 | 64 rows of w3
        FSB_TRY(tc_make_map_bf16(&maps[5 * i + 3], L.w3.ptr, lm->I, lm->D, 64));
        FSB_TRY(tc_make_map_bf16(&maps[5 * i + 4], L.w2.ptr, lm->D, lm->I, 128));
    }
    FSB_TRY(tc_make_map_bf16(&maps[5 * nl + 0], lm->out_w.ptr, lm->V, lm->D, 128));
    FSB_TRY(tc_make_map_bf16(&maps[5 * nl + 1], lm->fast_out.ptr, lm->CS, lm->D, 128));
    TcMap *d_maps = nullptr;
    FSB_TRY(dev_alloc(lm, &d_maps, maps.size()));
    FSB_CUDA_OK(cudaMemcpy(d_maps, maps.data(), maps.size() * sizeof(TcMap), cudaMemcpyHostToDevice));
    MegaBExtra &x = lm->mbx;
    x.maps = reinterpret_cast<const CUtensorMap *>(d_maps);
    const int B = std::min(lm->max_batch, 32);
    lm->megab_rows_cap = B;
    const int npad = B <= 16 ? 16 : 32;
    FSB_TRY(dev_alloc(lm, &x.ws, (size_t)64 * npad * lm->D));  // fused FFN: 64 column blocks x NPAD rows x dim
    FSB_TRY(dev_alloc(lm, &x.cnt, (size_t)2 * 32 + 4));
    x.att_cnt = x.cnt;
    x.go = x.att_cnt + 2 * 32;
    FSB_TRY(dev_alloc(lm, &x.att, (size_t)B * lm->D));
    FSB_TRY(dev_alloc(lm, &x.apart, (size_t)B * lm->H * kMBMaxSplit * (lm->hd + 4)));
    FSB_TRY(dev_alloc(lm, &x.ssq_x, (size_t)B * kMBSsq));
    FSB_TRY(dev_alloc(lm, &x.ssq_fx, (size_t)B * kMBSsq));
    {
        // operand images (sized for the 32-row tile; rows >= the batch stay zero)
        const size_t xop_bytes = (size_t)16 * 3 * 32 * 128;
        for (unsigned char **q : {&x.xop_x, &x.xop_fx, &x.xop_att}) {
            FSB_TRY(dev_alloc(lm, q, xop_bytes));
            FSB_CUDA_OK(cudaMemset(*q, 0, xop_bytes));
        }
    }
    x.head_tiles = (lm->n_slow_logits + 127) / 128;
    x.head_extra = lm->slow_row0 != lm->slow_rest_base - 1 ? 1 : 0;
    if (megab_smem_bytes(32, megab_max_stages(32)) > lm->smem_optin || megab_smem_bytes(16, megab_max_stages(16)) > lm->smem_optin)
        return FSB_OK;
    x.nstages = 0;  // per launch (depends on the tile width)
    lm->megab_ok = true;
    return FSB_OK;
}

// all rows of the batch in one launch (bsz <= 32)
static int megab_launch_rows(fsb_lm *lm, int nb, int nframes, bool first_is_tail = true) {
    MegaParams mp = lm->mp;
    mp.nb = nb;
    mp.nframes = nframes;
    mp.first_is_tail = first_is_tail ? 1 : 0;
    mp.row0 = 0;
    mp.st = lm->h_st;
    const size_t cnt_words = (size_t)2 * 32 + 4;
    FSB_CUDA_OK(cudaMemsetAsync(lm->mega_bar, 0, 4 * sizeof(unsigned int), lm->stream));
    FSB_CUDA_OK(cudaMemsetAsync(lm->mbx.cnt, 0, cnt_words * sizeof(unsigned), lm->stream));
    const int npad = nb <= 16 ? 16 : 32;  // the workspace is sized for the capacity; the tile width follows the batch
    MegaBExtra ex = lm->mbx;
    ex.nstages = megab_max_stages(npad);
    if (const char *v = getenv("FSB_MEGAB_STAGES")) ex.nstages = std::max(2, std::min(ex.nstages, atoi(v)));
    FSB_CUDA_OK(megab_launch(mp, ex, npad, lm->mega_grid, megab_smem_bytes(npad, ex.nstages), lm->stream));
    lm->launches++;
    return FSB_OK;
}

static void precompute_freqs(const fsb_model_args &c, int max_len, std::vector<float> *cosv, std::vector<float> *sinv) {
    // dual_ar.rs:168-186: theta_i = 1 / base^(i/n) in f32, idx_theta = pos * theta (f32 product), cos/sin.
    // Transcendentals in f64 rounded once to f32 (== correctly rounded f32), as in oracle/dual_ar.py.
    const int n_elem = c.dim / c.n_head;
    const int half = n_elem / 2;
    std::vector<float> theta(half);
    for (int i = 0; i < half; ++i) {
        const float expo = (float)(2 * i) / (float)n_elem;
        const float p = (float)std::pow((double)c.rope_base, (double)expo);
        theta[i] = 1.0f / p;
    }
    cosv->resize((size_t)max_len * half);
    sinv->resize((size_t)max_len * half);
    for (int pos = 0; pos < max_len; ++pos)
        for (int i = 0; i < half; ++i) {
            const float a = (float)pos * theta[i];
            (*cosv)[(size_t)pos * half + i] = (float)std::cos((double)a);
            (*sinv)[(size_t)pos * half + i] = (float)std::sin((double)a);
        }
}

static int load_block(fsb_lm *lm, const fsb_tensor *w, size_t n, const std::string &p, LayerW *L) {
    const int D = lm->D, I = lm->I, QKV = lm->QKV, wdt = lm->wdt;
    cudaStream_t st = lm->stream;
    FSB_TRY(upload_tensor(w, n, p + "attention.wqkv.weight", {QKV, D}, wdt, st, &L->wqkv, &lm->owned));
    FSB_TRY(upload_tensor(w, n, p + "attention.wo.weight", {D, lm->H * lm->hd}, wdt, st, &L->wo, &lm->owned));
    FSB_TRY(upload_tensor(w, n, p + "feed_forward.w1.weight", {I, D}, wdt, st, &L->w1, &lm->owned));
    FSB_TRY(upload_tensor(w, n, p + "feed_forward.w2.weight", {D, I}, wdt, st, &L->w2, &lm->owned));
    FSB_TRY(upload_tensor(w, n, p + "feed_forward.w3.weight", {I, D}, wdt, st, &L->w3, &lm->owned));
    FSB_TRY(upload_tensor(w, n, p + "ffn_norm.weight", {D}, FSB_F32, st, &L->ffn_norm, &lm->owned));
    FSB_TRY(upload_tensor(w, n, p + "attention_norm.weight", {D}, FSB_F32, st, &L->attn_norm, &lm->owned));
    return FSB_OK;
}

static int lm_create_impl(fsb_lm *lm, const fsb_tensor *w, size_t n) {
    const fsb_model_args &c = lm->cfg;
    FSB_TRY(select_device(lm->opt.device));
    if (lm->opt.stream) {
        lm->stream = (cudaStream_t)lm->opt.stream;
    } else {
        FSB_CUDA_OK(cudaStreamCreateWithFlags(&lm->stream, cudaStreamNonBlocking));
        lm->own_stream = true;
    }
    FSB_CUDA_OK(cudaEventCreate(&lm->ev0));
    FSB_CUDA_OK(cudaEventCreate(&lm->ev1));
    FSB_CUDA_OK(cudaEventCreate(&lm->ev2));
    const int D = lm->D, V = lm->V, C = lm->C, CS = lm->CS, wdt = lm->wdt;
    cudaStream_t st = lm->stream;
    FSB_TRY(upload_tensor(w, n, "embeddings.weight", {V, D}, wdt, st, &lm->emb, &lm->owned));
    FSB_TRY(upload_tensor(w, n, "codebook_embeddings.weight", {(int64_t)C * CS, D}, wdt, st, &lm->cb_emb, &lm->owned));
    lm->layers.resize(lm->NL);
    for (int l = 0; l < lm->NL; ++l) FSB_TRY(load_block(lm, w, n, "layers." + std::to_string(l) + ".", &lm->layers[l]));
    FSB_TRY(upload_tensor(w, n, "norm.weight", {D}, FSB_F32, st, &lm->norm, &lm->owned));
    if (c.tie_word_embeddings) lm->out_w = lm->emb;  // dual_ar.rs:486-490
    else FSB_TRY(upload_tensor(w, n, "output.weight", {V, D}, wdt, st, &lm->out_w, &lm->owned));
    FSB_TRY(upload_tensor(w, n, "fast_embeddings.weight", {CS, D}, wdt, st, &lm->fast_emb, &lm->owned));
    lm->fast_layers.resize(lm->NFL);
    for (int l = 0; l < lm->NFL; ++l)
        FSB_TRY(load_block(lm, w, n, "fast_layers." + std::to_string(l) + ".", &lm->fast_layers[l]));
    FSB_TRY(upload_tensor(w, n, "fast_norm.weight", {D}, FSB_F32, st, &lm->fast_norm, &lm->owned));
    FSB_TRY(upload_tensor(w, n, "fast_output.weight", {CS, D}, wdt, st, &lm->fast_out, &lm->owned));

    // RoPE tables (dual_ar.rs:168-186), built on the host exactly like the oracle
    std::vector<float> cosv, sinv;
    precompute_freqs(c, c.max_seq_len, &cosv, &sinv);
    FSB_TRY(dev_alloc(lm, &lm->cosT, cosv.size()));
    FSB_TRY(dev_alloc(lm, &lm->sinT, sinv.size()));
    FSB_CUDA_OK(cudaMemcpyAsync(lm->cosT, cosv.data(), cosv.size() * 4, cudaMemcpyHostToDevice, st));
    FSB_CUDA_OK(cudaMemcpyAsync(lm->sinT, sinv.data(), sinv.size() * 4, cudaMemcpyHostToDevice, st));
    FSB_CUDA_OK(cudaStreamSynchronize(st));

    const int B = lm->max_batch;
    const size_t kv_per_layer = (size_t)B * lm->KV * lm->max_len * lm->hd;
    FSB_TRY(dev_alloc(lm, &lm->kc, kv_per_layer * lm->NL));
    FSB_TRY(dev_alloc(lm, &lm->vc, kv_per_layer * lm->NL));
    const size_t fkv = (size_t)B * lm->KV * lm->fast_len * lm->hd;
    FSB_TRY(dev_alloc(lm, &lm->fkc, fkv * lm->NFL));
    FSB_TRY(dev_alloc(lm, &lm->fvc, fkv * lm->NFL));

    Scratch &s = lm->s;
    const int M = std::max(lm->prefill_rows, B);
    FSB_TRY(dev_alloc(lm, &s.x, (size_t)M * D));
    FSB_TRY(dev_alloc(lm, &s.xn, (size_t)M * D));
    FSB_TRY(dev_alloc(lm, &s.qkv, (size_t)M * lm->QKV));
    FSB_TRY(dev_alloc(lm, &s.q, (size_t)M * lm->H * lm->hd));
    FSB_TRY(dev_alloc(lm, &s.att, (size_t)M * lm->H * lm->hd));
    FSB_TRY(dev_alloc(lm, &s.g1, (size_t)M * lm->I));
    FSB_TRY(dev_alloc(lm, &s.g3, (size_t)M * lm->I));
    FSB_TRY(dev_alloc(lm, &s.partial, (size_t)B * lm->H * lm->nsplit * (lm->hd + 2)));
    FSB_TRY(dev_alloc(lm, &s.logits, (size_t)B * V));
    FSB_TRY(dev_alloc(lm, &s.slow_logits, (size_t)B * lm->n_slow_logits));
    FSB_TRY(dev_alloc(lm, &s.hidden, (size_t)B * D));
    FSB_TRY(dev_alloc(lm, &s.fast_x, (size_t)B * D));
    FSB_TRY(dev_alloc(lm, &s.fast_logits, (size_t)B * CS));
    lm->toks_cap = (size_t)(C + 1) * std::max(lm->max_len, 1) * B;  // every row's prompt at once (batched prefill)
    FSB_TRY(dev_alloc(lm, &s.toks, lm->toks_cap));
    FSB_TRY(dev_alloc(lm, &lm->d_segs, 256));

    GenState &g = lm->h_st;
    memset(&g, 0, sizeof(g));
    FSB_TRY(dev_alloc(lm, &g.pos, B));
    FSB_TRY(dev_alloc(lm, &g.active, B));
    FSB_TRY(dev_alloc(lm, &g.eos, B));
    FSB_TRY(dev_alloc(lm, &g.frame, B));
    FSB_TRY(dev_alloc(lm, &g.max_frames, B));
    FSB_TRY(dev_alloc(lm, &g.n_active, B));  // [0]: whole batch (per-op path); [g]: group g of 8 rows (megakernel)
    FSB_TRY(dev_alloc(lm, &g.cur, (size_t)B * (C + 1)));
    FSB_TRY(dev_alloc(lm, &g.prev, (size_t)B * (C + 1)));
    g.out_cap = lm->max_len + 2;
    FSB_TRY(dev_alloc(lm, &g.out, (size_t)B * g.out_cap * (C + 1)));
    FSB_TRY(dev_alloc(lm, &g.rep, (size_t)B * C));
    g.C = C;
    g.im_end_id = lm->tok.im_end_id;
    g.pad_id = lm->tok.pad_id;
    FSB_TRY(dev_alloc(lm, &lm->d_st, 1));
    FSB_CUDA_OK(cudaMemsetAsync(g.pos, 0, B * sizeof(int), st));
    FSB_CUDA_OK(cudaMallocHost((void **)&lm->h_pin, sizeof(int) * (4 * B + 16)));
    lm->kv_len.assign(B, 0);
    // sampler kernels may need > 48 KB dynamic smem
    {
        const int sm = (int)sample_smem(std::max(lm->n_slow_logits, CS));
        FSB_REQUIRE(std::max(lm->n_slow_logits, CS) <= kSampleMaxN, FSB_ERR_UNSUPPORTED,
                    "constrained head of %d rows exceeds the block sampler", lm->n_slow_logits);
        // the attribute is per kernel and per process, shared by every live handle: only ever raise it (a handle with a
        // smaller head must not lower the limit under another handle's launches)
        static int sm_max = 0;
        if (sm > sm_max) {
            FSB_CUDA_OK(cudaFuncSetAttribute(sample_slow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
            FSB_CUDA_OK(cudaFuncSetAttribute(sample_fast_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
            FSB_CUDA_OK(
                cudaFuncSetAttribute(sample_fast_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
            sm_max = sm;
        }
    }
    FSB_REQUIRE(std::max(lm->D, lm->I) * 32 <= 227 * 1024, FSB_ERR_UNSUPPORTED, "dim/intermediate_size too large");
    FSB_CUDA_OK(init_gemv_attrs(std::max(lm->D, lm->I)));
    FSB_TRY(mega_setup(lm));
    FSB_TRY(tc_setup(lm));
    FSB_TRY(megab_setup(lm));
    FSB_CUDA_OK(cudaStreamSynchronize(st));
    return FSB_OK;
}

static void lm_free(fsb_lm *lm) {
    if (!lm) return;
    for (auto &kv : lm->frame_graphs)
        if (kv.first >= 0 && kv.second) cudaGraphExecDestroy(kv.second);
    for (auto &kv : lm->tail_graphs)
        if (kv.first >= 0 && kv.second) cudaGraphExecDestroy(kv.second);
    for (void *p : lm->owned) cudaFree(p);
    for (cudaEvent_t e : lm->prof_ev) cudaEventDestroy(e);
    if (lm->h_pin) cudaFreeHost(lm->h_pin);
    if (lm->ev0) cudaEventDestroy(lm->ev0);
    if (lm->ev1) cudaEventDestroy(lm->ev1);
    if (lm->ev2) cudaEventDestroy(lm->ev2);
    if (lm->prof_m0) cudaEventDestroy(lm->prof_m0);
    if (lm->prof_m1) cudaEventDestroy(lm->prof_m1);
    if (lm->own_stream && lm->stream) cudaStreamDestroy(lm->stream);
    delete lm;
}

static int check_handle(fsb_lm *lm) {
    FSB_REQUIRE(lm != nullptr, FSB_ERR_INVALID, "null fsb_lm handle");
    FSB_REQUIRE(!lm->poisoned, FSB_ERR_CUDA, "handle poisoned by an earlier sticky CUDA error");
    cudaError_t e = cudaSetDevice(lm->opt.device);
    if (e != cudaSuccess) {
        set_error("cudaSetDevice(%d): %s", lm->opt.device, cudaGetErrorString(e));
        return FSB_ERR_CUDA;
    }
    return FSB_OK;
}

static int finish(fsb_lm *lm, int st) {
    if (st == FSB_ERR_CUDA) {
        // a sticky error (illegal address, launch failure) poisons the context
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess && e != cudaErrorNotReady) lm->poisoned = true;
        (void)cudaGetLastError();
    }
    return st;
}

static int upload_state(fsb_lm *lm) {
    FSB_CUDA_OK(cudaMemcpyAsync(lm->d_st, &lm->h_st, sizeof(GenState), cudaMemcpyHostToDevice, lm->stream));
    return FSB_OK;
}

static int sync_pos_to_device(fsb_lm *lm, int nb) {
    for (int b = 0; b < nb; ++b) lm->h_pin[b] = lm->kv_len[b];
    FSB_CUDA_OK(cudaMemcpyAsync(lm->h_st.pos, lm->h_pin, nb * sizeof(int), cudaMemcpyHostToDevice, lm->stream));
    return FSB_OK;
}

// ---------------------------------------------------------------- generate
static int generate_impl(fsb_lm *lm, const uint32_t *const *prompts, const int32_t *prompt_lens, int bsz,
                         size_t max_new_tokens, const fsb_sampling_args *sa, uint32_t flags, int32_t fixed_len,
                         uint32_t *const *out_codes, size_t cap, size_t *out_lens) {
    FSB_REQUIRE(prompts && prompt_lens && sa && out_codes && out_lens, FSB_ERR_INVALID, "null argument");
    FSB_REQUIRE(bsz >= 1 && bsz <= lm->max_batch, FSB_ERR_INVALID, "bsz %d outside [1, max_batch=%d]", bsz,
                lm->max_batch);
    const int C = lm->C;
    const bool fixed = (flags & FSB_GEN_FIXED_LEN) != 0;
    FSB_REQUIRE(!fixed || fixed_len >= 1, FSB_ERR_INVALID, "FSB_GEN_FIXED_LEN needs fixed_len >= 1");
    if (!(flags & FSB_GEN_KEEP_SLOW_KV))
        for (int b = 0; b < bsz; ++b) lm->kv_len[b] = 0;
    // sampling/mod.rs: WeightedIndex over an empty candidate set is an error in the reference too
    FSB_REQUIRE(std::isfinite(sa->temp) && sa->temp >= 0.0 && std::isfinite(sa->top_p), FSB_ERR_INVALID,
                "sampling: temp %g / top_p %g must be finite and temp >= 0", sa->temp, sa->top_p);
    FSB_REQUIRE(sa->temp <= 1e-7 || sa->top_k >= 1, FSB_ERR_INVALID, "sampling: top_k must be >= 1 when temp > 0");
    FSB_REQUIRE(std::isfinite(sa->repetition_penalty) && sa->repetition_penalty != 0.f, FSB_ERR_INVALID,
                "sampling: repetition_penalty must be finite and non-zero");
    std::vector<int> max_frames(bsz);
    for (int b = 0; b < bsz; ++b) {
        const int P = prompt_lens[b];
        FSB_REQUIRE(P >= 1 && prompts[b], FSB_ERR_INVALID, "prompt %d is empty", b);
        for (int i = 0; i < P; ++i)
            FSB_REQUIRE(prompts[b][i] < (uint32_t)lm->V, FSB_ERR_INVALID, "prompt %d: token id %u at column %d >= vocab_size %d",
                        b, prompts[b][i], i, lm->V);
        for (int c = 1; c <= C; ++c)
            for (int i = 0; i < P; ++i)
                FSB_REQUIRE(prompts[b][(size_t)c * P + i] < (uint32_t)lm->CS, FSB_ERR_INVALID,
                            "prompt %d: code %u (codebook %d, column %d) >= codebook_size %d", b,
                            prompts[b][(size_t)c * P + i], c - 1, i, lm->CS);
        // Q3: `input_pos > max_new_tokens + n_cached` stops the iterator (single_batch.rs:61,77)
        long long lim = (long long)max_new_tokens - P + 2;
        if (lim < 1) lim = 1;
        if (fixed) lim = std::min<long long>(lim, fixed_len);
        FSB_REQUIRE(lm->kv_len[b] + P + lim - 1 <= lm->max_len, FSB_ERR_STATE,
                    "row %d: %d cached + %d prompt + %lld frames exceed the KV arena (%d positions)", b, lm->kv_len[b],
                    P, lim, lm->max_len);
        FSB_REQUIRE(lim <= lm->h_st.out_cap, FSB_ERR_STATE, "frame budget %lld exceeds the output arena", lim);
        max_frames[b] = (int)lim;
    }
    // ---- state ----
    GenState &g = lm->h_st;
    g.fixed_len = fixed ? 1 : 0;
    g.legacy_slow = (lm->opt.fish_version != FSB_FISH_1_5) ? 1 : 0;
    g.sp.greedy = sa->temp <= 1e-7 ? 1 : 0;
    g.sp.inv_temp = g.sp.greedy ? 1.0f : (float)(1.0 / sa->temp);
    g.sp.top_p = (float)sa->top_p;
    g.sp.top_p_gate = top_p_gate_of(sa->top_p);
    g.sp.top_k = sa->top_k;
    g.sp.penalty = sa->repetition_penalty;
    g.sp.seed = sa->seed;
    FSB_TRY(upload_state(lm));
    int *hp = lm->h_pin;
    for (int b = 0; b < bsz; ++b) {
        hp[b] = 1;                    // active
        hp[bsz + b] = max_frames[b];  // max_frames
        hp[2 * bsz + b] = lm->kv_len[b] + prompt_lens[b];  // pos after prefill
    }
    hp[3 * bsz] = bsz;
    cudaStream_t st = lm->stream;
    FSB_CUDA_OK(cudaMemcpyAsync(g.active, hp, bsz * sizeof(int), cudaMemcpyHostToDevice, st));
    FSB_CUDA_OK(cudaMemcpyAsync(g.max_frames, hp + bsz, bsz * sizeof(int), cudaMemcpyHostToDevice, st));
    FSB_CUDA_OK(cudaMemcpyAsync(g.pos, hp + 2 * bsz, bsz * sizeof(int), cudaMemcpyHostToDevice, st));
    FSB_CUDA_OK(cudaMemcpyAsync(g.n_active, hp + 3 * bsz, sizeof(int), cudaMemcpyHostToDevice, st));
    FSB_CUDA_OK(cudaMemsetAsync(g.eos, 0, bsz * sizeof(int), st));
    FSB_CUDA_OK(cudaMemsetAsync(g.frame, 0, bsz * sizeof(int), st));
    FSB_CUDA_OK(cudaMemsetAsync(g.rep, 0, (size_t)bsz * C * sizeof(RepPenState), st));
    FSB_CUDA_OK(cudaStreamSynchronize(st));  // h_pin is reused below

    const uint64_t l0 = lm->launches;
    uint64_t graph_l = 0;
    // ---- prefill, row by row (independent utterances) ----
    FSB_CUDA_OK(cudaEventRecord(lm->ev0, st));
    {
        size_t tok_off = 0;
        for (int b = 0; b < bsz; ++b) {
            const size_t n = (size_t)(C + 1) * prompt_lens[b];
            FSB_REQUIRE(tok_off + n <= lm->toks_cap, FSB_ERR_STATE, "prompts exceed the token staging arena");
            FSB_CUDA_OK(cudaMemcpyAsync(lm->s.toks + tok_off, prompts[b], n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            tok_off += n;
        }
        FSB_TRY(prefill_batch(lm, lm->s.toks, prompt_lens, bsz));
    }
    int total_max = 0;
    for (int b = 0; b < bsz; ++b) total_max = std::max(total_max, max_frames[b]);
    bool use_mega = lm->mega_ok && lm->opt.decode_mode != 1 && !lm->profile && !lm->collect_hidden;
    int group = 8;  // rows per megakernel launch: the largest batch template whose shared memory fits
    // wide batches (cfg3 / cfg5): ONE tcgen05 megakernel launch for all rows (fsb_lm_megab.cuh)
    int megab_min = 9;
    if (const char *v = getenv("FSB_MEGAB_MIN_ROWS")) megab_min = std::max(2, atoi(v));
    const bool use_megab = use_mega && lm->megab_ok && bsz >= megab_min && bsz <= lm->megab_rows_cap &&
                           (g.sp.greedy || g.sp.top_k <= (uint32_t)kSelMaxK);
    if (use_megab) {
        group = bsz;
        FSB_CUDA_OK(cudaEventRecord(lm->ev1, st));
        FSB_TRY(megab_launch_rows(lm, bsz, total_max));
    } else if (use_mega) {
        int xf, vf;  // (the opt-in limit is cached at setup: cudaGetDeviceProperties costs ~200 ms per call)
        while (group > 1 && mega_smem_bytes(lm, group, &xf, &vf) > lm->smem_optin) group /= 2;
        use_mega = mega_smem_bytes(lm, group, &xf, &vf) <= lm->smem_optin;
        // auto mode: one launch only (rows beyond a group are better served by the per-op GEMV path)
        if (lm->opt.decode_mode == 0 && bsz > group) use_mega = false;
    }
    FSB_REQUIRE(use_mega || lm->opt.decode_mode != 2 || lm->profile || lm->collect_hidden, FSB_ERR_UNSUPPORTED,
                "decode_mode 2 (megakernel) needs bsz <= 8 (<= 32 with bf16 Fish shapes) and cooperative launch support");
    if (use_megab) {
        // launched above
    } else if (use_mega) {
        // frame 0 = tail on the prefilled hidden state, then whole frames; the kernel leaves its
        // loop by itself once every row is finished (no host polling).  The prefill event sits
        // right before the launch, so frame 0's tail is accounted to the frame loop here.
        FSB_CUDA_OK(cudaEventRecord(lm->ev1, st));
        // batches above 8 rows run as consecutive launches over groups of 8 rows (the kernel keeps one
        // activation row per batch row in shared memory); each group has its own live-row counter
        for (int r0 = 0, gi = 0; r0 < bsz; r0 += group, ++gi) {
            const int nbg = std::min(group, bsz - r0);
            int fmax_g = 0;
            for (int b = r0; b < r0 + nbg; ++b) fmax_g = std::max(fmax_g, max_frames[b]);
            hp[0] = nbg;
            FSB_CUDA_OK(cudaMemcpyAsync(g.n_active + gi, hp, sizeof(int), cudaMemcpyHostToDevice, st));
            FSB_CUDA_OK(cudaStreamSynchronize(st));  // hp is reused
            FSB_TRY(mega_launch(lm, r0, gi, nbg, fmax_g, true));
        }
    } else {
        {
            cudaGraphExec_t tg;
            FSB_TRY(get_graph(lm, lm->tail_graphs, bsz, false, &tg));
            FSB_CUDA_OK(cudaGraphLaunch(tg, st));
            graph_l += graph_launches(lm, lm->tail_graphs, bsz);
        }
        FSB_CUDA_OK(cudaEventRecord(lm->ev1, st));
        // ---- frame loop ----
        cudaGraphExec_t fg;
        FSB_TRY(get_graph(lm, lm->frame_graphs, bsz, true, &fg));
        const uint64_t per_frame = graph_launches(lm, lm->frame_graphs, bsz);
        const int kPoll = 16;
        int launched = 1;
        if (lm->profile) {
            // eager frames with per-launch events while the event pool lasts
            lm->prof_used = 0;
            lm->prof_bytes = 0;
            while (launched < total_max && lm->prof_used + 2 * per_frame <= lm->prof_ev.size()) {
                FSB_TRY(decode_frame(lm, bsz));
                ++launched;
            }
        }
        volatile int *flag = hp + 3 * bsz + 1;
        *flag = bsz;
        while (launched < total_max) {
            const int n = std::min(kPoll, total_max - launched);
            for (int i = 0; i < n; ++i) FSB_CUDA_OK(cudaGraphLaunch(fg, st));
            launched += n;
            graph_l += per_frame * n;
            if (fixed) continue;
            // early exit on EOS: one poll per kPoll frames (kernels no-op once n_active == 0)
            FSB_CUDA_OK(cudaMemcpyAsync((void *)flag, g.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
            FSB_CUDA_OK(cudaStreamSynchronize(st));
            if (*flag == 0) break;
        }
    }
    FSB_CUDA_OK(cudaEventRecord(lm->ev2, st));
    // ---- results ----
    std::vector<int> frames(bsz), pos(bsz);
    FSB_CUDA_OK(cudaMemcpyAsync(hp, g.frame, bsz * sizeof(int), cudaMemcpyDeviceToHost, st));
    FSB_CUDA_OK(cudaMemcpyAsync(hp + bsz, g.pos, bsz * sizeof(int), cudaMemcpyDeviceToHost, st));
    FSB_CUDA_OK(cudaStreamSynchronize(st));
    uint64_t total_frames = 0;
    std::vector<uint32_t> host_out;
    for (int b = 0; b < bsz; ++b) {
        frames[b] = hp[b];
        lm->kv_len[b] = hp[bsz + b];
    }
    for (int b = 0; b < bsz; ++b) {
        const int nf = frames[b];
        host_out.resize((size_t)nf * (C + 1));
        FSB_CUDA_OK(cudaMemcpyAsync(host_out.data(), g.out + (size_t)b * g.out_cap * (C + 1),
                                    host_out.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        FSB_CUDA_OK(cudaStreamSynchronize(st));
        // generate_blocking_with_hidden: keep frame 0 always (Q4), drop later <|im_end|> frames
        // (single_batch.rs:262-266), strip the semantic row (:280-282)
        size_t T = 0;
        for (int f = 0; f < nf; ++f) {
            const uint32_t *fr = &host_out[(size_t)f * (C + 1)];
            if (f > 0 && fr[0] == lm->tok.im_end_id) continue;
            FSB_REQUIRE(T < cap, FSB_ERR_INVALID, "row %d: output capacity %zu too small", b, cap);
            for (int c = 0; c < C; ++c) out_codes[b][(size_t)c * cap + T] = fr[1 + c];
            ++T;
        }
        out_lens[b] = T;
        total_frames += T;
    }
    float ms_pre = 0.f, ms_dec = 0.f;
    FSB_CUDA_OK(cudaEventElapsedTime(&ms_pre, lm->ev0, lm->ev1));
    FSB_CUDA_OK(cudaEventElapsedTime(&ms_dec, lm->ev1, lm->ev2));
    lm->stats.prefill_ms = ms_pre;
    lm->stats.decode_ms = ms_dec;
    lm->stats.frames = total_frames;
    lm->stats.kernel_launches = (lm->launches - l0) + graph_l;
    lm->stats.dominant_kernel_ms = 0;
    lm->stats.dominant_kernel_launches = 0;
    lm->stats.dominant_kernel_bytes = 0;
    if (use_mega && lm->mega_dbg) {
        unsigned long long h[256];
        FSB_CUDA_OK(cudaMemcpy(h, lm->mega_dbg, sizeof(h), cudaMemcpyDeviceToHost));
        FSB_CUDA_OK(cudaMemset(lm->mega_dbg, 0, sizeof(h)));
        static const char *kn[7] = {"qkv", "attn", "wo", "w13|ffn", "w2|fred", "head", "sample"};
        for (int k = 0; k < 7; ++k) {  // wide-batch kernel: sub-steps of a projection phase on its owner CTA
            const unsigned long long *u = h + 192 + k * 8;
            if (u[5] && k != 1)
                fprintf(stderr, "[megab owner] %-7s n=%6llu stage %5.2f  acc-wait %5.2f  epilogue|swiglu %5.2f  acc2-wait %5.2f  epilogue2 %5.2f us/phase\n",
                        kn[k], u[5], u[0] / 1965.0 / u[5], u[1] / 1965.0 / u[5], u[2] / 1965.0 / u[5], u[3] / 1965.0 / u[5],
                        u[4] / 1965.0 / u[5]);
        }
        {
            const unsigned long long *u = h + 192 + 1 * 8;  // attention on CTA 0: [slow range, slow tail, n, fast range, fast tail, n]
            if (u[2] || u[5])
                fprintf(stderr, "[megab attn cta 0] slow: range %5.2f tail %5.2f us (n=%llu)   fast: range %5.2f tail %5.2f us (n=%llu)\n",
                        u[2] ? u[0] / 1965.0 / u[2] : 0.0, u[2] ? u[1] / 1965.0 / u[2] : 0.0, u[2], u[5] ? u[3] / 1965.0 / u[5] : 0.0,
                        u[5] ? u[4] / 1965.0 / u[5] : 0.0, u[5]);
        }
        for (int k = 0; k < 7; ++k) {  // wide-batch kernel: third timed CTA (74: owns a WO tile)
            const unsigned long long *e = h + 128 + k * 4;
            if (e[3])
                fprintf(stderr, "[mega cta 74] %-7s n=%6llu work %7.2f  barrier %7.2f us/phase\n", kn[k], e[3], e[0] / 1965.0 / e[3],
                        e[2] / 1965.0 / e[3]);
        }
        if (h[103])
            fprintf(stderr, "[sample_fast] rep-pen %.2f  logits %.2f  block_sample %.2f  tail (next input) %.2f us (n=%llu)\n", h[100] / 1965.0 / h[103],
                    h[101] / 1965.0 / h[103], h[102] / 1965.0 / h[103], h[116] / 1965.0 / h[103], h[103]);
        if (h[107])
            fprintf(stderr, "[block_sample] max+softmax+keys %.2f  sort|select %.2f  scan|rank %.2f  (walk %.2f) us (n=%llu)\n",
                    h[104] / 1965.0 / h[107], h[105] / 1965.0 / h[107], h[106] / 1965.0 / h[107], h[108] / 1965.0 / h[107], h[107]);
        if (h[107])
            fprintf(stderr, "[sampler] per select round: ballots %.3f  atomic+barrier %.3f  scan %.3f us\n",
                    h[113] / 1965.0 / h[107] / 9, h[114] / 1965.0 / h[107] / 9, h[115] / 1965.0 / h[107] / 9);
        for (int c = 0; c < 2; ++c)
            for (int k = 0; k < 7; ++k) {
                const unsigned long long *e = h + c * 32 + k * 4;
                if (e[3])
                    fprintf(stderr, "[mega cta %s] %-6s n=%6llu work %7.2f  arrive+prep %7.2f  wait %7.2f us/phase\n",
                            c ? "last" : "0", kn[k], e[3], e[0] / 1965.0 / e[3], e[1] / 1965.0 / e[3], e[2] / 1965.0 / e[3]);
                if (c == 0 && e[3] && k != 1 && k != 6) {
                    const unsigned long long *u = h + 64 + k * 4;
                    fprintf(stderr, "             %-6s prologue %6.2f  sync %6.2f  norm %6.2f  tasks %6.2f us/phase\n", kn[k],
                            (double)(long long)u[3] / 1965.0 / e[3], u[0] / 1965.0 / e[3], u[1] / 1965.0 / e[3], u[2] / 1965.0 / e[3]);
                }
            }
    }
    if (use_mega) {
        // the single persistent launch IS the frame loop: algorithmic bytes = weights streamed per
        // executed frame (frame 0 runs no slow stack) + the KV rows attention had to read
        const size_t es = esize(lm->wdt);
        const size_t per_layer = (size_t)lm->QKV * lm->D + (size_t)lm->D * lm->H * lm->hd + 3ull * lm->I * lm->D;
        const uint64_t w_slow = es * (per_layer * lm->NL);
        const uint64_t w_tail = es * ((size_t)lm->n_slow_logits * lm->D +
                                      (size_t)lm->C * (per_layer * lm->NFL + (size_t)lm->CS * lm->D));
        uint64_t kv_bytes = 0, weight_bytes = 0;
        for (int r0 = 0; r0 < bsz; r0 += group) {  // every group streams the weights once per frame it runs
            int fmax = 0;
            for (int b = r0; b < std::min(bsz, r0 + group); ++b) {
                fmax = std::max(fmax, frames[b]);
                for (int f = 1; f < frames[b]; ++f)
                    kv_bytes += (uint64_t)(lm->kv_len[b] - (frames[b] - 1) + f) * lm->NL * 2 * lm->KV * lm->hd * sizeof(float);
            }
            weight_bytes += (uint64_t)fmax * w_tail + (uint64_t)std::max(fmax - 1, 0) * w_slow;
        }
        lm->stats.dominant_kernel_ms = ms_dec;
        lm->stats.dominant_kernel_launches = (bsz + group - 1) / group;
        lm->stats.dominant_kernel_bytes = weight_bytes + kv_bytes;
    }
    if (lm->profile) {
        double tot = 0;
        for (size_t i = 0; i + 1 < lm->prof_used; i += 2) {
            float ms = 0.f;
            FSB_CUDA_OK(cudaEventElapsedTime(&ms, lm->prof_ev[i], lm->prof_ev[i + 1]));
            tot += ms;
        }
        lm->stats.dominant_kernel_ms = tot;
        lm->stats.dominant_kernel_launches = lm->prof_used / 2;
        lm->stats.dominant_kernel_bytes = lm->prof_bytes;
    }
    return FSB_OK;
}

}  // namespace fsb

// =================================================================== C ABI
extern "C" {

int fsb_lm_create(const fsb_model_args *args, const fsb_token_config *tok, const fsb_tensor *weights,
                  size_t n_weights, const fsb_lm_options *opts, fsb_lm **out) {
    FSB_REQUIRE(args && tok && weights && opts && out, FSB_ERR_INVALID, "fsb_lm_create: null argument");
    *out = nullptr;
    FSB_REQUIRE(!args->attention_qkv_bias, FSB_ERR_UNSUPPORTED, "attention_qkv_bias is not supported");
    FSB_REQUIRE(args->head_dim == 64, FSB_ERR_UNSUPPORTED, "head_dim %d: kernels are specialised for 64",
                args->head_dim);
    FSB_REQUIRE(args->n_local_heads > 0 && args->n_head % args->n_local_heads == 0 &&
                    args->n_head / args->n_local_heads <= kAttnMaxRep,
                FSB_ERR_UNSUPPORTED, "n_head %d / n_local_heads %d unsupported", args->n_head, args->n_local_heads);
    FSB_REQUIRE(args->dim % 256 == 0, FSB_ERR_UNSUPPORTED, "dim %d must be a multiple of 256", args->dim);
    FSB_REQUIRE(args->dim == args->n_head * args->head_dim, FSB_ERR_UNSUPPORTED,
                "dim must equal n_head * head_dim (RoPE table uses dim / n_head, dual_ar.rs:173)");
    FSB_REQUIRE(opts->weight_dtype == FSB_F32 || opts->weight_dtype == FSB_BF16, FSB_ERR_INVALID,
                "weight_dtype must be FSB_F32 or FSB_BF16");
    FSB_REQUIRE(opts->max_batch >= 1, FSB_ERR_INVALID, "max_batch must be >= 1");
    FSB_REQUIRE(args->num_codebooks >= 1 && args->num_codebooks <= 16, FSB_ERR_UNSUPPORTED, "num_codebooks");
    FSB_REQUIRE(args->codebook_size <= 1024, FSB_ERR_UNSUPPORTED, "codebook_size > 1024 (rep-pen bitset)");
    std::unique_ptr<fsb_lm> lm(new fsb_lm());
    lm->cfg = *args;
    lm->tok = *tok;
    lm->opt = *opts;
    lm->D = args->dim;
    lm->I = args->intermediate_size ? args->intermediate_size : args->dim * 4;
    FSB_REQUIRE(lm->I % 256 == 0, FSB_ERR_UNSUPPORTED, "intermediate_size must be a multiple of 256");
    lm->H = args->n_head;
    lm->KV = args->n_local_heads;
    lm->hd = args->head_dim;
    lm->V = args->vocab_size;
    lm->C = args->num_codebooks;
    lm->CS = args->codebook_size;
    lm->QKV = (lm->H + 2 * lm->KV) * lm->hd;
    lm->NL = args->n_layer;
    lm->NFL = args->n_fast_layer;
    lm->wdt = opts->weight_dtype;
    lm->max_batch = opts->max_batch;
    lm->max_len = opts->max_seq_len > 0 ? std::min(opts->max_seq_len, args->max_seq_len) : args->max_seq_len;
    lm->fast_len = lm->C;
    // positions per prefill pass: one prompt for a single-row handle, several rows' prompts packed together otherwise
    lm->prefill_rows = lm->max_batch > 1 ? 4096 : 1024;
    lm->nsplit = 16;
    FSB_REQUIRE(tok->im_end_id < (uint32_t)lm->V && tok->semantic_start_id < (uint32_t)lm->V, FSB_ERR_INVALID,
                "token ids outside the vocabulary");
    if (opts->fish_version == FSB_FISH_1_5) {
        // generate/utils.rs:6-33: [im_end | semantic_start .. V)
        lm->slow_row0 = (int)tok->im_end_id;
        lm->slow_rest_base = (int)tok->semantic_start_id;
        lm->n_slow_logits = 1 + (lm->V - (int)tok->semantic_start_id);
    } else {
        // single_batch.rs:104-124: only the <|im_end|> and PAD logits are read
        lm->slow_row0 = (int)tok->im_end_id;
        lm->slow_rest_base = (int)tok->pad_id;
        lm->n_slow_logits = 2;
    }
    memset(&lm->stats, 0, sizeof(lm->stats));
    int st = lm_create_impl(lm.get(), weights, n_weights);
    if (st != FSB_OK) {
        lm_free(lm.release());
        (void)cudaGetLastError();
        return st;
    }
    const size_t es = esize(lm->wdt);
    const size_t per_layer = (size_t)lm->QKV * lm->D + (size_t)lm->D * lm->H * lm->hd + 3ull * lm->I * lm->D;
    lm->stats.weight_bytes_per_frame =
        es * (per_layer * lm->NL + (size_t)lm->n_slow_logits * lm->D +
              (size_t)lm->C * (per_layer * lm->NFL + (size_t)lm->CS * lm->D));
    *out = lm.release();
    return FSB_OK;
}

int fsb_lm_destroy(fsb_lm *lm) {
    if (!lm) return FSB_OK;
    cudaSetDevice(lm->opt.device);
    if (lm->stream) cudaStreamSynchronize(lm->stream);
    lm_free(lm);
    (void)cudaGetLastError();
    return FSB_OK;
}

int fsb_lm_forward_generate(fsb_lm *lm, const uint32_t *inp, int32_t bsz, int32_t seq_len, size_t input_pos,
                            float *logits, float *hidden) {
    FSB_TRY(check_handle(lm));
    FSB_REQUIRE(inp && bsz >= 1 && bsz <= lm->max_batch && seq_len >= 1, FSB_ERR_INVALID,
                "forward_generate: bad arguments (bsz %d, seq_len %d)", bsz, seq_len);
    const int C = lm->C, D = lm->D, V = lm->V;
    cudaStream_t st = lm->stream;
    auto body = [&]() -> int {
        for (int b = 0; b < bsz; ++b) {
            FSB_REQUIRE(lm->kv_len[b] + seq_len <= lm->max_len, FSB_ERR_STATE,
                        "KV overflow: %d cached + %d new > %d", lm->kv_len[b], seq_len, lm->max_len);
            FSB_REQUIRE(input_pos + seq_len <= (size_t)lm->cfg.max_seq_len, FSB_ERR_STATE,
                        "input_pos %zu + %d exceeds max_seq_len", input_pos, seq_len);
        }
        if (seq_len == 1) {
            FSB_CUDA_OK(cudaMemcpyAsync(lm->s.toks, inp, (size_t)bsz * (C + 1) * sizeof(uint32_t),
                                        cudaMemcpyHostToDevice, st));
            FSB_TRY(sync_pos_to_device(lm, bsz));
            FSB_TRY(embed(lm, lm->s.toks, bsz, 1, lm->s.hidden, nullptr));
            // all rows share input_pos (reference semantics); cache slot == rows cached so far
            for (int b = 1; b < bsz; ++b)
                FSB_REQUIRE(lm->kv_len[b] == lm->kv_len[0], FSB_ERR_STATE,
                            "step API needs equal KV lengths across rows");
            const int delta = (int)input_pos - lm->kv_len[0];
            for (int l = 0; l < lm->NL; ++l)
                FSB_TRY(decode_layer(lm, lm->layers[l], lm->s.hidden, bsz, slow_k(lm, l), slow_v(lm, l), lm->max_len,
                                     lm->h_st.pos, 0, delta, nullptr, lm->nsplit));
        } else {
            for (int b = 0; b < bsz; ++b) {
                FSB_CUDA_OK(cudaMemcpyAsync(lm->s.toks, inp + (size_t)b * (C + 1) * seq_len,
                                            (size_t)(C + 1) * seq_len * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
                FSB_TRY(prefill_row(lm, lm->s.toks, seq_len, b, lm->kv_len[b], (int)input_pos - lm->kv_len[b]));
            }
        }
        for (int b = 0; b < bsz; ++b) lm->kv_len[b] += seq_len;
        if (logits) {
            FSB_TRY(launch_gemv<EPI_STORE>(lm, lm->out_w, nullptr, lm->s.hidden, D, (const float *)lm->norm.ptr,
                                           nullptr, lm->s.logits, V, V, D, bsz, nullptr));
            FSB_CUDA_OK(cudaMemcpyAsync(logits, lm->s.logits, (size_t)bsz * V * sizeof(float),
                                        cudaMemcpyDeviceToHost, st));
        }
        if (hidden)
            FSB_CUDA_OK(cudaMemcpyAsync(hidden, lm->s.hidden, (size_t)bsz * D * sizeof(float),
                                        cudaMemcpyDeviceToHost, st));
        FSB_CUDA_OK(cudaStreamSynchronize(st));
        return FSB_OK;
    };
    return finish(lm, body());
}

int fsb_lm_forward_generate_fast(fsb_lm *lm, const float *x, int32_t bsz, size_t input_pos, float *logits) {
    FSB_TRY(check_handle(lm));
    FSB_REQUIRE(x && logits && bsz >= 1 && bsz <= lm->max_batch, FSB_ERR_INVALID, "forward_generate_fast: bad args");
    FSB_REQUIRE(input_pos < (size_t)lm->fast_len, FSB_ERR_STATE, "fast input_pos %zu >= num_codebooks %d", input_pos,
                lm->fast_len);
    cudaStream_t st = lm->stream;
    auto body = [&]() -> int {
        FSB_CUDA_OK(cudaMemcpyAsync(lm->s.fast_x, x, (size_t)bsz * lm->D * sizeof(float), cudaMemcpyHostToDevice, st));
        for (int l = 0; l < lm->NFL; ++l)
            FSB_TRY(decode_layer(lm, lm->fast_layers[l], lm->s.fast_x, bsz, fast_k(lm, l), fast_v(lm, l),
                                 lm->fast_len, nullptr, (int)input_pos, 0, nullptr, 1));
        FSB_TRY(launch_gemv<EPI_STORE>(lm, lm->fast_out, nullptr, lm->s.fast_x, lm->D,
                                       (const float *)lm->fast_norm.ptr, nullptr, lm->s.fast_logits, lm->CS, lm->CS,
                                       lm->D, bsz, nullptr));
        FSB_CUDA_OK(cudaMemcpyAsync(logits, lm->s.fast_logits, (size_t)bsz * lm->CS * sizeof(float),
                                    cudaMemcpyDeviceToHost, st));
        FSB_CUDA_OK(cudaStreamSynchronize(st));
        return FSB_OK;
    };
    return finish(lm, body());
}

int fsb_lm_fast_embeddings(fsb_lm *lm, const uint32_t *ids, int32_t n, float *out) {
    FSB_TRY(check_handle(lm));
    FSB_REQUIRE(ids && out && n >= 1, FSB_ERR_INVALID, "fast_embeddings: bad args");
    const int D = lm->D;
    const size_t es = esize(lm->wdt);
    std::vector<uint8_t> row(D * es);
    for (int i = 0; i < n; ++i) {
        FSB_REQUIRE(ids[i] < (uint32_t)lm->CS, FSB_ERR_INVALID, "fast_embeddings: id %u out of range", ids[i]);
        FSB_CUDA_OK(cudaMemcpy(row.data(), (const uint8_t *)lm->fast_emb.ptr + (size_t)ids[i] * D * es, D * es,
                               cudaMemcpyDeviceToHost));
        for (int d = 0; d < D; ++d) {
            if (lm->wdt == FSB_F32) {
                out[(size_t)i * D + d] = reinterpret_cast<const float *>(row.data())[d];
            } else {
                uint32_t u = (uint32_t)reinterpret_cast<const uint16_t *>(row.data())[d] << 16;
                float f;
                memcpy(&f, &u, 4);
                out[(size_t)i * D + d] = f;
            }
        }
    }
    return FSB_OK;
}

int fsb_lm_clear_fast_layer_caches(fsb_lm *lm) {
    FSB_TRY(check_handle(lm));
    return FSB_OK;  // fast KV slots are addressed by codebook index and always rewritten in order
}
int fsb_lm_clear_slow_layer_caches(fsb_lm *lm) {
    FSB_TRY(check_handle(lm));
    std::fill(lm->kv_len.begin(), lm->kv_len.end(), 0);
    return FSB_OK;
}
int fsb_lm_clear_slow_caches_until(fsb_lm *lm, size_t pos) {
    FSB_TRY(check_handle(lm));
    for (auto &v : lm->kv_len) v = (int)std::min<size_t>((size_t)v, pos);  // dual_ar.rs:392-404: truncate
    return FSB_OK;
}
int fsb_lm_curr_kv_size(fsb_lm *lm, size_t *out) {
    FSB_TRY(check_handle(lm));
    FSB_REQUIRE(out, FSB_ERR_INVALID, "null out");
    *out = (size_t)lm->kv_len[0];
    return FSB_OK;
}

int fsb_lm_generate_blocking(fsb_lm *lm, const uint32_t *prompt, int32_t prompt_len, size_t max_new_tokens,
                             const fsb_sampling_args *sampling, uint32_t flags, int32_t fixed_len,
                             uint32_t *out_codes, size_t cap, size_t *out_len) {
    FSB_TRY(check_handle(lm));
    const uint32_t *prompts[1] = {prompt};
    uint32_t *outs[1] = {out_codes};
    return finish(lm, generate_impl(lm, prompts, &prompt_len, 1, max_new_tokens, sampling, flags, fixed_len, outs, cap,
                                    out_len));
}

int fsb_lm_generate_blocking_with_hidden(fsb_lm *lm, const uint32_t *prompt, int32_t prompt_len, size_t max_new_tokens,
                                         const fsb_sampling_args *sampling, uint32_t flags, int32_t fixed_len,
                                         uint32_t *out_codes, size_t cap, size_t *out_len, float *hidden,
                                         size_t hidden_cap, size_t *n_hidden) {
    FSB_TRY(check_handle(lm));
    FSB_REQUIRE(hidden && n_hidden, FSB_ERR_INVALID, "generate_blocking_with_hidden: null hidden buffer");
    if (!lm->hid) FSB_TRY(dev_alloc(lm, &lm->hid, (size_t)lm->max_batch * lm->h_st.out_cap * lm->D));
    const uint32_t *prompts[1] = {prompt};
    uint32_t *outs[1] = {out_codes};
    lm->collect_hidden = true;
    int st = generate_impl(lm, prompts, &prompt_len, 1, max_new_tokens, sampling, flags, fixed_len, outs, cap, out_len);
    lm->collect_hidden = false;
    if (st == FSB_OK) {
        int nf = 0;
        FSB_CUDA_OK(cudaMemcpy(&nf, lm->h_st.frame, sizeof(int), cudaMemcpyDeviceToHost));
        if ((size_t)nf > hidden_cap) {
            set_error("generate_blocking_with_hidden: %d frames exceed the hidden capacity %zu", nf, hidden_cap);
            st = FSB_ERR_INVALID;
        } else {
            // every yielded frame, <|im_end|> frames included (single_batch.rs:268-270): (T, 1, D)
            FSB_CUDA_OK(cudaMemcpy(hidden, lm->hid, (size_t)nf * lm->D * sizeof(float), cudaMemcpyDeviceToHost));
            *n_hidden = (size_t)nf;
        }
    }
    return finish(lm, st);
}

int fsb_lm_generate_static_batch(fsb_lm *lm, const uint32_t *const *prompts, const int32_t *prompt_lens,
                                 int32_t bsz, size_t max_new_tokens, const fsb_sampling_args *sampling,
                                 uint32_t flags, int32_t fixed_len, uint32_t *const *out_codes, size_t cap,
                                 size_t *out_lens) {
    FSB_TRY(check_handle(lm));
    return finish(lm, generate_impl(lm, prompts, prompt_lens, bsz, max_new_tokens, sampling, flags, fixed_len,
                                    out_codes, cap, out_lens));
}

int fsb_lm_last_frames(fsb_lm *lm, int32_t row, uint32_t *out, size_t cap, size_t *out_len) {
    FSB_TRY(check_handle(lm));
    FSB_REQUIRE(out && out_len && row >= 0 && row < lm->max_batch, FSB_ERR_INVALID, "last_frames: bad arguments");
    const int C1 = lm->C + 1;
    int nf = 0;
    FSB_CUDA_OK(cudaMemcpy(&nf, lm->h_st.frame + row, sizeof(int), cudaMemcpyDeviceToHost));
    FSB_REQUIRE(nf >= 0 && nf <= lm->h_st.out_cap, FSB_ERR_STATE, "last_frames: no generation on row %d", row);
    FSB_REQUIRE((size_t)nf <= cap, FSB_ERR_INVALID, "last_frames: %d frames exceed the capacity %zu", nf, cap);
    std::vector<uint32_t> h((size_t)nf * C1);
    if (nf > 0)
        FSB_CUDA_OK(cudaMemcpy(h.data(), lm->h_st.out + (size_t)row * lm->h_st.out_cap * C1, h.size() * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost));
    for (int f = 0; f < nf; ++f)
        for (int c = 0; c < C1; ++c) out[(size_t)c * cap + f] = h[(size_t)f * C1 + c];
    *out_len = (size_t)nf;
    return FSB_OK;
}

// ---------------------------------------------------------------- per-voice KV snapshots
int fsb_lm_kv_snapshot_save(fsb_lm *lm, int32_t row, size_t n_positions, fsb_kv_snapshot **out) {
    FSB_TRY(check_handle(lm));
    FSB_REQUIRE(out && row >= 0 && row < lm->max_batch, FSB_ERR_INVALID, "kv_snapshot_save: bad arguments");
    FSB_REQUIRE(n_positions >= 1 && n_positions <= (size_t)lm->kv_len[row], FSB_ERR_STATE,
                "kv_snapshot_save: row %d caches %d positions, %zu requested", row, lm->kv_len[row], n_positions);
    std::unique_ptr<fsb_kv_snapshot> sn(new fsb_kv_snapshot());
    sn->n = n_positions;
    const size_t per_layer = (size_t)lm->KV * n_positions * lm->hd;
    FSB_CUDA_OK(cudaMalloc((void **)&sn->k, per_layer * lm->NL * sizeof(float)));
    if (cudaMalloc((void **)&sn->v, per_layer * lm->NL * sizeof(float)) != cudaSuccess) {
        cudaFree(sn->k);
        set_error("kv_snapshot_save: out of memory");
        return FSB_ERR_OOM;
    }
    const size_t width = n_positions * lm->hd * sizeof(float), pitch = (size_t)lm->max_len * lm->hd * sizeof(float);
    for (int l = 0; l < lm->NL; ++l) {
        const size_t src = ((size_t)row * lm->KV) * lm->max_len * lm->hd;
        FSB_CUDA_OK(cudaMemcpy2DAsync(sn->k + l * per_layer, width, slow_k(lm, l) + src, pitch, width, lm->KV, cudaMemcpyDeviceToDevice, lm->stream));
        FSB_CUDA_OK(cudaMemcpy2DAsync(sn->v + l * per_layer, width, slow_v(lm, l) + src, pitch, width, lm->KV, cudaMemcpyDeviceToDevice, lm->stream));
    }
    FSB_CUDA_OK(cudaStreamSynchronize(lm->stream));
    *out = sn.release();
    return FSB_OK;
}

int fsb_lm_kv_snapshot_restore(fsb_lm *lm, const fsb_kv_snapshot *sn, int32_t row) {
    FSB_TRY(check_handle(lm));
    FSB_REQUIRE(sn && row >= 0 && row < lm->max_batch, FSB_ERR_INVALID, "kv_snapshot_restore: bad arguments");
    FSB_REQUIRE(sn->n <= (size_t)lm->max_len, FSB_ERR_STATE, "kv_snapshot_restore: %zu positions exceed the KV arena", sn->n);
    const size_t per_layer = (size_t)lm->KV * sn->n * lm->hd;
    const size_t width = sn->n * lm->hd * sizeof(float), pitch = (size_t)lm->max_len * lm->hd * sizeof(float);
    for (int l = 0; l < lm->NL; ++l) {
        const size_t dst = ((size_t)row * lm->KV) * lm->max_len * lm->hd;
        FSB_CUDA_OK(cudaMemcpy2DAsync(slow_k(lm, l) + dst, pitch, sn->k + l * per_layer, width, width, lm->KV, cudaMemcpyDeviceToDevice, lm->stream));
        FSB_CUDA_OK(cudaMemcpy2DAsync(slow_v(lm, l) + dst, pitch, sn->v + l * per_layer, width, width, lm->KV, cudaMemcpyDeviceToDevice, lm->stream));
    }
    FSB_CUDA_OK(cudaStreamSynchronize(lm->stream));
    lm->kv_len[row] = (int)sn->n;
    return FSB_OK;
}

int fsb_lm_kv_snapshot_free(fsb_lm *lm, fsb_kv_snapshot *sn) {
    if (!sn) return FSB_OK;
    if (lm) cudaSetDevice(lm->opt.device);
    cudaFree(sn->k);
    cudaFree(sn->v);
    delete sn;
    return FSB_OK;
}

// ---------------------------------------------------------------- continuous batching
static int session_emit(fsb_lm *lm, int row, uint32_t *out_codes, size_t cap, size_t *out_len) {
    const int C = lm->C;
    int nf = 0;
    FSB_CUDA_OK(cudaMemcpy(&nf, lm->h_st.frame + row, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> h((size_t)nf * (C + 1));
    if (nf > 0)
        FSB_CUDA_OK(cudaMemcpy(h.data(), lm->h_st.out + (size_t)row * lm->h_st.out_cap * (C + 1), h.size() * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost));
    size_t T = 0;
    for (int f = 0; f < nf; ++f) {  // generate_blocking_with_hidden: frame 0 always kept (Q4), later <|im_end|> frames dropped
        const uint32_t *fr = &h[(size_t)f * (C + 1)];
        if (f > 0 && fr[0] == lm->tok.im_end_id) continue;
        FSB_REQUIRE(T < cap, FSB_ERR_INVALID, "slot %d: output capacity %zu too small", row, cap);
        for (int c = 0; c < C; ++c) out_codes[(size_t)c * cap + T] = fr[1 + c];
        ++T;
    }
    *out_len = T;
    return FSB_OK;
}

int fsb_lm_session_begin(fsb_lm *lm, const fsb_sampling_args *sa, uint32_t flags) {
    FSB_TRY(check_handle(lm));
    FSB_REQUIRE(sa, FSB_ERR_INVALID, "session_begin: null sampling args");
    FSB_REQUIRE(lm->megab_ok && lm->max_batch >= 2 && lm->max_batch <= lm->megab_rows_cap, FSB_ERR_UNSUPPORTED,
                "sessions need the wide-batch kernel (bf16 weights, Fish 1.4/1.5 shapes, 2 <= max_batch <= 32)");
    FSB_REQUIRE(std::isfinite(sa->temp) && sa->temp >= 0.0 && std::isfinite(sa->top_p) && (sa->temp <= 1e-7 || sa->top_k >= 1) &&
                    std::isfinite(sa->repetition_penalty) && sa->repetition_penalty != 0.f,
                FSB_ERR_INVALID, "session_begin: invalid sampling arguments");
    GenState &g = lm->h_st;
    g.fixed_len = (flags & FSB_GEN_FIXED_LEN) ? 1 : 0;
    g.legacy_slow = (lm->opt.fish_version != FSB_FISH_1_5) ? 1 : 0;
    g.sp.greedy = sa->temp <= 1e-7 ? 1 : 0;
    g.sp.inv_temp = g.sp.greedy ? 1.0f : (float)(1.0 / sa->temp);
    g.sp.top_p = (float)sa->top_p;
    g.sp.top_p_gate = top_p_gate_of(sa->top_p);
    g.sp.top_k = sa->top_k;
    g.sp.penalty = sa->repetition_penalty;
    g.sp.seed = sa->seed;
    FSB_REQUIRE(g.sp.greedy || g.sp.top_k <= (uint32_t)kSelMaxK, FSB_ERR_UNSUPPORTED, "sessions need top_k <= %d", kSelMaxK);
    const int B = lm->max_batch;
    cudaStream_t st = lm->stream;
    FSB_TRY(upload_state(lm));
    FSB_CUDA_OK(cudaMemsetAsync(g.active, 0, B * sizeof(int), st));
    FSB_CUDA_OK(cudaMemsetAsync(g.eos, 0, B * sizeof(int), st));
    FSB_CUDA_OK(cudaMemsetAsync(g.frame, 0, B * sizeof(int), st));
    FSB_CUDA_OK(cudaMemsetAsync(g.n_active, 0, B * sizeof(int), st));
    FSB_CUDA_OK(cudaStreamSynchronize(st));
    lm->slot_state.assign(B, 0);
    std::fill(lm->kv_len.begin(), lm->kv_len.end(), 0);
    lm->session = true;
    return FSB_OK;
}

int fsb_lm_session_admit(fsb_lm *lm, int32_t slot, const uint32_t *prompt, int32_t P, size_t max_new_tokens,
                         int32_t fixed_len) {
    FSB_TRY(check_handle(lm));
    auto body = [&]() -> int {
        FSB_REQUIRE(lm->session, FSB_ERR_STATE, "session_admit: no session (call fsb_lm_session_begin)");
        FSB_REQUIRE(slot >= 0 && slot < lm->max_batch && prompt && P >= 1, FSB_ERR_INVALID, "session_admit: bad arguments");
        FSB_REQUIRE(lm->slot_state[slot] == 0, FSB_ERR_STATE, "session_admit: slot %d is not free", slot);
        const int C = lm->C;
        GenState &g = lm->h_st;
        for (int i = 0; i < P; ++i)
            FSB_REQUIRE(prompt[i] < (uint32_t)lm->V, FSB_ERR_INVALID, "session_admit: token id %u >= vocab_size", prompt[i]);
        for (int c = 1; c <= C; ++c)
            for (int i = 0; i < P; ++i)
                FSB_REQUIRE(prompt[(size_t)c * P + i] < (uint32_t)lm->CS, FSB_ERR_INVALID, "session_admit: code out of range");
        long long lim = (long long)max_new_tokens - P + 2;  // Q3
        if (lim < 1) lim = 1;
        if (g.fixed_len) {
            FSB_REQUIRE(fixed_len >= 1, FSB_ERR_INVALID, "session_admit: FSB_GEN_FIXED_LEN needs fixed_len >= 1");
            lim = std::min<long long>(lim, fixed_len);
        }
        FSB_REQUIRE(P + lim <= lm->max_len && lim <= g.out_cap, FSB_ERR_STATE,
                    "session_admit: %d prompt + %lld frames exceed the KV arena (%d positions)", P, lim, lm->max_len);
        cudaStream_t st = lm->stream;
        FSB_CUDA_OK(cudaMemcpyAsync(lm->s.toks, prompt, (size_t)(C + 1) * P * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        lm->kv_len[slot] = 0;
        FSB_TRY(prefill_row(lm, lm->s.toks, P, slot, 0, 0));
        int *hp = lm->h_pin;
        hp[0] = 1; hp[1] = (int)lim; hp[2] = P; hp[3] = 0; hp[4] = 1;
        FSB_CUDA_OK(cudaMemcpyAsync(g.active + slot, hp, sizeof(int), cudaMemcpyHostToDevice, st));
        FSB_CUDA_OK(cudaMemcpyAsync(g.max_frames + slot, hp + 1, sizeof(int), cudaMemcpyHostToDevice, st));
        FSB_CUDA_OK(cudaMemcpyAsync(g.pos + slot, hp + 2, sizeof(int), cudaMemcpyHostToDevice, st));
        FSB_CUDA_OK(cudaMemcpyAsync(g.eos + slot, hp + 3, sizeof(int), cudaMemcpyHostToDevice, st));
        FSB_CUDA_OK(cudaMemcpyAsync(g.frame + slot, hp + 3, sizeof(int), cudaMemcpyHostToDevice, st));
        FSB_CUDA_OK(cudaMemcpyAsync(g.n_active + 1, hp + 4, sizeof(int), cudaMemcpyHostToDevice, st));  // live-row counter of the launch below
        FSB_CUDA_OK(cudaMemsetAsync(g.rep + (size_t)slot * C, 0, (size_t)C * sizeof(RepPenState), st));
        FSB_CUDA_OK(cudaStreamSynchronize(st));  // hp is reused
        // the frame produced from the prompt itself (slow head on the prefilled state + the C fast steps), by the
        // single-row kernel on this row alone; the slot then joins the wide launches in resume mode
        FSB_TRY(mega_launch(lm, slot, 1, 1, 1, true));
        FSB_CUDA_OK(cudaMemcpyAsync(hp, g.active + slot, sizeof(int), cudaMemcpyDeviceToHost, st));
        FSB_CUDA_OK(cudaStreamSynchronize(st));
        lm->slot_state[slot] = hp[0] ? 1 : 2;
        lm->kv_len[slot] = P;
        return FSB_OK;
    };
    return finish(lm, body());
}

int fsb_lm_session_run(fsb_lm *lm, int32_t max_frames, int32_t *active, int32_t *n_active) {
    FSB_TRY(check_handle(lm));
    auto body = [&]() -> int {
        FSB_REQUIRE(lm->session, FSB_ERR_STATE, "session_run: no session");
        FSB_REQUIRE(max_frames >= 1, FSB_ERR_INVALID, "session_run: max_frames must be >= 1");
        const int B = lm->max_batch;
        int live = 0;
        for (int b = 0; b < B; ++b) live += lm->slot_state[b] == 1;
        cudaStream_t st = lm->stream;
        int *hp = lm->h_pin;
        if (live > 0) {
            hp[0] = live;
            FSB_CUDA_OK(cudaMemcpyAsync(lm->h_st.n_active, hp, sizeof(int), cudaMemcpyHostToDevice, st));
            FSB_CUDA_OK(cudaStreamSynchronize(st));
            FSB_TRY(megab_launch_rows(lm, B, max_frames, false));  // resume mode: every frame starts with the slow stack
            FSB_CUDA_OK(cudaMemcpyAsync(hp, lm->h_st.active, B * sizeof(int), cudaMemcpyDeviceToHost, st));
            FSB_CUDA_OK(cudaMemcpyAsync(hp + B, lm->h_st.pos, B * sizeof(int), cudaMemcpyDeviceToHost, st));
            FSB_CUDA_OK(cudaStreamSynchronize(st));
            live = 0;
            for (int b = 0; b < B; ++b)
                if (lm->slot_state[b] == 1) {
                    lm->kv_len[b] = hp[B + b];
                    if (!hp[b]) lm->slot_state[b] = 2;
                    else ++live;
                }
        }
        if (active)
            for (int b = 0; b < B; ++b) active[b] = lm->slot_state[b] == 1 ? 1 : 0;
        if (n_active) *n_active = live;
        return FSB_OK;
    };
    return finish(lm, body());
}

int fsb_lm_session_collect(fsb_lm *lm, int32_t slot, uint32_t *out_codes, size_t cap, size_t *out_len) {
    FSB_TRY(check_handle(lm));
    FSB_REQUIRE(lm->session && slot >= 0 && slot < lm->max_batch && out_codes && out_len, FSB_ERR_INVALID, "session_collect: bad arguments");
    FSB_REQUIRE(lm->slot_state[slot] == 2, FSB_ERR_STATE, "session_collect: slot %d has not finished", slot);
    FSB_TRY(session_emit(lm, slot, out_codes, cap, out_len));
    lm->slot_state[slot] = 0;
    return FSB_OK;
}

int fsb_lm_set_profile(fsb_lm *lm, int on) {
    FSB_TRY(check_handle(lm));
    lm->profile = on != 0;
    if (lm->profile && lm->prof_ev.empty()) {
        lm->prof_ev.resize(2 * 4096);
        for (auto &e : lm->prof_ev) FSB_CUDA_OK(cudaEventCreate(&e));
    }
    return FSB_OK;
}

int fsb_lm_get_stats(fsb_lm *lm, fsb_lm_stats *out) {
    FSB_REQUIRE(lm && out, FSB_ERR_INVALID, "null argument");
    *out = lm->stats;
    return FSB_OK;
}

// ---------------------------------------------------------------- operator level
int fsb_op_repeat_kv(const void *src_dev, void *dst_dev, int32_t dtype, int32_t n_local_heads, int32_t n_rep,
                     int32_t seqlen, int32_t head_dim, void *stream) {
    FSB_REQUIRE(src_dev && dst_dev, FSB_ERR_INVALID, "repeat_kv: null pointer");
    FSB_REQUIRE(n_local_heads > 0 && n_rep > 0 && seqlen > 0 && head_dim > 0, FSB_ERR_INVALID, "repeat_kv: bad dims");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(seqlen, n_local_heads);
    const int threads = std::min(256, ((head_dim + 31) / 32) * 32);
    switch (dtype) {
        case FSB_F32:
        case FSB_U32:
            repeat_kv_kernel<uint32_t><<<grid, threads, 0, st>>>((const uint32_t *)src_dev, (uint32_t *)dst_dev, n_rep,
                                                                 seqlen, head_dim);
            break;
        case FSB_BF16:
        case FSB_F16:
            repeat_kv_kernel<uint16_t><<<grid, threads, 0, st>>>((const uint16_t *)src_dev, (uint16_t *)dst_dev, n_rep,
                                                                 seqlen, head_dim);
            break;
        case FSB_I64:
        case FSB_F64:
            repeat_kv_kernel<uint64_t><<<grid, threads, 0, st>>>((const uint64_t *)src_dev, (uint64_t *)dst_dev, n_rep,
                                                                 seqlen, head_dim);
            break;
        case FSB_U8:
            repeat_kv_kernel<uint8_t><<<grid, threads, 0, st>>>((const uint8_t *)src_dev, (uint8_t *)dst_dev, n_rep,
                                                                seqlen, head_dim);
            break;
        default:
            set_error("repeat_kv: unsupported dtype %d", dtype);
            return FSB_ERR_INVALID;
    }
    FSB_CUDA_OK(cudaGetLastError());
    return FSB_OK;
}

size_t fsb_op_gqa_decode_attn_scratch_bytes(int32_t bsz, int32_t n_head, int32_t head_dim) {
    return (size_t)bsz * n_head * 16 * (head_dim + 2) * sizeof(float) + (size_t)bsz * n_head * head_dim * sizeof(float);
}

int fsb_op_gqa_decode_attn(const float *qkv_dev, float *kcache_dev, float *vcache_dev, const float *cos_dev,
                           const float *sin_dev, const int32_t *pos_dev, int32_t bsz, int32_t n_head,
                           int32_t n_local_heads, int32_t head_dim, int32_t max_len, float *out_dev,
                           void *scratch_dev, size_t scratch_bytes, void *stream) {
    FSB_REQUIRE(qkv_dev && kcache_dev && vcache_dev && cos_dev && sin_dev && pos_dev && out_dev && scratch_dev,
                FSB_ERR_INVALID, "gqa_decode_attn: null pointer");
    FSB_REQUIRE(head_dim == 64, FSB_ERR_UNSUPPORTED, "gqa_decode_attn: head_dim must be 64");
    FSB_REQUIRE(n_local_heads > 0 && n_head % n_local_heads == 0 && n_head / n_local_heads <= kAttnMaxRep,
                FSB_ERR_UNSUPPORTED, "gqa_decode_attn: unsupported head configuration");
    FSB_REQUIRE(scratch_bytes >= fsb_op_gqa_decode_attn_scratch_bytes(bsz, n_head, head_dim), FSB_ERR_INVALID,
                "gqa_decode_attn: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int nsplit = 16, n_rep = n_head / n_local_heads;
    float *partial = (float *)scratch_dev;
    float *q = partial + (size_t)bsz * n_head * nsplit * (head_dim + 2);
    rope_append_kernel<<<bsz, 256, 0, st>>>(qkv_dev, q, kcache_dev, vcache_dev, cos_dev, sin_dev, pos_dev, 0, 0,
                                            n_head, n_local_heads, head_dim, max_len, nullptr);
    attn_decode_split_kernel<<<dim3(nsplit, n_local_heads, bsz), n_rep * 32, n_rep * head_dim * sizeof(float), st>>>(
        q, kcache_dev, vcache_dev, pos_dev, 0, n_head, n_local_heads, head_dim, max_len,
        1.0f / sqrtf((float)head_dim), partial, nullptr);
    attn_decode_combine_kernel<<<bsz * n_head, head_dim, 0, st>>>(partial, nsplit, head_dim, out_dev, nullptr);
    FSB_CUDA_OK(cudaGetLastError());
    return FSB_OK;
}

}  // extern "C"
