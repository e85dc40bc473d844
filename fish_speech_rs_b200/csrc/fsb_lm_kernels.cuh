// Per-op kernels of the dual-AR token loop (decode_mode 1) and of prefill.
// Reference call sites replaced are cited per kernel (paths relative to the
// reference root, file fish_speech_core/lib/lm/dual_ar.rs unless noted).
#pragma once
#include "fsb_common.cuh"
#include "fsb_sample.cuh"

namespace fsb {

// ------------------------------------------------------------------ embed
// DualARTransformer::embed, :532-567.  toks(b, c, s) = toks[(b*(C+1) + c)*S + s].
template <typename WT>
__global__ void embed_sum_kernel(const uint32_t *__restrict__ toks, int S, int C, int D, int codebook_size,
                                 const WT *__restrict__ emb, const WT *__restrict__ cb_emb, uint32_t sem_start,
                                 uint32_t sem_end, int has_end, float *__restrict__ x, const int *n_active) {
    if (n_active && *n_active == 0) return;
    const int row = blockIdx.x;  // b*S + s
    const int b = row / S, s = row % S;
    const uint32_t *t = toks + (size_t)b * (C + 1) * S + s;
    const uint32_t tok0 = t[0];
    const bool m = has_end ? (tok0 <= sem_end && tok0 >= sem_start) : (tok0 == sem_start);
    const float mf = m ? 1.f : 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = to_f32(emb[(size_t)tok0 * D + d]);
        for (int c = 0; c < C; ++c) {
            uint32_t code = t[(size_t)(c + 1) * S];
            acc = __fadd_rn(acc, __fmul_rn(to_f32(cb_emb[((size_t)c * codebook_size + code) * D + d]), mf));
        }
        x[(size_t)row * D + d] = acc;
    }
}

// ------------------------------------------------------------------ hidden-state collection
// generate_blocking_with_hidden, single_batch.rs:251,268-270: the pre-norm slow hidden state of every yielded frame
// (the one handed to the fast stack, Q1).  hid: (B, out_cap, D); frame index = frames emitted so far.
__global__ void store_hidden_kernel(const float *__restrict__ hidden, const GenState *st, float *__restrict__ hid, int D) {
    if (*st->n_active == 0) return;
    const int b = blockIdx.x;
    if (!st->active[b]) return;
    const int f = st->frame[b];
    if (f >= st->out_cap) return;
    for (int d = threadIdx.x; d < D; d += blockDim.x)
        hid[((size_t)b * st->out_cap + f) * D + d] = hidden[(size_t)b * D + d];
}

// ------------------------------------------------------------------ GEMV
// y[b, r] = epilogue( W[r, :] . xn[b, :] ),  xn = rms_norm(x[b]) * g  (optional)
// replaces RmsNorm + Linear (+ residual / silu*mul), :160-165,289,383,429-440.
// One warp streams two weight rows with 16-byte loads; x lives in shared memory.
enum { EPI_STORE = 0, EPI_RESID = 1, EPI_SWIGLU = 2 };

struct GemvArgs {
    const void *W;       // (rows, K)
    const void *W3;      // EPI_SWIGLU: second matrix (w3)
    const float *x;      // (NB, ldx)
    const float *norm_w; // (K) f32 or null
    const float *resid;  // (NB, ldy) for EPI_RESID
    float *y;            // (NB, ldy)
    int rows, K, ldx, ldy;
    float eps;
    int row0;            // logical row 0 reads weight row `row0`, row r>=1 reads `rest_base + r - 1`
    int rest_base;
    const int *n_active;
};

constexpr int kGemvThreads = 128;
constexpr int kGemvRowsPerWarp = 2;
constexpr int kGemvRowsPerCta = (kGemvThreads / 32) * kGemvRowsPerWarp;

template <typename WT>
__device__ __forceinline__ void load_w8(const WT *p, float (&w)[8]);
template <>
__device__ __forceinline__ void load_w8<float>(const float *p, float (&w)[8]) {
    float4 a = ldg_stream4(p), b = ldg_stream4(p + 4);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
template <>
__device__ __forceinline__ void load_w8<__nv_bfloat16>(const __nv_bfloat16 *p, float (&w)[8]) {
    uint4 u = ldg_stream_u4(p);
    w[0] = bf16lo(u.x); w[1] = bf16hi(u.x); w[2] = bf16lo(u.y); w[3] = bf16hi(u.y);
    w[4] = bf16lo(u.z); w[5] = bf16hi(u.z); w[6] = bf16lo(u.w); w[7] = bf16hi(u.w);
}

template <typename WT, int NB, int EPI>
__global__ void __launch_bounds__(kGemvThreads) gemv_kernel(GemvArgs a) {
    if (a.n_active && *a.n_active == 0) return;
    extern __shared__ float xs[];  // NB * K
    __shared__ float red[kGemvThreads / 32][NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = a.K;
    // ---- prologue: stage x, optional rms_norm ----
    float ss[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) ss[b] = 0.f;
    for (int k = tid; k < K; k += kGemvThreads) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float v = a.x[(size_t)b * a.ldx + k];
            xs[b * K + k] = v;
            ss[b] += v * v;
        }
    }
    if (a.norm_w) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float v = warp_sum(ss[b]);
            if (lane == 0) red[warp][b] = v;
        }
        __syncthreads();
        const float *g = a.norm_w;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < kGemvThreads / 32; ++w) tot += red[w][b];
            const float denom = sqrtf(tot / (float)K + a.eps);
            for (int k = tid; k < K; k += kGemvThreads)
                xs[b * K + k] = __fmul_rn(__fdiv_rn(xs[b * K + k], denom), g[k]);
        }
    }
    __syncthreads();
    // ---- main: 2 rows per warp ----
    const WT *W = reinterpret_cast<const WT *>(a.W);
    const WT *W3 = reinterpret_cast<const WT *>(a.W3);
    for (int r0 = (blockIdx.x * (kGemvThreads / 32) + warp) * kGemvRowsPerWarp; r0 < a.rows;
         r0 += gridDim.x * kGemvRowsPerCta) {
        const int r1 = r0 + 1;
        const bool has1 = r1 < a.rows;
        const size_t wr0 = (size_t)(r0 == 0 ? a.row0 : a.rest_base + r0 - 1);
        const size_t wr1 = (size_t)(has1 ? (a.rest_base + r1 - 1) : wr0);
        const WT *p0 = W + wr0 * K, *p1 = W + wr1 * K;
        float acc0[NB], acc1[NB], acc30[NB], acc31[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) acc0[b] = acc1[b] = acc30[b] = acc31[b] = 0.f;
#pragma unroll 4
        for (int k = lane * 8; k < K; k += 256) {
            float w0[8], w1[8];
            load_w8<WT>(p0 + k, w0);
            load_w8<WT>(p1 + k, w1);
            float v0[8], v1[8];
            if (EPI == EPI_SWIGLU) {
                load_w8<WT>(W3 + wr0 * K + k, v0);
                load_w8<WT>(W3 + wr1 * K + k, v1);
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const float4 xa = *reinterpret_cast<const float4 *>(&xs[b * K + k]);
                const float4 xb = *reinterpret_cast<const float4 *>(&xs[b * K + k + 4]);
                const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc0[b] = fmaf(w0[j], xv[j], acc0[b]);
                    acc1[b] = fmaf(w1[j], xv[j], acc1[b]);
                    if (EPI == EPI_SWIGLU) {
                        acc30[b] = fmaf(v0[j], xv[j], acc30[b]);
                        acc31[b] = fmaf(v1[j], xv[j], acc31[b]);
                    }
                }
            }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float s0 = warp_sum(acc0[b]), s1 = warp_sum(acc1[b]);
            float t0 = 0.f, t1 = 0.f;
            if (EPI == EPI_SWIGLU) { t0 = warp_sum(acc30[b]); t1 = warp_sum(acc31[b]); }
            if (lane == 0) {
                float *yo = a.y + (size_t)b * a.ldy;
                if (EPI == EPI_STORE) {
                    yo[r0] = s0;
                    if (has1) yo[r1] = s1;
                } else if (EPI == EPI_RESID) {
                    const float *rs = a.resid + (size_t)b * a.ldy;
                    yo[r0] = __fadd_rn(rs[r0], s0);
                    if (has1) yo[r1] = __fadd_rn(rs[r1], s1);
                } else {
                    yo[r0] = __fmul_rn(silu_f(s0), t0);
                    if (has1) yo[r1] = __fmul_rn(silu_f(s1), t1);
                }
            }
        }
    }
}

// ------------------------------------------------------------------ RoPE + KV append (decode, one token per row)
// rope_i (:246-247) on q and k, Tensor::cat replaced by an in-place write (:316-324).
// qkv (B, (H+2KV)*hd) -> q (B, H*hd) roped; K/V cache row `pos`.
// cache layout: (B, KV, max_len, hd).  pos = pos_ptr ? pos_ptr[b] : pos_imm.
__global__ void rope_append_kernel(const float *__restrict__ qkv, float *__restrict__ q, float *__restrict__ kc,
                                   float *__restrict__ vc, const float *__restrict__ cosT,
                                   const float *__restrict__ sinT, const int *pos_ptr, int pos_imm, int rope_delta,
                                   int H, int KV, int hd, int max_len, const int *n_active,
                                   const int *active = nullptr) {
    if (n_active && *n_active == 0) return;
    const int b = blockIdx.x;
    // a finished row of a ragged batch sits at pos == its budget (possibly == max_len): it must not append
    if (active && !active[b]) return;
    const int pos = pos_ptr ? pos_ptr[b] : pos_imm;  // cache slot
    const int rpos = pos + rope_delta;                // RoPE row (== input_pos)
    const int half = hd / 2;
    const float *src = qkv + (size_t)b * (H + 2 * KV) * hd;
    const int n_q = H * half, n_k = KV * half, n_v = KV * hd;
    for (int i = threadIdx.x; i < n_q + n_k + n_v; i += blockDim.x) {
        if (i < n_q + n_k) {
            const bool is_q = i < n_q;
            const int j = is_q ? i : i - n_q;
            const int h = j / half, p = j % half;
            const float *s = src + (is_q ? 0 : H * hd) + h * hd + 2 * p;
            const float c = cosT[(size_t)rpos * half + p], sn = sinT[(size_t)rpos * half + p];
            const float x0 = s[0], x1 = s[1];
            const float o0 = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, sn));
            const float o1 = __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, c));
            float *dst = is_q ? (q + (size_t)b * H * hd + h * hd + 2 * p)
                              : (kc + (((size_t)b * KV + h) * max_len + pos) * hd + 2 * p);
            dst[0] = o0;
            dst[1] = o1;
        } else {
            const int j = i - n_q - n_k;
            const int h = j / hd, d = j % hd;
            vc[(((size_t)b * KV + h) * max_len + pos) * hd + d] = src[(H + KV) * hd + h * hd + d];
        }
    }
}

// ------------------------------------------------------------------ decode attention (split-KV, GQA shared K/V)
// replaces repeat_kv + scaled_dot_product_attention (:252-279,327-376; unary.cu:8-58):
// the 8 query heads of a KV group read each cached K/V row once.
// grid (nsplit, KV, B), block = n_rep warps.  partial: (B, H, nsplit, hd + 2).
constexpr int kAttnMaxRep = 8;
__global__ void attn_decode_split_kernel(const float *__restrict__ q, const float *__restrict__ kc,
                                         const float *__restrict__ vc, const int *pos_ptr, int pos_imm, int H,
                                         int KV, int hd, int max_len, float scale, float *__restrict__ partial,
                                         const int *n_active, const int *active = nullptr) {
    if (n_active && *n_active == 0) return;
    if (active && !active[blockIdx.z]) return;  // finished row: nothing cached at `pos` (see rope_append_kernel)
    extern __shared__ float qs[];  // n_rep * hd
    const int split = blockIdx.x, nsplit = gridDim.x, kvh = blockIdx.y, b = blockIdx.z;
    const int n_rep = H / KV;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = kvh * n_rep + warp;
    const int len = (pos_ptr ? pos_ptr[b] : pos_imm) + 1;
    int chunk = (len + nsplit - 1) / nsplit;
    chunk = (chunk + 31) & ~31;
    const int j0 = split * chunk, j1 = min(len, j0 + chunk);
    for (int i = threadIdx.x; i < n_rep * hd; i += blockDim.x)
        qs[i] = q[(size_t)b * H * hd + (size_t)kvh * n_rep * hd + i];
    __syncthreads();
    const float *qh = qs + warp * hd;
    const float *kb = kc + ((size_t)b * KV + kvh) * max_len * hd;
    const float *vb = vc + ((size_t)b * KV + kvh) * max_len * hd;
    float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;  // lane owns dims lane, lane+32 (hd == 64)
    for (int t = j0; t < j1; t += 32) {
        const int j = t + lane;
        float sc = -INFINITY;
        if (j < j1) {
            const float4 *kr = reinterpret_cast<const float4 *>(kb + (size_t)j * hd);
            float acc = 0.f;
#pragma unroll
            for (int d4 = 0; d4 < 16; ++d4) {
                float4 kk = kr[d4];
                acc = fmaf(qh[4 * d4 + 0], kk.x * scale, acc);
                acc = fmaf(qh[4 * d4 + 1], kk.y * scale, acc);
                acc = fmaf(qh[4 * d4 + 2], kk.z * scale, acc);
                acc = fmaf(qh[4 * d4 + 3], kk.w * scale, acc);
            }
            sc = acc;
        }
        const float m_new = fmaxf(m, warp_max(sc));
        const float corr = expf(m - m_new);  // exp(-inf) == 0 on the first tile
        const float p = (j < j1) ? expf(sc - m_new) : 0.f;
        l = l * corr + warp_sum(p);
        o0 *= corr;
        o1 *= corr;
        const int cnt = min(32, j1 - t);
        for (int jj = 0; jj < cnt; ++jj) {
            const float pj = __shfl_sync(0xffffffffu, p, jj);
            const float *vr = vb + (size_t)(t + jj) * hd;
            o0 = fmaf(pj, vr[lane], o0);
            o1 = fmaf(pj, vr[lane + 32], o1);
        }
        m = m_new;
    }
    float *out = partial + (((size_t)b * H + h) * nsplit + split) * (hd + 2);
    out[lane] = o0;
    out[lane + 32] = o1;
    if (lane == 0) { out[hd] = m; out[hd + 1] = l; }
}

// grid (B*H), block hd.  y (B, H*hd).
__global__ void attn_decode_combine_kernel(const float *__restrict__ partial, int nsplit, int hd,
                                           float *__restrict__ y, const int *n_active) {
    if (n_active && *n_active == 0) return;
    const int bh = blockIdx.x, d = threadIdx.x;
    const float *p = partial + (size_t)bh * nsplit * (hd + 2);
    float M = -INFINITY;
    for (int s = 0; s < nsplit; ++s) M = fmaxf(M, p[s * (hd + 2) + hd]);
    float L = 0.f, o = 0.f;
    for (int s = 0; s < nsplit; ++s) {
        const float ms = p[s * (hd + 2) + hd];
        if (ms == -INFINITY) continue;
        const float w = expf(ms - M);
        L = fmaf(p[s * (hd + 2) + hd + 1], w, L);
        o = fmaf(p[s * (hd + 2) + d], w, o);
    }
    y[(size_t)bh * hd + d] = o / L;
}

// ------------------------------------------------------------------ prefill kernels (S > 1, one sequence)
// rms_norm over rows: one warp per row.
__global__ void rmsnorm_rows_kernel(const float *__restrict__ x, const float *__restrict__ g, float eps, int M, int D,
                                    float *__restrict__ y) {
    const int row = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float *xr = x + (size_t)row * D;
    float ss = 0.f;
    for (int k = lane; k < D; k += 32) ss += xr[k] * xr[k];
    ss = warp_sum(ss);
    const float denom = sqrtf(ss / (float)D + eps);
    for (int k = lane; k < D; k += 32) y[(size_t)row * D + k] = __fmul_rn(__fdiv_rn(xr[k], denom), g[k]);
}

// C[M, N] = A[M, K] . W[N, K]^T (+ resid), fp32 accumulate on CUDA cores.
// 64x64 tile, BK 16, 256 threads, 4x4 micro-tile.  Used for prefill and for
// decode batches too wide for the GEMV path; the fp32-parity companion of the
// tcgen05 path.
constexpr int kGemmBM = 64, kGemmBN = 64, kGemmBK = 16;
template <typename WT, int EPI>
__global__ void __launch_bounds__(256) gemm_nt_kernel(const float *__restrict__ A, const WT *__restrict__ W,
                                                      const float *__restrict__ resid, float *__restrict__ Cm, int M,
                                                      int N, int K) {
    __shared__ float As[kGemmBK][kGemmBM + 4];
    __shared__ float Ws[kGemmBK][kGemmBN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * kGemmBN;
    const int tx = tid % 16, ty = tid / 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid / 4, lk = (tid % 4) * 4;  // 64 rows x 16 k, 4 consecutive k per thread
    for (int k0 = 0; k0 < K; k0 += kGemmBK) {
        {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (m0 + lr < M) {
                const float4 t = *reinterpret_cast<const float4 *>(A + (size_t)(m0 + lr) * K + k0 + lk);
                v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) As[lk + i][lr] = v[i];
            float w[4] = {0.f, 0.f, 0.f, 0.f};
            if (n0 + lr < N) {
                const WT *wp = W + (size_t)(n0 + lr) * K + k0 + lk;
#pragma unroll
                for (int i = 0; i < 4; ++i) w[i] = to_f32(wp[i]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) Ws[lk + i][lr] = w[i];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kGemmBK; ++kk) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (EPI == EPI_RESID) v = __fadd_rn(resid[(size_t)m * N + n], v);
            Cm[(size_t)m * N + n] = v;
        }
    }
}

// h = silu(g1) * g3   (g1 = x . w1^T, g3 = x . w3^T), in place into g1 allowed
__global__ void swiglu_rows_kernel(const float *g1, const float *__restrict__ g3, size_t n, float *h) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    h[idx] = __fmul_rn(silu_f(g1[idx]), g3[idx]);
}

// RoPE + KV write for S consecutive positions of row `b` starting at pos0.
// qkv (S, (H+2KV)*hd) -> q (S, H*hd) roped, caches at pos0 + s.
__global__ void rope_append_rows_kernel(const float *__restrict__ qkv, float *__restrict__ q, float *__restrict__ kc,
                                        float *__restrict__ vc, const float *__restrict__ cosT,
                                        const float *__restrict__ sinT, int b, int pos0, int rope_delta, int H,
                                        int KV, int hd, int max_len, const int4 *segs = nullptr) {
    // batched form (prefill of several rows in one pass): blockIdx.y = segment {row b, cache pos0, positions n, first
    // position of the segment inside the pass}
    if (segs) {
        const int4 sg = segs[blockIdx.y];
        if ((int)blockIdx.x >= sg.z) return;
        b = sg.x;
        pos0 = sg.y;
        qkv += (size_t)sg.w * (H + 2 * KV) * hd;
        q += (size_t)sg.w * H * hd;
    }
    const int s = blockIdx.x;
    const int pos = pos0 + s;
    const int rpos = pos + rope_delta;
    const int half = hd / 2;
    const float *src = qkv + (size_t)s * (H + 2 * KV) * hd;
    const int n_q = H * half, n_k = KV * half, n_v = KV * hd;
    for (int i = threadIdx.x; i < n_q + n_k + n_v; i += blockDim.x) {
        if (i < n_q + n_k) {
            const bool is_q = i < n_q;
            const int j = is_q ? i : i - n_q;
            const int h = j / half, p = j % half;
            const float *sp = src + (is_q ? 0 : H * hd) + h * hd + 2 * p;
            const float c = cosT[(size_t)rpos * half + p], sn = sinT[(size_t)rpos * half + p];
            const float x0 = sp[0], x1 = sp[1];
            const float o0 = __fsub_rn(__fmul_rn(x0, c), __fmul_rn(x1, sn));
            const float o1 = __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, c));
            float *dst = is_q ? (q + (size_t)s * H * hd + h * hd + 2 * p)
                              : (kc + (((size_t)b * KV + h) * max_len + pos) * hd + 2 * p);
            dst[0] = o0;
            dst[1] = o1;
        } else {
            const int j = i - n_q - n_k;
            const int h = j / hd, d = j % hd;
            vc[(((size_t)b * KV + h) * max_len + pos) * hd + d] = src[(H + KV) * hd + h * hd + d];
        }
    }
}

// Causal attention with KV offset for prefill (get_mask_abs, :702-712: query s sees
// cached positions [0, pos0 + s]).  grid (ceil(S/8), KV), block 512: one CTA = 8 consecutive queries x the
// n_rep <= 8 query heads that share one KV head (GQA: replaces repeat_kv); warp w = (query w / n_rep... see below).
// K / V tiles of 32 positions are staged in shared memory once (coalesced) and reused by all 16 warps.
constexpr int kPrefQ = 8;          // queries per CTA
constexpr int kPrefStride = 65;    // floats per staged K row (conflict-free column reads)
constexpr int kPrefIPW = 4;        // (query, head) items per warp kept in registers while the K / V tiles stream by
__global__ void __launch_bounds__(512) attn_prefill_kernel(const float *__restrict__ q, const float *__restrict__ kc,
                                                           const float *__restrict__ vc, int b, int pos0, int S, int H, int KV,
                                                           int hd, int max_len, float scale, float *__restrict__ y,
                                                           const int4 *segs = nullptr) {
    // batched form: blockIdx.z = segment (see rope_append_rows_kernel); one launch covers every row of a prefill pass
    if (segs) {
        const int4 sg = segs[blockIdx.z];
        if ((int)blockIdx.x * kPrefQ >= sg.z) return;
        b = sg.x;
        pos0 = sg.y;
        S = sg.z;
        q += (size_t)sg.w * H * hd;
        y += (size_t)sg.w * H * hd;
    }
    // hd == 64 (host-checked).  Work items = kPrefQ queries x n_rep heads; 16 warps x kPrefIPW items per round.
    __shared__ float ks[32 * kPrefStride];
    __shared__ float vs[32 * 64];
    __shared__ float qsm[16 * kPrefIPW][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_rep = H / KV, kvh = blockIdx.y;
    const int s0 = blockIdx.x * kPrefQ;
    const int nq = min(kPrefQ, S - s0);
    const int nitems = nq * n_rep;
    const float *kb = kc + ((size_t)b * KV + kvh) * max_len * hd;
    const float *vb = vc + ((size_t)b * KV + kvh) * max_len * hd;
    const int len_max = pos0 + s0 + nq;  // the last query of the CTA sees this many positions
    for (int it0 = 0; it0 < nitems; it0 += 16 * kPrefIPW) {
        int sq[kPrefIPW], hh[kPrefIPW];
        float m[kPrefIPW], l[kPrefIPW], o0[kPrefIPW], o1[kPrefIPW];
        __syncthreads();  // qsm of the previous round consumed
#pragma unroll
        for (int u = 0; u < kPrefIPW; ++u) {
            const int item = it0 + u * 16 + warp;
            const bool active = item < nitems;
            sq[u] = active ? s0 + item / n_rep : -1;
            hh[u] = kvh * n_rep + (active ? item % n_rep : 0);
            m[u] = -INFINITY; l[u] = 0.f; o0[u] = 0.f; o1[u] = 0.f;
            if (active) {
                qsm[u * 16 + warp][lane] = q[((size_t)sq[u] * H + hh[u]) * hd + lane];
                qsm[u * 16 + warp][lane + 32] = q[((size_t)sq[u] * H + hh[u]) * hd + lane + 32];
            }
        }
        for (int t = 0; t < len_max; t += 32) {
            __syncthreads();  // previous tile consumed (and qsm written)
            for (int i = threadIdx.x; i < 32 * 16; i += 512) {
                const int j = i >> 4, sg = i & 15;
                float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
                if (t + j < len_max) {
                    kk = *reinterpret_cast<const float4 *>(kb + (size_t)(t + j) * hd + sg * 4);
                    vv = *reinterpret_cast<const float4 *>(vb + (size_t)(t + j) * hd + sg * 4);
                }
                float *kd = ks + j * kPrefStride + sg * 4;
                kd[0] = kk.x; kd[1] = kk.y; kd[2] = kk.z; kd[3] = kk.w;
                *reinterpret_cast<float4 *>(vs + j * 64 + sg * 4) = vv;
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < kPrefIPW; ++u) {
                const int len = pos0 + sq[u] + 1;  // sq < 0: inactive item
                if (sq[u] < 0 || t >= len) continue;
                const int j = t + lane;
                float sc = -INFINITY;
                if (j < len) {
                    const float *kr = ks + lane * kPrefStride, *qh = qsm[u * 16 + warp];
                    float acc = 0.f;
#pragma unroll
                    for (int d = 0; d < 64; ++d) acc = fmaf(qh[d], kr[d] * scale, acc);
                    sc = acc;
                }
                const float m_new = fmaxf(m[u], warp_max(sc));
                const float corr = expf(m[u] - m_new);
                const float pr = (j < len) ? expf(sc - m_new) : 0.f;
                l[u] = l[u] * corr + warp_sum(pr);
                o0[u] *= corr;
                o1[u] *= corr;
                const int cnt = min(32, len - t);
                for (int jj = 0; jj < cnt; ++jj) {
                    const float pj = __shfl_sync(0xffffffffu, pr, jj);
                    o0[u] = fmaf(pj, vs[jj * 64 + lane], o0[u]);
                    o1[u] = fmaf(pj, vs[jj * 64 + lane + 32], o1[u]);
                }
                m[u] = m_new;
            }
        }
#pragma unroll
        for (int u = 0; u < kPrefIPW; ++u) {
            if (sq[u] >= 0) {
                y[((size_t)sq[u] * H + hh[u]) * hd + lane] = o0[u] / l[u];
                y[((size_t)sq[u] * H + hh[u]) * hd + lane + 32] = o1[u] / l[u];
            }
        }
    }
}

// Register-blocked variant for the Fish shapes (head_dim 64, 8 query heads per KV head): one CTA = 16 consecutive queries x
// the 8 heads of a KV group = 128 (query, head) rows; K / V tiles of 32 positions in shared memory.  Thread (ty, tx) owns
// rows 4 ty .. 4 ty + 3 (ONE query, four heads) in both products: S = Q K^T as a 4 x 4 block (columns tx + 8 j), then
// O += P V as a 4 x 8 block (dims 4 tx .. and 32 + 4 tx ..), P handed over through shared memory within the 8 lanes that own
// the rows (no block barrier between the two products).  ~9 FMAs per 16-byte shared-memory load against ~0.7 in
// attn_prefill_kernel (one (query, head) item per warp, a full dot product per lane).
constexpr int kPf2Q = 16, kPf2Threads = 256, kPf2QS = 68, kPf2KS = 68, kPf2PS = 36;
constexpr int kPf2SmemFloats = 128 * kPf2QS + 32 * kPf2KS + 32 * 64 + 128 * kPf2PS;
__global__ void __launch_bounds__(kPf2Threads) attn_prefill8_kernel(const float *__restrict__ q, const float *__restrict__ kc,
                                                                    const float *__restrict__ vc, int b, int pos0, int S, int H,
                                                                    int KV, int max_len, float *__restrict__ y,
                                                                    const int4 *segs = nullptr) {
    if (segs) {
        const int4 sg = segs[blockIdx.z];
        if ((int)blockIdx.x * kPf2Q >= sg.z) return;
        b = sg.x;
        pos0 = sg.y;
        S = sg.z;
        q += (size_t)sg.w * H * 64;
        y += (size_t)sg.w * H * 64;
    }
    extern __shared__ float pf2_smem[];
    float *Qs = pf2_smem, *Ks = Qs + 128 * kPf2QS, *Vs = Ks + 32 * kPf2KS, *Ps = Vs + 32 * 64;
    const int tid = threadIdx.x, ty = tid >> 3, tx = tid & 7;
    const int kvh = blockIdx.y, s0 = blockIdx.x * kPf2Q;
    const int nq = min(kPf2Q, S - s0);
    const float *kb = kc + ((size_t)b * KV + kvh) * max_len * 64;
    const float *vb = vc + ((size_t)b * KV + kvh) * max_len * 64;
    // Q tile: row r = 8 qi + hh <- q[(s0 + qi) H + 8 kvh + hh], scaled by 1 / sqrt(64) (a power of two: bit-identical to
    // scaling the scores, dual_ar.rs:258-260)
    for (int i = tid; i < 128 * 16; i += kPf2Threads) {
        const int r = i >> 4, c4 = i & 15, qi = r >> 3, hh = r & 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (qi < nq) {
            v = *reinterpret_cast<const float4 *>(q + ((size_t)(s0 + qi) * H + kvh * 8 + hh) * 64 + c4 * 4);
            v.x *= 0.125f; v.y *= 0.125f; v.z *= 0.125f; v.w *= 0.125f;
        }
        *reinterpret_cast<float4 *>(Qs + r * kPf2QS + c4 * 4) = v;
    }
    const int qi = ty >> 1;                  // the thread's query (rows 4 ty .. 4 ty + 3 = heads 4 (ty & 1) ..)
    const int len = pos0 + s0 + qi + 1;      // positions this query sees (get_mask_abs, dual_ar.rs:702-712)
    const int len_max = pos0 + s0 + nq;      // ... and the last query of the CTA
    float m[4], l[4], o[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = -INFINITY;
        l[i] = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[i][e] = 0.f;
    }
    for (int t0 = 0; t0 < len_max; t0 += 32) {
        __syncthreads();  // previous tile consumed (first pass: Q tile written)
        for (int i = tid; i < 32 * 16; i += kPf2Threads) {
            const int j = i >> 4, c4 = i & 15;
            float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
            if (t0 + j < len_max) {
                kk = *reinterpret_cast<const float4 *>(kb + (size_t)(t0 + j) * 64 + c4 * 4);
                vv = *reinterpret_cast<const float4 *>(vb + (size_t)(t0 + j) * 64 + c4 * 4);
            }
            *reinterpret_cast<float4 *>(Ks + j * kPf2KS + c4 * 4) = kk;
            *reinterpret_cast<float4 *>(Vs + j * 64 + c4 * 4) = vv;
        }
        __syncthreads();
        // (a query that sees nothing of this tile still walks it with every column masked: m stays, corr = 1, p = 0 -- the
        // shuffles and __syncwarp below need the whole warp, and a warp holds two queries)
        // ---- S = Q K^T, rows 4 ty + i, columns tx + 8 j
        float sc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sc[i][j] = 0.f;
#pragma unroll 4
        for (int d = 0; d < 64; d += 4) {
            float4 qv[4], kv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4 *>(Qs + (4 * ty + i) * kPf2QS + d);
#pragma unroll
            for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4 *>(Ks + (tx + 8 * j) * kPf2KS + d);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    sc[i][j] = fmaf(qv[i].w, kv[j].w, fmaf(qv[i].z, kv[j].z, fmaf(qv[i].y, kv[j].y, fmaf(qv[i].x, kv[j].x, sc[i][j]))));
        }
        // ---- online softmax per row (the row's 32 scores live in the 8 lanes tx = 0..7 of this ty)
        bool ok[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) ok[j] = t0 + tx + 8 * j < len;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) mx = ok[j] ? fmaxf(mx, sc[i][j]) : mx;
#pragma unroll
            for (int off = 1; off < 8; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float m_new = fmaxf(m[i], mx);  // finite: position t0 is visible to this query
            const float corr = expf(m[i] - m_new);
            float rs = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float pr = ok[j] ? expf(sc[i][j] - m_new) : 0.f;
                Ps[(4 * ty + i) * kPf2PS + tx + 8 * j] = pr;
                rs += pr;
            }
#pragma unroll
            for (int off = 1; off < 8; off <<= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
            l[i] = fmaf(l[i], corr, rs);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[i][e] *= corr;
            m[i] = m_new;
        }
        __syncwarp();  // the 8 lanes of a ty sit in one warp: P is visible to its readers
        // ---- O += P V, rows 4 ty + i, dims 4 tx .. 4 tx + 3 and 32 + 4 tx ..
#pragma unroll 2
        for (int k = 0; k < 32; k += 4) {
            float4 pv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pv[i] = *reinterpret_cast<const float4 *>(Ps + (4 * ty + i) * kPf2PS + k);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 v0 = *reinterpret_cast<const float4 *>(Vs + (k + kk) * 64 + 4 * tx);
                const float4 v1 = *reinterpret_cast<const float4 *>(Vs + (k + kk) * 64 + 32 + 4 * tx);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float pr = kk == 0 ? pv[i].x : kk == 1 ? pv[i].y : kk == 2 ? pv[i].z : pv[i].w;
                    o[i][0] = fmaf(pr, v0.x, o[i][0]); o[i][1] = fmaf(pr, v0.y, o[i][1]);
                    o[i][2] = fmaf(pr, v0.z, o[i][2]); o[i][3] = fmaf(pr, v0.w, o[i][3]);
                    o[i][4] = fmaf(pr, v1.x, o[i][4]); o[i][5] = fmaf(pr, v1.y, o[i][5]);
                    o[i][6] = fmaf(pr, v1.z, o[i][6]); o[i][7] = fmaf(pr, v1.w, o[i][7]);
                }
            }
        }
        __syncwarp();  // P of this tile consumed before the next tile's rows overwrite it
    }
    if (qi < nq) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int hh = 4 * (ty & 1) + i;
            const float inv = 1.0f / l[i];
            float *dst = y + ((size_t)(s0 + qi) * H + kvh * 8 + hh) * 64;
            *reinterpret_cast<float4 *>(dst + 4 * tx) = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
            *reinterpret_cast<float4 *>(dst + 32 + 4 * tx) = make_float4(o[i][4] * inv, o[i][5] * inv, o[i][6] * inv, o[i][7] * inv);
        }
    }
}

// ------------------------------------------------------------------ samplers
// Slow head: constrained logits (generate/utils.rs:6-33) -> token (utils.rs:36-56),
// EOS bookkeeping (single_batch.rs:153-156,199-204).  grid B, block 1024.
// logits (B, ld) holds rows [im_end | semantic_start ..) i.e. n = V' entries
// (Fish <= 1.4: the two rows [im_end, pad], sampling/mod.rs:8-26).
__global__ void __launch_bounds__(kSampleThreads) sample_slow_kernel(const float *__restrict__ logits, int ld, int n,
                                                                      const GenState *__restrict__ stp,
                                                                      uint32_t sem_start, const float *hidden,
                                                                      float *fast_x, int D) {
    const GenState st = *stp;
    if (*st.n_active == 0) return;
    const int b = blockIdx.x;
    if (!st.active[b]) return;
    extern __shared__ unsigned char smem_raw[];
    int n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);
    float *vals = reinterpret_cast<float *>(keys + 2 * max(n_pad, (int)blockDim.x));
    float *red = vals + n_pad;
    const int frame = st.frame[b];
    const float u = philox_uniform(st.sp.seed, (uint64_t)frame * (st.C + 1), (uint32_t)b);
    uint32_t tok;
    if (st.legacy_slow) {
        // legacy_softmax_sample(pad, eos): thread_rng replaced by the Philox draw
        const float eos_l = logits[(size_t)b * ld + 0], pad_l = logits[(size_t)b * ld + 1];
        const float mx = fmaxf(pad_l, eos_l);
        const float e_pad = expf(pad_l - mx), e_eos = expf(eos_l - mx);
        const float p_pad = e_pad / (e_pad + e_eos);
        tok = (st.fixed_len || u < p_pad) ? st.pad_id : st.im_end_id;
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            float v = logits[(size_t)b * ld + i];
            if (i == 0 && st.fixed_len) v = -INFINITY;
            vals[i] = v;
        }
        __syncthreads();
        const int idx = block_sample(vals, keys, red, n, n_pad, st.sp, u);
        tok = (idx == 0) ? st.im_end_id : (sem_start + (uint32_t)idx - 1);
    }
    const bool eos = (tok == st.im_end_id);
    if (threadIdx.x == 0) {
        st.cur[b * (st.C + 1)] = tok;
        st.eos[b] = eos ? 1 : 0;
        if (eos)
            for (int c = 0; c < st.C; ++c) st.cur[b * (st.C + 1) + 1 + c] = 0;
    }
    // fast stack input = pre-norm hidden (Q1, :629-634)
    for (int d = threadIdx.x; d < D; d += blockDim.x) fast_x[(size_t)b * D + d] = hidden[(size_t)b * D + d];
}

// Fast head for codebook `cb`: rep-pen (from the 2nd frame, single_batch.rs:162-168)
// + sample + fast_embeddings gather (:176-182); after the last codebook, frame
// bookkeeping (:193-204).  grid B, block 1024.
template <typename WT>
__global__ void __launch_bounds__(kSampleThreads) sample_fast_kernel(const float *__restrict__ logits, int n, int cb,
                                                                      const GenState *__restrict__ stp,
                                                                      const WT *__restrict__ fast_emb,
                                                                      float *fast_x, int D) {
    const GenState st = *stp;
    if (*st.n_active == 0) return;
    const int b = blockIdx.x;
    if (!st.active[b]) return;
    const int C = st.C;
    const bool eos = st.eos[b] != 0;
    extern __shared__ unsigned char smem_raw[];
    int n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);
    float *vals = reinterpret_cast<float *>(keys + 2 * max(n_pad, (int)blockDim.x));
    float *red = vals + n_pad;
    const int frame = st.frame[b];
    if (!eos) {
        RepPenState *rp = st.rep + (size_t)b * C + cb;
        if (frame > 0) {
            if (threadIdx.x == 0) rep_pen_update(rp, st.prev[b * (C + 1) + 1 + cb]);
            __syncthreads();
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            float v = logits[(size_t)b * n + i];
            if (frame > 0 && ((rp->seen[i >> 5] >> (i & 31)) & 1u)) v = __fdiv_rn(v, st.sp.penalty);
            vals[i] = v;
        }
        __syncthreads();
        const float u = philox_uniform(st.sp.seed, (uint64_t)frame * (C + 1) + cb + 1, (uint32_t)b);
        const int a = block_sample(vals, keys, red, n, n_pad, st.sp, u);
        if (threadIdx.x == 0) st.cur[b * (C + 1) + 1 + cb] = (uint32_t)a;
        if (cb != C - 1)
            for (int d = threadIdx.x; d < D; d += blockDim.x)
                fast_x[(size_t)b * D + d] = to_f32(fast_emb[(size_t)a * D + d]);
    }
    if (cb == C - 1) {
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t *o = st.out + ((size_t)b * st.out_cap + frame) * (C + 1);
            for (int c = 0; c <= C; ++c) {
                const uint32_t v = st.cur[b * (C + 1) + c];
                o[c] = v;
                st.prev[b * (C + 1) + c] = v;
            }
            const int nf = frame + 1;
            st.frame[b] = nf;
            // the slow step that consumed the previous frame appended one KV row
            // (input_pos bookkeeping, single_batch.rs:193-197); prefill sets pos itself
            if (frame > 0) st.pos[b] += 1;
            if (eos || nf >= st.max_frames[b]) {
                st.active[b] = 0;
                atomicSub(st.n_active, 1);
            }
        }
    }
}

// repeat_kv (candle-gqa-kernels/src/unary.cu:8-58) for callers that still want the copy.
template <typename T>
__global__ void repeat_kv_kernel(const T *__restrict__ src, T *__restrict__ dst, int n_rep, int seqlen, int hd) {
    const int s = blockIdx.x, h = blockIdx.y;
    const T *in = src + ((size_t)h * seqlen + s) * hd;
    for (int r = 0; r < n_rep; ++r) {
        T *out = dst + (((size_t)h * n_rep + r) * seqlen + s) * hd;
        for (int d = threadIdx.x; d < hd; d += blockDim.x) out[d] = in[d];
    }
}

}  // namespace fsb
