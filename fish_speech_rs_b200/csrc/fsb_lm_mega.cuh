// Persistent decode megakernel (decode_mode 2): one cooperative launch runs whole
// frames of the dual-AR loop -- slow step (24 blocks + constrained head + sampler)
// and the C fast steps (4 blocks + head + rep-pen + sampler + embedding gather) --
// with grid-wide barriers between dependent phases instead of ~410 kernel launches
// per frame.  One CTA per SM, 16 warps.
//
// Why it is shaped this way (B200):
//   * batch-1 decode is a pure weight stream (1.69 GB / frame in bf16) cut into ~270
//     dependent phases per frame; what matters is (i) keeping HBM busy across the phase
//     boundaries and (ii) a short instruction path per phase (the first version of
//     this kernel was issue-bound: ~5 M warp instructions per SM per frame).
//   * weights are static, so every warp issues the 16-byte loads of its first task of
//     the NEXT phase into registers (and L2 prefetches for its later tasks) BEFORE it
//     waits on the grid barrier; DRAM latency is paid behind the barrier.  The same
//     holds for cached K/V rows: they are staged into shared memory with cp.async
//     before the barrier that precedes the attention phase.
//   * a CTA owns a contiguous block of output rows per phase; a task is one row (or a
//     2048-element slice of a long row) handled by one warp: <= 8 16-byte loads per
//     lane, one shuffle reduction, result into shared memory; the epilogue threads
//     combine the slices in a fixed order (deterministic).
//   * RMSNorm is a prologue (every CTA renormalises the 4 KB activation itself),
//     RoPE + KV append is the epilogue of the QKV phase, residual / SwiGLU are
//     epilogues of wo / w2 / w1,w3, the attention combine is the prologue of wo.
//   * GQA: one CTA per (row, kv head, 128-position chunk) reads each cached K/V row
//     once for the 8 query heads that share it (replaces repeat_kv,
//     candle-gqa-kernels/src/unary.cu).
//
// Reference call sites replaced: dual_ar.rs:160-165,239-384,429-440,574-673;
// generate/single_batch.rs:76-214; sampling/mod.rs; sampling/rep_pen.rs.
#pragma once
#include "fsb_lm_mega_params.cuh"

namespace fsb {

template <typename WT> struct WTraits;
template <> struct WTraits<float> { static constexpr int NE = 4; };
template <> struct WTraits<__nv_bfloat16> { static constexpr int NE = 8; };

template <typename WT> __device__ __forceinline__ void unpack16(const uint4 &v, float (&w)[WTraits<WT>::NE]);
template <> __device__ __forceinline__ void unpack16<float>(const uint4 &v, float (&w)[4]) {
    w[0] = __uint_as_float(v.x); w[1] = __uint_as_float(v.y); w[2] = __uint_as_float(v.z); w[3] = __uint_as_float(v.w);
}
template <> __device__ __forceinline__ void unpack16<__nv_bfloat16>(const uint4 &v, float (&w)[8]) {
    w[0] = bf16lo(v.x); w[1] = bf16hi(v.x); w[2] = bf16lo(v.y); w[3] = bf16hi(v.y);
    w[4] = bf16lo(v.z); w[5] = bf16hi(v.z); w[6] = bf16lo(v.w); w[7] = bf16hi(v.w);
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// debug timers: SM cycle counter (cheap to read, unlike %globaltimer); reported at 1.965 GHz
__device__ __forceinline__ unsigned long long globaltimer_ns() { return (unsigned long long)clock64(); }
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Grid-wide barrier, split in two so that the next phase's weight loads are issued between
// "arrive" and "wait": monotonically increasing arrival counter (zeroed by the host before the
// launch).  Release: fence + atomic add; acquire: relaxed polling, one fence after the last poll.
__device__ __forceinline__ void grid_arrive(unsigned int *bar, unsigned int &target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    }
}
__device__ __forceinline__ void grid_wait(unsigned int *bar, unsigned int target) {
    if (threadIdx.x == 0) {
        while ((int)(ld_relaxed_u32(bar) - target) < 0) {}
        asm volatile("fence.acquire.gpu;" ::: "memory");
    }
    __syncthreads();
}

// ---------------------------------------------------------------- weight-stream plan of one phase
// task t of a CTA: (mat, local row, k-slice) -> one warp, T 16-byte loads per lane.
template <typename WT>
struct GemvPlan {
    const WT *W0, *W1;  // W1 != null: second matrix (w3), tasks of matrix 1 follow those of matrix 0
    int K;
    int T;              // units (32 lanes x 16 B) per task, <= kMegaPre
    int ksplit;         // tasks per row
    int r0, nrows;      // CTA's logical rows [r0, r0 + nrows)
    int row_a, row_b;   // weight row of logical row r: r == 0 ? row_a : row_b + r - 1
    int ntasks;
};

template <typename WT>
__device__ __forceinline__ GemvPlan<WT> make_plan(const void *W0, const void *W1, int rows_total, int K, int align,
                                                  int row_a, int row_b) {
    GemvPlan<WT> pl;
    pl.W0 = reinterpret_cast<const WT *>(W0);
    pl.W1 = reinterpret_cast<const WT *>(W1);
    pl.K = K;
    const int upr = K / (32 * WTraits<WT>::NE);
    pl.ksplit = (upr + kMegaPre - 1) / kMegaPre;
    pl.T = upr / pl.ksplit;  // host guarantees divisibility
    const unsigned groups = rows_total / align;  // blockIdx * groups < 2^32 for any vocabulary
    const unsigned g0 = (blockIdx.x * groups) / gridDim.x, g1 = ((blockIdx.x + 1) * groups) / gridDim.x;
    pl.r0 = (int)g0 * align;
    pl.nrows = (blockIdx.x + 1 == gridDim.x ? rows_total : (int)g1 * align) - pl.r0;
    pl.row_a = row_a;
    pl.row_b = row_b;
    pl.ntasks = pl.nrows * pl.ksplit * (W1 ? 2 : 1);
    return pl;
}

// pointer to lane's first 16 bytes of task t, and the task's offset into the activation row
template <typename WT>
__device__ __forceinline__ const WT *task_ptr(const GemvPlan<WT> &pl, int t, int lane, int *xoff) {
    const int per_mat = pl.nrows * pl.ksplit;
    const WT *W = pl.W0;
    if (t >= per_mat) { W = pl.W1; t -= per_mat; }
    const int rl = t / pl.ksplit, kc = t - rl * pl.ksplit;
    const int r = pl.r0 + rl;
    const int wrow = r == 0 ? pl.row_a : pl.row_b + r - 1;
    const int ko = kc * pl.T * 32 * WTraits<WT>::NE + lane * WTraits<WT>::NE;
    *xoff = ko;
    return W + (size_t)wrow * pl.K + ko;
}

template <typename WT>
__device__ __forceinline__ void task_load(const GemvPlan<WT> &pl, const WT *ptr, uint4 (&v)[kMegaPre]) {
#pragma unroll
    for (int i = 0; i < kMegaPre; ++i)
        if (i < pl.T) v[i] = ldg_stream_u4(ptr + (size_t)i * 32 * WTraits<WT>::NE);
}

// dot products of one task against NB activation rows (row stride K floats in smem)
template <typename WT, int NB>
__device__ __forceinline__ void task_dot(const GemvPlan<WT> &pl, const uint4 (&v)[kMegaPre], const float *xs, int xoff,
                                         float (&acc)[NB]) {
    constexpr int NE = WTraits<WT>::NE;
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.f;
#pragma unroll
    for (int i = 0; i < kMegaPre; ++i) {
        if (i < pl.T) {
            float w[NE];
            unpack16<WT>(v[i], w);
            const float *xp = xs + xoff + i * 32 * NE;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
#pragma unroll
                for (int j4 = 0; j4 < NE / 4; ++j4) {
                    const float4 xv = *reinterpret_cast<const float4 *>(xp + (size_t)b * pl.K + j4 * 4);
                    acc[b] = fmaf(w[j4 * 4 + 0], xv.x, acc[b]);
                    acc[b] = fmaf(w[j4 * 4 + 1], xv.y, acc[b]);
                    acc[b] = fmaf(w[j4 * 4 + 2], xv.z, acc[b]);
                    acc[b] = fmaf(w[j4 * 4 + 3], xv.w, acc[b]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- the kernel
template <typename WT, int NB>
struct Mega {
    const MegaParams &p;
    float *xs, *val, *red, *kvs;
    float *cs_s, *csf_s;  // cos|sin rows: per batch row at its slow position; per codebook index (fast stack)
    int *pos_s;           // slow positions of the current frame
    int tid, lane, warp;
    unsigned int target;
    uint4 pre[kMegaPre];
    int pre_xoff;
    float gpre[4];   // norm weights of the coming phase (k = tid + i * 512), loaded with the weights
    int n_staged;    // later tasks of this warp staged in smem by cp.async (same lane writes and reads)
    GemvPlan<WT> plan;
    // sampler state, authoritative copy in CTA 0's shared memory (global copies are write-through)
    int *s_active, *s_eos, *s_frame, *s_maxf;
    uint32_t *s_cur, *s_prev;
    RepPenState *s_rep;
    // attention item staged for the coming K_ATT phase
    int att_item, att_n;  // item id (or -1), positions staged

    __device__ Mega(const MegaParams &pp, float *smem) : p(pp) {
        xs = smem;
        val = smem + pp.xs_floats;
        red = val + pp.val_floats;
        cs_s = red + 64 * NB;
        csf_s = cs_s + NB * 64;
        pos_s = reinterpret_cast<int *>(csf_s + 8 * 64);
        s_active = pos_s + 4 * ((NB + 3) / 4);
        s_eos = s_active + NB;
        s_frame = s_eos + NB;
        s_maxf = s_frame + NB;
        s_cur = reinterpret_cast<uint32_t *>(s_maxf + NB);
        s_prev = s_cur + NB * 20;
        s_rep = reinterpret_cast<RepPenState *>(s_prev + NB * 20);
        kvs = reinterpret_cast<float *>(s_rep + NB * 8);
        n_staged = 0;
        tid = threadIdx.x;
        lane = tid & 31;
        warp = tid >> 5;
        target = 0;
        att_item = -1;
        att_n = 0;
        pre_xoff = 0;
    }


    // plan + register preload of the next phase's weights; called between barrier arrive and wait.
    // First task of the warp -> registers; later tasks -> this warp's slots of the smem staging area
    // (cp.async, lane-private: the lane that copies a 16-byte piece is the lane that consumes it);
    // what does not fit is pulled into L2.  norm_w: rms_norm weights of the phase (or null).
    __device__ __forceinline__ void prep(const void *W0, const void *W1, int rows, int K, const float *norm_w,
                                         int align = 1, int row_a = 0, int row_b = 1, bool allow_stage = true) {
        constexpr int NE = WTraits<WT>::NE;
        plan = make_plan<WT>(W0, W1, rows, K, align, row_a, row_b);
        if (warp < plan.ntasks) {
            const WT *ptr = task_ptr<WT>(plan, warp, lane, &pre_xoff);
            task_load<WT>(plan, ptr, pre);
        }
        if (norm_w) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (tid + i * kMegaThreads < p.D) gpre[i] = __ldg(norm_w + tid + i * kMegaThreads);
        }
        n_staged = 0;
        (void)allow_stage;
        // later tasks of the phase are NOT requested here: their lines were pulled into L2 one phase
        // earlier (l2_prefetch_rows), and anything outstanding towards this SM would delay the
        // activation reload that follows the barrier (the SM's return path is a FIFO)
    }

    // Pull this CTA's whole weight slice of a FUTURE phase into L2 (prefetch.global.L2 returns no data to
    // the SM).  Issued two weight phases ahead, so HBM streams in the background while the CTA computes,
    // sits in barriers or reloads activations; the loads that finally feed the FMAs are L2 hits.
    __device__ __forceinline__ void l2_prefetch_rows(const void *W0, const void *W1, int rows, int K, int align,
                                                     int row_a, int row_b) {
        const GemvPlan<WT> pl = make_plan<WT>(W0, W1, rows, K, align, row_a, row_b);
        const int lines = K * (int)sizeof(WT) / 128;
        const int nr = pl.nrows * (W1 ? 2 : 1);
        for (int fr = warp; fr < nr; fr += kMegaWarps) {
            const WT *W = fr < pl.nrows ? pl.W0 : pl.W1;
            const int r = pl.r0 + (fr < pl.nrows ? fr : fr - pl.nrows);
            const size_t wrow = (size_t)(r == 0 ? pl.row_a : pl.row_b + r - 1);
            const char *base = reinterpret_cast<const char *>(W + wrow * K);
            for (int ln = lane; ln < lines; ln += 32) prefetch_l2(base + (size_t)ln * 128);
        }
    }

    // all tasks of the CTA: val[t * NB + b] = dot(task t, activation row b); the next task's loads are
    // issued before the current one is consumed
    __device__ __forceinline__ void run_tasks() {
        uint4 nxt[kMegaPre];
        int nxoff = 0;
        for (int t = warp; t < plan.ntasks; t += kMegaWarps) {
            const bool more = t + kMegaWarps < plan.ntasks;
            if (more) {
                const WT *nptr = task_ptr<WT>(plan, t + kMegaWarps, lane, &nxoff);
                task_load<WT>(plan, nptr, nxt);
            }
            float acc[NB];
            task_dot<WT, NB>(plan, pre, xs, pre_xoff, acc);
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const float sres = warp_sum(acc[b]);
                if (lane == 0) val[t * NB + b] = sres;
            }
            if (more) {
#pragma unroll
                for (int i = 0; i < kMegaPre; ++i) pre[i] = nxt[i];
                pre_xoff = nxoff;
            }
        }
    }

    // value of (matrix m, local row rl, batch row b): slices summed in a fixed order
    __device__ __forceinline__ float row_val(int m, int rl, int b) const {
        const int t0 = (m * plan.nrows + rl) * plan.ksplit;
        float s = val[t0 * NB + b];
        for (int k = 1; k < plan.ksplit; ++k) s += val[(t0 + k) * NB + b];
        return s;
    }

    // rms_norm of NB rows staged in xs (row stride D): x / sqrt(mean(x^2) + eps) * g  (candle_nn::RmsNorm);
    // g comes from the registers `prep` filled (D <= 4 * 512), else from global memory
    __device__ __forceinline__ void norm_in_smem(int K, const float *g) {
        float ss[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            ss[b] = 0.f;
            for (int k = tid; k < K; k += kMegaThreads) {
                const float v = xs[b * K + k];
                ss[b] = fmaf(v, v, ss[b]);
            }
            ss[b] = warp_sum(ss[b]);
            if (lane == 0) red[warp * NB + b] = ss[b];
        }
        __syncthreads();
        const bool in_regs = K <= 4 * kMegaThreads;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < kMegaWarps; ++w) tot += red[w * NB + b];
            const float denom = sqrtf(tot / (float)K + p.eps);
            if (in_regs) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = tid + i * kMegaThreads;
                    if (k < K) xs[b * K + k] = __fmul_rn(__fdiv_rn(xs[b * K + k], denom), gpre[i]);
                }
            } else {
                for (int k = tid; k < K; k += kMegaThreads)
                    xs[b * K + k] = __fmul_rn(__fdiv_rn(xs[b * K + k], denom), g[k]);
            }
        }
    }

    // rows of a (B, K) global activation -> xs; K % 4 == 0
    __device__ __forceinline__ void stage_rows(const float *src, int K) {
        const int n4 = p.nb * K / 4;
        for (int i = tid; i < n4; i += kMegaThreads)
            reinterpret_cast<float4 *>(xs)[i] = __ldcg(reinterpret_cast<const float4 *>(src) + i);
        for (int i = p.nb * K + tid; i < NB * K; i += kMegaThreads) xs[i] = 0.f;
    }

    // ------------------------------------------------------------ split-KV GQA attention (slow blocks)
    // item = ((b * KV + kvh) * n_chunks_max + chunk); positions [chunk*128, min(len, chunk*128 + 128))
    __device__ __forceinline__ bool att_decode(int item, int *b, int *kvh, int *chunk, int *len) const {
        *chunk = item % p.n_chunks_max;
        *kvh = (item / p.n_chunks_max) % p.KV;
        *b = item / (p.n_chunks_max * p.KV);
        *len = pos_s[*b] + 1;
        return *chunk * kMegaChunk < *len;
    }

    // rows [j0 + from, j0 + to) of the item's K and V -> smem with cp.async (no wait)
    __device__ __forceinline__ void att_stage(const float *kcache, const float *vcache, int b, int kvh, int j0,
                                              int from, int to) {
        const float *kb = kcache + (((size_t)b * p.KV + kvh) * p.max_len + j0) * p.hd;
        const float *vb = vcache + (((size_t)b * p.KV + kvh) * p.max_len + j0) * p.hd;
        float *ks = kvs, *vs = kvs + kMegaChunk * kMegaKvStride;
        const int segs = p.hd / 4;  // 16-byte segments per row
        for (int i = tid; i < (to - from) * segs; i += kMegaThreads) {
            const int j = from + i / segs, sg = i % segs;
            cp_async16(ks + j * kMegaKvStride + sg * 4, kb + (size_t)j * p.hd + sg * 4);
            cp_async16(vs + j * kMegaKvStride + sg * 4, vb + (size_t)j * p.hd + sg * 4);
        }
    }

    // called before the barrier that precedes K_ATT: stage everything of this CTA's first item that
    // is already in the cache (all positions but the one the QKV phase is writing right now)
    __device__ __forceinline__ void att_prefetch(int layer) {
        att_item = -1;
        const int nitems = p.nb * p.KV * p.n_chunks_max;
        const size_t slow_kv = (size_t)p.max_batch * p.KV * p.max_len * p.hd;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            int b, kvh, chunk, len;
            if (!att_decode(item, &b, &kvh, &chunk, &len)) continue;
            const int j0 = chunk * kMegaChunk, j1 = min(len - 1, j0 + kMegaChunk);  // exclude position len-1
            att_item = item;
            att_n = max(j1 - j0, 0);
            if (att_n > 0) att_stage(p.kc + layer * slow_kv, p.vc + layer * slow_kv, b, kvh, j0, 0, att_n);
            break;
        }
        cp_async_commit();
    }

    __device__ __forceinline__ void phase_attn_slow(int layer) {
        const size_t slow_kv = (size_t)p.max_batch * p.KV * p.max_len * p.hd;
        const float *kcache = p.kc + layer * slow_kv, *vcache = p.vc + layer * slow_kv;
        const int n_rep = p.H / p.KV;
        const int hq = warp & 7, hf = warp >> 3;
        const int g = lane >> 2, sub = lane & 3;
        const float scale = 1.0f / sqrtf((float)p.hd);
        const int nitems = p.nb * p.KV * p.n_chunks_max;
        const float *ks = kvs, *vs = kvs + kMegaChunk * kMegaKvStride;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            int b, kvh, chunk, len;
            if (!att_decode(item, &b, &kvh, &chunk, &len)) continue;
            const int j0 = chunk * kMegaChunk, j1 = min(len, j0 + kMegaChunk);
            // rows not staged before the barrier (the new position; everything for later items)
            const int have = item == att_item ? att_n : 0;
            __syncthreads();  // previous item's smem reads are done
            if (j0 + have < j1) att_stage(kcache, vcache, b, kvh, j0, have, j1 - j0);
            cp_async_commit();
            cp_async_wait_all();
            __syncthreads();
            if (hq < n_rep) {
                const int n = j1 - j0;
                const int mid = (n + 1) / 2;
                const int a0 = hf == 0 ? 0 : mid, a1 = hf == 0 ? mid : n;
                const int h = kvh * n_rep + hq;
                float4 qv[4];
                const float *qp = p.q + (size_t)b * p.H * p.hd + (size_t)h * p.hd + sub * 4;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) qv[jj] = __ldcg(reinterpret_cast<const float4 *>(qp + jj * 16));
                float m = -INFINITY, l = 0.f;
                float o[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = 0.f;
                for (int jb = a0; jb < a1; jb += 8) {  // warp-uniform trip count (the shuffles need all lanes)
                    const int j = jb + g;
                    const bool valid = j < a1;
                    const int jc = valid ? j : a0;
                    const float *kr = ks + jc * kMegaKvStride + sub * 4, *vr = vs + jc * kMegaKvStride + sub * 4;
                    float dot = 0.f;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float4 kk = *reinterpret_cast<const float4 *>(kr + jj * 16);
                        dot = fmaf(qv[jj].x, kk.x * scale, dot);
                        dot = fmaf(qv[jj].y, kk.y * scale, dot);
                        dot = fmaf(qv[jj].z, kk.z * scale, dot);
                        dot = fmaf(qv[jj].w, kk.w * scale, dot);
                    }
                    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                    if (valid) {
                        const float m_new = fmaxf(m, dot);
                        const float corr = expf(m - m_new);
                        const float pj = expf(dot - m_new);
                        l = fmaf(l, corr, pj);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const float4 vv = *reinterpret_cast<const float4 *>(vr + jj * 16);
                            o[jj * 4 + 0] = fmaf(o[jj * 4 + 0], corr, pj * vv.x);
                            o[jj * 4 + 1] = fmaf(o[jj * 4 + 1], corr, pj * vv.y);
                            o[jj * 4 + 2] = fmaf(o[jj * 4 + 2], corr, pj * vv.z);
                            o[jj * 4 + 3] = fmaf(o[jj * 4 + 3], corr, pj * vv.w);
                        }
                        m = m_new;
                    }
                }
                // merge the 8 position groups (lanes with equal `sub`)
#pragma unroll
                for (int off = 4; off < 32; off <<= 1) {
                    const float mo = __shfl_xor_sync(0xffffffffu, m, off), lo = __shfl_xor_sync(0xffffffffu, l, off);
                    const float M = fmaxf(m, mo);
                    const float wa = (m == -INFINITY) ? 0.f : expf(m - M), wb = (mo == -INFINITY) ? 0.f : expf(mo - M);
                    l = l * wa + lo * wb;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float oo = __shfl_xor_sync(0xffffffffu, o[i], off);
                        o[i] = o[i] * wa + oo * wb;
                    }
                    m = M;
                }
                if (g == 0) {
                    float *out = p.partial + (((size_t)b * p.H + h) * (2 * p.n_chunks_max) + (chunk * 2 + hf)) * (p.hd + 4);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        *reinterpret_cast<float4 *>(out + jj * 16 + sub * 4) =
                            make_float4(o[jj * 4 + 0], o[jj * 4 + 1], o[jj * 4 + 2], o[jj * 4 + 3]);
                    if (sub == 0) { out[p.hd] = m; out[p.hd + 1] = l; }
                }
            }
        }
        att_item = -1;
    }

    // prologue of wo (slow): combine the chunk partials into xs (row stride H*hd).  Loads are issued in
    // batches of 8 slots (m, l, o of every slot in the batch are independent loads: one L2 round trip).
    __device__ __forceinline__ void combine_attn() {
        const int Hhd = p.H * p.hd;
        for (int i = tid; i < NB * Hhd; i += kMegaThreads) {
            const int b = i / Hhd, hd_i = i - b * Hhd, h = hd_i / p.hd, d = hd_i - h * p.hd;
            float v = 0.f;
            if (b < p.nb) {
                const int len = pos_s[b] + 1;
                const int ns = 2 * ((len + kMegaChunk - 1) / kMegaChunk);
                const float *pp = p.partial + ((size_t)b * p.H + h) * (2 * p.n_chunks_max) * (p.hd + 4);
                float M = -INFINITY, Lsum = 0.f, o = 0.f;
                for (int s0 = 0; s0 < ns; s0 += 8) {
                    float ms[8], ls[8], os[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (s0 + j < ns) {
                            const float *sp = pp + (s0 + j) * (p.hd + 4);
                            ms[j] = __ldcg(sp + p.hd);
                            ls[j] = __ldcg(sp + p.hd + 1);
                            os[j] = __ldcg(sp + d);
                        } else {
                            ms[j] = -INFINITY; ls[j] = 0.f; os[j] = 0.f;
                        }
                    }
                    float Mn = M;
#pragma unroll
                    for (int j = 0; j < 8; ++j) Mn = fmaxf(Mn, ms[j]);
                    const float c0 = (M == -INFINITY) ? 0.f : expf(M - Mn);
                    Lsum *= c0;
                    o *= c0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float w = (ms[j] == -INFINITY) ? 0.f : expf(ms[j] - Mn);
                        Lsum = fmaf(ls[j], w, Lsum);
                        o = fmaf(os[j], w, o);
                    }
                    M = Mn;
                }
                v = ns > 0 ? o / Lsum : 0.f;
            }
            xs[i] = v;
        }
    }

    // prologue of wo (fast): the whole attention over <= C cached positions, recomputed by every CTA.
    // q and the cached rows are staged in smem by one coalesced pass (one L2 round trip).
    __device__ __forceinline__ void fast_attn(const float *kcache, const float *vcache, int cb) {
        const int Hhd = p.H * p.hd, n_rep = p.H / p.KV;
        const float scale = 1.0f / sqrtf((float)p.hd);
        const int npos = cb + 1;
        float *qs = xs + NB * Hhd;                         // NB * Hhd, behind the output rows (max K >= 2 * Hhd)
        float *kss = kvs;                                  // NB * KV * fast_len * hd
        float *vss = kss + NB * p.KV * p.fast_len * p.hd;  // same
        for (int i = tid; i < p.nb * Hhd / 4; i += kMegaThreads)
            reinterpret_cast<float4 *>(qs)[i] = __ldcg(reinterpret_cast<const float4 *>(p.q) + i);
        const int rows = p.nb * p.KV;  // (b, kvh) pairs; each holds fast_len rows of hd floats, npos of them valid
        const int seg = p.hd / 4;
        for (int i = tid; i < rows * npos * seg; i += kMegaThreads) {
            const int r = i / (npos * seg), rem = i - r * npos * seg;
            const size_t off = (size_t)r * p.fast_len * p.hd + rem * 4;
            *reinterpret_cast<float4 *>(kss + off) = __ldcg(reinterpret_cast<const float4 *>(kcache + off));
            *reinterpret_cast<float4 *>(vss + off) = __ldcg(reinterpret_cast<const float4 *>(vcache + off));
        }
        __syncthreads();
        // one warp per (row, head): lane = (position j = lane / 4, 16 of the 64 dims); all positions at once
        const int j = lane >> 2, sub = lane & 3;
        const bool valid = j < npos;
        for (int item = warp; item < NB * p.H; item += kMegaWarps) {
            const int b = item / p.H, h = item - b * p.H;
            if (b >= p.nb) {
                xs[b * Hhd + h * p.hd + lane] = 0.f;
                xs[b * Hhd + h * p.hd + lane + 32] = 0.f;
                continue;
            }
            const int kvh = h / n_rep;
            const float *qp = qs + b * Hhd + h * p.hd + sub * 4;
            const float *kr = kss + (((size_t)b * p.KV + kvh) * p.fast_len + (valid ? j : 0)) * p.hd + sub * 4;
            const float *vr = vss + (((size_t)b * p.KV + kvh) * p.fast_len + (valid ? j : 0)) * p.hd + sub * 4;
            float dot = 0.f;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const float4 qv = *reinterpret_cast<const float4 *>(qp + jj * 16);
                const float4 kk = *reinterpret_cast<const float4 *>(kr + jj * 16);
                dot = fmaf(qv.x, kk.x * scale, dot);
                dot = fmaf(qv.y, kk.y * scale, dot);
                dot = fmaf(qv.z, kk.z * scale, dot);
                dot = fmaf(qv.w, kk.w * scale, dot);
            }
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            const float sc = valid ? dot : -INFINITY;
            float m = sc;
#pragma unroll
            for (int off = 4; off < 32; off <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
            const float pj = valid ? expf(sc - m) : 0.f;
            float l = pj;
#pragma unroll
            for (int off = 4; off < 32; off <<= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
            float o[16];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const float4 vv = *reinterpret_cast<const float4 *>(vr + jj * 16);
                o[jj * 4 + 0] = pj * vv.x; o[jj * 4 + 1] = pj * vv.y; o[jj * 4 + 2] = pj * vv.z; o[jj * 4 + 3] = pj * vv.w;
            }
#pragma unroll
            for (int off = 4; off < 32; off <<= 1)
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] += __shfl_xor_sync(0xffffffffu, o[i], off);
            if (j == 0) {
                float *dst = xs + b * Hhd + h * p.hd + sub * 4;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    *reinterpret_cast<float4 *>(dst + jj * 16) =
                        make_float4(o[jj * 4 + 0] / l, o[jj * 4 + 1] / l, o[jj * 4 + 2] / l, o[jj * 4 + 3] / l);
            }
        }
    }

    // ------------------------------------------------------------ the weight-streaming phases (single code path)
    enum { K_QKV = 0, K_ATT = 1, K_WO = 2, K_W13 = 3, K_W2 = 4, K_HEAD = 5, K_SAMPLE = 6, K_END = 7 };
    struct Step { int frame, pass, l, kind; };  // pass 0 = slow stack, pass c+1 = fast step of codebook c

    __device__ __forceinline__ const MegaLayer &layer_of(const Step &s) const { return (s.pass == 0 ? p.slow : p.fast)[s.l]; }

    __device__ __forceinline__ Step advance(const Step &s) const {
        Step n = s;
        const bool slow = s.pass == 0;
        switch (s.kind) {
            case K_QKV: n.kind = slow ? K_ATT : K_WO; break;
            case K_ATT: n.kind = K_WO; break;
            case K_WO: n.kind = K_W13; break;
            case K_W13: n.kind = K_W2; break;
            case K_W2:
                if (s.l + 1 < (slow ? p.NL : p.NFL)) { n.l = s.l + 1; n.kind = K_QKV; }
                else n.kind = K_HEAD;
                break;
            case K_HEAD: n.kind = K_SAMPLE; break;
            default:  // K_SAMPLE
                n.l = 0;
                if (s.pass < p.C) { n.pass = s.pass + 1; n.kind = p.NFL > 0 ? K_QKV : K_HEAD; }
                else {
                    n.pass = 0;
                    n.frame = s.frame + 1;
                    n.kind = n.frame < p.nframes ? (p.NL > 0 ? K_QKV : K_HEAD) : K_END;
                }
        }
        return n;
    }

    __device__ __forceinline__ void prep_step(const Step &s) {
        const bool slow = s.pass == 0;
        switch (s.kind) {
            case K_QKV: prep(layer_of(s).wqkv, nullptr, p.QKV, p.D, layer_of(s).attn_norm, 2); break;
            case K_WO:  // the fast wo prologue stages K/V in the same smem area: no weight staging there
                prep(layer_of(s).wo, nullptr, p.D, p.H * p.hd, nullptr, 1, 0, 1, slow);
                break;
            case K_W13: prep(layer_of(s).w1, layer_of(s).w3, p.I, p.D, layer_of(s).ffn_norm); break;
            case K_W2: prep(layer_of(s).w2, nullptr, p.D, p.I, nullptr); break;
            case K_HEAD:
                if (slow) prep(p.out_w, nullptr, p.n_slow_logits, p.D, p.norm, 1, p.slow_row0, p.slow_rest_base);
                else prep(p.fast_out, nullptr, p.CS, p.D, p.fast_norm);
                break;
            default: break;
        }
    }

    __device__ __forceinline__ void l2_prefetch_step(const Step &s) {
        const bool slow = s.pass == 0;
        switch (s.kind) {
            case K_QKV: l2_prefetch_rows(layer_of(s).wqkv, nullptr, p.QKV, p.D, 2, 0, 1); break;
            case K_WO: l2_prefetch_rows(layer_of(s).wo, nullptr, p.D, p.H * p.hd, 1, 0, 1); break;
            case K_W13: l2_prefetch_rows(layer_of(s).w1, layer_of(s).w3, p.I, p.D, 1, 0, 1); break;
            case K_W2: l2_prefetch_rows(layer_of(s).w2, nullptr, p.D, p.I, 1, 0, 1); break;
            case K_HEAD:
                if (slow) l2_prefetch_rows(p.out_w, nullptr, p.n_slow_logits, p.D, 1, p.slow_row0, p.slow_rest_base);
                else l2_prefetch_rows(p.fast_out, nullptr, p.CS, p.D, 1, 0, 1);
                break;
            default: break;
        }
    }
    // the next step (after s) that streams weights, or K_END
    __device__ __forceinline__ Step next_weight_step(Step s) const {
        do { s = advance(s); } while (s.kind == K_ATT || s.kind == K_SAMPLE);
        return s;
    }

    __device__ __forceinline__ void gemv_phase(const Step &s) {
        const bool slow = s.pass == 0;
        const int cb = s.pass - 1;
        const int kind = s.kind;
        float *xg = slow ? p.x : p.fx;
        const size_t kv_stride = (size_t)p.max_batch * p.KV * (slow ? p.max_len : p.fast_len) * p.hd;
        float *kcl = (slow ? p.kc : p.fkc) + s.l * kv_stride, *vcl = (slow ? p.vc : p.fvc) + s.l * kv_stride;
        const int cache_len = slow ? p.max_len : p.fast_len;
        // ---- prologue: stage the activation rows in smem
        const float *norm_w = nullptr;
        if (kind == K_QKV) {
            const int D = p.D;
            if (s.l > 0) {
                stage_rows(xg, D);
            } else if (slow) {
                // once per frame: this frame's slow positions and their RoPE rows -> smem
                for (int i = tid; i < NB * p.hd; i += kMegaThreads) {
                    const int b = i / p.hd, d = i - b * p.hd, half = p.hd / 2;
                    // a finished row of a ragged batch (pos == its budget, possibly == max_len) is parked at
                    // pos -1: no attention item, no K/V append, RoPE row 0
                    const bool live = b < p.nb && __ldcg(p.st.active + b) != 0;
                    const int pos = live ? __ldcg(p.st.pos + b) : 0;
                    cs_s[i] = d < half ? p.cosT[(size_t)pos * half + d] : p.sinT[(size_t)pos * half + d - half];
                    if (d == 0) pos_s[b] = live ? pos : -1;
                }
                // DualARTransformer::embed, dual_ar.rs:532-567, on the previous frame's codes
                const WT *emb = reinterpret_cast<const WT *>(p.emb), *cbe = reinterpret_cast<const WT *>(p.cb_emb);
                for (int i = tid; i < NB * D; i += kMegaThreads) {
                    const int b = i / D, d = i - b * D;
                    float acc = 0.f;
                    if (b < p.nb) {
                        const uint32_t *t = p.st.prev + b * (p.C + 1);
                        const uint32_t tok0 = __ldcg(t);
                        const bool m = p.has_end ? (tok0 <= p.sem_end && tok0 >= p.sem_start) : (tok0 == p.sem_start);
                        const float mf = m ? 1.f : 0.f;
                        acc = to_f32(emb[(size_t)tok0 * D + d]);
                        for (int c = 0; c < p.C; ++c) {
                            const uint32_t code = __ldcg(t + 1 + c);
                            acc = __fadd_rn(acc, __fmul_rn(to_f32(cbe[((size_t)c * p.CS + code) * D + d]), mf));
                        }
                        if (blockIdx.x == 0) xg[i] = acc;
                    }
                    xs[i] = acc;
                }
            } else {
                // fast stack input: pre-norm slow hidden (Q1) for codebook 0, else fast_embeddings[previous code]
                const WT *fe = reinterpret_cast<const WT *>(p.fast_emb);
                for (int i = tid; i < NB * D; i += kMegaThreads) {
                    const int b = i / D, d = i - b * D;
                    float v = 0.f;
                    if (b < p.nb) {
                        if (cb == 0) v = __ldcg(p.x + i);
                        else v = to_f32(fe[(size_t)__ldcg(p.st.cur + b * (p.C + 1) + cb) * D + d]);
                        if (blockIdx.x == 0) xg[i] = v;
                    }
                    xs[i] = v;
                }
            }
            norm_w = layer_of(s).attn_norm;
        } else if (kind == K_WO) {
            if (slow) combine_attn();
            else fast_attn(kcl, vcl, cb);
        } else if (kind == K_W13) {
            stage_rows(xg, p.D);
            norm_w = layer_of(s).ffn_norm;
        } else if (kind == K_W2) {
            stage_rows(p.h, p.I);
        } else {  // K_HEAD
            stage_rows(xg, p.D);
            norm_w = slow ? p.norm : p.fast_norm;
        }
        const bool sub_timed = p.dbg != nullptr && tid == 0 && blockIdx.x == 0;
        unsigned long long ta = 0, tb = 0, tc = 0, td = 0;
        if (sub_timed) ta = globaltimer_ns();
        __syncthreads();
        if (sub_timed) tb = globaltimer_ns();
        if (norm_w) {
            norm_in_smem(p.D, norm_w);
            __syncthreads();
        }
        if (sub_timed) tc = globaltimer_ns();
        run_tasks();
        __syncthreads();
        if (sub_timed) {
            td = globaltimer_ns();
            p.dbg[64 + kind * 4 + 0] += tb - ta;
            p.dbg[64 + kind * 4 + 1] += tc - tb;
            p.dbg[64 + kind * 4 + 2] += td - tc;
            p.dbg[64 + kind * 4 + 3] += ta;  // absolute start, combined with the phase start below
        }
        // ---- epilogue
        if (kind == K_QKV) {
            // pairs of rows -> rope_i (dual_ar.rs:246-247) -> q buffer / K cache; V rows -> V cache (Tensor::cat, :316-324)
            const int half = p.hd / 2, Hhd = p.H * p.hd, KVhd = p.KV * p.hd;
            for (int i = tid; i < (plan.nrows / 2) * NB; i += kMegaThreads) {
                const int pr = i / NB, b = i - pr * NB;
                if (b >= p.nb) continue;
                const int r = plan.r0 + 2 * pr;
                const float v0 = row_val(0, 2 * pr, b), v1 = row_val(0, 2 * pr + 1, b);
                const int pos = slow ? pos_s[b] : cb;
                if (pos < 0) continue;  // finished row
                if (r < Hhd + KVhd) {
                    const int pi = (r % p.hd) / 2;
                    const float *cs = slow ? cs_s + b * p.hd : csf_s + cb * p.hd;
                    const float c = cs[pi], sn = cs[half + pi];
                    const float o0 = __fsub_rn(__fmul_rn(v0, c), __fmul_rn(v1, sn));
                    const float o1 = __fadd_rn(__fmul_rn(v0, sn), __fmul_rn(v1, c));
                    if (r < Hhd) {
                        p.q[(size_t)b * Hhd + r] = o0;
                        p.q[(size_t)b * Hhd + r + 1] = o1;
                    } else {
                        const int rk = r - Hhd, kvh = rk / p.hd, d = rk % p.hd;
                        float *dst = kcl + (((size_t)b * p.KV + kvh) * cache_len + pos) * p.hd + d;
                        dst[0] = o0;
                        dst[1] = o1;
                    }
                } else {
                    const int rv = r - Hhd - KVhd, kvh = rv / p.hd, d = rv % p.hd;
                    float *dst = vcl + (((size_t)b * p.KV + kvh) * cache_len + pos) * p.hd + d;
                    dst[0] = v0;
                    dst[1] = v1;
                }
            }
        } else {
            for (int i = tid; i < plan.nrows * NB; i += kMegaThreads) {
                const int rl = i / NB, b = i - rl * NB;
                if (b >= p.nb) continue;
                const float s1 = row_val(0, rl, b);
                const int r = plan.r0 + rl;
                if (kind == K_W13) {
                    // silu(w1 x) * (w3 x), dual_ar.rs:160-165
                    p.h[(size_t)b * p.I + r] = __fmul_rn(silu_f(s1), row_val(1, rl, b));
                } else if (kind == K_HEAD) {
                    p.logits[(size_t)b * p.ldl + r] = s1;
                } else {  // K_WO / K_W2: residual add, dual_ar.rs:436-440
                    float *dst = xg + (size_t)b * p.D + r;
                    *dst = __fadd_rn(__ldcg(dst), s1);
                }
            }
        }
    }

    // ------------------------------------------------------------ samplers (CTA 0 only; scratch aliased on xs/val)
    // The per-row generator state lives in CTA 0's shared memory for the whole launch; every update is
    // also written through to the global copy (other CTAs read cur / prev / pos / n_active from there,
    // the host reads frame / pos / out after the launch).
    __device__ __forceinline__ void load_sampler_state() {
        const GenState &st = p.st;
        const int C1 = st.C + 1;
        for (int i = tid; i < p.nb; i += kMegaThreads) {
            s_active[i] = st.active[i];
            s_eos[i] = st.eos[i];
            s_frame[i] = st.frame[i];
            s_maxf[i] = st.max_frames[i];
        }
        for (int i = tid; i < p.nb * C1; i += kMegaThreads) {
            const int b = i / C1, c = i - b * C1;
            s_cur[b * 20 + c] = st.cur[i];
            s_prev[b * 20 + c] = st.prev[i];
        }
        const int words = (int)(sizeof(RepPenState) / 4);
        for (int i = tid; i < p.nb * st.C * words; i += kMegaThreads)
            reinterpret_cast<uint32_t *>(s_rep)[i] = reinterpret_cast<const uint32_t *>(st.rep)[i];
        __syncthreads();
    }

    __device__ __forceinline__ void sample_slow() {
        const GenState &st = p.st;
        const int n = p.n_slow_logits;
        int n_pad = 1;
        while (n_pad < n) n_pad <<= 1;
        unsigned long long *keys = reinterpret_cast<unsigned long long *>(xs);
        float *vals = reinterpret_cast<float *>(keys + 2 * max(n_pad, kMegaThreads));
        float *sred = vals + n_pad;
        for (int b = 0; b < p.nb; ++b) {
            if (!s_active[b]) continue;
            const int frame = s_frame[b];
            const float u = philox_uniform(st.sp.seed, (uint64_t)frame * (st.C + 1), (uint32_t)(p.row0 + b));
            uint32_t tok;
            if (st.legacy_slow) {
                const float eos_l = __ldcg(p.logits + (size_t)b * p.ldl), pad_l = __ldcg(p.logits + (size_t)b * p.ldl + 1);
                const float mx = fmaxf(pad_l, eos_l);
                const float e_pad = expf(pad_l - mx), e_eos = expf(eos_l - mx);
                tok = (st.fixed_len || u < e_pad / (e_pad + e_eos)) ? st.pad_id : st.im_end_id;
            } else {
                for (int i = tid; i < n; i += kMegaThreads) {
                    float v = __ldcg(p.logits + (size_t)b * p.ldl + i);
                    if (i == 0 && st.fixed_len) v = -INFINITY;
                    vals[i] = v;
                }
                __syncthreads();
                const int idx = block_sample(vals, keys, sred, n, n_pad, st.sp, u);
                tok = (idx == 0) ? st.im_end_id : (p.sem_start + (uint32_t)idx - 1);
            }
            if (tid == 0) {
                const bool eos = tok == st.im_end_id;
                s_cur[b * 20] = tok;
                st.cur[b * (st.C + 1)] = tok;
                s_eos[b] = eos ? 1 : 0;
                st.eos[b] = eos ? 1 : 0;
                if (eos)
                    for (int c = 0; c < st.C; ++c) {
                        s_cur[b * 20 + 1 + c] = 0;
                        st.cur[b * (st.C + 1) + 1 + c] = 0;
                    }
            }
            __syncthreads();
        }
    }

    __device__ __forceinline__ void sample_fast(int cb) {
        const GenState &st = p.st;
        const int n = p.CS, C = st.C;
        int n_pad = 1;
        while (n_pad < n) n_pad <<= 1;
        unsigned long long *keys = reinterpret_cast<unsigned long long *>(xs);
        float *vals = reinterpret_cast<float *>(keys + 2 * max(n_pad, kMegaThreads));
        float *sred = vals + n_pad;
        for (int b = 0; b < p.nb; ++b) {
            if (!s_active[b]) continue;
            const bool eos = s_eos[b] != 0;
            const int frame = s_frame[b];
            if (!eos) {
                const bool tm = p.dbg != nullptr && tid == 0;
                const long long c0 = tm ? clock64() : 0;
                RepPenState *rp = s_rep + (size_t)b * C + cb;
                if (frame > 0) {
                    if (tid == 0) rep_pen_update(rp, s_prev[b * 20 + 1 + cb]);
                    __syncthreads();
                }
                const long long c1 = tm ? clock64() : 0;
                for (int i = tid; i < n; i += kMegaThreads) {
                    float v = __ldcg(p.logits + (size_t)b * p.ldl + i);
                    if (frame > 0 && ((rp->seen[i >> 5] >> (i & 31)) & 1u)) v = __fdiv_rn(v, st.sp.penalty);
                    vals[i] = v;
                }
                __syncthreads();
                const long long c2 = tm ? clock64() : 0;
                const float u = philox_uniform(st.sp.seed, (uint64_t)frame * (C + 1) + cb + 1, (uint32_t)(p.row0 + b));
                const int a = block_sample(vals, keys, sred, n, n_pad, st.sp, u);
                if (tm) {
                    const long long c3 = clock64();
                    p.dbg[100] += c1 - c0; p.dbg[101] += c2 - c1; p.dbg[102] += c3 - c2; p.dbg[103] += 1;
                }
                if (tid == 0) {
                    s_cur[b * 20 + 1 + cb] = (uint32_t)a;
                    st.cur[b * (C + 1) + 1 + cb] = (uint32_t)a;
                }
            }
            if (cb == C - 1) {
                __syncthreads();
                // frame bookkeeping (single_batch.rs:193-204), write-through
                if (tid <= C) {
                    const uint32_t v = s_cur[b * 20 + tid];
                    st.out[((size_t)b * st.out_cap + frame) * (C + 1) + tid] = v;
                    s_prev[b * 20 + tid] = v;
                    st.prev[b * (C + 1) + tid] = v;
                }
                if (tid == 0) {
                    const int nf = frame + 1;
                    s_frame[b] = nf;
                    st.frame[b] = nf;
                    if (frame > 0) {
                        pos_s[b] += 1;  // CTA 0's cached copy is reloaded at the next frame start anyway
                        st.pos[b] = pos_s[b];
                    }
                    if (eos || nf >= s_maxf[b]) {
                        s_active[b] = 0;
                        st.active[b] = 0;
                        atomicSub(st.n_active, 1);
                    }
                }
                // the rep-pen windows only matter to a later launch on the same rows: write them back
                {
                    const int words = (int)(sizeof(RepPenState) / 4);
                    const uint32_t *src = reinterpret_cast<const uint32_t *>(s_rep + (size_t)b * C);
                    uint32_t *dst = reinterpret_cast<uint32_t *>(st.rep + (size_t)b * C);
                    for (int i = tid; i < C * words; i += kMegaThreads) dst[i] = src[i];
                }
            }
            __syncthreads();
        }
    }

    // ------------------------------------------------------------ frame loop
    __device__ __forceinline__ void run() {
        Step cur;
        cur.frame = 0; cur.pass = 0; cur.l = 0;
        cur.kind = (p.first_is_tail || p.NL == 0) ? K_HEAD : K_QKV;
        if (p.nframes <= 0 || __ldcg(p.st.n_active) == 0) return;
        for (int i = tid; i < p.C * p.hd; i += kMegaThreads) {
            const int row = i / p.hd, d = i - row * p.hd, half = p.hd / 2;
            csf_s[i] = d < half ? p.cosT[(size_t)row * half + d] : p.sinT[(size_t)row * half + d - half];
        }
        if (blockIdx.x == 0 && tid == 0 && p.dbg) g_sample_dbg = p.dbg + 104;
        if (blockIdx.x == 0) {
            for (int i = tid; i < p.nb; i += kMegaThreads) pos_s[i] = p.st.pos[i];
            load_sampler_state();
        }
        __syncthreads();
        l2_prefetch_step(cur);
        l2_prefetch_step(next_weight_step(cur));
        prep_step(cur);
        const bool timed = p.dbg != nullptr && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
        unsigned long long *dbg = p.dbg + (blockIdx.x == 0 ? 0 : 32);
        while (cur.kind != K_END) {
            unsigned long long t0 = 0, t1 = 0, t2 = 0;
            if (timed) {
                t0 = globaltimer_ns();
                if (blockIdx.x == 0) p.dbg[64 + cur.kind * 4 + 3] -= t0;  // += (prologue end - phase start)
            }
            if (cur.kind == K_ATT) phase_attn_slow(cur.l);
            else if (cur.kind == K_SAMPLE) {
                if (blockIdx.x == 0) {
                    if (cur.pass == 0) sample_slow();
                    else sample_fast(cur.pass - 1);
                }
            } else gemv_phase(cur);
            Step nxt = advance(cur);
            if (timed) t1 = globaltimer_ns();
            grid_arrive(p.bar, target);
            // between arrive and wait: the weights of the next streaming phase start moving (a HEAD is
            // followed by a SAMPLE that only CTA 0 runs: prefetch the phase after it instead); cached
            // K/V rows of the next attention phase likewise
            if (nxt.kind == K_SAMPLE) prep_step(advance(nxt));
            else if (cur.kind != K_SAMPLE) prep_step(nxt);
            // HBM -> L2 for the weight phase after the one just prepped (once per weight phase)
            if (cur.kind != K_ATT && cur.kind != K_SAMPLE) l2_prefetch_step(next_weight_step(next_weight_step(cur)));
            if (nxt.kind == K_ATT) att_prefetch(nxt.l);
            if (timed) t2 = globaltimer_ns();
            grid_wait(p.bar, target);
            if (timed) {
                const unsigned long long t3 = globaltimer_ns();
                dbg[cur.kind * 4 + 0] += t1 - t0;
                dbg[cur.kind * 4 + 1] += t2 - t1;
                dbg[cur.kind * 4 + 2] += t3 - t2;
                dbg[cur.kind * 4 + 3] += 1;
            }
            if (nxt.frame != cur.frame && nxt.kind != K_END && __ldcg(p.st.n_active) == 0) break;
            cur = nxt;
        }
    }
};

template <typename WT, int NB>
__global__ void __launch_bounds__(kMegaThreads, 1) mega_decode_kernel(const __grid_constant__ MegaParams p) {
    extern __shared__ __align__(16) float mega_smem[];
    Mega<WT, NB> m(p, mega_smem);
    m.run();
}

template <typename WT>
static cudaError_t mega_launch_impl(int NB, const MegaParams &mp, int grid, size_t smem, cudaStream_t st) {
    const void *kern = nullptr;
    switch (NB) {
        case 1: kern = (const void *)mega_decode_kernel<WT, 1>; break;
        case 2: kern = (const void *)mega_decode_kernel<WT, 2>; break;
        case 4: kern = (const void *)mega_decode_kernel<WT, 4>; break;
        case 8: kern = (const void *)mega_decode_kernel<WT, 8>; break;
        default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    void *args[] = {(void *)&mp};
    return cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kMegaThreads), args, smem, st);
}

}  // namespace fsb
