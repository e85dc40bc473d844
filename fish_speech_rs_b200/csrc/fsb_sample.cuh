// Device-side sampler: constrained slow head / fast codebook heads.
// Replaces the host round trip of candle_transformers LogitsProcessor +
// SingleBatchedRepPenProcessor (single_batch.rs:126-169, sampling/mod.rs:51-132,
// rep_pen.rs:37-65).  Semantics are documented in oracle/sampling.py, which is
// the checker for this file.
#pragma once
#include "fsb_common.cuh"

namespace fsb {

// Philox4x32-10, identical to oracle/rng.py.
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t draw, uint32_t row) {
    uint32_t c0 = (uint32_t)draw, c1 = (uint32_t)(draw >> 32), c2 = row, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return (float)(c0 >> 8) * (1.0f / 16777216.0f);
}

// Repetition-penalty window of one (row, codebook): rep_pen.rs:4-72 bug-for-bug
// (a token is un-penalised as soon as any occurrence of it leaves the window).
struct RepPenState {
    uint32_t seen[32];   // 1024-bit membership of `tokens_seen`
    uint16_t ring[16];   // context deque, oldest at `head`
    uint32_t len;
    uint32_t head;
};
constexpr int kRepPenWindow = 16;  // single_batch.rs:51

__device__ __forceinline__ void rep_pen_update(RepPenState *st, uint32_t tok) {
    st->seen[tok >> 5] |= 1u << (tok & 31);
    if (st->len < kRepPenWindow) {
        st->ring[(st->head + st->len) % kRepPenWindow] = (uint16_t)tok;
        st->len++;
    } else {
        uint32_t dropped = st->ring[st->head];
        st->ring[st->head] = (uint16_t)tok;
        st->head = (st->head + 1) % kRepPenWindow;
        // pop_back after push_front: the just-pushed token can be the one un-penalised
        st->seen[dropped >> 5] &= ~(1u << (dropped & 31));
    }
}

struct SampleParams {
    float inv_temp;     // f32(1 / temp)
    float top_p;
    uint32_t top_k;
    int greedy;         // temp <= 1e-7
    float penalty;      // f32 repetition penalty
    uint64_t seed;
};

constexpr int kSampleThreads = 1024;
constexpr int kSampleMaxN = 4096;

// Block-wide sampler (any block size that is a multiple of 32, n_pad <= 8 * blockDim.x).
// `vals` (smem, n floats) holds the adjusted logits; returns the chosen index to every thread.
// `keys`: smem scratch of 2 * max(n_pad, blockDim.x) u64.
// top-k -> top-p -> multinomial as in sampling/mod.rs:51-75,113-132.  The candidates are ordered by a
// bitonic sort of (~prob bits, index) keys held in REGISTERS: exchange distances >= blockDim.x stay
// inside a thread, distances < 32 are warp shuffles, only the distances in between go through shared
// memory (double buffered, one __syncthreads each).  The running sums over the sorted candidates are a
// 32-lane segmented scan (each lane sums a contiguous segment sequentially, lane totals are combined
// by a warp scan): same distribution as the reference's sequential f32 chain, last-bit differences
// in the CDF only.
constexpr int kSampleMaxE = 8;

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
    unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    lo = __shfl_xor_sync(0xffffffffu, lo, m);
    hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return ((unsigned long long)hi << 32) | lo;
}

__device__ unsigned long long *g_sample_dbg = nullptr;  // optional cycle counters (debug builds of the megakernel)

template <int E>
__device__ __forceinline__ int block_sample_t(float *vals, unsigned long long *keys, float *red, int n, int n_pad,
                                              const SampleParams &sp, float u) {
    const bool tm_ = g_sample_dbg != nullptr && threadIdx.x == 0;
    const long long q0 = tm_ ? clock64() : 0;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int n_eff = E * nthreads;  // elements per thread E: index e * nthreads + tid
    float v[E];
    // ---- max (and first argmax) ----
    float best = -INFINITY;
    int best_i = 0x7fffffff;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * nthreads + tid;
        v[e] = (i < n) ? vals[i] : -INFINITY;
        if (v[e] > best) { best = v[e]; best_i = i; }  // ascending i: the first maximum wins
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    int *red_i = reinterpret_cast<int *>(red + 32);
    if (lane == 0) { red[warp] = best; red_i[warp] = best_i; }
    __syncthreads();
    {
        float b2 = lane < nwarps ? red[lane] : -INFINITY;
        int i2 = lane < nwarps ? red_i[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, b2, o);
            int oi = __shfl_xor_sync(0xffffffffu, i2, o);
            if (ov > b2 || (ov == b2 && oi < i2)) { b2 = ov; i2 = oi; }
        }
        best = b2;
        best_i = i2;
    }
    if (sp.greedy) {
        __syncthreads();
        return best_i;
    }
    // ---- softmax(logits * inv_temp) ----
    const float mx = __fmul_rn(best, sp.inv_temp);
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        v[e] = (e * nthreads + tid < n) ? expf(__fsub_rn(__fmul_rn(v[e], sp.inv_temp), mx)) : 0.f;
        s += v[e];
    }
    s = warp_sum(s);
    __syncthreads();  // red[] reads above are done
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float denom = lane < nwarps ? red[lane] : 0.f;
    denom = warp_sum(denom);
    // keys: (~prob bits) << 32 | index; ascending sort == prob desc, index asc
    unsigned long long key[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * nthreads + tid;
        key[e] = ~0ull;
        if (i < n) {
            const float p = v[e] / denom;
            key[e] = ((unsigned long long)(~__float_as_uint(p)) << 32) | (unsigned)i;
        }
    }
    const long long q1 = tm_ ? clock64() : 0;
    // ---- bitonic sort over n_eff keys ----
    int buf = 0;
    for (int k = 2; k <= n_eff; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= nthreads) {
                const int je = j / nthreads;  // partner element inside the same thread
#pragma unroll
                for (int jj = 1; jj < E; jj <<= 1) {  // compile-time register indices
                    if (je == jj) {
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            if ((e & jj) == 0) {
                                const int i = e * nthreads + tid;
                                const bool up = (i & k) == 0;
                                const unsigned long long a = key[e], b = key[e | jj];
                                if ((a > b) == up) { key[e] = b; key[e | jj] = a; }
                            }
                        }
                    }
                }
            } else if (j >= 32) {
                unsigned long long *kb = keys + (size_t)buf * n_eff;
#pragma unroll
                for (int e = 0; e < E; ++e) kb[e * nthreads + tid] = key[e];
                __syncthreads();
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = e * nthreads + tid;
                    const unsigned long long o = kb[i ^ j];
                    const bool up = (i & k) == 0, lower = (i & j) == 0;
                    const bool take_min = up == lower;
                    key[e] = take_min ? (o < key[e] ? o : key[e]) : (o > key[e] ? o : key[e]);
                }
                buf ^= 1;
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = e * nthreads + tid;
                    const unsigned long long o = shfl_xor_u64(key[e], j);
                    const bool up = (i & k) == 0, lower = (i & j) == 0;
                    const bool take_min = up == lower;
                    key[e] = take_min ? (o < key[e] ? o : key[e]) : (o > key[e] ? o : key[e]);
                }
            }
        }
    }
    // sorted position of key[e] is e * nthreads + tid; the scan below only walks the first <= n entries
    {
        unsigned long long *kb = keys + (size_t)buf * n_eff;
#pragma unroll
        for (int e = 0; e < E; ++e) kb[e * nthreads + tid] = key[e];
        __syncthreads();
        keys = kb;
    }
    const long long q2 = tm_ ? clock64() : 0;
    // ---- top-k -> top-p -> multinomial on warp 0 ----
    if (warp == 0) {
        auto wgt = [&](int i) { return __uint_as_float(~(uint32_t)(keys[i] >> 32)); };
        const int k = (sp.top_k >= (uint32_t)n) ? n : (int)sp.top_k;
        const int seg = (k + 31) / 32;
        const int i0 = min(k, lane * seg), i1 = min(k, i0 + seg);
        float local = 0.f;
        for (int i = i0; i < i1; ++i) local += wgt(i);
        float incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 0.f;
        const float sum_p = __shfl_sync(0xffffffffu, incl, 31);
        int kept = k;
        float total = sum_p;
        const bool do_topp = !(sp.top_p <= 0.f || sp.top_p >= sum_p) || (sp.top_k >= (uint32_t)n);
        if (do_topp) {
            // an entry survives iff the running sum BEFORE it is still below top_p (mod.rs:119-129)
            int kl = 0;
            float tl = 0.f, c = excl;
            for (int i = i0; i < i1; ++i) {
                const bool keep = c < sp.top_p;
                c += wgt(i);
                if (keep) { kl = i + 1; tl = c; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int ok = __shfl_xor_sync(0xffffffffu, kl, o);
                const float ot = __shfl_xor_sync(0xffffffffu, tl, o);
                if (ok > kl) { kl = ok; tl = ot; }
            }
            kept = kl;
            total = tl;
        }
        // WeightedIndex (rand 0.8.5): first i with cum[i] > u * total
        const float chosen = u * total;
        int cand = 0x7fffffff;
        {
            float c = excl;
            for (int i = i0; i < min(i1, kept); ++i) {
                const float w = wgt(i);
                c += w;
                if (c > chosen && w > 0.f) { cand = i; break; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        if (lane == 0) {
            int pick = cand;
            if (pick == 0x7fffffff) {  // nothing crossed `chosen`: last kept entry with non-zero weight
                pick = max(kept - 1, 0);
                while (pick > 0 && wgt(pick) <= 0.f) --pick;
            }
            red_i[0] = (int)(keys[pick] & 0xffffffffu);
        }
    }
    __syncthreads();
    int r = red_i[0];
    __syncthreads();
    if (tm_) {
        const long long q3 = clock64();
        g_sample_dbg[0] += q1 - q0; g_sample_dbg[1] += q2 - q1; g_sample_dbg[2] += q3 - q2; g_sample_dbg[3] += 1;
    }
    return r;
}

// elements per thread is a compile-time constant of the sort network (a generic loop over "up to 8"
// slots made the sampler issue-bound: 17 us of a 21 us draw)
__device__ inline int block_sample(float *vals, unsigned long long *keys, float *red, int n, int n_pad,
                                   const SampleParams &sp, float u) {
    const int E = max(n_pad, (int)blockDim.x) / (int)blockDim.x;
    switch (E) {
        case 1: return block_sample_t<1>(vals, keys, red, n, n_pad, sp, u);
        case 2: return block_sample_t<2>(vals, keys, red, n, n_pad, sp, u);
        case 4: return block_sample_t<4>(vals, keys, red, n, n_pad, sp, u);
        default: return block_sample_t<8>(vals, keys, red, n, n_pad, sp, u);
    }
}

// Device-resident state of the frame loop (single_batch.rs:19-28 fields that the
// reference keeps on the host: input_pos, previous_codes, prompt/None, rep-pen).
// The struct itself lives in device memory so that a captured frame graph stays
// valid across generate calls; kernels read it through a pointer.
struct GenState {
    int *pos;            // (B) cached positions == position of the next token
    int *active;         // (B) 1 while the row still generates
    int *eos;            // (B) slow token of the current frame was <|im_end|>
    int *frame;          // (B) frames emitted so far
    int *max_frames;     // (B) frame budget of the row (Q3 / fixed_len)
    int *n_active;       // (1) rows still active; kernels no-op when 0
    uint32_t *cur;       // (B, C+1) codes of the frame being built
    uint32_t *prev;      // (B, C+1) codes of the previous frame (previous_codes)
    uint32_t *out;       // (B, out_cap, C+1) every emitted frame
    RepPenState *rep;    // (B, C)
    int out_cap;
    int fixed_len;       // FSB_GEN_FIXED_LEN: <|im_end|> not eligible
    int legacy_slow;     // Fish <= 1.4: slow token is a 2-way PAD/EOS draw (single_batch.rs:104-124)
    int C;
    uint32_t im_end_id;
    uint32_t pad_id;
    SampleParams sp;
};

}  // namespace fsb
