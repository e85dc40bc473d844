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

// Block-wide sampler.  `vals` (smem, n_pad floats) holds the adjusted logits;
// returns the chosen index to every thread.  keys: smem n_pad u64.
__device__ int block_sample(float *vals, unsigned long long *keys, float *red, int n, int n_pad,
                            const SampleParams &sp, float u) {
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    // ---- max (and first argmax) ----
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    for (int i = tid; i < n; i += kSampleThreads) {
        float v = vals[i];
        if (v > best || (v == best && i < best_i)) { best = v; best_i = i; }
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
    if (warp == 0) {
        best = red[lane]; best_i = red_i[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
        }
        if (lane == 0) { red[0] = best; red_i[0] = best_i; }
    }
    __syncthreads();
    best = red[0]; best_i = red_i[0];
    __syncthreads();
    if (sp.greedy) return best_i;

    // ---- softmax(logits * inv_temp) ----
    const float mx = __fmul_rn(best, sp.inv_temp);
    float s = 0.f;
    for (int i = tid; i < n; i += kSampleThreads) {
        float e = expf(__fsub_rn(__fmul_rn(vals[i], sp.inv_temp), mx));
        vals[i] = e;
        s += e;
    }
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (warp == 0) {
        float t = red[lane];
        t = warp_sum(t);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    const float denom = red[0];
    // keys: (~prob bits) << 32 | index, ascending sort == prob desc, index asc
    for (int i = tid; i < n_pad; i += kSampleThreads) {
        unsigned long long k = ~0ull;
        if (i < n) {
            float p = vals[i] / denom;
            k = ((unsigned long long)(~__float_as_uint(p)) << 32) | (unsigned)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n_pad; i += kSampleThreads) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = keys[i], b = keys[ixj];
                    bool up = ((i & k) == 0);
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    // ---- top-k -> top-p -> multinomial, sequential f32 like the reference ----
    if (tid == 0) {
        const int k = (sp.top_k >= (uint32_t)n) ? n : (int)sp.top_k;
        float sum_p = 0.f;
        for (int i = 0; i < k; ++i) sum_p += __uint_as_float(~(uint32_t)(keys[i] >> 32));
        int kept = k;
        float total = sum_p;
        const bool do_topp = !(sp.top_p <= 0.f || sp.top_p >= sum_p) || (sp.top_k >= (uint32_t)n);
        if (do_topp) {
            float cumsum = 0.f;
            kept = 0;
            for (int i = 0; i < k; ++i) {
                if (cumsum >= sp.top_p) break;
                cumsum += __uint_as_float(~(uint32_t)(keys[i] >> 32));
                kept = i + 1;
            }
            total = cumsum;  // == sequential sum of the kept weights
        }
        const float chosen = u * total;
        float cum = 0.f;
        int pick = kept - 1;
        for (int i = 0; i < kept; ++i) {
            float w = __uint_as_float(~(uint32_t)(keys[i] >> 32));
            cum += w;
            if (cum > chosen && w > 0.f) { pick = i; break; }
        }
        // last kept entry with non-zero weight if nothing crossed `chosen`
        while (pick > 0 && __uint_as_float(~(uint32_t)(keys[pick] >> 32)) <= 0.f) --pick;
        red_i[0] = (int)(keys[pick] & 0xffffffffu);
    }
    __syncthreads();
    int r = red_i[0];
    __syncthreads();
    return r;
}

}  // namespace fsb
