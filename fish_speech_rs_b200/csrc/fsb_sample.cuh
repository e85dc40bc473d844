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
    float top_p;        // f32(top_p): the threshold sample_topp compares running sums with (mod.rs:70 `top_p as f32`)
    float top_p_gate;   // largest f32 <= the caller's f64 top_p (+inf if top_p <= 0): `sum_p <= top_p_gate` is exactly the
                        // reference's f64 gate `top_p <= 0.0 || top_p >= sum_p as f64` (mod.rs:67) for an f32 sum_p
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

// Block-synchronisation policy of the sampler: the whole CTA (stand-alone sampler kernels) or the
// compute warps of a warp-specialised kernel (named barrier; the TMA producer warp does not take part).
struct SyncAll {
    static __device__ __forceinline__ void sync() { __syncthreads(); }
    static __device__ __forceinline__ int nthreads() { return (int)blockDim.x; }
};
template <int NT, int BAR>
struct SyncNamed {
    static __device__ __forceinline__ void sync() { asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NT) : "memory"); }
    static __device__ __forceinline__ int nthreads() { return NT; }
};

template <int E, class SY>
__device__ __forceinline__ int block_sample_t(float *vals, unsigned long long *keys, float *red, int n, int n_pad,
                                              const SampleParams &sp, float u) {
    const bool tm_ = g_sample_dbg != nullptr && threadIdx.x == 0 && blockIdx.x == 0;
    const long long q0 = tm_ ? clock64() : 0;
    const int tid = threadIdx.x, nthreads = SY::nthreads();
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
    SY::sync();
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
        SY::sync();
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
    SY::sync();  // red[] reads above are done
    if (lane == 0) red[warp] = s;
    SY::sync();
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
                SY::sync();
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
        SY::sync();
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
        const bool do_topp = !(sum_p <= sp.top_p_gate) || (sp.top_k >= (uint32_t)n);
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
    SY::sync();
    int r = red_i[0];
    SY::sync();
    if (tm_) {
        const long long q3 = clock64();
        g_sample_dbg[0] += q1 - q0; g_sample_dbg[1] += q2 - q1; g_sample_dbg[2] += q3 - q2; g_sample_dbg[3] += 1;
    }
    return r;
}

// elements per thread is a compile-time constant of the sort network (a generic loop over "up to 8"
// slots made the sampler issue-bound: 17 us of a 21 us draw)
template <class SY = SyncAll>
__device__ inline int block_sample(float *vals, unsigned long long *keys, float *red, int n, int n_pad,
                                   const SampleParams &sp, float u) {
    const int E = max(n_pad, SY::nthreads()) / SY::nthreads();
    switch (E) {
        case 1: return block_sample_t<1, SY>(vals, keys, red, n, n_pad, sp, u);
        case 2: return block_sample_t<2, SY>(vals, keys, red, n, n_pad, sp, u);
        case 4: return block_sample_t<4, SY>(vals, keys, red, n, n_pad, sp, u);
        default: return block_sample_t<8, SY>(vals, keys, red, n, n_pad, sp, u);
    }
}

// ---------------------------------------------------------------- selection sampler (top_k <= 256)
// Same candidates, same order, same running sums as block_sample_t, without sorting all n entries:
//   1. softmax, composite keys  (~prob bits) << 12 | index  (unique, ascending == prob desc / index asc)
//   2. radix select of the k-th smallest key: 9 rounds of 5-bit digits, per-warp 32-bin histograms filled
//      with match_any, one block barrier per round (every warp redoes the 32-lane scan itself)
//   3. compaction of the k keys <= threshold, rank of each by counting (k^2 / 2 broadcast compares),
//      scatter into sorted order
//   4. top-p cut and multinomial walk on warp 0 (identical arithmetic to block_sample_t)
// Shared-memory scratch: kSelScratchBytes(nthreads).
constexpr int kSelMaxK = 256;
constexpr int kSelIdxBits = 12;  // n <= 4096
__host__ __device__ constexpr int sel_list_entries(int nthreads) { return 256 + 2 * (nthreads / 256); }
__host__ __device__ constexpr int sel_scratch_bytes(int nthreads) {
    return sel_list_entries(nthreads) * 8 + 256 * 8 + 3 * (nthreads / 32) * 32 * 4 + 64;
}

template <int E, class SY>
__device__ __noinline__ int block_sample_sel_t(const float *vals, unsigned char *scratch, float *red, int n,
                                               const SampleParams &sp, float u) {
    const int tid = threadIdx.x, nthreads = SY::nthreads();
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int G = nthreads / 256, seglen = 256 / G;  // threads per candidate in the rank step
    unsigned long long *list = reinterpret_cast<unsigned long long *>(scratch);   // G segments of seglen + 2
    unsigned long long *sorted = list + sel_list_entries(nthreads);               // 256
    unsigned *hist = reinterpret_cast<unsigned *>(sorted + 256);                  // 3 x nwarps x 32
    unsigned *counter = hist + 3 * nwarps * 32;
    int *red_i = reinterpret_cast<int *>(red + 32);
    const bool tm_ = g_sample_dbg != nullptr && threadIdx.x == 0 && blockIdx.x == 0;
    const long long q0 = tm_ ? clock64() : 0;
    float v[E];
    // ---- max (and first argmax) ----
    float best = -INFINITY;
    int best_i = 0x7fffffff;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * nthreads + tid;
        v[e] = (i < n) ? vals[i] : -INFINITY;
        if (v[e] > best) { best = v[e]; best_i = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) { red[warp] = best; red_i[warp] = best_i; }
    // scratch initialisation rides on the same barrier
    for (int i = tid; i < sel_list_entries(nthreads); i += nthreads) list[i] = ~0ull;
    if (warp == 0) { hist[lane] = 0; hist[32 + lane] = 0; }
    if (tid == 0) *counter = 0;
    SY::sync();
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
        SY::sync();
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
    SY::sync();  // red[] reads above are done
    if (lane == 0) red[warp] = s;
    SY::sync();
    float denom = lane < nwarps ? red[lane] : 0.f;
    denom = warp_sum(denom);
    unsigned long long key[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * nthreads + tid;
        key[e] = ~0ull;
        if (i < n) {
            const float p = v[e] / denom;
            key[e] = ((unsigned long long)(~__float_as_uint(p)) << kSelIdxBits) | (unsigned)i;
        }
    }
    const long long q1 = tm_ ? clock64() : 0;
    // ---- radix select: k-th smallest key ----
    // Round r looks at a 5-bit digit: lane b of every warp counts the warp's matching elements whose digit
    // is b from six ballots per element (no match_any, no per-element shared-memory traffic), adds the
    // count to the block histogram (32 bins, one shared atomic per lane), and after one barrier every warp
    // scans the 32 totals itself.
    const int k = (sp.top_k >= (uint32_t)n) ? n : (int)sp.top_k;
    unsigned long long prefix = 0;
    int krem = k;
#pragma unroll 1
    for (int r = 0; r < 9; ++r) {
        const long long r0 = tm_ ? clock64() : 0;
        const int shift = 40 - 5 * r;
        unsigned *hb = hist + (r % 3) * 32, *hn = hist + ((r + 1) % 3) * 32;
        unsigned cnt = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const bool part = (key[e] >> (shift + 5)) == prefix;
            const unsigned d = (unsigned)(key[e] >> shift) & 31u;
            unsigned m = __ballot_sync(0xffffffffu, part);
#pragma unroll
            for (int bit = 0; bit < 5; ++bit) {
                const unsigned bb = __ballot_sync(0xffffffffu, part && ((d >> bit) & 1u));
                m &= ((lane >> bit) & 1) ? bb : ~bb;
            }
            cnt += __popc(m);
        }
        const long long r1 = tm_ ? clock64() : 0;
        if (cnt) atomicAdd(hb + lane, cnt);
        if (warp == 0 && r >= 1) hn[lane] = 0;  // buffers 0 and 1 were cleared up front
        SY::sync();
        const unsigned tot = hb[lane];
        const long long r2 = tm_ ? clock64() : 0;
        unsigned cum = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, cum, o);
            if (lane >= o) cum += t;
        }
        const unsigned ball = __ballot_sync(0xffffffffu, cum >= (unsigned)krem);
        const int d = __ffs(ball) - 1;  // ball != 0: the matching elements number >= krem by construction
        const unsigned below = __shfl_sync(0xffffffffu, cum - tot, d);
        const unsigned in_bucket = __shfl_sync(0xffffffffu, tot, d);
        krem -= (int)below;
        prefix = (prefix << 5) | (unsigned)d;
        if (tm_) {
            const long long r3 = clock64();
            g_sample_dbg[9] += r1 - r0; g_sample_dbg[10] += r2 - r1; g_sample_dbg[11] += r3 - r2;
        }
        // every key of this bucket is among the k smallest: the threshold is the largest key with this prefix and the
        // remaining digits need no refinement (block-uniform: all threads read the same totals).  Typically 4-5 rounds
        // instead of 9 for 1024 keys.
        if ((unsigned)krem == in_bucket) {
            prefix = (prefix << shift) | ((1ull << shift) - 1ull);
            break;
        }
    }
    const unsigned long long kth = prefix;
    const long long q2 = tm_ ? clock64() : 0;
    // ---- compaction of the k candidates (order irrelevant) ----
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const bool cand = key[e] <= kth;
        const unsigned m = __ballot_sync(0xffffffffu, cand);
        unsigned base = 0;
        if (lane == 0 && m) base = atomicAdd(counter, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (cand) {
            const unsigned pos = base + __popc(m & ((1u << lane) - 1u));
            list[(pos / seglen) * (seglen + 2) + (pos % seglen)] = key[e];
        }
    }
    SY::sync();
    // ---- rank by counting: thread (candidate c, segment g); unused slots hold ~0 ----
    {
        const int c = tid % 256, g = tid / 256;
        const unsigned long long mine = list[(c / seglen) * (seglen + 2) + (c % seglen)];
        const int nj = min(seglen, max(0, k - g * seglen));
        const uint4 *seg = reinterpret_cast<const uint4 *>(list + g * (seglen + 2));
        unsigned cnt = 0;
#pragma unroll 8
        for (int j = 0; j < (nj + 1) / 2; ++j) {  // (unrolled: the 16-byte loads of 8 steps are in flight together)
            const uint4 q = seg[j];
            const unsigned long long a = ((unsigned long long)q.y << 32) | q.x, b = ((unsigned long long)q.w << 32) | q.z;
            cnt += (a < mine) ? 1u : 0u;
            cnt += (b < mine) ? 1u : 0u;
        }
        // partner threads of one candidate sit 256 threads apart: combine through shared memory
        unsigned *rank = reinterpret_cast<unsigned *>(sorted);  // reused below only after the next barrier
        if (G > 1) {
            unsigned *part = hist;  // histograms are dead; G * 256 counters fit (3 * nwarps * 32 >= G * 256)
            part[g * 256 + c] = cnt;
            SY::sync();
            if (g == 0) {
                for (int gg = 1; gg < G; ++gg) cnt += part[gg * 256 + c];
            }
        }
        (void)rank;
        SY::sync();
        if (g == 0 && c < k) sorted[cnt] = mine;
    }
    SY::sync();
    const long long q3 = tm_ ? clock64() : 0;
    // ---- top-k -> top-p -> multinomial on warp 0 (arithmetic identical to block_sample_t) ----
    if (warp == 0) {
        const int seg = (k + 31) / 32;  // <= 8
        const int i0 = min(k, lane * seg), i1 = min(k, i0 + seg);
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            w[j] = (i0 + j < i1) ? __uint_as_float(~(uint32_t)(sorted[i0 + j] >> kSelIdxBits)) : 0.f;
        float local = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (i0 + j < i1) local += w[j];
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
        const bool do_topp = !(sum_p <= sp.top_p_gate) || (sp.top_k >= (uint32_t)n);
        if (do_topp) {
            int kl = 0;
            float tl = 0.f, c = excl;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (i0 + j < i1) {
                    const bool keep = c < sp.top_p;
                    c += w[j];
                    if (keep) { kl = i0 + j + 1; tl = c; }
                }
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
        const float chosen = u * total;
        int cand = 0x7fffffff;
        {
            float c = excl;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (i0 + j < min(i1, kept) && cand == 0x7fffffff) {
                    c += w[j];
                    if (c > chosen && w[j] > 0.f) cand = i0 + j;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        if (lane == 0) {
            int pick = cand;
            if (pick == 0x7fffffff) {
                pick = max(kept - 1, 0);
                while (pick > 0 && __uint_as_float(~(uint32_t)(sorted[pick] >> kSelIdxBits)) <= 0.f) --pick;
            }
            red_i[0] = (int)(sorted[pick] & ((1u << kSelIdxBits) - 1u));
        }
    }
    SY::sync();
    const int rsel = red_i[0];
    SY::sync();
    if (tm_) {
        const long long q4 = clock64();
        g_sample_dbg[0] += q1 - q0; g_sample_dbg[1] += q2 - q1; g_sample_dbg[2] += q3 - q2; g_sample_dbg[3] += 1;
        g_sample_dbg[4] += q4 - q3;
    }
    return rsel;
}

// n <= 8 * nthreads, n <= 4096, top_k <= kSelMaxK (callers fall back to block_sample otherwise)
template <class SY>
__device__ __forceinline__ int block_sample_sel(const float *vals, unsigned char *scratch, float *red, int n,
                                                const SampleParams &sp, float u) {
    const int E = (n + SY::nthreads() - 1) / SY::nthreads();
    if (E <= 2) return block_sample_sel_t<2, SY>(vals, scratch, red, n, sp, u);
    if (E <= 4) return block_sample_sel_t<4, SY>(vals, scratch, red, n, sp, u);
    return block_sample_sel_t<8, SY>(vals, scratch, red, n, sp, u);
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
