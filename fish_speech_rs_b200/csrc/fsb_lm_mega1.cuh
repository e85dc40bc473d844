// Persistent decode megakernel, single-utterance specialisation (decode_mode 2, one batch row).
//
// Same phase structure as fsb_lm_mega.cuh (one cooperative launch runs whole frames of the dual-AR
// loop of single_batch.rs:76-214 with grid barriers between dependent phases), rebuilt around what the
// per-phase timers of that kernel showed on B200:
//
//   * the weight stream is decoupled from the compute warps: a dedicated producer warp walks the
//     (static) phase schedule ahead of the consumers and moves each CTA's contiguous weight slice with
//     TMA bulk copies (cp.async.bulk global -> shared, mbarrier complete_tx) into a ring of 32 KB
//     chunks.  HBM keeps streaming through grid barriers, prologues and samplers; a phase's dot
//     products read shared memory only.
//   * the activation vector lives in REGISTERS: a task is a 1024-element K-slice of one weight row,
//     lane l always owns the same 32 columns, so x is read from shared memory once per phase instead
//     of once per task (the old kernel's w13 phase was bound by those LDS, not by HBM).
//   * RMSNorm is folded: the staged vector is x * g, sum(x^2) is reduced on the side and 1 / denom
//     scales the finished dot products -- no block-wide norm pass before the first FMA.
//   * warp reductions are deferred and interleaved (up to 8 tasks per warp in flight).
//   * the residual stream of the CTA's own rows is kept in shared memory (no L2 round trip in the
//     wo / w2 epilogues); attention items cover 64 positions (twice the CTAs of the old kernel).
//
// Reference call sites replaced: dual_ar.rs:160-165,239-384,429-440,574-673;
// generate/single_batch.rs:76-214; sampling/mod.rs; sampling/rep_pen.rs.
#pragma once
#include "fsb_lm_mega.cuh"

namespace fsb {

typedef SyncNamed<kM1Threads, 1> M1Sync;

__device__ __forceinline__ uint32_t m1_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void m1_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m1_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void m1_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m1_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void m1_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(m1_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void m1_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(m1_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// contiguous global -> shared bulk copy (TMA engine, no tensor map); bytes % 16 == 0
__device__ __forceinline__ void m1_bulk_g2s(void *smem, const void *gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     m1_smem_u32(smem)),
                 "l"(gmem), "r"(bytes), "r"(m1_smem_u32(bar))
                 : "memory");
}

// The CTA's share of one weight phase: a stream of tasks (1024-element K-slices, row-major inside
// the CTA's contiguous row block), made of <= 3 contiguous global segments.
struct M1Plan {
    const char *seg[3];
    unsigned seg_bytes[3];
    int nseg;
    int r0, nrows, ksplit, nmat, ntasks;
};

template <typename WT>
struct Mega1 {
    static constexpr int NE = WTraits<WT>::NE;                 // elements per 16 bytes
    static constexpr int TB = kM1Slice * (int)sizeof(WT);      // task bytes
    static constexpr int U = TB / 512;                         // 16-byte units per lane per task
    static constexpr int CT = kM1ChunkBytes / TB;              // tasks (= warps) per chunk
    static constexpr int NG = kM1Warps / CT;                   // warp groups working on different chunks

    const MegaParams &p;
    // shared memory
    unsigned char *ring;
    float *xs, *xres, *val, *red, *kvs, *cs_s, *csf_s, *hmerge;
    int *pos_s, *rtab;
    const MegaLayer *ltab;
    int *s_active, *s_eos, *s_frame, *s_maxf;
    uint32_t *s_cur, *s_prev;
    RepPenState *s_rep;
    uint64_t *full, *empty;
    volatile int *go_frames, *done_flag;
    int tid, lane, warp;
    unsigned int target;
    unsigned int gchunk;  // chunks consumed (consumers) / issued (producer) so far
    unsigned int cslot, cpar;  // ring slot and mbarrier parity of chunk `gchunk` (consumers)
    float xr[32];         // this lane's columns of the staged activation slice (x * g where a norm applies)
    float inv_denom;      // 1 / sqrt(mean(x^2) + eps) of the phase (1 without a norm)
    float2 gpre;          // norm weights of the coming phase for elements 2 * tid, 2 * tid + 1
    struct { int r0, nrows, ksplit, ntasks; } plan;  // consumers' view of the phase (the producer builds full M1Plans)
    int att_item, att_n;

    __device__ Mega1(const MegaParams &pp, unsigned char *smem) : p(pp) {
        ring = smem;
        float *f = reinterpret_cast<float *>(smem + (size_t)pp.ring_depth * kM1ChunkBytes);
        xs = f; f += pp.xs_floats;
        xres = f; f += pp.D;
        val = f; f += kM1ValFloats;
        red = f; f += 64;
        cs_s = f; f += 64;
        csf_s = f; f += 8 * 64;
        hmerge = xs;  // attention phases only: xs is idle then
        kvs = f; f += pp.kvs_floats;
        pos_s = reinterpret_cast<int *>(f); f += 4;
        rtab = reinterpret_cast<int *>(f); f += 12;
        ltab = reinterpret_cast<const MegaLayer *>(f); f += (sizeof(MegaLayer) / 4) * (pp.NL + pp.NFL);
        s_active = reinterpret_cast<int *>(f); f += 1;
        s_eos = reinterpret_cast<int *>(f); f += 1;
        s_frame = reinterpret_cast<int *>(f); f += 1;
        s_maxf = reinterpret_cast<int *>(f); f += 1;
        s_cur = reinterpret_cast<uint32_t *>(f); f += 20;
        s_prev = reinterpret_cast<uint32_t *>(f); f += 20;
        s_rep = reinterpret_cast<RepPenState *>(f); f += (sizeof(RepPenState) / 4) * 8;
        full = reinterpret_cast<uint64_t *>(f); f += 2 * kM1MaxDepth;
        empty = reinterpret_cast<uint64_t *>(f); f += 2 * kM1MaxDepth;
        go_frames = reinterpret_cast<volatile int *>(f); f += 1;
        done_flag = reinterpret_cast<volatile int *>(f); f += 1;
        tid = threadIdx.x;
        lane = tid & 31;
        warp = tid >> 5;
        target = 0;
        gchunk = 0;
        cslot = 0;
        cpar = 0;
        att_item = -1;
        att_n = 0;
        inv_denom = 1.f;
        gpre = make_float2(1.f, 1.f);
    }

    static __device__ __forceinline__ void csync() { M1Sync::sync(); }

    __device__ __forceinline__ void grid_arrive1() {
        csync();
        if (tid == 0) {
            target += gridDim.x;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.bar) : "memory");
        }
    }
    __device__ __forceinline__ void grid_wait1() {
        if (tid == 0) {
            while ((int)(ld_relaxed_u32(p.bar) - target) < 0) {}
            asm volatile("fence.acquire.gpu;" ::: "memory");
        }
        csync();
    }

    // ------------------------------------------------------------ schedule (shared by producer and consumers)
    enum { K_QKV = 0, K_ATT = 1, K_WO = 2, K_W13 = 3, K_W2 = 4, K_HEAD = 5, K_SAMPLE = 6, K_END = 7 };
    struct Step { int frame, pass, l, kind; };  // pass 0 = slow stack, pass c+1 = fast step of codebook c

    // layer tables are copied to shared memory once (prep and the producer chase them every phase)
    __device__ __forceinline__ const MegaLayer &layer_of(const Step &s) const { return ltab[s.pass == 0 ? s.l : p.NL + s.l]; }

    __device__ __forceinline__ Step first_step() const {
        Step s;
        s.frame = 0; s.pass = 0; s.l = 0;
        s.kind = (p.first_is_tail || p.NL == 0) ? K_HEAD : K_QKV;
        return s;
    }
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
    __device__ __forceinline__ Step next_weight_step(Step s) const {
        do { s = advance(s); } while (s.kind == K_ATT || s.kind == K_SAMPLE);
        return s;
    }

    // row block of this CTA per phase kind (constant for the launch): {r0, nrows} pairs in shared memory
    enum { R_QKV = 0, R_WO = 1, R_W13 = 2, R_W2 = 3, R_HEAD_SLOW = 4, R_HEAD_FAST = 5, R_COUNT = 6 };
    // The LAST CTA of the grid is the sampler: it owns no weight rows, runs no attention item and no TMA
    // producer -- its instruction working set is the sampler alone (warm in its instruction caches), and
    // the 147 streaming CTAs never fetch sampler code.
    // (p.sampler_cta < 0: CTA 0 samples and streams like everybody else)
    __device__ __forceinline__ bool is_sampler() const { return (int)blockIdx.x == p.sampler_cta; }
    __device__ __forceinline__ int n_compute() const { return (int)gridDim.x - (p.sampler_cta >= 0 ? 1 : 0); }

    __device__ __forceinline__ void init_row_ranges() {
        if (threadIdx.x < R_COUNT) {
            const int k = threadIdx.x;
            const int rows_total = k == R_QKV ? p.QKV : k == R_W13 ? p.I : k == R_HEAD_SLOW ? p.n_slow_logits
                                 : k == R_HEAD_FAST ? p.CS : p.D;
            const int align = k == R_QKV ? 2 : 1;
            const unsigned groups = rows_total / align, nc = (unsigned)n_compute();
            const unsigned g0 = (blockIdx.x * groups) / nc, g1 = ((blockIdx.x + 1) * groups) / nc;
            const int r0 = (int)g0 * align;
            rtab[2 * k] = is_sampler() ? 0 : r0;
            rtab[2 * k + 1] = is_sampler() ? 0 : (blockIdx.x + 1 == nc ? rows_total : (int)g1 * align) - r0;
        }
    }

    __device__ __forceinline__ M1Plan make_plan(int rk, const void *W0, const void *W1, int K, int row_a, int row_b) const {
        M1Plan pl;
        pl.r0 = rtab[2 * rk];
        pl.nrows = rtab[2 * rk + 1];
        pl.ksplit = K / kM1Slice;
        pl.nmat = W1 ? 2 : 1;
        pl.ntasks = pl.nrows * pl.ksplit * pl.nmat;
        const size_t rowb = (size_t)K * sizeof(WT);
        const char *w0 = reinterpret_cast<const char *>(W0);
        pl.nseg = 0;
        if (pl.nrows > 0) {
            if (W1) {
                pl.seg[0] = w0 + (size_t)pl.r0 * rowb;
                pl.seg[1] = reinterpret_cast<const char *>(W1) + (size_t)pl.r0 * rowb;
                pl.seg_bytes[0] = pl.seg_bytes[1] = (unsigned)(pl.nrows * rowb);
                pl.nseg = 2;
            } else if (pl.r0 == 0 && row_b != row_a + 1) {
                // logical row 0 -> weight row row_a, logical row r >= 1 -> row_b + r - 1 (constrained slow head, Q5)
                pl.seg[0] = w0 + (size_t)row_a * rowb;
                pl.seg_bytes[0] = (unsigned)rowb;
                pl.nseg = 1;
                if (pl.nrows > 1) {
                    pl.seg[1] = w0 + (size_t)row_b * rowb;
                    pl.seg_bytes[1] = (unsigned)((pl.nrows - 1) * rowb);
                    pl.nseg = 2;
                }
            } else {
                const int wrow = pl.r0 == 0 ? row_a : row_b + pl.r0 - 1;
                pl.seg[0] = w0 + (size_t)wrow * rowb;
                pl.seg_bytes[0] = (unsigned)(pl.nrows * rowb);
                pl.nseg = 1;
            }
        }
        return pl;
    }

    __device__ __forceinline__ M1Plan plan_of(const Step &s) const {
        const bool slow = s.pass == 0;
        int rk = R_HEAD_FAST, K = p.D, row_a = 0, row_b = 1;
        const void *W0 = p.fast_out, *W1 = nullptr;
        if (s.kind == K_HEAD) {
            if (slow) { rk = R_HEAD_SLOW; W0 = p.out_w; row_a = p.slow_row0; row_b = p.slow_rest_base; }
        } else {
            const MegaLayer &L = layer_of(s);
            switch (s.kind) {
                case K_QKV: rk = R_QKV; W0 = L.wqkv; break;
                case K_WO: rk = R_WO; W0 = L.wo; K = p.H * p.hd; break;
                case K_W13: rk = R_W13; W0 = L.w1; W1 = L.w3; break;
                default: rk = R_W2; W0 = L.w2; K = p.I; break;
            }
        }
        return make_plan(rk, W0, W1, K, row_a, row_b);  // one body: the plan code stays warm in the I-cache
    }

    // ------------------------------------------------------------ producer warp (lane 0)
    __device__ __forceinline__ void producer() {
        Step s = first_step();
        if (s.kind == K_ATT || s.kind == K_SAMPLE) s = next_weight_step(s);
        const int depth = p.ring_depth;
        while (s.kind != K_END) {
            // frames beyond the last confirmed one are not streamed (a finished utterance must not
            // leave bulk copies in flight when the CTA exits)
            while (s.frame >= *go_frames) {
                if (*done_flag) return;
                __nanosleep(64);
            }
            const M1Plan pl = plan_of(s);
            const unsigned total = (unsigned)pl.ntasks * TB;
            for (unsigned off = 0; off < total; off += kM1ChunkBytes) {
                const unsigned bytes = min((unsigned)kM1ChunkBytes, total - off);
                const unsigned slot = gchunk % depth, use = gchunk / depth;
                if (use > 0) m1_mbar_wait(empty + slot, (use - 1) & 1);
                m1_mbar_expect_tx(full + slot, bytes);
                unsigned char *dst = ring + (size_t)slot * kM1ChunkBytes;
                // intersect [off, off + bytes) with the segments
                unsigned seg_lo = 0;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    if (i < pl.nseg) {
                        const unsigned seg_hi = seg_lo + pl.seg_bytes[i];
                        const unsigned lo = max(off, seg_lo), hi = min(off + bytes, seg_hi);
                        if (lo < hi) m1_bulk_g2s(dst + (lo - off), pl.seg[i] + (lo - seg_lo), hi - lo, full + slot);
                        seg_lo = seg_hi;
                    }
                }
                ++gchunk;
            }
            s = next_weight_step(s);
        }
    }

    // ------------------------------------------------------------ consumers: dot products of one phase
    // Layout of a staged 1024-element slice in xs.  bf16 weights: a lane's 16-byte weight unit covers 8
    // consecutive columns = two float4 of x; the two halves are kept in separate 512-float planes so that
    // the 32 lanes of a warp read 32 consecutive float4 (conflict-free LDS.128).  Index of the float4 that
    // holds columns [4 * e4, 4 * e4 + 4) of the slice:
    static __device__ __forceinline__ int x4_index(int e4) {
        if (NE == 8) return ((e4 & 1) << 7) | (e4 >> 1);
        return e4;
    }
    static __device__ __forceinline__ int x_index(int e) { return (e & ~1023) | (x4_index((e & 1023) >> 2) << 2) | (e & 3); }

    // xr <- staged slice `sl` of xs
    __device__ __forceinline__ void load_xr(int sl) {
        const float4 *x4 = reinterpret_cast<const float4 *>(xs + sl * kM1Slice);
        if (NE == 8) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 lo = x4[i * 32 + lane], hi = x4[128 + i * 32 + lane];
                xr[i * 8 + 0] = lo.x; xr[i * 8 + 1] = lo.y; xr[i * 8 + 2] = lo.z; xr[i * 8 + 3] = lo.w;
                xr[i * 8 + 4] = hi.x; xr[i * 8 + 5] = hi.y; xr[i * 8 + 6] = hi.z; xr[i * 8 + 7] = hi.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 a = x4[i * 32 + lane];
                xr[i * 4 + 0] = a.x; xr[i * 4 + 1] = a.y; xr[i * 4 + 2] = a.z; xr[i * 4 + 3] = a.w;
            }
        }
    }

    __device__ __forceinline__ float task_dot1(const unsigned char *src) const {
        const uint4 *w4 = reinterpret_cast<const uint4 *>(src) + lane;
        uint4 v[U];
#pragma unroll
        for (int i = 0; i < U; ++i) v[i] = w4[i * 32];
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < U; ++i) {
            float w[NE];
            unpack16<WT>(v[i], w);
#pragma unroll
            for (int j = 0; j < NE; j += 2) {
                a0 = fmaf(w[j], xr[i * NE + j], a0);
                a1 = fmaf(w[j + 1], xr[i * NE + j + 1], a1);
            }
        }
        return a0 + a1;
    }

    // all tasks of the CTA: val[t] = dot(task t, x slice) * inv_denom.  Warp w handles task (w % CT) of the
    // chunks c == (w / CT) mod NG.  Deliberately a rolled loop: every phase runs this code once, so the
    // instruction footprint (not the shuffle latency, which the other warps hide) is what matters.
    __device__ __forceinline__ void run_tasks1() {
        const int depth = p.ring_depth;
        const int nchunks = (plan.ntasks + CT - 1) / CT;
        const int tw = warp % CT;
        // slot / parity of chunk gchunk + c, advanced without divisions (cslot, cpar track gchunk itself)
        unsigned slot = cslot + (unsigned)(warp / CT), par = cpar;
        if (slot >= (unsigned)depth) { slot -= depth; par ^= 1; }
#pragma unroll 1
        for (int c = warp / CT; c < nchunks; c += NG, slot += NG) {
            if (slot >= (unsigned)depth) { slot -= depth; par ^= 1; }
            m1_mbar_wait(full + slot, par);
            const int t = c * CT + tw;
            float a = 0.f;
            if (t < plan.ntasks) a = task_dot1(ring + (size_t)slot * kM1ChunkBytes + (size_t)tw * TB);
            __syncwarp();
            if (lane == 0) m1_mbar_arrive(empty + slot);
            a = warp_sum(a);
            if (lane == 0 && t < plan.ntasks) val[t] = a * inv_denom;
        }
        gchunk += (unsigned)nchunks;
        cslot += (unsigned)nchunks;
        while (cslot >= (unsigned)depth) { cslot -= depth; cpar ^= 1; }
    }

    __device__ __forceinline__ float row_val(int m, int rl) const {
        const int t0 = (m * plan.nrows + rl) * plan.ksplit;
        float s = val[t0];
        for (int k = 1; k < plan.ksplit; ++k) s += val[t0 + k];
        return s;
    }

    // ------------------------------------------------------------ activation staging
    // x (D floats, global) -> xres (raw) and xs (x * g); every warp ends with the full sum(x^2)
    __device__ __forceinline__ void stage_norm(const float *src, int K, bool with_norm) {
        float ss = 0.f;
        for (int k2 = tid; k2 < K / 2; k2 += kM1Threads) {
            const float2 v = __ldcg(reinterpret_cast<const float2 *>(src) + k2);
            reinterpret_cast<float2 *>(xres)[k2] = v;
            float2 o = v;
            if (with_norm) {
                o.x = __fmul_rn(v.x, gpre.x);
                o.y = __fmul_rn(v.y, gpre.y);
                ss = fmaf(v.x, v.x, fmaf(v.y, v.y, ss));
            }
            *reinterpret_cast<float2 *>(xs + x_index(2 * k2)) = o;
        }
        if (with_norm) {
            ss = warp_sum(ss);
            if (lane == 0) red[warp] = ss;
        }
    }
    // after the csync that follows stage_norm
    __device__ __forceinline__ void finish_norm(int K, bool with_norm) {
        if (with_norm) {
            float t = lane < kM1Warps ? red[lane] : 0.f;
            t = warp_sum(t);
            inv_denom = rsqrtf(t / (float)K + p.eps);
        } else {
            inv_denom = 1.f;
        }
    }
    __device__ __forceinline__ void stage_plain(const float *src, int K) {
        const int n4 = K / 4;
        for (int i = tid; i < n4; i += kM1Threads)
            reinterpret_cast<float4 *>(xs)[(i & ~255) | x4_index(i & 255)] = __ldcg(reinterpret_cast<const float4 *>(src) + i);
    }

    // ------------------------------------------------------------ split-KV GQA attention (slow blocks)
    // item = kvh * n_chunks_max + chunk; positions [chunk*64, min(len, chunk*64 + 64))
    __device__ __forceinline__ void att_stage(const float *kcache, const float *vcache, int kvh, int j0, int from, int to) {
        const float *kb = kcache + ((size_t)kvh * p.max_len + j0) * p.hd;
        const float *vb = vcache + ((size_t)kvh * p.max_len + j0) * p.hd;
        float *ks = kvs, *vs = kvs + kM1AttChunk * kM1KvStride;
        // hd == 64 (host-checked): 16 16-byte segments per row
        for (int i = tid; i < (to - from) * 16; i += kM1Threads) {
            const int j = from + (i >> 4), sg = i & 15;
            cp_async16(ks + j * kM1KvStride + sg * 4, kb + j * 64 + sg * 4);
            cp_async16(vs + j * kM1KvStride + sg * 4, vb + j * 64 + sg * 4);
        }
    }
    __device__ __forceinline__ int att_chunks() const { return (pos_s[0] + 1 + kM1AttChunk - 1) / kM1AttChunk; }

    // before the barrier that precedes K_ATT: stage what is already cached of this CTA's first item
    __device__ __forceinline__ void att_prefetch(int layer) {
        att_item = -1;
        const int len = pos_s[0] + 1, nch = att_chunks();
        const int nitems = p.KV * nch;
        const size_t slow_kv = (size_t)p.max_batch * p.KV * p.max_len * p.hd;
        if ((int)blockIdx.x < nitems && !is_sampler()) {
            const int item = blockIdx.x, kvh = item / nch, chunk = item - kvh * nch;
            const int j0 = chunk * kM1AttChunk, j1 = min(len - 1, j0 + kM1AttChunk);  // exclude position len-1
            att_item = item;
            att_n = max(j1 - j0, 0);
            if (att_n > 0) att_stage(p.kc + layer * slow_kv, p.vc + layer * slow_kv, kvh, j0, 0, att_n);
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
        const int len = pos_s[0] + 1, nch = att_chunks();
        const int nitems = p.KV * nch;
        const float *ks = kvs, *vs = kvs + kM1AttChunk * kM1KvStride;
        for (int item = is_sampler() ? nitems : (int)blockIdx.x; item < nitems; item += n_compute()) {
            const int kvh = item / nch, chunk = item - kvh * nch;
            const int j0 = chunk * kM1AttChunk, j1 = min(len, j0 + kM1AttChunk);
            const int have = item == att_item ? att_n : 0;
            const int h = kvh * n_rep + hq;
            float4 qv[4];
            if (hq < n_rep) {
                const float *qp = p.q + (size_t)h * p.hd + sub * 4;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) qv[jj] = __ldcg(reinterpret_cast<const float4 *>(qp + jj * 16));
            }
            csync();  // previous item's smem reads are done
            if (j0 + have < j1) att_stage(kcache, vcache, kvh, j0, have, j1 - j0);
            cp_async_commit();
            cp_async_wait_all();
            csync();
            float m = -INFINITY, l = 0.f;
            float o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = 0.f;
            if (hq < n_rep) {
                const int n = j1 - j0;
                const int mid = (n + 1) / 2;
                const int a0 = hf == 0 ? 0 : mid, a1 = hf == 0 ? mid : n;
                for (int jb = a0; jb < a1; jb += 8) {  // warp-uniform trip count (the shuffles need all lanes)
                    const int j = jb + g;
                    const bool valid = j < a1;
                    const int jc = valid ? j : a0;
                    const float *kr = ks + jc * kM1KvStride + sub * 4, *vr = vs + jc * kM1KvStride + sub * 4;
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
                // second half of the positions -> shared memory, merged by the first half's warp
                if (hf == 1 && g == 0) {
                    float *dst = hmerge + hq * 68;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        *reinterpret_cast<float4 *>(dst + jj * 16 + sub * 4) =
                            make_float4(o[jj * 4 + 0], o[jj * 4 + 1], o[jj * 4 + 2], o[jj * 4 + 3]);
                    if (sub == 0) { dst[64] = m; dst[65] = l; }
                }
            }
            csync();
            if (hq < n_rep && hf == 0 && g == 0) {
                const float *src = hmerge + hq * 68;
                const float mo = src[64], lo = src[65];
                const float M = fmaxf(m, mo);
                const float wa = (m == -INFINITY) ? 0.f : expf(m - M), wb = (mo == -INFINITY) ? 0.f : expf(mo - M);
                l = l * wa + lo * wb;
                float *out = p.partial + ((size_t)h * (2 * p.n_chunks_max) + chunk) * (p.hd + 4);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float4 oo = *reinterpret_cast<const float4 *>(src + jj * 16 + sub * 4);
                    *reinterpret_cast<float4 *>(out + jj * 16 + sub * 4) =
                        make_float4(o[jj * 4 + 0] * wa + oo.x * wb, o[jj * 4 + 1] * wa + oo.y * wb,
                                    o[jj * 4 + 2] * wa + oo.z * wb, o[jj * 4 + 3] * wa + oo.w * wb);
                }
                if (sub == 0) { out[p.hd] = M; out[p.hd + 1] = l; }
            }
        }
        att_item = -1;
    }

    // prologue of wo (slow): combine the chunk partials into xs.  Thread = (head, two adjacent dims); the
    // (m, l) pair of a slot is one 8-byte load shared by the warp; 8 slots are in flight at once.
    __device__ __forceinline__ void combine_attn() {
        const int ns = att_chunks();
        const int h = tid >> 5, d2 = (tid & 31) * 2;  // H * hd == 2 * kM1Threads (host-checked: H = 16, hd = 64)
        const float *pp = p.partial + (size_t)h * (2 * p.n_chunks_max) * (p.hd + 4);
        float M = -INFINITY, Lsum = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll 1
        for (int s0 = 0; s0 < ns; s0 += 8) {
            float2 ml[8], ov[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int sj = min(s0 + j, ns - 1);  // clamp: the duplicate gets weight 0 below
                const float *sp = pp + sj * (p.hd + 4);
                ml[j] = __ldcg(reinterpret_cast<const float2 *>(sp + p.hd));
                ov[j] = __ldcg(reinterpret_cast<const float2 *>(sp + d2));
                if (s0 + j >= ns) ml[j].x = -INFINITY;
            }
            float Mn = M;
#pragma unroll
            for (int j = 0; j < 8; ++j) Mn = fmaxf(Mn, ml[j].x);
            const float c0 = (M == -INFINITY) ? 0.f : expf(M - Mn);
            Lsum *= c0;
            o0 *= c0;
            o1 *= c0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float w = (ml[j].x == -INFINITY) ? 0.f : expf(ml[j].x - Mn);
                Lsum = fmaf(ml[j].y, w, Lsum);
                o0 = fmaf(ov[j].x, w, o0);
                o1 = fmaf(ov[j].y, w, o1);
            }
            M = Mn;
        }
        *reinterpret_cast<float2 *>(xs + x_index(h * p.hd + d2)) = make_float2(o0 / Lsum, o1 / Lsum);
    }

    // prologue of wo (fast): the whole attention over <= C cached positions, recomputed by every CTA
    __device__ __forceinline__ void fast_attn(const float *kcache, const float *vcache, int cb) {
        const int Hhd = p.H * p.hd, n_rep = p.H / p.KV;
        const float scale = 1.0f / sqrtf((float)p.hd);
        const int npos = cb + 1;
        float *qs = xs + Hhd;                          // behind the output row (xs holds >= 2 * Hhd floats)
        float *kss = kvs;                              // KV * fast_len * hd
        float *vss = kss + p.KV * p.fast_len * p.hd;   // same
        for (int i = tid; i < Hhd / 4; i += kM1Threads)
            reinterpret_cast<float4 *>(qs)[i] = __ldcg(reinterpret_cast<const float4 *>(p.q) + i);
        const int seg = p.hd / 4;
        for (int i = tid; i < p.KV * npos * seg; i += kM1Threads) {
            const int r = i / (npos * seg), rem = i - r * npos * seg;
            const size_t off = (size_t)r * p.fast_len * p.hd + rem * 4;
            *reinterpret_cast<float4 *>(kss + off) = __ldcg(reinterpret_cast<const float4 *>(kcache + off));
            *reinterpret_cast<float4 *>(vss + off) = __ldcg(reinterpret_cast<const float4 *>(vcache + off));
        }
        csync();
        // one warp per head: lane = (position j = lane / 4, 16 of the 64 dims) for the scores, then every
        // lane owns two output dims for P.V (the probabilities are broadcast by shuffles)
        const int j = lane >> 2, sub = lane & 3;
        const bool valid = j < npos;
        for (int h = warp; h < p.H; h += kM1Warps) {
            const int kvh = h / n_rep;
            const float *qp = qs + h * p.hd + sub * 4;
            const float *kr = kss + ((size_t)kvh * p.fast_len + (valid ? j : 0)) * p.hd + sub * 4;
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
            const float *vbase = vss + (size_t)kvh * p.fast_len * p.hd + lane * 2;
            float o0 = 0.f, o1 = 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float pw = __shfl_sync(0xffffffffu, pj, jj * 4);
                if (jj < npos) {
                    const float2 vv = *reinterpret_cast<const float2 *>(vbase + jj * p.hd);
                    o0 = fmaf(pw, vv.x, o0);
                    o1 = fmaf(pw, vv.y, o1);
                }
            }
            *reinterpret_cast<float2 *>(xs + x_index(h * p.hd + lane * 2)) = make_float2(o0 / l, o1 / l);
        }
    }

    // ------------------------------------------------------------ replicated activation vectors
    // Right after a grid barrier all CTAs read the same 4-16 KB vector: 148 requests per 128-byte line
    // queue up at the few L2 slices that hold it.  Producers therefore store every element kM1Rep times
    // (replica r at its own address, i.e. its own slices) and CTA b reads replica b % kM1Rep.  (Measured:
    // -5 % decode time for x / fx / h; replicating q, the attention partials and the fast K/V did not pay
    // for the extra stores of the few CTAs that produce them.)
    __device__ __forceinline__ float *stream_rep(bool slow, int rep) const {
        if (rep == 0) return slow ? p.x : p.fx;
        return p.rep + (slow ? 0 : (kM1Rep - 1) * p.D) + (rep - 1) * p.D;
    }
    __device__ __forceinline__ float *h_rep(int rep) const {
        return rep == 0 ? p.h : p.rep + 2 * (kM1Rep - 1) * p.D + (rep - 1) * p.I;
    }
    __device__ __forceinline__ int my_rep() const { return (int)(blockIdx.x % kM1Rep); }

    // ------------------------------------------------------------ one weight phase
    __device__ __forceinline__ const float *norm_of(const Step &s) const {
        const bool slow = s.pass == 0;
        switch (s.kind) {
            case K_QKV: return layer_of(s).attn_norm;
            case K_W13: return layer_of(s).ffn_norm;
            case K_HEAD: return slow ? p.norm : p.fast_norm;
            default: return nullptr;
        }
    }
    // between barrier arrive and wait: the coming phase's plan and its norm weights
    __device__ __forceinline__ void prep_step(const Step &s) {
        if (s.kind == K_END || s.kind == K_ATT || s.kind == K_SAMPLE) return;
        {
            const int rk = s.kind == K_HEAD ? (s.pass == 0 ? R_HEAD_SLOW : R_HEAD_FAST)
                         : s.kind == K_QKV ? R_QKV : s.kind == K_WO ? R_WO : s.kind == K_W13 ? R_W13 : R_W2;
            plan.r0 = rtab[2 * rk];
            plan.nrows = rtab[2 * rk + 1];
            plan.ksplit = s.kind == K_W2 ? p.I / kM1Slice : 1;
            plan.ntasks = plan.nrows * plan.ksplit * (s.kind == K_W13 ? 2 : 1);
        }
        const float *g = norm_of(s);
        if (g && 2 * tid < p.D) gpre = __ldg(reinterpret_cast<const float2 *>(g) + tid);
    }

    __device__ __forceinline__ void gemv_phase(const Step &s) {
        const bool slow = s.pass == 0;
        const int cb = s.pass - 1;
        const int kind = s.kind;
        // the slow stream of frame 0 comes from the prefill (canonical copy only) when the launch starts at the tail
        const bool prefilled = s.frame == 0 && p.first_is_tail != 0;
        const float *xg = stream_rep(slow, (slow && prefilled) ? 0 : my_rep());
        const size_t kv_stride = (size_t)p.max_batch * p.KV * (slow ? p.max_len : p.fast_len) * p.hd;
        float *kcl = (slow ? p.kc : p.fkc) + s.l * kv_stride, *vcl = (slow ? p.vc : p.fvc) + s.l * kv_stride;
        const int cache_len = slow ? p.max_len : p.fast_len;
        const int D = p.D;
        bool with_norm = false;
        int K = D;
        // ---- prologue: the activation vector -> xs (scaled by the norm weights where a norm applies)
        if (kind == K_QKV && s.l == 0) {
            with_norm = true;
            if (slow) {
                // once per frame: the slow position and its RoPE row
                if (tid < p.hd) {
                    const int half = p.hd / 2;
                    const int pos = __ldcg(p.st.pos);
                    cs_s[tid] = tid < half ? p.cosT[(size_t)pos * half + tid] : p.sinT[(size_t)pos * half + tid - half];
                    if (tid == 0) pos_s[0] = pos;
                }
                // DualARTransformer::embed, dual_ar.rs:532-567, on the previous frame's codes
                const WT *emb = reinterpret_cast<const WT *>(p.emb), *cbe = reinterpret_cast<const WT *>(p.cb_emb);
                const uint32_t *t = p.st.prev;
                const uint32_t tok0 = __ldcg(t);
                const bool msk = p.has_end ? (tok0 <= p.sem_end && tok0 >= p.sem_start) : (tok0 == p.sem_start);
                const float mf = msk ? 1.f : 0.f;
                float ss = 0.f;
                for (int d = tid; d < D; d += kM1Threads) {
                    float acc = to_f32(emb[(size_t)tok0 * D + d]);
                    for (int c = 0; c < p.C; ++c) {
                        const uint32_t code = __ldcg(t + 1 + c);
                        acc = __fadd_rn(acc, __fmul_rn(to_f32(cbe[((size_t)c * p.CS + code) * D + d]), mf));
                    }
                    xres[d] = acc;  // scaled into xs below (gpre is laid out for elements 2 * tid, 2 * tid + 1)
                    ss = fmaf(acc, acc, ss);
                }
                ss = warp_sum(ss);
                if (lane == 0) red[warp] = ss;
            } else {
                // fast stack input: pre-norm slow hidden (Q1) for codebook 0, else fast_embeddings[previous code]
                const WT *fe = reinterpret_cast<const WT *>(p.fast_emb);
                const uint32_t code = cb == 0 ? 0u : __ldcg(p.st.cur + cb);
                float ss = 0.f;
                for (int d = tid; d < D; d += kM1Threads) {
                    const float v = cb == 0 ? __ldcg(stream_rep(true, prefilled ? 0 : my_rep()) + d) : to_f32(fe[(size_t)code * D + d]);
                    xres[d] = v;
                    ss = fmaf(v, v, ss);
                }
                ss = warp_sum(ss);
                if (lane == 0) red[warp] = ss;
            }
            csync();
            if (2 * tid < D) {
                float2 v = reinterpret_cast<float2 *>(xres)[tid];
                v.x = __fmul_rn(v.x, gpre.x);
                v.y = __fmul_rn(v.y, gpre.y);
                *reinterpret_cast<float2 *>(xs + x_index(2 * tid)) = v;
            }
        } else if (kind == K_QKV || kind == K_W13 || kind == K_HEAD) {
            with_norm = true;
            stage_norm(xg, D, true);
        } else if (kind == K_WO) {
            K = p.H * p.hd;
            if (slow) combine_attn();
            else fast_attn(kcl, vcl, cb);
        } else {  // K_W2
            K = p.I;
            stage_plain(h_rep(my_rep()), p.I);
        }
        csync();
        load_xr(K > kM1Slice ? (warp % CT) % (K / kM1Slice) : 0);
        finish_norm(D, with_norm);
        run_tasks1();
        csync();
        // ---- epilogue
        if (kind == K_QKV) {
            // pairs of rows -> rope_i (dual_ar.rs:246-247) -> q buffer / K cache; V rows -> V cache (Tensor::cat, :316-324)
            const int half = p.hd / 2, Hhd = p.H * p.hd, KVhd = p.KV * p.hd;
            for (int pr = tid; pr < plan.nrows / 2; pr += kM1Threads) {
                const int r = plan.r0 + 2 * pr;
                const float v0 = row_val(0, 2 * pr), v1 = row_val(0, 2 * pr + 1);
                const int pos = slow ? pos_s[0] : cb;
                if (r < Hhd + KVhd) {
                    const int pi = (r % p.hd) / 2;
                    const float *cs = slow ? cs_s : csf_s + cb * p.hd;
                    const float c = cs[pi], sn = cs[half + pi];
                    const float o0 = __fsub_rn(__fmul_rn(v0, c), __fmul_rn(v1, sn));
                    const float o1 = __fadd_rn(__fmul_rn(v0, sn), __fmul_rn(v1, c));
                    if (r < Hhd) {
                        p.q[r] = o0;
                        p.q[r + 1] = o1;
                    } else {
                        const int rk = r - Hhd, kvh = rk / p.hd, d = rk % p.hd;
                        float *dst = kcl + ((size_t)kvh * cache_len + pos) * p.hd + d;
                        dst[0] = o0;
                        dst[1] = o1;
                    }
                } else {
                    const int rv = r - Hhd - KVhd, kvh = rv / p.hd, d = rv % p.hd;
                    float *dst = vcl + ((size_t)kvh * cache_len + pos) * p.hd + d;
                    dst[0] = v0;
                    dst[1] = v1;
                }
            }
        } else {
            if (kind == K_HEAD) {
                for (int rl = tid; rl < plan.nrows; rl += kM1Threads) p.logits[plan.r0 + rl] = row_val(0, rl);
            } else {
                // thread = (row, replica): every replica of the result vector gets the value
                for (int i = tid; i < plan.nrows * kM1Rep; i += kM1Threads) {
                    const int rl = i / kM1Rep, rep = i % kM1Rep;
                    const int r = plan.r0 + rl;
                    const float s1 = row_val(0, rl);
                    if (kind == K_W13) {
                        h_rep(rep)[r] = __fmul_rn(silu_f(s1), row_val(1, rl));  // silu(w1 x) * (w3 x), dual_ar.rs:160-165
                    } else {
                        // residual add, dual_ar.rs:436-440 (xres: the stream as staged by the qkv / w13 prologue)
                        stream_rep(slow, rep)[r] = __fadd_rn(xres[r], s1);
                    }
                }
            }
        }
    }

    // ------------------------------------------------------------ samplers (CTA 0 only; scratch on the K/V staging area)
    __device__ __forceinline__ void load_sampler_state() {
        const GenState &st = p.st;
        const int C1 = st.C + 1;
        if (tid == 0) {
            s_active[0] = st.active[0];
            s_eos[0] = st.eos[0];
            s_frame[0] = st.frame[0];
            s_maxf[0] = st.max_frames[0];
        }
        for (int i = tid; i < C1; i += kM1Threads) {
            s_cur[i] = st.cur[i];
            s_prev[i] = st.prev[i];
        }
        const int words = (int)(sizeof(RepPenState) / 4);
        for (int i = tid; i < st.C * words; i += kM1Threads)
            reinterpret_cast<uint32_t *>(s_rep)[i] = reinterpret_cast<const uint32_t *>(st.rep)[i];
        csync();
    }

    // sampler scratch on the K/V staging area: [selection scratch | logits (n floats) | reduction scratch (64 floats)]
    __device__ __forceinline__ void sampler_scratch(int n, unsigned char **scratch, float **vals, float **sred) {
        *scratch = reinterpret_cast<unsigned char *>(kvs);
        *vals = kvs + (sel_scratch_bytes(kM1Threads) + 15) / 16 * 4;
        *sred = *vals + ((n + 3) & ~3);
    }

    __device__ __forceinline__ void sample_slow() {
        const GenState &st = p.st;
        const int n = p.n_slow_logits;
        unsigned char *scratch;
        float *vals, *sred;
        sampler_scratch(n, &scratch, &vals, &sred);
        if (s_active[0]) {
            const int frame = s_frame[0];
            const float u = philox_uniform(st.sp.seed, (uint64_t)frame * (st.C + 1), (uint32_t)p.row0);
            uint32_t tok;
            if (st.legacy_slow) {
                const float eos_l = __ldcg(p.logits), pad_l = __ldcg(p.logits + 1);
                const float mx = fmaxf(pad_l, eos_l);
                const float e_pad = expf(pad_l - mx), e_eos = expf(eos_l - mx);
                tok = (st.fixed_len || u < e_pad / (e_pad + e_eos)) ? st.pad_id : st.im_end_id;
            } else {
                for (int i = tid; i < n; i += kM1Threads) {
                    float v = __ldcg(p.logits + i);
                    if (i == 0 && st.fixed_len) v = -INFINITY;
                    vals[i] = v;
                }
                csync();
                const int idx = block_sample_sel<M1Sync>(vals, scratch, sred, n, st.sp, u);
                tok = (idx == 0) ? st.im_end_id : (p.sem_start + (uint32_t)idx - 1);
            }
            if (tid == 0) {
                const bool eos = tok == st.im_end_id;
                s_cur[0] = tok;
                st.cur[0] = tok;
                s_eos[0] = eos ? 1 : 0;
                st.eos[0] = eos ? 1 : 0;
                if (eos)
                    for (int c = 0; c < st.C; ++c) {
                        s_cur[1 + c] = 0;
                        st.cur[1 + c] = 0;
                    }
                // the next frame runs iff this row goes on (single_batch.rs:193-204): tell every CTA's
                // producer now, 8 fast steps ahead of the frame boundary
                if (!eos && frame + 1 < s_maxf[0]) p.bar[1] = (unsigned)(frame + 2);
            }
            csync();
        }
    }

    __device__ __forceinline__ void sample_fast(int cb) {
        const GenState &st = p.st;
        const int n = p.CS, C = st.C;
        unsigned char *scratch;
        float *vals, *sred;
        sampler_scratch(n, &scratch, &vals, &sred);
        if (!s_active[0]) return;
        const bool eos = s_eos[0] != 0;
        const int frame = s_frame[0];
        if (!eos) {
            RepPenState *rp = s_rep + cb;
            if (frame > 0) {
                if (tid == 0) rep_pen_update(rp, s_prev[1 + cb]);
                csync();
            }
            for (int i = tid; i < n; i += kM1Threads) {
                float v = __ldcg(p.logits + i);
                if (frame > 0 && ((rp->seen[i >> 5] >> (i & 31)) & 1u)) v = __fdiv_rn(v, st.sp.penalty);
                vals[i] = v;
            }
            csync();
            const float u = philox_uniform(st.sp.seed, (uint64_t)frame * (C + 1) + cb + 1, (uint32_t)p.row0);
            const int a = block_sample_sel<M1Sync>(vals, scratch, sred, n, st.sp, u);
            if (tid == 0) {
                s_cur[1 + cb] = (uint32_t)a;
                st.cur[1 + cb] = (uint32_t)a;
            }
        }
        if (cb == C - 1) {
            csync();
            // frame bookkeeping (single_batch.rs:193-204), write-through
            if (tid <= C) {
                const uint32_t v = s_cur[tid];
                st.out[(size_t)frame * (C + 1) + tid] = v;
                s_prev[tid] = v;
                st.prev[tid] = v;
            }
            if (tid == 0) {
                const int nf = frame + 1;
                s_frame[0] = nf;
                st.frame[0] = nf;
                if (frame > 0) {
                    pos_s[0] += 1;
                    st.pos[0] = pos_s[0];
                }
                if (eos || nf >= s_maxf[0]) {
                    s_active[0] = 0;
                    st.active[0] = 0;
                    atomicSub(st.n_active, 1);
                }
            }
            const int words = (int)(sizeof(RepPenState) / 4);
            const uint32_t *src = reinterpret_cast<const uint32_t *>(s_rep);
            uint32_t *dst = reinterpret_cast<uint32_t *>(st.rep);
            for (int i = tid; i < C * words; i += kM1Threads) dst[i] = src[i];
        }
        csync();
    }

    // ------------------------------------------------------------ frame loop (compute warps)
    __device__ __forceinline__ void run() {
        Step cur = first_step();
        for (int i = tid; i < p.C * p.hd; i += kM1Threads) {
            const int row = i / p.hd, d = i - row * p.hd, half = p.hd / 2;
            csf_s[i] = d < half ? p.cosT[(size_t)row * half + d] : p.sinT[(size_t)row * half + d - half];
        }
        if (tid == 0) pos_s[0] = p.st.pos[0];
        const bool sampler = is_sampler();                                  // streams nothing
        const bool samples = p.sampler_cta >= 0 ? sampler : blockIdx.x == 0;  // runs the K_SAMPLE phases
        if (samples) load_sampler_state();
        if (samples && tid == 0 && p.dbg) g_sample_dbg = p.dbg + 104;
        csync();
        if (!sampler) prep_step(cur);
        const bool timed = p.dbg != nullptr && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
        unsigned long long *dbg = p.dbg + (blockIdx.x == 0 ? 0 : 32);
        while (cur.kind != K_END) {
            unsigned long long t0 = 0, t1 = 0, t2 = 0;
            if (timed) t0 = clock64();
            if (cur.kind == K_SAMPLE) {
                if (samples) {
                    if (cur.pass == 0) sample_slow();
                    else sample_fast(cur.pass - 1);
                }
            } else if (!sampler) {
                if (cur.kind == K_ATT) phase_attn_slow(cur.l);
                else gemv_phase(cur);
            }
            const Step nxt = advance(cur);
            if (timed) t1 = clock64();
            grid_arrive1();
            if (!sampler) {
                if (nxt.kind == K_SAMPLE) prep_step(advance(nxt));
                else if (cur.kind != K_SAMPLE) prep_step(nxt);
                if (nxt.kind == K_ATT) att_prefetch(nxt.l);
            }
            if (timed) t2 = clock64();
            grid_wait1();
            if (timed) {
                const unsigned long long t3 = clock64();
                dbg[cur.kind * 4 + 0] += t1 - t0;
                dbg[cur.kind * 4 + 1] += t2 - t1;
                dbg[cur.kind * 4 + 2] += t3 - t2;
                dbg[cur.kind * 4 + 3] += 1;
            }
            if (cur.kind == K_SAMPLE && cur.pass == 0) {
                // CTA 0 published whether frame cur.frame + 1 runs before it arrived at this barrier
                if (tid == 0) *go_frames = (int)ld_relaxed_u32(p.bar + 1);
                csync();
            }
            if (nxt.frame != cur.frame && nxt.kind != K_END && nxt.frame >= *go_frames) break;
            cur = nxt;
        }
        if (tid == 0) *done_flag = 1;
    }
};

template <typename WT>
__global__ void __launch_bounds__(kM1AllThreads, 1) mega1_decode_kernel(const __grid_constant__ MegaParams p) {
    extern __shared__ __align__(128) unsigned char mega1_smem[];
    Mega1<WT> m(p, mega1_smem);
    if (p.nframes <= 0 || __ldcg(p.st.n_active) == 0) return;
    m.init_row_ranges();
    {
        const int nl = p.NL + p.NFL, words = (int)(sizeof(MegaLayer) / 4);
        uint32_t *dst = reinterpret_cast<uint32_t *>(const_cast<MegaLayer *>(m.ltab));
        for (int i = threadIdx.x; i < nl * words; i += blockDim.x) {
            const int l = i / words, w = i - l * words;
            const MegaLayer *src = l < p.NL ? p.slow + l : p.fast + (l - p.NL);
            dst[i] = reinterpret_cast<const uint32_t *>(src)[w];
        }
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.ring_depth; ++i) {
            m1_mbar_init(m.full + i, 1);
            m1_mbar_init(m.empty + i, Mega1<WT>::CT);
        }
        *m.go_frames = 1;
        *m.done_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= kM1Threads) {
        if (threadIdx.x == kM1Threads && !m.is_sampler()) m.producer();
        return;
    }
    m.run();
}

template <typename WT>
static cudaError_t mega1_launch_impl(const MegaParams &mp, int grid, size_t smem, cudaStream_t st) {
    const void *kern = (const void *)mega1_decode_kernel<WT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    void *args[] = {(void *)&mp};
    return cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kM1AllThreads), args, smem, st);
}

}  // namespace fsb
