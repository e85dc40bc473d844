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

// ---------------------------------------------------------------- flag-in-data synchronisation ("LL")
// Every value that crosses CTAs travels as one 8-byte word {fp32 value, 32-bit tag}; the tag names the
// phase that produced it (frame, pass, layer, kind), so a consumer that sees the tag has the value -- no
// release fence, no atomic counter, no separate reload after a barrier.  Per-CTA progress flags (plain
// relaxed stores of the tag, monotonic) are only a hint that keeps the polling traffic small; correctness
// rests on the tags.  8-byte scalar accesses are single-copy atomic.
__device__ __forceinline__ void st_ll(unsigned long long *p, float v, unsigned tag) {
    const unsigned long long w = ((unsigned long long)tag << 32) | __float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_ll_raw(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
constexpr int kLLSpinLimit = 1 << 24;  // ~seconds: a tag that never arrives traps instead of hanging the GPU
__device__ __forceinline__ float ld_ll(const unsigned long long *p, unsigned tag) {
    unsigned long long v = ld_ll_raw(p);
    int spins = 0;
    while ((unsigned)(v >> 32) != tag) {
        if (++spins > kLLSpinLimit) __trap();
        v = ld_ll_raw(p);
    }
    return __uint_as_float((unsigned)v);
}
__device__ __forceinline__ float ll_take(unsigned long long w, const unsigned long long *p, unsigned tag) {
    return ((unsigned)(w >> 32) == tag) ? __uint_as_float((unsigned)w) : ld_ll(p, tag);
}
// The CTA's share of one weight phase: a stream of tasks (1024-element K-slices, row-major inside
// the CTA's contiguous row block), made of <= 3 contiguous global segments.
struct M1Plan {
    const char *seg[3];
    unsigned seg_bytes[3];
    int nseg;
    int r0, nrows, ksplit, nmat, ntasks;
};

template <typename WT, bool LL>
struct Mega1 {
    // shapes the kernel is specialised for (host-checked in mega1_eligible): compile-time constants keep divisions and
    // index arithmetic out of the per-phase code
    static constexpr int kD = kM1Slice, kHd = 64, kHhd = kM1Slice, kH = kHhd / kHd, kKV = 2, kI = 4 * kM1Slice, kC = 8, kCS = 1024;
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
    int pos0;      // LL: slow position at launch

    __device__ Mega1(const MegaParams &pp, unsigned char *smem)
        : p(pp) {
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
        pos0 = 0;
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

    // ---- LL: tags, progress flags
    // tag of the phase (frame, pass, layer, kind); HEAD / SAMPLE count as layer 31 so that tags grow along the schedule
    static __device__ __forceinline__ unsigned tag_of(int frame, int pass, int l, int kind) {
        return 1u + ((((unsigned)frame * 16u + (unsigned)pass) * 32u + (unsigned)l) * 8u + (unsigned)kind);
    }
    // call after the phase's tagged stores: "CTA blockIdx.x has produced phase `tag`" (hint only)
    __device__ __forceinline__ void ll_publish(unsigned tag) {
        csync();
        if (tid < kM1Rep) st_ll(p.ll_fl + (size_t)tid * kM1FlagStride + blockIdx.x, 0.f, tag);
    }
    // wait until CTAs [0, nprod) have published a phase >= tag (tags are monotonic per CTA)
    __device__ __forceinline__ void ll_wait(unsigned tag, int nprod) {
        if (warp == 0) {
            // all flags of a polling round are in flight together: one L2 round trip per round
            const unsigned long long *f = p.ll_fl + (size_t)my_rep() * kM1FlagStride;
            int spins = 0;
            while (true) {
                unsigned lo = 0xffffffffu;
#pragma unroll
                for (int i = 0; i < kM1FlagStride / 32; ++i) {
                    const int c = lane + 32 * i;
                    if (c < nprod) lo = min(lo, (unsigned)(ld_ll_raw(f + c) >> 32));
                }
                if (__all_sync(0xffffffffu, lo >= tag)) break;
                if (++spins > kLLSpinLimit) __trap();
            }
        }
        csync();
    }
    __device__ __forceinline__ unsigned long long *xt_rep(int par, int rep) const { return p.ll_xt + ((size_t)par * kM1Rep + rep) * kD; }
    __device__ __forceinline__ unsigned long long *ht_rep(int rep) const { return p.ll_ht + (size_t)rep * kI; }

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
                if (s.pass < kC) { n.pass = s.pass + 1; n.kind = p.NFL > 0 ? K_QKV : K_HEAD; }
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
            const int rows_total = k == R_QKV ? p.QKV : k == R_W13 ? kI : k == R_HEAD_SLOW ? p.n_slow_logits
                                 : k == R_HEAD_FAST ? kCS : kD;
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
        int rk = R_HEAD_FAST, K = kD, row_a = 0, row_b = 1;
        const void *W0 = p.fast_out, *W1 = nullptr;
        if (s.kind == K_HEAD) {
            if (slow) { rk = R_HEAD_SLOW; W0 = p.out_w; row_a = p.slow_row0; row_b = p.slow_rest_base; }
        } else {
            const MegaLayer &L = layer_of(s);
            switch (s.kind) {
                case K_QKV: rk = R_QKV; W0 = L.wqkv; break;
                case K_WO: rk = R_WO; W0 = L.wo; K = kHhd; break;
                case K_W13: rk = R_W13; W0 = L.w1; W1 = L.w3; break;
                default: rk = R_W2; W0 = L.w2; K = kI; break;
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

    // LL variants: the vector arrives as tagged words (replica my_rep()); same smem results as above
    __device__ __forceinline__ void stage_norm_ll(const unsigned long long *src, unsigned tag) {
        float ss = 0.f;
        if (2 * tid < kD) {
            const unsigned long long a = ld_ll_raw(src + 2 * tid), b = ld_ll_raw(src + 2 * tid + 1);
            float2 v;
            v.x = ((unsigned)(a >> 32) == tag) ? __uint_as_float((unsigned)a) : ld_ll(src + 2 * tid, tag);
            v.y = ((unsigned)(b >> 32) == tag) ? __uint_as_float((unsigned)b) : ld_ll(src + 2 * tid + 1, tag);
            reinterpret_cast<float2 *>(xres)[tid] = v;
            ss = fmaf(v.x, v.x, v.y * v.y);
            v.x = __fmul_rn(v.x, gpre.x);
            v.y = __fmul_rn(v.y, gpre.y);
            *reinterpret_cast<float2 *>(xs + x_index(2 * tid)) = v;
        }
        ss = warp_sum(ss);
        if (lane == 0) red[warp] = ss;
    }
    __device__ __forceinline__ void stage_plain_ll(const unsigned long long *src, int K, unsigned tag) {
        for (int i = tid; i < K / 4; i += kM1Threads) {
            unsigned long long w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = ld_ll_raw(src + 4 * i + j);
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = ((unsigned)(w[j] >> 32) == tag) ? __uint_as_float((unsigned)w[j]) : ld_ll(src + 4 * i + j, tag);
            reinterpret_cast<float4 *>(xs)[(i & ~255) | x4_index(i & 255)] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }

    // ------------------------------------------------------------ split-KV GQA attention (slow blocks)
    // item = kvh * n_chunks_max + chunk; positions [chunk*64, min(len, chunk*64 + 64))
    __device__ __forceinline__ void att_stage(const float *kcache, const float *vcache, int kvh, int j0, int from, int to) {
        const float *kb = kcache + ((size_t)kvh * p.max_len + j0) * kHd;
        const float *vb = vcache + ((size_t)kvh * p.max_len + j0) * kHd;
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
        const int nitems = kKV * nch;
        const size_t slow_kv = p.slow_kv_stride;
        if ((int)blockIdx.x < nitems && !is_sampler()) {
            const int item = blockIdx.x, kvh = item / nch, chunk = item - kvh * nch;
            const int j0 = chunk * kM1AttChunk, j1 = min(len - 1, j0 + kM1AttChunk);  // exclude position len-1
            att_item = item;
            att_n = max(j1 - j0, 0);
            if (att_n > 0) att_stage(p.kc + layer * slow_kv, p.vc + layer * slow_kv, kvh, j0, 0, att_n);
        }
        cp_async_commit();
    }

    __device__ __forceinline__ void phase_attn_slow(int layer, int att_frame) {
        const size_t slow_kv = p.slow_kv_stride;
        const float *kcache = p.kc + layer * slow_kv, *vcache = p.vc + layer * slow_kv;
        const int n_rep = kH / kKV;
        const int hq = warp & 7, hf = warp >> 3;
        const int g = lane >> 2, sub = lane & 3;
        const float scale = 1.0f / sqrtf((float)kHd);
        const int len = pos_s[0] + 1, nch = att_chunks();
        const int nitems = kKV * nch;
        const float *ks = kvs, *vs = kvs + kM1AttChunk * kM1KvStride;
        for (int item = is_sampler() ? nitems : (int)blockIdx.x; item < nitems; item += n_compute()) {
            const int kvh = item / nch, chunk = item - kvh * nch;
            const int j0 = chunk * kM1AttChunk, j1 = min(len, j0 + kM1AttChunk);
            const int have = item == att_item ? att_n : 0;
            const int h = kvh * n_rep + hq;
            float4 qv[4];
            if (LL) {
                // q and the new K/V row come tagged from the QKV phase of this layer; cached rows (older
                // positions, fenced once per frame) through cp.async
                const unsigned qtag = tag_of(att_frame, 0, layer, K_QKV);
                if (hq < n_rep) {
                    const unsigned long long *qp = p.ll_qt + (size_t)h * kHd + sub * 4;
                    unsigned long long qw[16];
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) qw[jj] = ld_ll_raw(qp + (jj >> 2) * 16 + (jj & 3));
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        qv[jj].x = ll_take(qw[jj * 4 + 0], qp + jj * 16 + 0, qtag); qv[jj].y = ll_take(qw[jj * 4 + 1], qp + jj * 16 + 1, qtag);
                        qv[jj].z = ll_take(qw[jj * 4 + 2], qp + jj * 16 + 2, qtag); qv[jj].w = ll_take(qw[jj * 4 + 3], qp + jj * 16 + 3, qtag);
                    }
                }
                csync();  // previous item's smem reads are done
                const int jold = min(j1, len - 1);  // rows below len - 1 are in the cache
                if (j0 + have < jold) att_stage(kcache, vcache, kvh, j0, have, jold - j0);
                cp_async_commit();
                if (j1 == len && tid < 2 * kHd) {  // this chunk holds the new position
                    const int which = tid / kHd, d = tid - which * kHd;
                    const float v = ld_ll(p.ll_nkv + (size_t)which * kKV * kHd + kvh * kHd + d, qtag);
                    kvs[which * kM1AttChunk * kM1KvStride + (len - 1 - j0) * kM1KvStride + d] = v;
                }
                cp_async_wait_all();
                csync();
            } else {
                if (hq < n_rep) {
                    const float *qp = p.q + (size_t)h * kHd + sub * 4;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) qv[jj] = __ldcg(reinterpret_cast<const float4 *>(qp + jj * 16));
                }
                csync();  // previous item's smem reads are done
                if (j0 + have < j1) att_stage(kcache, vcache, kvh, j0, have, j1 - j0);
                cp_async_commit();
                cp_async_wait_all();
                csync();
            }
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
                const size_t slot_off = ((size_t)h * (2 * p.n_chunks_max) + chunk) * (kHd + 4);
                float *out = p.partial + slot_off;
                unsigned long long *outl = p.ll_pt + slot_off;
                const unsigned ptag = tag_of(att_frame, 0, layer, K_ATT);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float4 oo = *reinterpret_cast<const float4 *>(src + jj * 16 + sub * 4);
                    const float4 r4 = make_float4(o[jj * 4 + 0] * wa + oo.x * wb, o[jj * 4 + 1] * wa + oo.y * wb,
                                                  o[jj * 4 + 2] * wa + oo.z * wb, o[jj * 4 + 3] * wa + oo.w * wb);
                    if (LL) {
                        st_ll(outl + jj * 16 + sub * 4 + 0, r4.x, ptag); st_ll(outl + jj * 16 + sub * 4 + 1, r4.y, ptag);
                        st_ll(outl + jj * 16 + sub * 4 + 2, r4.z, ptag); st_ll(outl + jj * 16 + sub * 4 + 3, r4.w, ptag);
                    } else {
                        *reinterpret_cast<float4 *>(out + jj * 16 + sub * 4) = r4;
                    }
                }
                if (sub == 0) {
                    if (LL) { st_ll(outl + kHd, M, ptag); st_ll(outl + kHd + 1, l, ptag); }
                    else { out[kHd] = M; out[kHd + 1] = l; }
                }
            }
        }
        if (LL && (int)blockIdx.x < nitems && !is_sampler()) ll_publish(tag_of(att_frame, 0, layer, K_ATT));
        att_item = -1;
    }

    // prologue of wo (slow): combine the chunk partials into xs.  Thread = (head, two adjacent dims); the
    // (m, l) pair of a slot is one 8-byte load shared by the warp; 8 slots are in flight at once.
    __device__ __forceinline__ void combine_attn(unsigned ll_tag) {
        const int ns = att_chunks();
        const int h = tid >> 5, d2 = (tid & 31) * 2;  // H * hd == 2 * kM1Threads (host-checked: H = 16, hd = 64)
        const float *pp = p.partial + (size_t)h * (2 * p.n_chunks_max) * (kHd + 4);
        constexpr int BS = LL ? 4 : 8;  // slots per batch of loads (LL words take two registers each)
        float M = -INFINITY, Lsum = 0.f, o0 = 0.f, o1 = 0.f;
        const unsigned long long *ppl = p.ll_pt + (size_t)h * (2 * p.n_chunks_max) * (kHd + 4);
#pragma unroll 1
        for (int s0 = 0; s0 < ns; s0 += BS) {
            float2 ml[BS], ov[BS];
            unsigned long long raw[LL ? 4 * BS : 1];
#pragma unroll
            for (int j = 0; j < BS; ++j) {
                const int sj = min(s0 + j, ns - 1);  // clamp: the duplicate gets weight 0 below
                if (LL) {
                    // (raw words first: all 32 loads of the batch are in flight before the first tag check)
                    const unsigned long long *sp = ppl + sj * (kHd + 4);
                    raw[j * 4 + 0] = ld_ll_raw(sp + kHd); raw[j * 4 + 1] = ld_ll_raw(sp + kHd + 1);
                    raw[j * 4 + 2] = ld_ll_raw(sp + d2); raw[j * 4 + 3] = ld_ll_raw(sp + d2 + 1);
                } else {
                    const float *sp = pp + sj * (kHd + 4);
                    ml[j] = __ldcg(reinterpret_cast<const float2 *>(sp + kHd));
                    ov[j] = __ldcg(reinterpret_cast<const float2 *>(sp + d2));
                }
                if (!LL && s0 + j >= ns) ml[j].x = -INFINITY;
            }
            if (LL) {
#pragma unroll
                for (int j = 0; j < BS; ++j) {
                    const int sj = min(s0 + j, ns - 1);
                    const unsigned long long *sp = ppl + sj * (kHd + 4);
                    ml[j].x = ll_take(raw[j * 4 + 0], sp + kHd, ll_tag); ml[j].y = ll_take(raw[j * 4 + 1], sp + kHd + 1, ll_tag);
                    ov[j].x = ll_take(raw[j * 4 + 2], sp + d2, ll_tag); ov[j].y = ll_take(raw[j * 4 + 3], sp + d2 + 1, ll_tag);
                    if (s0 + j >= ns) ml[j].x = -INFINITY;
                }
            }
            float Mn = M;
#pragma unroll
            for (int j = 0; j < BS; ++j) Mn = fmaxf(Mn, ml[j].x);
            const float c0 = (M == -INFINITY) ? 0.f : expf(M - Mn);
            Lsum *= c0;
            o0 *= c0;
            o1 *= c0;
#pragma unroll
            for (int j = 0; j < BS; ++j) {
                const float w = (ml[j].x == -INFINITY) ? 0.f : expf(ml[j].x - Mn);
                Lsum = fmaf(ml[j].y, w, Lsum);
                o0 = fmaf(ov[j].x, w, o0);
                o1 = fmaf(ov[j].y, w, o1);
            }
            M = Mn;
        }
        *reinterpret_cast<float2 *>(xs + x_index(h * kHd + d2)) = make_float2(o0 / Lsum, o1 / Lsum);
    }

    // prologue of wo (fast): the whole attention over <= C cached positions, recomputed by every CTA
    __device__ __forceinline__ void fast_attn(const float *kcache, const float *vcache, int cb, int ll_frame = 0, int ll_layer = 0) {
        const int Hhd = kHhd, n_rep = kH / kKV;
        const float scale = 1.0f / sqrtf((float)kHd);
        const int npos = cb + 1;
        float *qs = xs + Hhd;                          // behind the output row (xs holds >= 2 * Hhd floats)
        float *kss = kvs;                              // KV * fast_len * hd
        float *vss = kss + kKV * kC * kHd;   // same
        if (LL) {
            // q: this phase's QKV; K / V row j: the QKV phase of pass j + 1 of this frame (tagged fast cache)
            const unsigned qtag = tag_of(ll_frame, cb + 1, ll_layer, K_QKV);
            // all of a thread's words are requested before the first tag check (H * hd == 2 * kM1Threads,
            // KV * fast_len * hd <= 2 * kM1Threads: host-checked)
            const unsigned long long *ql = p.ll_qt;
            const size_t lsz = (size_t)kKV * kC * kHd;
            const unsigned long long *kl = p.ll_fkv + (size_t)ll_layer * 2 * lsz, *vl = kl + lsz;
            const int nkv = kKV * npos * kHd;
            unsigned long long w[6];
            size_t off[2];
            unsigned rtag[2];
            w[0] = ld_ll_raw(ql + tid);
            w[1] = ld_ll_raw(ql + tid + kM1Threads);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = tid + u * kM1Threads;
                const int r = i / (npos * kHd), rem = i - r * npos * kHd;
                off[u] = (size_t)r * kC * kHd + rem;
                rtag[u] = tag_of(ll_frame, rem / kHd + 1, ll_layer, K_QKV);
                if (i < nkv) { w[2 + 2 * u] = ld_ll_raw(kl + off[u]); w[3 + 2 * u] = ld_ll_raw(vl + off[u]); }
            }
            qs[tid] = ll_take(w[0], ql + tid, qtag);
            qs[tid + kM1Threads] = ll_take(w[1], ql + tid + kM1Threads, qtag);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (tid + u * kM1Threads < nkv) {
                    kss[off[u]] = ll_take(w[2 + 2 * u], kl + off[u], rtag[u]);
                    vss[off[u]] = ll_take(w[3 + 2 * u], vl + off[u], rtag[u]);
                }
            }
        } else {
            for (int i = tid; i < Hhd / 4; i += kM1Threads)
                reinterpret_cast<float4 *>(qs)[i] = __ldcg(reinterpret_cast<const float4 *>(p.q) + i);
            const int seg = kHd / 4;
            for (int i = tid; i < kKV * npos * seg; i += kM1Threads) {
                const int r = i / (npos * seg), rem = i - r * npos * seg;
                const size_t off = (size_t)r * kC * kHd + rem * 4;
                *reinterpret_cast<float4 *>(kss + off) = __ldcg(reinterpret_cast<const float4 *>(kcache + off));
                *reinterpret_cast<float4 *>(vss + off) = __ldcg(reinterpret_cast<const float4 *>(vcache + off));
            }
        }
        csync();
        // one warp per head: lane = (position j = lane / 4, 16 of the 64 dims) for the scores, then every
        // lane owns two output dims for P.V (the probabilities are broadcast by shuffles)
        const int j = lane >> 2, sub = lane & 3;
        const bool valid = j < npos;
        for (int h = warp; h < kH; h += kM1Warps) {
            const int kvh = h / n_rep;
            const float *qp = qs + h * kHd + sub * 4;
            const float *kr = kss + ((size_t)kvh * kC + (valid ? j : 0)) * kHd + sub * 4;
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
            const float *vbase = vss + (size_t)kvh * kC * kHd + lane * 2;
            float o0 = 0.f, o1 = 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float pw = __shfl_sync(0xffffffffu, pj, jj * 4);
                if (jj < npos) {
                    const float2 vv = *reinterpret_cast<const float2 *>(vbase + jj * kHd);
                    o0 = fmaf(pw, vv.x, o0);
                    o1 = fmaf(pw, vv.y, o1);
                }
            }
            *reinterpret_cast<float2 *>(xs + x_index(h * kHd + lane * 2)) = make_float2(o0 / l, o1 / l);
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
        return p.rep + (slow ? 0 : (kM1Rep - 1) * kD) + (rep - 1) * kD;
    }
    __device__ __forceinline__ float *h_rep(int rep) const {
        return rep == 0 ? p.h : p.rep + 2 * (kM1Rep - 1) * kD + (rep - 1) * kI;
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
            plan.ksplit = s.kind == K_W2 ? kI / kM1Slice : 1;
            plan.ntasks = plan.nrows * plan.ksplit * (s.kind == K_W13 ? 2 : 1);
        }
        const float *g = norm_of(s);
        if (g && 2 * tid < kD) gpre = __ldg(reinterpret_cast<const float2 *>(g) + tid);
    }

    // one specialisation per phase kind: the prologue / epilogue variants become straight-line code
    __device__ __forceinline__ void gemv_phase(const Step &s) {
        switch (s.kind) {
            case K_QKV: gemv_phase_t<K_QKV>(s); break;
            case K_WO: gemv_phase_t<K_WO>(s); break;
            case K_W13: gemv_phase_t<K_W13>(s); break;
            case K_W2: gemv_phase_t<K_W2>(s); break;
            default: gemv_phase_t<K_HEAD>(s); break;
        }
    }
    template <int KIND>
    __device__ __forceinline__ void gemv_phase_t(const Step &s) {
        const bool slow = s.pass == 0;
        const int cb = s.pass - 1;
        constexpr int kind = KIND;
        // the slow stream of frame 0 comes from the prefill (canonical copy only) when the launch starts at the tail
        const bool prefilled = s.frame == 0 && p.first_is_tail != 0;
        const float *xg = stream_rep(slow, (slow && prefilled) ? 0 : my_rep());
        const size_t kv_stride = slow ? p.slow_kv_stride : p.fast_kv_stride;
        float *kcl = (slow ? p.kc : p.fkc) + s.l * kv_stride, *vcl = (slow ? p.vc : p.fvc) + s.l * kv_stride;
        const int cache_len = slow ? p.max_len : kC;
        const int D = kD;
        bool with_norm = false;
        int K = D;
        // ---- prologue: the activation vector -> xs (scaled by the norm weights where a norm applies)
        if (kind == K_QKV && s.l == 0) {
            with_norm = true;
            if (slow) {
                // once per frame: the slow position and its RoPE row
                int *codes_s = reinterpret_cast<int *>(red + 32);
                if (LL) {
                    // previous frame's codes: tagged words from the sampler (the launch starts at the tail, so
                    // frame >= 1 here); the position advances by one per slow step
                    if (tid <= kC) codes_s[tid] = (int)__float_as_uint(ld_ll(p.ll_ct + tid, tag_of(s.frame - 1, tid, 31, K_SAMPLE)));
                    csync();
                }
                if (tid < kHd) {
                    const int half = kHd / 2;
                    const int pos = LL ? pos0 + s.frame - 1 : __ldcg(p.st.pos);
                    cs_s[tid] = tid < half ? p.cosT[(size_t)pos * half + tid] : p.sinT[(size_t)pos * half + tid - half];
                    if (tid == 0) pos_s[0] = pos;
                }
                // DualARTransformer::embed, dual_ar.rs:532-567, on the previous frame's codes
                const WT *emb = reinterpret_cast<const WT *>(p.emb), *cbe = reinterpret_cast<const WT *>(p.cb_emb);
                const uint32_t *t = p.st.prev;
                const uint32_t tok0 = LL ? (uint32_t)codes_s[0] : __ldcg(t);
                const bool msk = p.has_end ? (tok0 <= p.sem_end && tok0 >= p.sem_start) : (tok0 == p.sem_start);
                const float mf = msk ? 1.f : 0.f;
                float ss = 0.f;
                for (int d = tid; d < D; d += kM1Threads) {
                    float acc = to_f32(emb[(size_t)tok0 * D + d]);
                    for (int c = 0; c < kC; ++c) {
                        const uint32_t code = LL ? (uint32_t)codes_s[1 + c] : __ldcg(t + 1 + c);
                        acc = __fadd_rn(acc, __fmul_rn(to_f32(cbe[((size_t)c * kCS + code) * D + d]), mf));
                    }
                    xres[d] = acc;  // scaled into xs below (gpre is laid out for elements 2 * tid, 2 * tid + 1)
                    ss = fmaf(acc, acc, ss);
                }
                ss = warp_sum(ss);
                if (lane == 0) red[warp] = ss;
            } else {
                // fast stack input: pre-norm slow hidden (Q1) for codebook 0, else fast_embeddings[previous code]
                const WT *fe = reinterpret_cast<const WT *>(p.fast_emb);
                uint32_t code = 0u;
                if (LL) {
                    // the code sampled by the previous pass (and, after the slow sample, whether the next frame runs)
                    int *codes_s = reinterpret_cast<int *>(red + 32);
                    if (tid == 0) {
                        if (cb > 0) codes_s[0] = (int)__float_as_uint(ld_ll(p.ll_ct + cb, tag_of(s.frame, cb, 31, K_SAMPLE)));
                        else *go_frames = (int)__float_as_uint(ld_ll(p.ll_ct + kC + 1, tag_of(s.frame, 0, 31, K_SAMPLE)));
                    }
                    csync();
                    code = (uint32_t)codes_s[0];
                } else if (cb > 0) {
                    code = __ldcg(p.st.cur + cb);
                }
                const float *xslow = kvs + 6144;  // LL: raw slow hidden stashed by the slow HEAD prologue
                float ss = 0.f;
                for (int d = tid; d < D; d += kM1Threads) {
                    const float v = cb == 0 ? (LL ? xslow[d] : __ldcg(stream_rep(true, prefilled ? 0 : my_rep()) + d))
                                            : to_f32(fe[(size_t)code * D + d]);
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
            if (LL && !(kind == K_HEAD && slow && prefilled)) {
                // producer of the stream: w2 of the previous layer (qkv), wo of this layer (w13), w2 of the last layer (head)
                const int pl = kind == K_QKV ? s.l - 1 : kind == K_W13 ? s.l : (slow ? p.NL : p.NFL) - 1;
                const int pk = kind == K_W13 ? K_WO : K_W2;
                const unsigned tg = tag_of(s.frame, s.pass, pl, pk);
                ll_wait(tg, (int)gridDim.x);
                stage_norm_ll(xt_rep(pk == K_WO ? 0 : 1, my_rep()), tg);
            } else {
                stage_norm(xg, D, true);
            }
            if (LL && kind == K_HEAD && slow) {
                csync();  // xres holds the raw slow hidden: keep it for codebook 0 of the fast stack (Q1)
                if (2 * tid < D) reinterpret_cast<float2 *>(kvs + 6144)[tid] = reinterpret_cast<const float2 *>(xres)[tid];
            }
        } else if (kind == K_WO) {
            K = kHhd;
            if (slow) {
                const unsigned tg = tag_of(s.frame, 0, s.l, K_ATT);
                if (LL) ll_wait(tg, min(kKV * att_chunks(), (int)gridDim.x));
                combine_attn(tg);
            } else {
                if (LL) ll_wait(tag_of(s.frame, s.pass, s.l, K_QKV), (int)gridDim.x);
                fast_attn(kcl, vcl, cb, s.frame, s.l);
            }
        } else {  // K_W2
            K = kI;
            if (LL) {
                const unsigned tg = tag_of(s.frame, s.pass, s.l, K_W13);
                ll_wait(tg, (int)gridDim.x);
                stage_plain_ll(ht_rep(my_rep()), kI, tg);
            } else {
                stage_plain(h_rep(my_rep()), kI);
            }
        }
        csync();
        load_xr(K > kM1Slice ? (warp % CT) % (K / kM1Slice) : 0);
        finish_norm(D, with_norm);
        run_tasks1();
        csync();
        // ---- epilogue
        const unsigned mytag = tag_of(s.frame, s.pass, kind == K_HEAD ? 31 : s.l, kind);
        if (kind == K_QKV) {
            // pairs of rows -> rope_i (dual_ar.rs:246-247) -> q buffer / K cache; V rows -> V cache (Tensor::cat, :316-324)
            const int half = kHd / 2, Hhd = kHhd, KVhd = kKV * kHd;
            for (int pr = tid; pr < plan.nrows / 2; pr += kM1Threads) {
                const int r = plan.r0 + 2 * pr;
                const float v0 = row_val(0, 2 * pr), v1 = row_val(0, 2 * pr + 1);
                const int pos = slow ? pos_s[0] : cb;
                if (r < Hhd + KVhd) {
                    const int pi = (r % kHd) / 2;
                    const float *cs = slow ? cs_s : csf_s + cb * kHd;
                    const float c = cs[pi], sn = cs[half + pi];
                    const float o0 = __fsub_rn(__fmul_rn(v0, c), __fmul_rn(v1, sn));
                    const float o1 = __fadd_rn(__fmul_rn(v0, sn), __fmul_rn(v1, c));
                    if (r < Hhd) {
                        if (LL) { st_ll(p.ll_qt + r, o0, mytag); st_ll(p.ll_qt + r + 1, o1, mytag); }
                        else { p.q[r] = o0; p.q[r + 1] = o1; }
                    } else {
                        const int rk = r - Hhd, kvh = rk / kHd, d = rk % kHd;
                        if (!LL || slow) {
                            float *dst = kcl + ((size_t)kvh * cache_len + pos) * kHd + d;
                            dst[0] = o0;
                            dst[1] = o1;
                        }
                        if (LL) {
                            unsigned long long *dl = slow ? p.ll_nkv + rk
                                : p.ll_fkv + (size_t)s.l * 2 * KVhd * kC + ((size_t)kvh * kC + pos) * kHd + d;
                            st_ll(dl, o0, mytag);
                            st_ll(dl + 1, o1, mytag);
                        }
                    }
                } else {
                    const int rv = r - Hhd - KVhd, kvh = rv / kHd, d = rv % kHd;
                    if (!LL || slow) {
                        float *dst = vcl + ((size_t)kvh * cache_len + pos) * kHd + d;
                        dst[0] = v0;
                        dst[1] = v1;
                    }
                    if (LL) {
                        unsigned long long *dl = slow ? p.ll_nkv + KVhd + rv
                            : p.ll_fkv + ((size_t)s.l * 2 + 1) * KVhd * kC + ((size_t)kvh * kC + pos) * kHd + d;
                        st_ll(dl, v0, mytag);
                        st_ll(dl + 1, v1, mytag);
                    }
                }
            }
        } else {
            if (kind == K_HEAD) {
                for (int rl = tid; rl < plan.nrows; rl += kM1Threads) {
                    if (LL) st_ll(p.ll_lt + plan.r0 + rl, row_val(0, rl), mytag);
                    else p.logits[plan.r0 + rl] = row_val(0, rl);
                }
            } else {
                // thread = (row, replica): every replica of the result vector gets the value
                for (int i = tid; i < plan.nrows * kM1Rep; i += kM1Threads) {
                    const int rl = i / kM1Rep, rep = i % kM1Rep;
                    const int r = plan.r0 + rl;
                    const float s1 = row_val(0, rl);
                    if (kind == K_W13) {
                        const float hv = __fmul_rn(silu_f(s1), row_val(1, rl));  // silu(w1 x) * (w3 x), dual_ar.rs:160-165
                        if (LL) st_ll(ht_rep(rep) + r, hv, mytag);
                        else h_rep(rep)[r] = hv;
                    } else {
                        // residual add, dual_ar.rs:436-440 (xres: the stream as staged by the qkv / w13 prologue)
                        const float xv = __fadd_rn(xres[r], s1);
                        if (LL) st_ll(xt_rep(kind == K_WO ? 0 : 1, rep) + r, xv, mytag);
                        else stream_rep(slow, rep)[r] = xv;
                    }
                }
            }
        }
        if (LL) ll_publish(mytag);
    }

    // ------------------------------------------------------------ samplers (CTA 0 only; scratch on the K/V staging area)
    __device__ __forceinline__ void load_sampler_state() {
        const GenState &st = p.st;
        const int C1 = kC + 1;
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
        for (int i = tid; i < kC * words; i += kM1Threads)
            reinterpret_cast<uint32_t *>(s_rep)[i] = reinterpret_cast<const uint32_t *>(st.rep)[i];
        csync();
    }

    // sampler scratch on the K/V staging area: [selection scratch | logits (n floats) | reduction scratch (64 floats)]
    __device__ __forceinline__ void sampler_scratch(int n, unsigned char **scratch, float **vals, float **sred) {
        *scratch = reinterpret_cast<unsigned char *>(kvs);
        *vals = kvs + (sel_scratch_bytes(kM1Threads) + 15) / 16 * 4;
        *sred = *vals + ((n + 3) & ~3);
    }

    __device__ __forceinline__ void sample_slow(int kframe) {
        const GenState &st = p.st;
        const int n = p.n_slow_logits;
        unsigned char *scratch;
        float *vals, *sred;
        sampler_scratch(n, &scratch, &vals, &sred);
        if (s_active[0]) {
            const int frame = s_frame[0];
            const float u = philox_uniform(st.sp.seed, (uint64_t)frame * (kC + 1), (uint32_t)p.row0);
            uint32_t tok;
            const unsigned htag = tag_of(kframe, 0, 31, K_HEAD);
            if (LL) ll_wait(htag, (int)gridDim.x);
            if (st.legacy_slow) {
                const float eos_l = LL ? ld_ll(p.ll_lt, htag) : __ldcg(p.logits);
                const float pad_l = LL ? ld_ll(p.ll_lt + 1, htag) : __ldcg(p.logits + 1);
                const float mx = fmaxf(pad_l, eos_l);
                const float e_pad = expf(pad_l - mx), e_eos = expf(eos_l - mx);
                tok = (st.fixed_len || u < e_pad / (e_pad + e_eos)) ? st.pad_id : st.im_end_id;
            } else {
                for (int i = tid; i < n; i += kM1Threads) {
                    float v = LL ? ld_ll(p.ll_lt + i, htag) : __ldcg(p.logits + i);
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
                    for (int c = 0; c < kC; ++c) {
                        s_cur[1 + c] = 0;
                        st.cur[1 + c] = 0;
                    }
                // the next frame runs iff this row goes on (single_batch.rs:193-204): tell every CTA's
                // producer now, 8 fast steps ahead of the frame boundary
                const bool cont = !eos && frame + 1 < s_maxf[0];
                if (LL) {
                    // the slow token, and whether frame kframe + 1 runs, as tagged words for every CTA
                    const unsigned stag = tag_of(kframe, 0, 31, K_SAMPLE);
                    st_ll(p.ll_ct, __uint_as_float(tok), stag);
                    st_ll(p.ll_ct + kC + 1, __uint_as_float((unsigned)(cont ? kframe + 2 : kframe + 1)), stag);
                } else if (cont) {
                    p.bar[1] = (unsigned)(frame + 2);
                }
            }
            csync();
        }
    }

    __device__ __forceinline__ void sample_fast(int cb, int kframe) {
        const GenState &st = p.st;
        constexpr int n = kCS, C = kC;
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
            const unsigned htag = tag_of(kframe, cb + 1, 31, K_HEAD);
            if (LL) ll_wait(htag, (int)gridDim.x);
            for (int i = tid; i < n; i += kM1Threads) {
                float v = LL ? ld_ll(p.ll_lt + i, htag) : __ldcg(p.logits + i);
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
        if (LL && tid == 0)  // published for every pass (0 after <|im_end|>): the next prologue waits for it
            st_ll(p.ll_ct + 1 + cb, __uint_as_float(s_cur[1 + cb]), tag_of(kframe, cb + 1, 31, K_SAMPLE));
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
                    if (LL) st.pos[0] = pos0 + kframe;  // the launch starts at the tail: frame k >= 1 ran the slow step at pos0 + k - 1
                    else { pos_s[0] += 1; st.pos[0] = pos_s[0]; }
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
        for (int i = tid; i < kC * kHd; i += kM1Threads) {
            const int row = i / kHd, d = i - row * kHd, half = kHd / 2;
            csf_s[i] = d < half ? p.cosT[(size_t)row * half + d] : p.sinT[(size_t)row * half + d - half];
        }
        if (tid == 0) pos_s[0] = p.st.pos[0];
        pos0 = p.st.pos[0];
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
                // LL: once per pass every CTA makes its plain stores (slow K/V cache rows) visible before its next
                // tagged store, and orders its later cache reads after the tagged data it has seen (hidden behind
                // the sampler for the 147 CTAs that wait for the code anyway)
                if (LL) __threadfence();
                if (samples) {
                    if (cur.pass == 0) sample_slow(cur.frame);
                    else sample_fast(cur.pass - 1, cur.frame);
                }
            } else if (!sampler) {
                if (cur.kind == K_ATT) phase_attn_slow(cur.l, cur.frame);
                else gemv_phase(cur);
            }
            const Step nxt = advance(cur);
            if (timed) t1 = clock64();
            if (!LL) grid_arrive1();
            if (!sampler) {
                if (nxt.kind == K_SAMPLE) prep_step(advance(nxt));
                else if (cur.kind != K_SAMPLE) prep_step(nxt);
                if (nxt.kind == K_ATT) att_prefetch(nxt.l);
            }
            if (timed) t2 = clock64();
            if (!LL) grid_wait1();
            if (timed) {
                const unsigned long long t3 = clock64();
                dbg[cur.kind * 4 + 0] += t1 - t0;
                dbg[cur.kind * 4 + 1] += t2 - t1;
                dbg[cur.kind * 4 + 2] += t3 - t2;
                dbg[cur.kind * 4 + 3] += 1;
            }
            if (!LL && cur.kind == K_SAMPLE && cur.pass == 0) {
                // CTA 0 published whether frame cur.frame + 1 runs before it arrived at this barrier
                // (LL: the flag travels with the slow token and is picked up by the first fast prologue)
                if (tid == 0) *go_frames = (int)ld_relaxed_u32(p.bar + 1);
                csync();
            }
            if (nxt.frame != cur.frame && nxt.kind != K_END && nxt.frame >= *go_frames) break;
            cur = nxt;
        }
        if (tid == 0) *done_flag = 1;
    }
};

template <typename WT, bool LL>
__global__ void __launch_bounds__(kM1AllThreads, 1) mega1_decode_kernel(const __grid_constant__ MegaParams p) {
    extern __shared__ __align__(128) unsigned char mega1_smem[];
    Mega1<WT, LL> m(p, mega1_smem);
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
            m1_mbar_init(m.empty + i, Mega1<WT, LL>::CT);
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
    const void *kern = mp.ll ? (const void *)mega1_decode_kernel<WT, true> : (const void *)mega1_decode_kernel<WT, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    void *args[] = {(void *)&mp};
    return cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kM1AllThreads), args, smem, st);
}

}  // namespace fsb
