// Persistent decode megakernel for WIDE batches (9..32 utterances per GPU, bf16 weights): the cfg3 / cfg5 path.
//
// Same contract as fsb_lm_mega1.cuh (one cooperative launch runs every frame of the dual-AR loop of
// generate/static_batch.rs:117-274 with "independent rows" semantics, grid barriers between dependent phases),
// but the weight stream feeds the 5th-generation tensor cores instead of FMA dot products:
//
//   * every projection is Y^T[R, B] = W[R, K] . X^T[K, B]: 128 weight rows fill the MMA M dimension, the batch
//     rows are the MMA N dimension (NPAD = 16 or 32), so the weights cross HBM ONCE per step for all rows;
//   * warp 8 (one lane) is the TMA producer: it walks the static phase schedule ahead of everybody and moves
//     [128 rows x 64 k] bf16 tiles (cp.async.bulk.tensor.2d, 128-byte swizzle) into a ring of 16 KB stages, so HBM
//     keeps streaming through grid barriers, attention and the samplers;
//   * warp 9 (one lane) issues tcgen05.mma.cta_group::1.kind::f16 with the accumulator in TMEM (two buffers:
//     the epilogue of unit i overlaps the MMAs of unit i + 1);
//   * warps 0-7 are workers: they stage the activation slice (fp32 from L2 -> x * g -> three bf16 terms hi + mid +
//     lo = 24 mantissa bits, written straight into the UMMA K-major 128B-swizzle layout), drain TMEM, and run the
//     fused epilogues, attention and the samplers;
//   * no split-K on QKV / WO / heads: one CTA owns a 128-row weight tile and streams its full K; the activation
//     arrives as four 256-column slices that alternate between two shared-memory operand buffers (the MMAs of
//     slice s run while slice s + 1 is staged), and the epilogue runs straight out of TMEM (no partial buffers,
//     no cross-CTA reduction, no second synchronisation per phase);
//   * the feed-forward is ONE phase: CTA t owns 64 columns of the intermediate dimension, computes w1/w3 for them
//     (one M = 128 tile: 64 w1 rows | 64 w3 rows), applies SwiGLU in shared memory, and multiplies the 64-wide h
//     slice with its K-slice of w2 right away (8 more accumulators in TMEM); a short reduce phase sums the 64
//     partial outputs in a fixed order (deterministic, no atomics on data), adds the residual and leaves sum(x^2);
//   * RMSNorm is folded: the staged operand is x * g, sum(x^2) is produced by whoever writes x (residual fixups,
//     embedding gathers) and 1 / sqrt(mean + eps) scales the reduced sums; RoPE + KV append, SwiGLU, residual adds
//     and the constrained head are the fixups of their phases;
//   * GQA attention: an item is (row, kv head, range of positions); the 8 query heads of a KV group share the
//     staged K/V chunk; ranges of one (row, kv head) are merged by the last CTA to arrive (no spinning);
//   * row b is sampled by CTA b (rep-pen window, top-k -> top-p -> multinomial, Philox), which then gathers the
//     next step's input row (fast_embeddings / embed) and its sum of squares.
//
// Reference call sites replaced: dual_ar.rs:160-165,239-384,429-440,574-673; generate/static_batch.rs:117-274
// (per-row semantics of single_batch.rs:76-214); sampling/mod.rs; sampling/rep_pen.rs.
#pragma once
#include <cuda.h>

#include "fsb_lm_mega1.cuh"

namespace fsb {

typedef SyncNamed<kMBWorkers, 1> MBSync;

__device__ __forceinline__ bool mb_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(m1_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mb_wait(uint64_t *bar, uint32_t parity) {
    if (mb_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mb_try_wait(bar, parity))
        if (clock64() - t0 > kMBSpinLimit) __trap();
}
__device__ __forceinline__ unsigned mb_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mb_red_release(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void mb_tma_2d(void *smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            m1_smem_u32(smem)),
        "l"(map), "r"(m1_smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// K-major operand tile [rows][64 bf16], 128-byte swizzle: 8-row x 128 B atoms, 1024 B between atoms
__device__ __forceinline__ uint64_t mb_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mb_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void mb_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(m1_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mb_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void mb_tmem_ld8(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// v -> hi + mid + lo (each the bf16 rounding of what is left): 24 mantissa bits in three bf16 terms
__device__ __forceinline__ void mb_split3(float v, unsigned short &hi, unsigned short &mid, unsigned short &lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const __nv_bfloat16 l = __float2bfloat16_rn(r1 - __bfloat162float(m));
    hi = __bfloat16_as_ushort(h);
    mid = __bfloat16_as_ushort(m);
    lo = __bfloat16_as_ushort(l);
}

template <int NPAD>
struct MegaB {
    static constexpr int kD = 1024, kHd = 64, kH = 16, kKV = 2, kRep = 8, kI = 4096, kC = 8, kCS = 1024, kQKV = 1280;
    static constexpr int kKs = 256;                   // columns of one activation slice
    // activation operand of one 64-wide k-block: [3 * NPAD rows][64 k] -- rows [0, NPAD) hold the hi terms of the batch
    // rows, [NPAD, 2 NPAD) the mid terms, [2 NPAD, 3 NPAD) the lo terms, so ONE MMA of N = 3 * NPAD multiplies a weight
    // tile with all three terms; the three column groups of the accumulator are added when it is drained
    static constexpr int kXKb = 3 * NPAD * 128;       // bytes per k-block
    static constexpr int kXBuf = 4 * kXKb;            // one slice buffer (4 k-blocks); two of them alternate
    static constexpr int kAccCols = 3 * NPAD;         // accumulator columns of one term-stacked tile
    // second GEMM of the fused FFN: 8 accumulators (one per 128 output columns).  NPAD 16: term-stacked (8 x 48 columns);
    // NPAD 32: the three terms accumulate into the same 32 columns (8 x 96 would not fit the 512 TMEM columns)
    static constexpr bool kStack2 = NPAD == 16;
    static constexpr int kAcc2Cols = kStack2 ? kAccCols : NPAD;
    static constexpr int kTmemCols = 512;
    static constexpr int kXsBytes = NPAD == 16 ? kMBXsBytes16 : kMBXsBytes32;
    static_assert(2 * kXBuf <= kXsBytes, "two activation slice buffers do not fit");
        static_assert(kAccCols + 8 * kAcc2Cols <= kTmemCols, "TMEM columns");

    enum { K_QKV = 0, K_ATT = 1, K_WO = 2, K_FFN = 3, K_FRED = 4, K_HEAD = 5, K_SAMPLE = 6, K_END = 7 };
    struct Step { int frame, pass, l, kind; };  // pass 0 = slow stack, pass c + 1 = fast step of codebook c

    const MegaParams &p;
    const MegaBExtra &e;
    // shared memory
    unsigned char *ring, *xs;
    uint64_t *full, *empty, *cmd_ready, *xr, *xf, *acc_full, *hr, *a2f;
    uint32_t *tmem_slot;
    volatile int *cmd, *go_frames, *done_flag;
    int *pos_s, *act_s, *ibase_s, *nsplit_s, *prange_s, *bcast_s;
    float *invd, *red;
    const float **normtab;
    int *s_active, *s_eos, *s_frame, *s_maxf;
    uint32_t *s_cur, *s_prev;
    RepPenState *s_rep;
    int tid, lane, warp;
    unsigned int target;     // grid barrier
    unsigned int pacc, pa2f; // parities of acc_full / a2f (workers)

    __device__ MegaB(const MegaParams &pp, const MegaBExtra &ee, unsigned char *smem) : p(pp), e(ee) {
        ring = smem;
        xs = smem + (size_t)ee.nstages * kMBStage;
        unsigned char *q = xs + kXsBytes;
        full = reinterpret_cast<uint64_t *>(q); q += 8 * kMBMaxStages;
        empty = reinterpret_cast<uint64_t *>(q); q += 8 * kMBMaxStages;
        cmd_ready = reinterpret_cast<uint64_t *>(q); q += 8;
        xr = reinterpret_cast<uint64_t *>(q); q += 16;
        xf = reinterpret_cast<uint64_t *>(q); q += 16;
        acc_full = reinterpret_cast<uint64_t *>(q); q += 8;
        hr = reinterpret_cast<uint64_t *>(q); q += 8;
        a2f = reinterpret_cast<uint64_t *>(q); q += 8 * (kD / 128);  // one per accumulator tile of the second FFN GEMM
        tmem_slot = reinterpret_cast<uint32_t *>(q); q += 8;
        cmd = reinterpret_cast<volatile int *>(q); q += 4;
        go_frames = reinterpret_cast<volatile int *>(q); q += 4;
        done_flag = reinterpret_cast<volatile int *>(q); q += 8;
        pos_s = reinterpret_cast<int *>(q); q += 4 * 32;
        act_s = reinterpret_cast<int *>(q); q += 4 * 32;
        ibase_s = reinterpret_cast<int *>(q); q += 4 * 36;
        nsplit_s = reinterpret_cast<int *>(q); q += 4 * 32;
        prange_s = reinterpret_cast<int *>(q); q += 4 * 32;
        bcast_s = reinterpret_cast<int *>(q); q += 4 * 8;
        invd = reinterpret_cast<float *>(q); q += 4 * 32;
        red = reinterpret_cast<float *>(q); q += 4 * 128;
        normtab = reinterpret_cast<const float **>(q); q += 8 * (2 * (pp.NL + pp.NFL) + 2);
        s_active = reinterpret_cast<int *>(q); q += 4;
        s_eos = reinterpret_cast<int *>(q); q += 4;
        s_frame = reinterpret_cast<int *>(q); q += 4;
        s_maxf = reinterpret_cast<int *>(q); q += 4;
        s_cur = reinterpret_cast<uint32_t *>(q); q += 4 * 12;
        s_prev = reinterpret_cast<uint32_t *>(q); q += 4 * 12;
        s_rep = reinterpret_cast<RepPenState *>(q); q += sizeof(RepPenState) * 8;
        tid = threadIdx.x;
        lane = tid & 31;
        warp = tid >> 5;
        target = 0;
        pacc = 0;
        pa2f = 0;
    }

    static __device__ __forceinline__ void wsync() { MBSync::sync(); }

    // ------------------------------------------------------------ schedule
    __device__ __forceinline__ Step first_step() const {
        Step s;
        s.frame = 0; s.pass = 0; s.l = 0;
        // first_is_tail: the launch starts at the slow head of frame 0 (the hidden rows come from the prefill).  Otherwise
        // it RESUMES rows whose previous frame is complete: every frame of the launch runs the slow stack first
        s.kind = p.first_is_tail ? K_HEAD : K_QKV;
        return s;
    }
    __device__ __forceinline__ Step advance(const Step &s) const {
        Step n = s;
        const bool slow = s.pass == 0;
        switch (s.kind) {
            case K_QKV: n.kind = K_ATT; break;
            case K_ATT: n.kind = K_WO; break;
            case K_WO: n.kind = K_FFN; break;
            case K_FFN: n.kind = K_FRED; break;
            case K_FRED:
                if (s.l + 1 < (slow ? p.NL : p.NFL)) { n.l = s.l + 1; n.kind = K_QKV; }
                else n.kind = K_HEAD;
                break;
            case K_HEAD: n.kind = K_SAMPLE; break;
            default:  // K_SAMPLE
                n.l = 0;
                if (s.pass < kC) { n.pass = s.pass + 1; n.kind = K_QKV; }
                else {
                    n.pass = 0;
                    n.frame = s.frame + 1;
                    n.kind = n.frame < p.nframes ? K_QKV : K_END;
                }
        }
        return n;
    }
    __device__ __forceinline__ Step next_weight_step(Step s) const {
        do { s = advance(s); } while (s.kind == K_ATT || s.kind == K_SAMPLE || s.kind == K_FRED);
        return s;
    }
    // The 128-row weight tile this CTA owns in a projection phase (-1: none).  No split-K: the owner streams the
    // tile's full K.  Disjoint CTA ranges per phase kind, so a CTA's share of the weight stream stays even over a layer:
    // FFN column blocks on CTAs [0, 64), QKV tiles on [64, 74), WO tiles on [74, 82), head tiles from 82.
    __device__ __forceinline__ int unit_of(const Step &s) const {
        const int G = (int)gridDim.x;
        int T, off;
        switch (s.kind) {
            case K_QKV: T = kQKV / 128; off = 64; break;
            case K_WO: T = kD / 128; off = 74; break;
            case K_FFN: T = kI / 64; off = 0; break;
            case K_HEAD: T = s.pass == 0 ? e.head_tiles + e.head_extra : kCS / 128; off = 82; break;
            default: return -1;
        }
        int v = (int)blockIdx.x - (off % G);
        if (v < 0) v += G;
        return v < T ? v : -1;
    }
    __device__ __forceinline__ int map_of(const Step &s) const {
        const int lb = 5 * (s.pass == 0 ? s.l : p.NL + s.l);
        switch (s.kind) {
            case K_QKV: return lb;
            case K_WO: return lb + 1;
            case K_FFN: return lb + 2;  // w1; w3 = +1, w2 = +2
            default: return 5 * (p.NL + p.NFL) + (s.pass == 0 ? 0 : 1);
        }
    }
    // first weight row of tile t (TMA coordinate)
    __device__ __forceinline__ int tile_row(const Step &s, int t) const {
        if (s.kind == K_HEAD && s.pass == 0) return t < e.head_tiles ? p.slow_rest_base - 1 + 128 * t : p.slow_row0;
        return 128 * t;
    }

    // ------------------------------------------------------------ TMA producer (warp 8, lane 0)
    static __device__ __forceinline__ void tma_prefetch_l2(const CUtensorMap *map, int c0, int c1) {
        asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
    }
    __device__ __forceinline__ void prefetch_unit(const Step &s) const {
        const int t = unit_of(s);
        const CUtensorMap *m0 = e.maps + map_of(s);
        if (s.kind == K_FFN) {
            for (int kb = 0; kb < kD / 64; ++kb) {
                tma_prefetch_l2(m0, kb * 64, 64 * t);
                tma_prefetch_l2(m0 + 1, kb * 64, 64 * t);
            }
            for (int r = 0; r < kD / 128; ++r) tma_prefetch_l2(m0 + 2, 64 * t, 128 * r);
        } else {
            const int row = tile_row(s, t);
            for (int kb = 0; kb < kD / 64; ++kb) tma_prefetch_l2(m0, kb * 64, row);
        }
    }
    __device__ __forceinline__ void producer() {
        Step s = first_step();
        const int depth = e.nstages;
        unsigned int slot = 0, epar = 1;  // parity of the `empty` phase to wait for (first lap: passes immediately)
        auto next_slot = [&]() {
            mb_wait(empty + slot, epar);  // a fresh mbarrier reports the phase before its first one as complete
            m1_mbar_expect_tx(full + slot, kMBStage);
            return ring + (size_t)slot * kMBStage;
        };
        auto advance_slot = [&]() {
            if (++slot == (unsigned)depth) { slot = 0; epar ^= 1; }
        };
        while (s.kind != K_END) {
            // frames beyond the last confirmed one are not streamed (no bulk copy may be in flight at exit)
            if (s.frame >= *go_frames) {
                const long long t0 = clock64();
                while (s.frame >= *go_frames) {
                    if (*done_flag) return;
                    __nanosleep(64);
                    if (clock64() - t0 > kMBSpinLimit) __trap();
                }
            }
            const int t = unit_of(s);
            if (t >= 0) {
                // the unit AFTER this one goes to L2 now (TMA prefetch, no shared memory involved): the ring only holds
                // half a tile, so the second half of every tile is fetched while the phase runs -- from L2, not from HBM
                {
                    Step q = next_weight_step(s);
                    for (int guard = 0; guard < 12 && q.kind != K_END && unit_of(q) < 0; ++guard) q = next_weight_step(q);
                    if (q.kind != K_END && unit_of(q) >= 0) prefetch_unit(q);
                }
                const CUtensorMap *m0 = e.maps + map_of(s);
                if (s.kind == K_FFN) {
                    // GEMM 1: rows [64 t, +64) of w1 and of w3, K = 1024 -> 16 stages of (64 + 64 rows) x 64 k
                    for (int kb = 0; kb < kD / 64; ++kb) {
                        unsigned char *dst = next_slot();
                        mb_tma_2d(dst, m0, full + slot, kb * 64, 64 * t);
                        mb_tma_2d(dst + kMBStage / 2, m0 + 1, full + slot, kb * 64, 64 * t);
                        advance_slot();
                    }
                    // GEMM 2: columns [64 t, +64) of w2, all 1024 rows -> 8 stages of 128 rows x 64 k
                    for (int r = 0; r < kD / 128; ++r) {
                        unsigned char *dst = next_slot();
                        mb_tma_2d(dst, m0 + 2, full + slot, 64 * t, 128 * r);
                        advance_slot();
                    }
                } else {
                    const int row = tile_row(s, t);
                    for (int kb = 0; kb < kD / 64; ++kb) {
                        unsigned char *dst = next_slot();
                        mb_tma_2d(dst, m0, full + slot, kb * 64, row);
                        advance_slot();
                    }
                }
            }
            s = next_weight_step(s);
        }
    }

    // ------------------------------------------------------------ MMA issuer (warp 9, converged; one elected lane issues)
    static __device__ __forceinline__ bool elect_one() {
        uint32_t pred;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "elect.sync _|p, 0xffffffff;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(pred));
        return pred != 0;
    }
    __device__ __forceinline__ void mma_loop() {
        const uint32_t tmem_base = *tmem_slot;
        const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kAccCols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kAcc2Cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int depth = e.nstages;
        unsigned int slot = 0, spar = 0, pcmd = 0, pxr0 = 0, pxr1 = 0, phr = 0;  // ring slot + barrier parities
        // descriptors differ only in the 14-bit start-address field (bytes >> 4): base + constant offsets
        const uint64_t desc_hi = ((uint64_t)((1024 >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        const uint64_t b_base = desc_hi | (uint64_t)((m1_smem_u32(xs) >> 4) & 0x3FFF);
        const uint64_t a_base = desc_hi | (uint64_t)((m1_smem_u32(ring) >> 4) & 0x3FFF);
        for (;;) {
            // (no spin limit here: a CTA that owns no weight tile idles on this barrier for the whole launch)
            while (!mb_try_wait(cmd_ready, pcmd)) {}
            pcmd ^= 1;
            const int c = *cmd;
            if (c < 0) break;
            // GEMM 1 (every projection): four 256-column activation slices alternate between the two operand buffers
#pragma unroll 1
            for (int s = 0; s < 4; ++s) {
                const int i = s & 1;
                if (i == 0) { mb_wait(xr, pxr0); pxr0 ^= 1; }
                else { mb_wait(xr + 1, pxr1); pxr1 ^= 1; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
                    mb_wait(full + slot, spar);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint64_t ad = a_base + (uint64_t)(slot * (kMBStage >> 4));
                        const uint64_t bd = b_base + (uint64_t)((i * 4 + kb) * (kXKb >> 4));
#pragma unroll
                        for (int k = 0; k < 4; ++k) mb_umma(tmem_base, ad + 2 * k, bd + 2 * k, idesc1, (s | kb | k) != 0 ? 1u : 0u);
                        mb_commit(empty + slot);  // frees the ring stage once these MMAs have read it
                        if (kb == 3) mb_commit(xf + i);  // ... and the slice buffer
                        if (kb == 3 && s == 3) mb_commit(acc_full);
                    }
                    __syncwarp();
                    if (++slot == (unsigned)depth) { slot = 0; spar ^= 1; }
                }
            }
            if (c == 2) {
                // GEMM 2 of the fused FFN: y[:, 128 r .. +128) += w2[128 r .. , 64 t .. +64) . h^T, h staged by the workers
                mb_wait(hr, phr);
                phr ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int r = 0; r < kD / 128; ++r) {
                    mb_wait(full + slot, spar);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint64_t ad = a_base + (uint64_t)(slot * (kMBStage >> 4));
                        const uint32_t acc = tmem_base + kAccCols + r * kAcc2Cols;
                        if (kStack2) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) mb_umma(acc, ad + 2 * k, b_base + 2 * k, idesc2, k != 0 ? 1u : 0u);
                        } else {
#pragma unroll
                            for (int term = 0; term < 3; ++term)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    mb_umma(acc, ad + 2 * k, b_base + (uint64_t)(term * (NPAD * 128 >> 4)) + 2 * k, idesc2,
                                            (term | k) != 0 ? 1u : 0u);
                        }
                        mb_commit(empty + slot);
                        mb_commit(a2f + r);  // tile r can be drained while the MMAs of the later tiles run
                    }
                    __syncwarp();
                    if (++slot == (unsigned)depth) { slot = 0; spar ^= 1; }
                }
            }
        }
    }

    // ------------------------------------------------------------ grid barrier (workers)
    __device__ __forceinline__ void grid_arrive() {
        wsync();
        if (tid == 0) {
            target += gridDim.x;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.bar) : "memory");
        }
    }
    __device__ __forceinline__ void grid_wait() {
        if (tid == 0) {
            if ((int)(ld_relaxed_u32(p.bar) - target) < 0) {
                const long long t0 = clock64();
                while ((int)(ld_relaxed_u32(p.bar) - target) < 0)
                    if (clock64() - t0 > kMBSpinLimit) __trap();
            }
            asm volatile("fence.acquire.gpu;" ::: "memory");
        }
        wsync();
    }

    // ------------------------------------------------------------ activation operand
    // A 256-column slice of rows [0, nb) of a row-major fp32 activation -> three bf16 terms in the UMMA K-major layout:
    // k-block kb = [3 NPAD rows][64 k] with 128-byte rows (hi | mid | lo row groups), 16-byte chunk c of row r stored
    // at chunk c ^ (r & 7) (the 128-byte swizzle TMA would produce).  Rows >= nb are zero.  The global loads of slice
    // s + 1 are in flight while slice s is converted, and the MMAs of slice s run while slice s + 1 is staged.
    // byte offset of element (row b, column col) of term `term` in an operand image
    static __device__ __forceinline__ uint32_t xop_off(int term, int b, int col) {
        return (uint32_t)((col >> 6) * kXKb + (term * NPAD + b) * 128 + ((((col >> 3) & 7) ^ (b & 7)) << 4) + (col & 7) * 2);
    }
    static __device__ __forceinline__ void store_xop1(unsigned char *xop, int b, int col, float v) {
        unsigned short h, m, l;
        mb_split3(v, h, m, l);
        *reinterpret_cast<unsigned short *>(xop + xop_off(0, b, col)) = h;
        *reinterpret_cast<unsigned short *>(xop + xop_off(1, b, col)) = m;
        *reinterpret_cast<unsigned short *>(xop + xop_off(2, b, col)) = l;
    }
    // four consecutive columns (col % 4 == 0): one 8-byte store per term
    static __device__ __forceinline__ void store_xop4(unsigned char *xop, int b, int col, float4 v) {
        unsigned short h[4], m[4], l[4];
        mb_split3(v.x, h[0], m[0], l[0]);
        mb_split3(v.y, h[1], m[1], l[1]);
        mb_split3(v.z, h[2], m[2], l[2]);
        mb_split3(v.w, h[3], m[3], l[3]);
        auto pk = [](const unsigned short (&q)[4]) {
            return make_uint2((uint32_t)q[0] | ((uint32_t)q[1] << 16), (uint32_t)q[2] | ((uint32_t)q[3] << 16));
        };
        *reinterpret_cast<uint2 *>(xop + xop_off(0, b, col)) = pk(h);
        *reinterpret_cast<uint2 *>(xop + xop_off(1, b, col)) = pk(m);
        *reinterpret_cast<uint2 *>(xop + xop_off(2, b, col)) = pk(l);
    }
    // the norm weight vector the CONSUMER of the activation produced by step s applies (null: none)
    __device__ __forceinline__ const float *consumer_norm(const Step &s) const {
        const bool slow = s.pass == 0;
        const int li = slow ? s.l : p.NL + s.l;
        if (s.kind == K_WO) return normtab[2 * li + 1];  // -> ffn_norm of the same block
        // K_FRED: -> attention_norm of the next block, or the final norm in front of the head
        if (s.l + 1 < (slow ? p.NL : p.NFL)) return normtab[2 * (li + 1)];
        return normtab[2 * (p.NL + p.NFL) + (slow ? 0 : 1)];
    }
    // consumer side: the four 256-column slices of an operand image alternate between the two operand buffers; the
    // MMAs of slice s run while slice s + 1 streams in (L2 -> shared memory, no thread touches the data)
    __device__ __forceinline__ void stage_bulk(const unsigned char *xop) {
        if (tid == 0) {
            // the region was last touched through the generic proxy (K/V staging, sampler scratch, h tile) and the image
            // was written with ordinary stores by other CTAs: order both before the async-proxy copies
            asm volatile("fence.proxy.async;" ::: "memory");
#pragma unroll 1
            for (int sl = 0; sl < 4; ++sl) {
                const int i = sl & 1;
                // buffer i was read by the MMAs of slice sl - 2: its first completion of this phase (the barrier
                // completes exactly twice per phase, so that one always has parity 0)
                if (sl >= 2) mb_wait(xf + i, 0);
                m1_mbar_expect_tx(xr + i, kXBuf);
                m1_bulk_g2s(xs + i * kXBuf, xop + (size_t)sl * kXBuf, kXBuf, xr + i);
            }
        }
    }
    // 1 / sqrt(mean(x^2) + eps) of every row (candle_nn::RmsNorm), from the partial sums the producer of x left
    __device__ __forceinline__ void compute_invd(const float *ssq) {
        if (tid < p.nb) {
            const float4 *q4 = reinterpret_cast<const float4 *>(ssq + (size_t)tid * kMBSsq);
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < kMBSsq / 4; ++i) {
                const float4 q = __ldcg(q4 + i);
                t += (q.x + q.y) + (q.z + q.w);
            }
            invd[tid] = 1.0f / sqrtf(t / (float)kD + p.eps);
        }
    }
    // this thread's accumulator row (TMEM lane quadrant * 32 + lane) x its half of the batch columns, terms added
    __device__ __forceinline__ void drain_stacked(uint32_t col_base, float (&r)[NPAD / 2]) const {
        const uint32_t tmem_base = *tmem_slot;
        const int qd = warp & 3, cc0 = (warp >> 2) * (NPAD / 2);
#pragma unroll
        for (int term = 0; term < 3; ++term) {
            uint32_t q[16];
            const uint32_t taddr = tmem_base + col_base + term * NPAD + ((uint32_t)(qd * 32) << 16) + (uint32_t)cc0;
            if (NPAD == 32) mb_tmem_ld16(taddr, q);
            else mb_tmem_ld8(taddr, q);
#pragma unroll
            for (int j = 0; j < NPAD / 2; ++j) r[j] = term == 0 ? __uint_as_float(q[j]) : r[j] + __uint_as_float(q[j]);
        }
    }

    // ------------------------------------------------------------ projection with a direct epilogue (QKV, WO, heads)
    __device__ __forceinline__ void fullk_phase(const Step &s) {
        const int t = unit_of(s);
        if (t < 0) return;
        const bool slow = s.pass == 0;
        const int cb = s.pass - 1;
        const int kind = s.kind;
        float *stream = slow ? p.x : p.fx;
        float *ssq = slow ? e.ssq_x : e.ssq_fx;
        if (tid == 0) {
            *cmd = 1;
            m1_mbar_arrive(cmd_ready);
        }
        const bool tm = p.dbg != nullptr && tid == 0 && (blockIdx.x == 64 || blockIdx.x == 74 || blockIdx.x == 82);
        unsigned long long *td = p.dbg + 192 + kind * 8;
        long long c0 = tm ? clock64() : 0, c1 = 0;
        const int li = slow ? s.l : p.NL + s.l;
        const float *gw = kind == K_QKV ? normtab[2 * li] : kind == K_HEAD ? normtab[2 * (p.NL + p.NFL) + (slow ? 0 : 1)] : nullptr;
        stage_bulk(kind == K_WO ? e.xop_att : (slow ? e.xop_x : e.xop_fx));
        if (gw) compute_invd(ssq);
        const float gnext = kind == K_WO ? __ldg(consumer_norm(s) + 128 * t + (warp & 3) * 32 + lane) : 0.f;
        if (tm) { c1 = clock64(); td[0] += c1 - c0; c0 = c1; }
        const int qd = warp & 3, cc0 = (warp >> 2) * (NPAD / 2);
        const int row = qd * 32 + lane;
        // operands of the epilogue that do not depend on the accumulator are requested BEFORE waiting for it: the residual
        // rows (WO) / the RoPE table entries (QKV) arrive while the MMAs run instead of costing an L2 round trip afterwards
        float xo[NPAD / 2], sn[NPAD / 2];  // WO: xo = residual;  QKV: xo = cos, sn = sin
        if (kind == K_WO) {
#pragma unroll
            for (int j = 0; j < NPAD / 2; ++j)
                xo[j] = cc0 + j < p.nb ? __ldcg(stream + (size_t)(cc0 + j) * kD + 128 * t + row) : 0.f;
        } else if (kind == K_QKV) {
            const int rr = 128 * t + row;
            const int pi = (rr & 63) >> 1;
            const bool roped = rr < kD + kKV * kHd;
#pragma unroll
            for (int j = 0; j < NPAD / 2; ++j) {
                const int b = min(cc0 + j, p.nb - 1);
                const int pos = slow ? (act_s[b] ? pos_s[b] : 0) : cb;  // a finished row may sit at max_len
                xo[j] = roped ? __ldg(p.cosT + (size_t)pos * 32 + pi) : 1.f;
                sn[j] = roped ? __ldg(p.sinT + (size_t)pos * 32 + pi) : 0.f;
            }
        }
        wsync();  // invd visible
        mb_wait(acc_full, pacc);
        pacc ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tm) { c1 = clock64(); td[1] += c1 - c0; c0 = c1; }
        float r[NPAD / 2];
        drain_stacked(0, r);
        if (kind == K_WO) {
            // residual add (dual_ar.rs:436-440) + this warp's share of sum(x^2) for the next RMSNorm.  All batch columns
            // are computed unconditionally (padding columns hold zeros) so that the NPAD / 2 butterfly chains are
            // independent instruction streams; only the stores are predicated.  (With the work inside `if (b < nb)`
            // the columns ran one after the other: 3.5-4.6 us per phase at 32 rows.)
            float nv[NPAD / 2], sq[NPAD / 2];
#pragma unroll
            for (int j = 0; j < NPAD / 2; ++j) {
                nv[j] = __fadd_rn(xo[j], r[j]);
                sq[j] = nv[j] * nv[j];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {  // same tree as warp_sum
#pragma unroll
                for (int j = 0; j < NPAD / 2; ++j) sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], o);
            }
            unsigned char *xop = slow ? e.xop_x : e.xop_fx;
            float mine = 0.f;  // lane j keeps column j's sum: one store instruction for all columns
#pragma unroll
            for (int j = 0; j < NPAD / 2; ++j) {
                const int b = cc0 + j;
                if (lane == j) mine = sq[j];
                if (b < p.nb) {
                    stream[(size_t)b * kD + 128 * t + row] = nv[j];
                    store_xop1(xop, b, 128 * t + row, __fmul_rn(nv[j], gnext));
                }
            }
            if (lane < NPAD / 2 && cc0 + lane < p.nb) ssq[(size_t)(cc0 + lane) * kMBSsq + t * 4 + qd] = mine;
        } else if (kind == K_QKV) {
            // rope_i on row pairs (dual_ar.rs:246-247; adjacent rows = adjacent lanes) -> q buffer / K cache; V rows ->
            // V cache (Tensor::cat, :316-324).  Values for all columns first (independent chains), predicated stores after.
            const size_t kv_stride = slow ? p.slow_kv_stride : p.fast_kv_stride;
            float *kcl = (slow ? p.kc : p.fkc) + s.l * kv_stride, *vcl = (slow ? p.vc : p.fvc) + s.l * kv_stride;
            const int cache_len = slow ? p.max_len : kC;
            const int rr = 128 * t + row;
            const bool roped = rr < kD + kKV * kHd;
            const float(&cs)[NPAD / 2] = xo;
            float ov[NPAD / 2];
#pragma unroll
            for (int j = 0; j < NPAD / 2; ++j) {
                const float v = r[j] * invd[min(cc0 + j, p.nb - 1)];
                const float w = __shfl_xor_sync(0xffffffffu, v, 1);
                const float o = (lane & 1) ? __fadd_rn(__fmul_rn(w, sn[j]), __fmul_rn(v, cs[j]))
                                           : __fsub_rn(__fmul_rn(v, cs[j]), __fmul_rn(w, sn[j]));
                ov[j] = roped ? o : v;
            }
            if (rr < kD) {
#pragma unroll
                for (int j = 0; j < NPAD / 2; ++j)
                    if (cc0 + j < p.nb) p.q[(size_t)(cc0 + j) * kD + rr] = ov[j];
            } else {
                const int rk = roped ? rr - kD : rr - kD - kKV * kHd, kvh = rk >> 6, d = rk & 63;
                float *cl = roped ? kcl : vcl;
#pragma unroll
                for (int j = 0; j < NPAD / 2; ++j) {
                    const int b = cc0 + j;
                    if (b < p.nb && act_s[b]) {
                        const int pos = slow ? pos_s[b] : cb;
                        cl[(((size_t)b * kKV + kvh) * cache_len + pos) * kHd + d] = ov[j];
                    }
                }
            }
        } else {  // K_HEAD: logits of the constrained slow head (generate/utils.rs:6-33) or of fast_output
            int rr = 128 * t + row;
            bool ok = rr < (slow ? p.n_slow_logits : kCS);
            if (slow && e.head_extra) {
                if (t == e.head_tiles) { rr = 0; ok = row == 0; }  // the extra tile starts at the <|im_end|> row
                else if (rr == 0) ok = false;
            }
#pragma unroll
            for (int j = 0; j < NPAD / 2; ++j) {
                const int b = cc0 + j;
                if (b < p.nb && ok) p.logits[(size_t)b * p.ldl + rr] = r[j] * invd[b];
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (tm) { td[2] += clock64() - c0; td[5] += 1; }
    }

    // ------------------------------------------------------------ fused feed-forward (dual_ar.rs:160-165)
    // CTA t owns 64 columns of the intermediate dimension: h[:, 64 t .. +64) = silu(w1 x) * (w3 x) stays in shared
    // memory and is multiplied with w2[:, 64 t .. +64) right away; the 64 partial outputs are summed by fred_phase.
    __device__ __forceinline__ void ffn_phase(const Step &s) {
        const int t = unit_of(s);
        if (t < 0) return;
        const bool slow = s.pass == 0;
        float *stream = slow ? p.x : p.fx;
        if (tid == 0) {
            *cmd = 2;
            m1_mbar_arrive(cmd_ready);
        }
        const bool tm = p.dbg != nullptr && tid == 0 && blockIdx.x == 0;
        unsigned long long *td = p.dbg + 192 + K_FFN * 8;
        long long c0 = tm ? clock64() : 0, c1 = 0;
        const int li = slow ? s.l : p.NL + s.l;
        stage_bulk(slow ? e.xop_x : e.xop_fx);
        compute_invd(slow ? e.ssq_x : e.ssq_fx);
        wsync();  // invd visible
        (void)li;
        (void)stream;
        if (tm) { c1 = clock64(); td[0] += c1 - c0; c0 = c1; }
        mb_wait(acc_full, pacc);
        pacc ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tm) { c1 = clock64(); td[1] += c1 - c0; c0 = c1; }
        const int qd = warp & 3, cc0 = (warp >> 2) * (NPAD / 2);
        {
            // accumulator lanes [0, 64) = w1 rows, [64, 128) = w3 rows of the block: pair them through shared memory
            float r[NPAD / 2];
            drain_stacked(0, r);
            // all eight warps take part in SwiGLU: a w1 warp (qd < 2) keeps the first half of its batch columns and hands
            // the second half to the w3 warp of the same rows, which hands over its first half: [2][NPAD][64] floats
            float *exch = reinterpret_cast<float *>(xs + kXBuf);  // both slice buffers are idle now
            constexpr int kHalf = NPAD / 4;
            const int i = (qd & 1) * 32 + lane;  // column of the block == k index of the second GEMM
            const int jg = qd < 2 ? kHalf : 0;   // the columns this thread gives away
            float *mine = exch + (qd < 2 ? 0 : NPAD * 64), *theirs = exch + (qd < 2 ? NPAD * 64 : 0);
#pragma unroll
            for (int j = 0; j < kHalf; ++j) mine[(cc0 + jg + j) * 64 + i] = qd < 2 ? r[kHalf + j] : r[j];
            wsync();
            {
                const int jk = kHalf - jg;  // the columns this thread keeps
#pragma unroll
                for (int j = 0; j < kHalf; ++j) {
                    const int b = cc0 + jk + j;
                    const float other = theirs[b * 64 + i], own = qd < 2 ? r[j] : r[kHalf + j];
                    const float sc = invd[b < p.nb ? b : 0];
                    const float g1 = (qd < 2 ? own : other) * sc, g3 = (qd < 2 ? other : own) * sc;
                    const float hv = b < p.nb ? __fmul_rn(silu_f(g1), g3) : 0.f;
                    unsigned short hh, hm, hl;
                    mb_split3(hv, hh, hm, hl);
                    // k-block layout: row (term * NPAD + b), 16-byte chunk (i / 8) ^ (row & 7), element i % 8
                    const uint32_t off = (uint32_t)(b * 128 + (((i >> 3) ^ (b & 7)) << 4) + (i & 7) * 2);
                    *reinterpret_cast<unsigned short *>(xs + off) = hh;
                    *reinterpret_cast<unsigned short *>(xs + NPAD * 128 + off) = hm;
                    *reinterpret_cast<unsigned short *>(xs + 2 * NPAD * 128 + off) = hl;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            wsync();
            if (tid == 0) m1_mbar_arrive(hr);
        }
        if (tm) { c1 = clock64(); td[2] += c1 - c0; c0 = c1; }
        mb_wait(a2f, pa2f);
        if (tm) { c1 = clock64(); td[3] += c1 - c0; c0 = c1; }
        // partial y[t][b][128 r + row] (summed over the 64 blocks by fred_phase)
        // (Measured and NOT kept: staging the tiles in shared memory and handing each 512-byte batch row to the bulk-copy
        // engine -- 11 us instead of 2.8; the WO epilogue transposed through shared memory for 16-byte stores -- no change.)
        float *yp = e.ws + (size_t)t * NPAD * kD + qd * 32 + lane;
#pragma unroll 1
        for (int r8 = 0; r8 < kD / 128; ++r8) {
            // the store-bound drain of tile r8 (16 KB per tile and CTA) overlaps the MMAs of the tiles behind it
            mb_wait(a2f + r8, pa2f);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float r[NPAD / 2];
            if (kStack2) {
                drain_stacked(kAccCols + r8 * kAcc2Cols, r);
            } else {
                uint32_t q[16];
                const uint32_t taddr = *tmem_slot + kAccCols + r8 * kAcc2Cols + ((uint32_t)(qd * 32) << 16) + (uint32_t)cc0;
                mb_tmem_ld16(taddr, q);
#pragma unroll
                for (int j = 0; j < NPAD / 2; ++j) r[j] = __uint_as_float(q[j]);
            }
#pragma unroll
            for (int j = 0; j < NPAD / 2; ++j)
                if (cc0 + j < p.nb) yp[(size_t)(cc0 + j) * kD + 128 * r8] = r[j];
        }
        pa2f ^= 1;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (tm) { td[4] += clock64() - c0; td[5] += 1; }
    }

    // x += sum over the 64 blocks of the FFN partials (fixed order), and the sums of squares the next RMSNorm needs.
    // CTA c < 4 nb: batch row c / 4, columns [(c % 4) * 256, +256); warp = 32 columns; 4 lanes per column quad, each
    // summing 16 blocks, combined by a fixed shuffle tree.
    __device__ __forceinline__ void fred_phase(const Step &s) {
        const int c = blockIdx.x;
        const bool slow = s.pass == 0;
        float *stream = slow ? p.x : p.fx;
        float *ssq = slow ? e.ssq_x : e.ssq_fx;
        if constexpr (NPAD == 16) {
            // at most 16 rows: EIGHT CTAs per row (128 columns each, 8 lanes per column quad, 8 blocks per lane) so that
            // 128 instead of 64 SMs pull the 4 MB of partials
            if (c >= 8 * p.nb) return;
            const int b = c >> 3, sub = lane & 7;
            const int col = (c & 7) * 128 + warp * 16 + (lane >> 3) * 4;
            const float *yp = e.ws + ((size_t)(sub * 8) * NPAD + b) * kD + col;
            float4 *xp = reinterpret_cast<float4 *>(stream + (size_t)b * kD + col);
            float4 xo = make_float4(0.f, 0.f, 0.f, 0.f), g4 = xo;
            if (sub == 0) {
                xo = __ldcg(xp);
                g4 = __ldg(reinterpret_cast<const float4 *>(consumer_norm(s) + col));
            }
            float4 a = __ldcg(reinterpret_cast<const float4 *>(yp));
#pragma unroll
            for (int j = 1; j < 8; ++j) {
                const float4 q = __ldcg(reinterpret_cast<const float4 *>(yp + (size_t)j * NPAD * kD));
                a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
            }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
                a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
                a.z += __shfl_xor_sync(0xffffffffu, a.z, o);
                a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
            }
            float sq = 0.f;
            if (sub == 0) {
                const float4 nv = make_float4(__fadd_rn(xo.x, a.x), __fadd_rn(xo.y, a.y), __fadd_rn(xo.z, a.z), __fadd_rn(xo.w, a.w));
                *xp = nv;
                sq = nv.x * nv.x + nv.y * nv.y + nv.z * nv.z + nv.w * nv.w;
                store_xop4(slow ? e.xop_x : e.xop_fx, b, col,
                           make_float4(__fmul_rn(nv.x, g4.x), __fmul_rn(nv.y, g4.y), __fmul_rn(nv.z, g4.z), __fmul_rn(nv.w, g4.w)));
            }
            sq = warp_sum(sq);
            // 8 CTAs x 8 warps share the row's kMBSsq = 32 slots: warp pairs are combined through shared memory
            if (lane == 0) red[64 + warp] = sq;
            wsync();
            if (tid < 4) ssq[(size_t)b * kMBSsq + (c & 7) * 4 + tid] = red[64 + 2 * tid] + red[64 + 2 * tid + 1];
            return;
        }
        if (c >= 4 * p.nb) return;
        const int b = c >> 2, sub = lane & 3;
        const int col = (c & 3) * 256 + warp * 32 + (lane >> 2) * 4;
        const float *yp = e.ws + ((size_t)(sub * 16) * NPAD + b) * kD + col;
        // (the residual row and the consumer's norm weight are requested together with the partials)
        float4 *xp = reinterpret_cast<float4 *>(stream + (size_t)b * kD + col);
        float4 xo = make_float4(0.f, 0.f, 0.f, 0.f), g4 = xo;
        if (sub == 0) {
            xo = __ldcg(xp);
            g4 = __ldg(reinterpret_cast<const float4 *>(consumer_norm(s) + col));
        }
        float4 a = __ldcg(reinterpret_cast<const float4 *>(yp));
#pragma unroll
        for (int j = 1; j < 16; ++j) {
            const float4 q = __ldcg(reinterpret_cast<const float4 *>(yp + (size_t)j * NPAD * kD));
            a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
        }
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
            a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
            a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
            a.z += __shfl_xor_sync(0xffffffffu, a.z, o);
            a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
        }
        float sq = 0.f;
        if (sub == 0) {
            const float4 nv = make_float4(__fadd_rn(xo.x, a.x), __fadd_rn(xo.y, a.y), __fadd_rn(xo.z, a.z), __fadd_rn(xo.w, a.w));
            *xp = nv;
            sq = nv.x * nv.x + nv.y * nv.y + nv.z * nv.z + nv.w * nv.w;
            store_xop4(slow ? e.xop_x : e.xop_fx, b, col,
                       make_float4(__fmul_rn(nv.x, g4.x), __fmul_rn(nv.y, g4.y), __fmul_rn(nv.z, g4.z), __fmul_rn(nv.w, g4.w)));
        }
        sq = warp_sum(sq);
        if (lane == 0) ssq[(size_t)b * kMBSsq + (c & 3) * 8 + warp] = sq;
    }

    // ------------------------------------------------------------ attention
    // positions [j0, j1) of (row b, kv head kvh) for the 8 query heads of the group (warp = head); returns this
    // lane's (m, l, o[16]) already merged over the 8 position groups of the warp
    __device__ __forceinline__ void att_range(const float *kcache, const float *vcache, int cache_len, int b, int kvh,
                                              int j0, int j1, float &m, float &l, float (&o)[16]) {
        const int g = lane >> 2, sub = lane & 3;
        const int h = kvh * kRep + warp;
        // K/V rows stream through a ring of kNB buffers of kCh positions (K rows then V rows, padded row stride):
        // kNB - 1 chunks are in flight while one is consumed -- the cache stream is latency-bound (HBM round trip per
        // chunk), so what matters is the number of bytes in flight per SM
        constexpr int kCh = 32, kNB = 4;
        constexpr int kBuf = 2 * kCh * kMBKvStride;
        static_assert(kNB * kBuf * 4 <= kXsBytes, "K/V ring does not fit");
        float *kvb = reinterpret_cast<float *>(xs);
        float4 qv[4];
        const float *qp = p.q + (size_t)b * kD + (size_t)h * kHd + sub * 4;
        // 1 / sqrt(head_dim) is a power of two: scaling q instead of every K row (dual_ar.rs:258-260) is bit-identical
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            qv[jj] = __ldcg(reinterpret_cast<const float4 *>(qp + jj * 16));
            qv[jj].x *= 0.125f; qv[jj].y *= 0.125f; qv[jj].z *= 0.125f; qv[jj].w *= 0.125f;
        }
        m = -INFINITY;
        l = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = 0.f;
        const float *kb = kcache + ((size_t)b * kKV + kvh) * cache_len * kHd;
        const float *vb = vcache + ((size_t)b * kKV + kvh) * cache_len * kHd;
        const int nchunks = (j1 - j0 + kCh - 1) / kCh;
        auto stage = [&](int c) {  // chunk c -> buffer c % kNB; always commits a group (possibly empty)
            if (c < nchunks) {
                const int c0 = j0 + c * kCh, n = min(kCh, j1 - c0);
                float *ks = kvb + (c % kNB) * kBuf, *vs = ks + kCh * kMBKvStride;
                for (int i = tid; i < n * 16; i += kMBWorkers) {
                    const int j = i >> 4, sg = i & 15;
                    cp_async16(ks + j * kMBKvStride + sg * 4, kb + (size_t)(c0 + j) * kHd + sg * 4);
                    cp_async16(vs + j * kMBKvStride + sg * 4, vb + (size_t)(c0 + j) * kHd + sg * 4);
                }
            }
            cp_async_commit();
        };
        wsync();  // the ring is free (previous item / previous user of the region)
#pragma unroll
        for (int c = 0; c < kNB - 1; ++c) stage(c);
        for (int c = 0; c < nchunks; ++c) {
            // ONE barrier per chunk: it publishes chunk c (every thread waited for its own copies) and it proves that every
            // warp is done with chunk c - 1, whose buffer takes chunk c + kNB - 1 right behind it
            asm volatile("cp.async.wait_group %0;" ::"n"(kNB - 2) : "memory");  // chunk c has landed
            wsync();
            stage(c + kNB - 1);
            const int n = min(kCh, j1 - (j0 + c * kCh));
            const float *ks = kvb + (c % kNB) * kBuf, *vs = ks + kCh * kMBKvStride;
            // The whole chunk in ONE step: lane group g takes positions g, g + 8, g + 16, g + 24.  A warp issues in order and
            // only two warps share a scheduler, so the loop runs at the pace of its dependent chain (LDS -> 4 FMA -> add
            // tree -> 2 shuffles -> max -> ex2 -> FMA): four independent chains per step instead of two bought 9 % (17.3 ->
            // 15.7 us per 576-position range).  Measured and NOT kept: -30 % instructions (ex2.approx, rescale only when the
            // maximum moves) alone changed nothing; an L2 bulk prefetch of the range one phase ahead changed nothing; four
            // heads per warp (4x less LDS traffic) was slower (19.3 us), with TMA bulk staging slower still (22.8 us).
            // Head PAIRS per warp on half of a chunk's positions (2x less LDS traffic, same four chains per lane): 15.3 -> 14.7 us
            // per range, but the cross-warp merge it needs costs the 32 short fast-layer phases of a frame 1.2 us each: net loss.
            {
                bool vld[4];
                int row[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    vld[u] = g + 8 * u < n;
                    row[u] = vld[u] ? g + 8 * u : 0;
                }
                float sc[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float *kr = ks + row[u] * kMBKvStride + sub * 4;
                    float d[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float4 x = *reinterpret_cast<const float4 *>(kr + jj * 16);
                        d[jj] = fmaf(qv[jj].w, x.w, fmaf(qv[jj].z, x.z, fmaf(qv[jj].y, x.y, qv[jj].x * x.x)));
                    }
                    sc[u] = (d[0] + d[1]) + (d[2] + d[3]);
                }
                float4 vv[4][4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        vv[u][jj] = *reinterpret_cast<const float4 *>(vs + row[u] * kMBKvStride + sub * 4 + jj * 16);
#pragma unroll
                for (int u = 0; u < 4; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 1);
#pragma unroll
                for (int u = 0; u < 4; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 2);
                if (vld[0]) {  // (an empty lane group skips the update altogether)
                    // the running maximum rarely moves after the first chunks: the output is rescaled only when it does,
                    // and the exponentials run on ex2.approx (2^-22 relative; the weights are renormalised by l anyway)
                    float mx = sc[0];
#pragma unroll
                    for (int u = 1; u < 4; ++u) mx = vld[u] ? fmaxf(mx, sc[u]) : mx;
                    if (mx > m) {
                        const float corr = __expf(m - mx);  // exp(-inf) == 0 on the first chunk
                        l *= corr;
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] *= corr;
                        m = mx;
                    }
                    float pr[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) pr[u] = vld[u] ? __expf(sc[u] - m) : 0.f;
                    l += (pr[0] + pr[1]) + (pr[2] + pr[3]);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        o[jj * 4 + 0] += fmaf(pr[3], vv[3][jj].x, fmaf(pr[2], vv[2][jj].x, fmaf(pr[1], vv[1][jj].x, pr[0] * vv[0][jj].x)));
                        o[jj * 4 + 1] += fmaf(pr[3], vv[3][jj].y, fmaf(pr[2], vv[2][jj].y, fmaf(pr[1], vv[1][jj].y, pr[0] * vv[0][jj].y)));
                        o[jj * 4 + 2] += fmaf(pr[3], vv[3][jj].z, fmaf(pr[2], vv[2][jj].z, fmaf(pr[1], vv[1][jj].z, pr[0] * vv[0][jj].z)));
                        o[jj * 4 + 3] += fmaf(pr[3], vv[3][jj].w, fmaf(pr[2], vv[2][jj].w, fmaf(pr[1], vv[1][jj].w, pr[0] * vv[0][jj].w)));
                    }
                }
            }
        }
        cp_async_wait_all();
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
    }

    // per-frame table of the slow attention items: row b is cut into nsplit_s[b] ranges of prange_s[b] positions
    __device__ __forceinline__ void frame_prep() {
        if (tid < 32) {
            const int b = tid;
            pos_s[b] = b < p.nb ? __ldcg(p.st.pos + b) : 0;
            act_s[b] = b < p.nb ? __ldcg(p.st.active + b) : 0;
        }
        wsync();
        if (tid == 0) {
            // split every live (row, kv head) into ns position ranges, ns proportional to the row's length, such that
            // the items never outnumber the CTAs (an item is sequential; a second round would double the phase)
            int total = 0, live = 0;
            for (int b = 0; b < p.nb; ++b)
                if (act_s[b]) { total += pos_s[b] + 1; ++live; }
            const int G = (int)gridDim.x;
            const int slots = max(G / kKV, live);  // ranges to hand out over all rows (per kv head)
            int base = 0;
            for (int b = 0; b < p.nb; ++b) {
                const int len = pos_s[b] + 1;
                int ns = (int)(((long long)len * slots) / max(total, 1));  // floor: sum over live rows <= slots
                ns = max(1, min(min(ns, kMBMaxSplit), (len + kMBChunk - 1) / kMBChunk));
                int pr = ((len + ns - 1) / ns + kMBChunk - 1) / kMBChunk * kMBChunk;
                ns = (len + pr - 1) / pr;
                nsplit_s[b] = ns;
                prange_s[b] = pr;
                ibase_s[b] = base;
                base += act_s[b] ? kKV * ns : 0;
            }
            ibase_s[p.nb] = base;
        }
        wsync();
    }

    __device__ __forceinline__ void attn_phase(const Step &s) {
        const bool slow = s.pass == 0;
        const size_t kv_stride = slow ? p.slow_kv_stride : p.fast_kv_stride;
        const float *kcl = (slow ? p.kc : p.fkc) + s.l * kv_stride, *vcl = (slow ? p.vc : p.fvc) + s.l * kv_stride;
        const int cache_len = slow ? p.max_len : kC;
        const int nitems = slow ? ibase_s[p.nb] : p.nb * kKV;
        const int g = lane >> 2, sub = lane & 3;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            int b, kvh, sp, ns, j0, j1;
            if (slow) {
                b = 0;
                while (item >= ibase_s[b + 1]) ++b;
                const int r = item - ibase_s[b];
                ns = nsplit_s[b];
                kvh = r / ns;
                sp = r - kvh * ns;
                j0 = sp * prange_s[b];
                j1 = min(pos_s[b] + 1, j0 + prange_s[b]);
            } else {
                b = item / kKV;
                kvh = item - b * kKV;
                sp = 0;
                ns = 1;
                j0 = 0;
                j1 = s.pass;  // cb + 1 cached positions
            }
            float m, l, o[16];
            const bool tm = p.dbg != nullptr && tid == 0 && blockIdx.x == 0;
            unsigned long long *td = p.dbg + 192 + K_ATT * 8 + (slow ? 0 : 3);
            const long long c0 = tm ? clock64() : 0;
            att_range(kcl, vcl, cache_len, b, kvh, j0, j1, m, l, o);
            const long long c1 = tm ? clock64() : 0;
            if (tm) { td[0] += c1 - c0; td[2] += 1; }
            const int h = kvh * kRep + warp;
            if (ns == 1) {
                if (g == 0) {
                    const float inv = 1.0f / l;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        store_xop4(e.xop_att, b, h * kHd + sub * 4 + jj * 16,
                                   make_float4(o[jj * 4 + 0] * inv, o[jj * 4 + 1] * inv, o[jj * 4 + 2] * inv, o[jj * 4 + 3] * inv));
                }
                continue;
            }
            if (g == 0) {
                float *out = e.apart + (((size_t)b * kH + h) * kMBMaxSplit + sp) * (kHd + 4);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    *reinterpret_cast<float4 *>(out + jj * 16 + sub * 4) =
                        make_float4(o[jj * 4 + 0], o[jj * 4 + 1], o[jj * 4 + 2], o[jj * 4 + 3]);
                if (sub == 0) { out[kHd] = m; out[kHd + 1] = l; }
            }
            // the last range of (b, kvh) to finish merges all of them (no waiting: whoever arrives last does it)
            __threadfence();
            wsync();
            if (tid == 0) {
                const unsigned old = atomicAdd(e.att_cnt + b * kKV + kvh, 1u);
                const int last = old == (unsigned)(ns - 1);
                if (last) e.att_cnt[b * kKV + kvh] = 0;  // next use is behind a grid barrier
                bcast_s[0] = last;
            }
            wsync();
            if (bcast_s[0]) {
                __threadfence();
                const float *pp = e.apart + ((size_t)b * kH + h) * kMBMaxSplit * (kHd + 4);
                // every partial (m, l, o) of the row is requested at once: one L2 round trip instead of two dependent ones
                float2 ml[kMBMaxSplit], ov[kMBMaxSplit];
#pragma unroll
                for (int i = 0; i < kMBMaxSplit; ++i) {
                    ml[i] = make_float2(-INFINITY, 0.f);
                    ov[i] = make_float2(0.f, 0.f);
                    if (i < ns) {
                        ml[i] = __ldcg(reinterpret_cast<const float2 *>(pp + i * (kHd + 4) + kHd));
                        ov[i] = __ldcg(reinterpret_cast<const float2 *>(pp + i * (kHd + 4) + lane * 2));
                    }
                }
                float M = -INFINITY;
#pragma unroll
                for (int i = 0; i < kMBMaxSplit; ++i) M = fmaxf(M, ml[i].x);
                float L = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int i = 0; i < kMBMaxSplit; ++i) {
                    const float w = (ml[i].x == -INFINITY) ? 0.f : expf(ml[i].x - M);
                    L = fmaf(ml[i].y, w, L);
                    a0 = fmaf(ov[i].x, w, a0);
                    a1 = fmaf(ov[i].y, w, a1);
                }
                store_xop1(e.xop_att, b, h * kHd + lane * 2, a0 / L);
                store_xop1(e.xop_att, b, h * kHd + lane * 2 + 1, a1 / L);
            }
            if (tm) td[1] += clock64() - c1;
        }
    }

    // ------------------------------------------------------------ samplers: CTA b owns row b
    __device__ __forceinline__ void load_sampler_state() {
        const GenState &st = p.st;
        const int b = blockIdx.x, C1 = kC + 1;
        if (tid == 0) {
            s_active[0] = st.active[b];
            s_eos[0] = st.eos[b];
            s_frame[0] = st.frame[b];
            s_maxf[0] = st.max_frames[b];
        }
        for (int i = tid; i < C1; i += kMBWorkers) {
            s_cur[i] = st.cur[b * C1 + i];
            s_prev[i] = st.prev[b * C1 + i];
        }
        const int words = (int)(sizeof(RepPenState) / 4);
        for (int i = tid; i < kC * words; i += kMBWorkers)
            reinterpret_cast<uint32_t *>(s_rep)[i] = reinterpret_cast<const uint32_t *>(st.rep + (size_t)b * kC)[i];
        wsync();
    }
    __device__ __forceinline__ void sampler_scratch(int n, unsigned char **scratch, float **vals, float **sred) {
        *scratch = xs;
        *vals = reinterpret_cast<float *>(xs) + (sel_scratch_bytes(kMBWorkers) + 15) / 16 * 4;
        *sred = *vals + ((n + 3) & ~3);
    }
    // row-owner helper: dst[0..D) = src row, its sum of squares -> ssq slots (slot 0 = total, the rest 0)
    __device__ __forceinline__ void finish_row(float acc_sq, float *ssq_row) {
        acc_sq = warp_sum(acc_sq);
        if (lane == 0) red[64 + warp] = acc_sq;
        wsync();
        if (tid < kMBSsq) {
            float t = 0.f;
            if (tid == 0)
                for (int w = 0; w < kMBWorkers / 32; ++w) t += red[64 + w];
            ssq_row[tid] = t;
        }
    }

    // slow input of the row's next frame: DualARTransformer::embed (dual_ar.rs:532-567) on the codes of the frame just
    // emitted (s_prev), plus its operand image (first block's attention_norm folded in) and sum of squares
    __device__ __forceinline__ void embed_next_input() {
        const int b = blockIdx.x;
        const __nv_bfloat16 *emb = reinterpret_cast<const __nv_bfloat16 *>(p.emb);
        const __nv_bfloat16 *cbe = reinterpret_cast<const __nv_bfloat16 *>(p.cb_emb);
        const uint32_t tok0 = s_prev[0];
        const bool msk = p.has_end ? (tok0 <= p.sem_end && tok0 >= p.sem_start) : (tok0 == p.sem_start);
        const float mf = msk ? 1.f : 0.f;
        const uint2 r0 = __ldg(reinterpret_cast<const uint2 *>(emb + (size_t)tok0 * kD) + tid);
        float4 acc = make_float4(bf16lo(r0.x), bf16hi(r0.x), bf16lo(r0.y), bf16hi(r0.y));
        uint2 rc[kC];
#pragma unroll
        for (int c = 0; c < kC; ++c)
            rc[c] = __ldg(reinterpret_cast<const uint2 *>(cbe + ((size_t)c * kCS + s_prev[1 + c]) * kD) + tid);
#pragma unroll
        for (int c = 0; c < kC; ++c) {
            acc.x = __fadd_rn(acc.x, __fmul_rn(bf16lo(rc[c].x), mf));
            acc.y = __fadd_rn(acc.y, __fmul_rn(bf16hi(rc[c].x), mf));
            acc.z = __fadd_rn(acc.z, __fmul_rn(bf16lo(rc[c].y), mf));
            acc.w = __fadd_rn(acc.w, __fmul_rn(bf16hi(rc[c].y), mf));
        }
        reinterpret_cast<float4 *>(p.x + (size_t)b * kD)[tid] = acc;
        const float4 g4 = __ldg(reinterpret_cast<const float4 *>(normtab[0]) + tid);  // first slow block's attention_norm
        store_xop4(e.xop_x, b, 4 * tid, make_float4(__fmul_rn(acc.x, g4.x), __fmul_rn(acc.y, g4.y), __fmul_rn(acc.z, g4.z), __fmul_rn(acc.w, g4.w)));
        finish_row(acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w, e.ssq_x + (size_t)b * kMBSsq);
    }

    __device__ __forceinline__ void sample_slow(int kframe) {
        const GenState &st = p.st;
        const int b = blockIdx.x, C1 = kC + 1;
        const int n = p.n_slow_logits;
        unsigned char *scratch;
        float *vals, *sred;
        sampler_scratch(n, &scratch, &vals, &sred);
        if (tid == 0 && b == 0) e.go[(kframe + 2) & 3] = 0;
        if (!s_active[0]) return;
        const int frame = s_frame[0];
        const float u = philox_uniform(st.sp.seed, (uint64_t)frame * C1, (uint32_t)(p.row0 + b));
        const float *lg = p.logits + (size_t)b * p.ldl;
        uint32_t tok;
        if (st.legacy_slow) {
            const float eos_l = __ldcg(lg), pad_l = __ldcg(lg + 1);
            const float mx = fmaxf(pad_l, eos_l);
            const float e_pad = expf(pad_l - mx), e_eos = expf(eos_l - mx);
            tok = (st.fixed_len || u < e_pad / (e_pad + e_eos)) ? st.pad_id : st.im_end_id;
        } else {
            for (int i = tid; i < n; i += kMBWorkers) {
                float v = __ldcg(lg + i);
                if (i == 0 && st.fixed_len) v = -INFINITY;
                vals[i] = v;
            }
            wsync();
            const int idx = block_sample_sel<MBSync>(vals, scratch, sred, n, st.sp, u);
            tok = (idx == 0) ? st.im_end_id : (p.sem_start + (uint32_t)idx - 1);
        }
        const bool eos = tok == st.im_end_id;
        if (tid == 0) {
            s_cur[0] = tok;
            st.cur[b * C1] = tok;
            s_eos[0] = eos ? 1 : 0;
            st.eos[b] = eos ? 1 : 0;
            if (eos)
                for (int c = 0; c < kC; ++c) {
                    s_cur[1 + c] = 0;
                    st.cur[b * C1 + 1 + c] = 0;
                }
            // the next frame runs iff some row goes on (single_batch.rs:193-204): every CTA's producer learns it
            // 8 fast steps ahead of the frame boundary
            if (!eos && frame + 1 < s_maxf[0]) atomicOr(e.go + ((kframe + 1) & 3), 1u);
        }
        // fast stack input of codebook 0: the pre-norm slow hidden state (Q1, dual_ar.rs:629-634)
        {
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(p.x + (size_t)b * kD) + tid);
            reinterpret_cast<float4 *>(p.fx + (size_t)b * kD)[tid] = v;
            const float4 g4 = __ldg(reinterpret_cast<const float4 *>(normtab[2 * p.NL]) + tid);  // first fast block's attention_norm
            store_xop4(e.xop_fx, b, 4 * tid, make_float4(__fmul_rn(v.x, g4.x), __fmul_rn(v.y, g4.y), __fmul_rn(v.z, g4.z), __fmul_rn(v.w, g4.w)));
            if (tid < kMBSsq) e.ssq_fx[(size_t)b * kMBSsq + tid] = __ldcg(e.ssq_x + (size_t)b * kMBSsq + tid);
        }
        wsync();
    }

    __device__ __forceinline__ void sample_fast(int cb) {
        const GenState &st = p.st;
        const int b = blockIdx.x, C1 = kC + 1;
        constexpr int n = kCS;
        unsigned char *scratch;
        float *vals, *sred;
        sampler_scratch(n, &scratch, &vals, &sred);
        if (!s_active[0]) return;
        const bool eos = s_eos[0] != 0;
        const int frame = s_frame[0];
        const bool tms = p.dbg != nullptr && tid == 0 && blockIdx.x == 0;
        long long c0 = tms ? clock64() : 0, c1 = 0, c2 = 0, c3 = 0;
        if (!eos) {
            RepPenState *rp = s_rep + cb;
            if (frame > 0) {
                if (tid == 0) rep_pen_update(rp, s_prev[1 + cb]);
                wsync();
            }
            if (tms) c1 = clock64();
            const float *lg = p.logits + (size_t)b * p.ldl;
            for (int i = tid; i < n; i += kMBWorkers) {
                float v = __ldcg(lg + i);
                if (frame > 0 && ((rp->seen[i >> 5] >> (i & 31)) & 1u)) v = __fdiv_rn(v, st.sp.penalty);
                vals[i] = v;
            }
            wsync();
            if (tms) c2 = clock64();
            const float u = philox_uniform(st.sp.seed, (uint64_t)frame * C1 + cb + 1, (uint32_t)(p.row0 + b));
            const int a = block_sample_sel<MBSync>(vals, scratch, sred, n, st.sp, u);
            if (tms) c3 = clock64();
            if (tid == 0) {
                s_cur[1 + cb] = (uint32_t)a;
                st.cur[b * C1 + 1 + cb] = (uint32_t)a;
            }
            if (cb + 1 < kC) {
                // next fast step's input: fast_embeddings[code] (single_batch.rs:176-182) and its sum of squares
                const __nv_bfloat16 *fe = reinterpret_cast<const __nv_bfloat16 *>(p.fast_emb) + (size_t)a * kD;
                const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(fe) + tid);
                const float4 v = make_float4(bf16lo(raw.x), bf16hi(raw.x), bf16lo(raw.y), bf16hi(raw.y));
                reinterpret_cast<float4 *>(p.fx + (size_t)b * kD)[tid] = v;
                const float4 g4 = __ldg(reinterpret_cast<const float4 *>(normtab[2 * p.NL]) + tid);
                store_xop4(e.xop_fx, b, 4 * tid, make_float4(__fmul_rn(v.x, g4.x), __fmul_rn(v.y, g4.y), __fmul_rn(v.z, g4.z), __fmul_rn(v.w, g4.w)));
                finish_row(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w, e.ssq_fx + (size_t)b * kMBSsq);
            }
        }
        if (cb == kC - 1) {
            wsync();
            // frame bookkeeping (single_batch.rs:193-204), write-through
            if (tid <= kC) {
                const uint32_t v = s_cur[tid];
                st.out[((size_t)b * st.out_cap + frame) * C1 + tid] = v;
                s_prev[tid] = v;
                st.prev[b * C1 + tid] = v;
            }
            bool cont = true;
            const int nf = frame + 1;
            if (eos || nf >= s_maxf[0]) cont = false;
            wsync();
            if (tid == 0) {
                s_frame[0] = nf;
                st.frame[b] = nf;
                if (frame > 0) st.pos[b] = st.pos[b] + 1;  // only this CTA writes the row's position
                if (!cont) {
                    s_active[0] = 0;
                    st.active[b] = 0;
                    atomicSub(st.n_active, 1);
                }
            }
            const int words = (int)(sizeof(RepPenState) / 4);
            const uint32_t *src = reinterpret_cast<const uint32_t *>(s_rep);
            uint32_t *dst = reinterpret_cast<uint32_t *>(st.rep + (size_t)b * kC);
            for (int i = tid; i < kC * words; i += kMBWorkers) dst[i] = src[i];
            if (cont) embed_next_input();
        }
        wsync();
        if (tms && !eos) {
            const long long c4 = clock64();
            p.dbg[100] += c1 - c0; p.dbg[101] += c2 - c1; p.dbg[102] += c3 - c2; p.dbg[103] += 1; p.dbg[116] += c4 - c3;
        }
    }

    // ------------------------------------------------------------ frame loop (workers)
    __device__ __forceinline__ void run() {
        Step cur = first_step();
        const bool samples = (int)blockIdx.x < p.nb;
        if (samples) load_sampler_state();
        if (samples && !p.first_is_tail) {
            if (s_active[0]) embed_next_input();  // resume: the row's input is rebuilt from the codes of its last frame
        }
        // sum of squares of the prefilled hidden rows (the launch starts at the slow head of frame 0)
        if (samples && p.first_is_tail) {
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(p.x + (size_t)blockIdx.x * kD) + tid);
            const float4 g4 = __ldg(reinterpret_cast<const float4 *>(p.norm) + tid);  // consumer: the slow head
            store_xop4(e.xop_x, blockIdx.x, 4 * tid, make_float4(__fmul_rn(v.x, g4.x), __fmul_rn(v.y, g4.y), __fmul_rn(v.z, g4.z), __fmul_rn(v.w, g4.w)));
            finish_row(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w, e.ssq_x + (size_t)blockIdx.x * kMBSsq);
        }
        if (blockIdx.x == 0 && tid == 0 && p.dbg) g_sample_dbg = p.dbg + 104;
        frame_prep();
        grid_arrive();
        grid_wait();
        const bool timed = p.dbg != nullptr && tid == 0 && (blockIdx.x == 0 || blockIdx.x == 64 || blockIdx.x == 74);
        unsigned long long *dbg = p.dbg + (blockIdx.x == 0 ? 0 : blockIdx.x == 64 ? 32 : 128);
        while (cur.kind != K_END) {
            unsigned long long t0 = 0, t1 = 0;
            if (timed) t0 = clock64();
            switch (cur.kind) {
                case K_SAMPLE:
                    if (samples) {
                        if (cur.pass == 0) sample_slow(cur.frame);
                        else sample_fast(cur.pass - 1);
                    }
                    break;
                case K_ATT: attn_phase(cur); break;
                case K_FFN: ffn_phase(cur); break;
                case K_FRED: fred_phase(cur); break;
                default: fullk_phase(cur);
            }
            const Step nxt = advance(cur);
            if (timed) t1 = clock64();
            grid_arrive();
            grid_wait();
            if (timed) {
                const unsigned long long t2 = clock64();
                dbg[cur.kind * 4 + 0] += t1 - t0;
                dbg[cur.kind * 4 + 2] += t2 - t1;
                dbg[cur.kind * 4 + 3] += 1;
            }
            if (cur.kind == K_SAMPLE && cur.pass == 0) {
                // the sampling CTAs said whether frame cur.frame + 1 runs before they arrived at this barrier
                if (tid == 0) *go_frames = ld_relaxed_u32(e.go + ((cur.frame + 1) & 3)) ? cur.frame + 2 : cur.frame + 1;
                wsync();
            }
            if (nxt.frame != cur.frame && nxt.kind != K_END) {
                if (nxt.frame >= *go_frames) break;
                frame_prep();  // positions / live rows / attention items of the new frame
            }
            cur = nxt;
        }
        wsync();
        if (tid == 0) {
            *done_flag = 1;
            *cmd = -1;
            m1_mbar_arrive(cmd_ready);
        }
    }
};

template <int NPAD>
__global__ void __launch_bounds__(kMBThreads, 1)
megab_decode_kernel(const __grid_constant__ MegaParams p, const __grid_constant__ MegaBExtra e) {
    extern __shared__ unsigned char megab_smem_raw[];
    // the 128-byte swizzle atoms need 1024-byte aligned tiles
    unsigned char *smem = megab_smem_raw + ((1024u - (m1_smem_u32(megab_smem_raw) & 1023u)) & 1023u);
    MegaB<NPAD> m(p, e, smem);
    const bool idle = p.nframes <= 0 || __ldcg(p.st.n_active) == 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < e.nstages; ++i) {
            m1_mbar_init(m.full + i, 1);
            m1_mbar_init(m.empty + i, 1);
        }
        m1_mbar_init(m.cmd_ready, 1);
        for (int i = 0; i < 2; ++i) {
            m1_mbar_init(m.xr + i, 1);
            m1_mbar_init(m.xf + i, 1);
        }
        m1_mbar_init(m.acc_full, 1);
        m1_mbar_init(m.hr, 1);
        for (int i = 0; i < 8; ++i) m1_mbar_init(m.a2f + i, 1);
        *m.go_frames = 1;
        *m.done_flag = 0;
        *m.cmd = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        // norm-weight pointers of every layer (chased every phase)
        const int nl = p.NL + p.NFL;
        for (int i = threadIdx.x; i < nl; i += blockDim.x) {
            const MegaLayer *L = i < p.NL ? p.slow + i : p.fast + (i - p.NL);
            m.normtab[2 * i] = L->attn_norm;
            m.normtab[2 * i + 1] = L->ffn_norm;
        }
        if (threadIdx.x == 0) {
            m.normtab[2 * nl] = p.norm;
            m.normtab[2 * nl + 1] = p.fast_norm;
        }
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(m1_smem_u32(m.tmem_slot)),
                     "n"(MegaB<NPAD>::kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!idle) {
        if (warp < 8) m.run();
        else if (warp == 8) { if ((threadIdx.x & 31) == 0) m.producer(); }
        else m.mma_loop();  // whole warp, converged: an elected lane issues
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 9) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*m.tmem_slot), "n"(MegaB<NPAD>::kTmemCols)
                     : "memory");
    }
}

}  // namespace fsb
