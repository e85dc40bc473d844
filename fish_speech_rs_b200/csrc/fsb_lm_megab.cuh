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
//   * a work unit is (128-row tile, 256-wide K slice): a CTA only ever stages a 256-column slice of the
//     activations (B x 256 x 4 B from L2 instead of B x 1024), and the K-slices of a tile are summed in a FIXED
//     order by a split-K fixup that is spread over the CTAs of the tile (each reduces its share of the batch rows
//     after a per-tile arrival counter says all partials have landed) -- deterministic, no atomics on data;
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
    static constexpr int kKs = 256;                   // K slice of a work unit
    // activation operand of one 64-wide k-block: [3 * NPAD rows][64 k] -- rows [0, NPAD) hold the hi terms of the batch
    // rows, [NPAD, 2 NPAD) the mid terms, [2 NPAD, 3 NPAD) the lo terms, so ONE MMA of N = 3 * NPAD multiplies a weight
    // tile with all three terms; the three column groups of the accumulator are added when it is drained
    static constexpr int kXKb = 3 * NPAD * 128;       // bytes per k-block
    static constexpr int kAccCols = 3 * NPAD;         // accumulator columns of one buffer
    static constexpr int kTmemCols = NPAD == 16 ? 128 : 256;  // two buffers, power of two
    static_assert(4 * kXKb <= kMBXsBytes, "activation operand does not fit");
    static_assert(2 * 2 * kMBChunk * kMBKvStride * 4 <= kMBXsBytes, "two K/V chunk buffers do not fit");

    enum { K_QKV = 0, K_ATT = 1, K_WO = 2, K_W13 = 3, K_W2 = 4, K_HEAD = 5, K_SAMPLE = 6, K_END = 7 };
    enum { G_QKV = 0, G_WO = 1, G_W13 = 2, G_W2 = 3, G_HEAD_S = 4, G_HEAD_F = 5, G_COUNT = 6 };
    struct Step { int frame, pass, l, kind; };  // pass 0 = slow stack, pass c + 1 = fast step of codebook c
    struct Gemm { int map0, map1, T, S, off, gk; };
    struct Units { int s, j, ns, n; };

    const MegaParams &p;
    const MegaBExtra &e;
    // shared memory
    unsigned char *ring, *xs;
    uint64_t *full, *empty, *x_ready, *acc_full, *acc_empty;
    uint32_t *tmem_slot;
    volatile int *cmd, *go_frames, *done_flag;
    int *pos_s, *act_s, *ibase_s, *nsplit_s, *prange_s, *epoch_s, *bcast_s;
    float *invd, *red;
    const float **normtab;
    int *s_active, *s_eos, *s_frame, *s_maxf;
    uint32_t *s_cur, *s_prev;
    RepPenState *s_rep;
    int tid, lane, warp;
    unsigned int target;     // grid barrier
    unsigned int ucount;     // units processed so far (accumulator buffer / parity)
    float gpre[8];           // norm weights of the coming phase for this thread's 8 columns

    __device__ MegaB(const MegaParams &pp, const MegaBExtra &ee, unsigned char *smem) : p(pp), e(ee) {
        ring = smem;
        xs = smem + (size_t)ee.nstages * kMBStage;
        unsigned char *q = xs + kMBXsBytes;
        full = reinterpret_cast<uint64_t *>(q); q += 8 * kMBMaxStages;
        empty = reinterpret_cast<uint64_t *>(q); q += 8 * kMBMaxStages;
        x_ready = reinterpret_cast<uint64_t *>(q); q += 8;
        acc_full = reinterpret_cast<uint64_t *>(q); q += 16;
        acc_empty = reinterpret_cast<uint64_t *>(q); q += 16;
        tmem_slot = reinterpret_cast<uint32_t *>(q); q += 8;
        cmd = reinterpret_cast<volatile int *>(q); q += 4;
        go_frames = reinterpret_cast<volatile int *>(q); q += 4;
        done_flag = reinterpret_cast<volatile int *>(q); q += 8;
        pos_s = reinterpret_cast<int *>(q); q += 4 * 32;
        act_s = reinterpret_cast<int *>(q); q += 4 * 32;
        ibase_s = reinterpret_cast<int *>(q); q += 4 * 36;
        nsplit_s = reinterpret_cast<int *>(q); q += 4 * 32;
        prange_s = reinterpret_cast<int *>(q); q += 4 * 32;
        epoch_s = reinterpret_cast<int *>(q); q += 4 * 8;
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
        ucount = 0;
    }
    static __host__ __device__ size_t smem_bytes(int nstages, int NL, int NFL) {
        return (size_t)nstages * kMBStage + kMBXsBytes + 8 * kMBMaxStages * 2 + 8 + 16 + 16 + 8 + 4 + 4 + 8 +
               4 * (32 + 32 + 36 + 32 + 32 + 8 + 8 + 32 + 128) + 8 * (2 * (NL + NFL) + 2) + 16 + 4 * 24 +
               sizeof(RepPenState) * 8 + 1024 /* alignment slack */;
    }

    static __device__ __forceinline__ void wsync() { MBSync::sync(); }

    // ------------------------------------------------------------ schedule
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
            case K_QKV: n.kind = K_ATT; break;
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
    __device__ __forceinline__ Gemm gemm_of(const Step &s) const {
        Gemm g;
        const bool slow = s.pass == 0;
        const int G = (int)gridDim.x;
        const int lb = 5 * (slow ? s.l : p.NL + s.l);
        g.map1 = -1;
        g.S = 4;
        g.off = 0;
        switch (s.kind) {
            case K_QKV: g.map0 = lb; g.T = kQKV / 128; g.gk = G_QKV; g.off = G > 40 ? G - 40 : 0; break;
            case K_WO: g.map0 = lb + 1; g.T = kD / 128; g.gk = G_WO; g.off = G > 20 ? G - 20 : 0; break;
            case K_W13: g.map0 = lb + 2; g.map1 = lb + 3; g.T = kI / 64; g.gk = G_W13; break;
            case K_W2: g.map0 = lb + 4; g.T = kD / 128; g.S = kI / kKs; g.gk = G_W2; break;
            default:
                g.map0 = 5 * (p.NL + p.NFL) + (slow ? 0 : 1);
                g.T = slow ? e.head_tiles + e.head_extra : kCS / 128;
                g.gk = slow ? G_HEAD_S : G_HEAD_F;
        }
        return g;
    }
    // units of this CTA: all of one K slice s, tiles j, j + ns, ... < T
    __device__ __forceinline__ Units units_of(const Gemm &g) const {
        Units u;
        const int G = (int)gridDim.x;
        int v = (int)blockIdx.x - g.off;
        if (v < 0) v += G;
        u.s = v % g.S;
        u.j = v / g.S;
        u.ns = (G - u.s + g.S - 1) / g.S;
        u.n = u.j < g.T ? (g.T - 1 - u.j) / u.ns + 1 : 0;
        return u;
    }
    // first weight row of tile t (TMA coordinate)
    __device__ __forceinline__ int tile_row(const Step &s, int t) const {
        if (s.kind == K_W13) return 64 * t;
        if (s.kind == K_HEAD && s.pass == 0) return t < e.head_tiles ? p.slow_rest_base - 1 + 128 * t : p.slow_row0;
        return 128 * t;
    }

    // ------------------------------------------------------------ TMA producer (warp 8, lane 0)
    __device__ __forceinline__ void producer() {
        Step s = first_step();
        if (s.kind == K_ATT || s.kind == K_SAMPLE) s = next_weight_step(s);
        const int depth = e.nstages;
        unsigned int slot = 0, epar = 1;  // parity of the `empty` phase to wait for (first lap: passes immediately)
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
            const Gemm g = gemm_of(s);
            const Units u = units_of(g);
            for (int i = 0; i < u.n; ++i) {
                const int row = tile_row(s, u.j + i * u.ns);
                for (int kb = 0; kb < kKs / 64; ++kb) {
                    mb_wait(empty + slot, epar);  // a fresh mbarrier reports the phase before its first one as complete
                    m1_mbar_expect_tx(full + slot, kMBStage);
                    unsigned char *dst = ring + (size_t)slot * kMBStage;
                    const int k0 = u.s * kKs + kb * 64;
                    if (g.map1 >= 0) {
                        mb_tma_2d(dst, e.maps + g.map0, full + slot, k0, row);
                        mb_tma_2d(dst + kMBStage / 2, e.maps + g.map1, full + slot, k0, row);
                    } else {
                        mb_tma_2d(dst, e.maps + g.map0, full + slot, k0, row);
                    }
                    if (++slot == (unsigned)depth) { slot = 0; epar ^= 1; }
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
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kAccCols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int depth = e.nstages;
        unsigned int slot = 0, spar = 0, units = 0, xphase = 0;  // ring slot / parity advance without divisions
        // descriptors differ only in the 14-bit start-address field (bytes >> 4): base + constant offsets
        const uint64_t desc_hi = ((uint64_t)((1024 >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        const uint64_t b_base = desc_hi | (uint64_t)((m1_smem_u32(xs) >> 4) & 0x3FFF);
        const uint64_t a_base = desc_hi | (uint64_t)((m1_smem_u32(ring) >> 4) & 0x3FFF);
        for (;;) {
            mb_wait(x_ready, xphase & 1);
            ++xphase;
            const int n = *cmd;
            if (n < 0) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int i = 0; i < n; ++i) {
                const unsigned buf = units & 1, use = units >> 1;
                if (use > 0) {
                    mb_wait(acc_empty + buf, (use - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t acc = tmem_base + buf * kAccCols;
#pragma unroll
                for (int kb = 0; kb < kKs / 64; ++kb) {
                    mb_wait(full + slot, spar);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint64_t ad = a_base + (uint64_t)(slot * (kMBStage >> 4));
                        const uint64_t bd = b_base + (uint64_t)(kb * (kXKb >> 4));
#pragma unroll
                        for (int k = 0; k < 4; ++k) mb_umma(acc, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        mb_commit(empty + slot);  // frees the ring stage once these MMAs have read it
                        if (kb == kKs / 64 - 1) mb_commit(acc_full + buf);
                    }
                    __syncwarp();
                    if (++slot == (unsigned)depth) { slot = 0; spar ^= 1; }
                }
                ++units;
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
    // rows [0, nb) x columns [k0, k0 + 256) of a row-major fp32 activation (leading dimension ld) -> three bf16
    // terms in the UMMA K-major layout: k-block kb = [3 NPAD rows][64 k] with 128-byte rows (hi | mid | lo row groups),
    // 16-byte chunk c of row r stored at chunk c ^ (r & 7) (the 128-byte swizzle TMA would produce).  Rows >= nb are zero.
    __device__ __forceinline__ void stage_x(const float *src, int ld, bool with_norm) {
        const int c32 = tid & 31;           // 16-byte chunk (8 columns) of the 256-column slice
        const int kb = c32 >> 3, c = c32 & 7;
#pragma unroll
        for (int i = 0; i < NPAD / 8; ++i) {
            const int b = (tid >> 5) + i * 8;
            float v[8];
            if (b < p.nb) {
                const float4 a0 = __ldcg(reinterpret_cast<const float4 *>(src + (size_t)b * ld + c32 * 8));
                const float4 a1 = __ldcg(reinterpret_cast<const float4 *>(src + (size_t)b * ld + c32 * 8 + 4));
                v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
                if (with_norm) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = __fmul_rn(v[j], gpre[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = 0.f;
            }
            unsigned short h[8], m[8], l[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) mb_split3(v[j], h[j], m[j], l[j]);
            const uint32_t off = (uint32_t)(kb * kXKb + b * 128 + ((c ^ (b & 7)) << 4));
            auto pack = [](const unsigned short (&q)[8]) {
                return make_uint4((uint32_t)q[0] | ((uint32_t)q[1] << 16), (uint32_t)q[2] | ((uint32_t)q[3] << 16),
                                  (uint32_t)q[4] | ((uint32_t)q[5] << 16), (uint32_t)q[6] | ((uint32_t)q[7] << 16));
            };
            *reinterpret_cast<uint4 *>(xs + off) = pack(h);
            *reinterpret_cast<uint4 *>(xs + NPAD * 128 + off) = pack(m);
            *reinterpret_cast<uint4 *>(xs + 2 * NPAD * 128 + off) = pack(l);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> tensor-core (async proxy) reads
    }

    // norm weights of the coming phase for this thread's 8 columns (issued before the grid-barrier wait)
    __device__ __forceinline__ void prep_step(const Step &s) {
        if (s.kind == K_END || s.kind == K_ATT || s.kind == K_SAMPLE || s.kind == K_WO || s.kind == K_W2) return;
        const Gemm g = gemm_of(s);
        const Units u = units_of(g);
        if (u.n == 0) return;
        const bool slow = s.pass == 0;
        const int li = slow ? s.l : p.NL + s.l;
        const float *gw = s.kind == K_QKV ? normtab[2 * li] : s.kind == K_W13 ? normtab[2 * li + 1]
                                                                              : normtab[2 * (p.NL + p.NFL) + (slow ? 0 : 1)];
        const float4 a = __ldg(reinterpret_cast<const float4 *>(gw + u.s * kKs + (tid & 31) * 8));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(gw + u.s * kKs + (tid & 31) * 8 + 4));
        gpre[0] = a.x; gpre[1] = a.y; gpre[2] = a.z; gpre[3] = a.w; gpre[4] = b.x; gpre[5] = b.y; gpre[6] = b.z; gpre[7] = b.w;
    }

    // sum of the S partials of (tile, batch row b, accumulator lane), fixed order
    __device__ __forceinline__ float psum(const float *wt, int S, int b, int ln) const {
        float a = __ldcg(wt + b * 128 + ln);
#pragma unroll 4
        for (int sp = 1; sp < S; ++sp) a += __ldcg(wt + ((size_t)sp * NPAD + b) * 128 + ln);
        return a;
    }
    __device__ __forceinline__ float2 psum2(const float *wt, int S, int b, int ln) const {
        float2 a = __ldcg(reinterpret_cast<const float2 *>(wt + b * 128 + ln));
#pragma unroll 4
        for (int sp = 1; sp < S; ++sp) {
            const float2 q = __ldcg(reinterpret_cast<const float2 *>(wt + ((size_t)sp * NPAD + b) * 128 + ln));
            a.x += q.x;
            a.y += q.y;
        }
        return a;
    }

    // ------------------------------------------------------------ one projection phase
    __device__ __forceinline__ void gemm_phase(const Step &s) {
        const Gemm g = gemm_of(s);
        const Units u = units_of(g);
        const bool slow = s.pass == 0;
        const int cb = s.pass - 1;
        const int kind = s.kind;
        float *stream = slow ? p.x : p.fx;
        float *ssq = slow ? e.ssq_x : e.ssq_fx;
        const bool with_norm = kind == K_QKV || kind == K_W13 || kind == K_HEAD;
        const int epoch = ++epoch_reg[g.gk];
        const bool tm = p.dbg != nullptr && tid == 0 && blockIdx.x == 0 && u.n > 0;
        unsigned long long *td = p.dbg + 128 + kind * 8;
        long long c0 = tm ? clock64() : 0, c1 = 0;
        if (u.n > 0) {
            const float *src = kind == K_WO ? e.att : kind == K_W2 ? p.h : stream;
            stage_x(src + u.s * kKs, kind == K_W2 ? kI : kD, with_norm);
            wsync();
            if (tid == 0) {
                *cmd = u.n;
                m1_mbar_arrive(x_ready);
            }
            if (tm) { c1 = clock64(); td[0] += c1 - c0; c0 = c1; }
            // 1 / sqrt(mean(x^2) + eps) of every row (candle_nn::RmsNorm), from the partial sums the producer of x left
            if (with_norm && tid < p.nb) {
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
        // ---- drain the accumulators: partial[tile][slice][b][lane]
        const uint32_t tmem_base = *tmem_slot;
        for (int i = 0; i < u.n; ++i) {
            const int t = u.j + i * u.ns;
            const unsigned buf = ucount & 1, use = ucount >> 1;
            mb_wait(acc_full + buf, use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tm) { c1 = clock64(); td[1] += c1 - c0; c0 = c1; }
            const int qd = warp & 3, cc0 = (warp >> 2) * (NPAD / 2);
            float r[NPAD / 2];
#pragma unroll
            for (int term = 0; term < 3; ++term) {
                uint32_t q[16];
                const uint32_t taddr = tmem_base + buf * kAccCols + term * NPAD + ((uint32_t)(qd * 32) << 16) + (uint32_t)cc0;
                if (NPAD == 32) mb_tmem_ld16(taddr, q);
                else mb_tmem_ld8(taddr, q);
#pragma unroll
                for (int j = 0; j < NPAD / 2; ++j) r[j] = term == 0 ? __uint_as_float(q[j]) : r[j] + __uint_as_float(q[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            m1_mbar_arrive(acc_empty + buf);
            float *dst = e.ws + (((size_t)t * g.S + u.s) * NPAD + cc0) * 128 + qd * 32 + lane;
#pragma unroll
            for (int j = 0; j < NPAD / 2; ++j)
                if (cc0 + j < p.nb) dst[j * 128] = r[j];
            ++ucount;
            wsync();
            if (tid == 0) mb_red_release(e.cnt + g.gk * kMBCntStride + t, 1u);
            if (tm) { c1 = clock64(); td[2] += c1 - c0; c0 = c1; }
        }
        // ---- split-K fixup: the CTA of slice s reduces batch rows [b_lo, b_hi) of each of its tiles
        const int per = (p.nb + g.S - 1) / g.S;
        const int b_lo = u.s * per, b_hi = min(p.nb, b_lo + per);
        if (tm) td[5] += 1;
        if (b_lo >= b_hi) return;
        for (int i = 0; i < u.n; ++i) {
            const int t = u.j + i * u.ns;
            if (tm) c0 = clock64();
            if (tid == 0) {
                const unsigned want = (unsigned)(g.S * epoch);
                const unsigned *c = e.cnt + g.gk * kMBCntStride + t;
                if (mb_ld_acquire(c) < want) {
                    const long long t0 = clock64();
                    while (mb_ld_acquire(c) < want)
                        if (clock64() - t0 > kMBSpinLimit) __trap();
                }
            }
            wsync();
            if (tm) { c1 = clock64(); td[3] += c1 - c0; c0 = c1; }
            const float *wt = e.ws + (size_t)t * g.S * NPAD * 128;
            if (kind == K_WO || kind == K_W2) {
                // residual add (dual_ar.rs:436-440) + this tile's share of sum(x^2) for the next RMSNorm
                const int ln = tid & 127;
                for (int b = b_lo + (tid >> 7); b < b_hi; b += 2) {
                    float *xp = stream + (size_t)b * kD + 128 * t + ln;
                    const float nv = __fadd_rn(__ldcg(xp), psum(wt, g.S, b, ln));
                    *xp = nv;
                    const float q = warp_sum(nv * nv);
                    if (lane == 0) ssq[(size_t)b * kMBSsq + t * 4 + (ln >> 5)] = q;
                }
            } else if (kind == K_W13) {
                // silu(w1 x) * (w3 x), dual_ar.rs:160-165: accumulator lanes [0, 64) = w1 rows, [64, 128) = w3 rows
                const int ii = tid & 63;
                for (int b = b_lo + (tid >> 6); b < b_hi; b += 4) {
                    const float g1 = psum(wt, g.S, b, ii) * invd[b], g3 = psum(wt, g.S, b, 64 + ii) * invd[b];
                    p.h[(size_t)b * kI + 64 * t + ii] = __fmul_rn(silu_f(g1), g3);
                }
            } else if (kind == K_QKV) {
                // rope_i on row pairs (dual_ar.rs:246-247) -> q buffer / K cache; V rows -> V cache (Tensor::cat, :316-324)
                const int pr = tid & 63;
                const size_t kv_stride = slow ? p.slow_kv_stride : p.fast_kv_stride;
                float *kcl = (slow ? p.kc : p.fkc) + s.l * kv_stride, *vcl = (slow ? p.vc : p.fvc) + s.l * kv_stride;
                const int cache_len = slow ? p.max_len : kC;
                for (int b = b_lo + (tid >> 6); b < b_hi; b += 4) {
                    float2 v = psum2(wt, g.S, b, 2 * pr);
                    v.x *= invd[b];
                    v.y *= invd[b];
                    const int pos = slow ? (act_s[b] ? pos_s[b] : 0) : cb;  // a finished row may sit at max_len
                    const int r = 128 * t + 2 * pr;
                    if (r < kD + kKV * kHd) {
                        const int pi = (r & 63) >> 1;
                        const float c = __ldg(p.cosT + (size_t)pos * 32 + pi), sn = __ldg(p.sinT + (size_t)pos * 32 + pi);
                        const float o0 = __fsub_rn(__fmul_rn(v.x, c), __fmul_rn(v.y, sn));
                        const float o1 = __fadd_rn(__fmul_rn(v.x, sn), __fmul_rn(v.y, c));
                        if (r < kD) {
                            *reinterpret_cast<float2 *>(p.q + (size_t)b * kD + r) = make_float2(o0, o1);
                        } else if (act_s[b]) {
                            const int rk = r - kD, kvh = rk >> 6, d = rk & 63;
                            *reinterpret_cast<float2 *>(kcl + (((size_t)b * kKV + kvh) * cache_len + pos) * kHd + d) = make_float2(o0, o1);
                        }
                    } else if (act_s[b]) {
                        const int rv = r - kD - kKV * kHd, kvh = rv >> 6, d = rv & 63;
                        *reinterpret_cast<float2 *>(vcl + (((size_t)b * kKV + kvh) * cache_len + pos) * kHd + d) = v;
                    }
                }
            } else {  // K_HEAD: logits of the constrained slow head (generate/utils.rs:6-33) or of fast_output
                const int ln = tid & 127;
                int r = 128 * t + ln;
                bool ok = r < (slow ? p.n_slow_logits : kCS);
                if (slow && e.head_extra) {
                    if (t == e.head_tiles) { r = 0; ok = ln == 0; }   // the extra tile starts at the <|im_end|> row
                    else if (r == 0) ok = false;
                }
                for (int b = b_lo + (tid >> 7); b < b_hi; b += 2)
                    if (ok) p.logits[(size_t)b * p.ldl + r] = psum(wt, g.S, b, ln) * invd[b];
            }
            if (tm) td[4] += clock64() - c0;
        }
    }
    int epoch_reg[G_COUNT];

    // ------------------------------------------------------------ attention
    // positions [j0, j1) of (row b, kv head kvh) for the 8 query heads of the group (warp = head); returns this
    // lane's (m, l, o[16]) already merged over the 8 position groups of the warp
    __device__ __forceinline__ void att_range(const float *kcache, const float *vcache, int cache_len, int b, int kvh,
                                              int j0, int j1, float &m, float &l, float (&o)[16]) {
        const int g = lane >> 2, sub = lane & 3;
        const int h = kvh * kRep + warp;
        constexpr int kBuf = 2 * kMBChunk * kMBKvStride;  // floats of one chunk buffer (K rows then V rows)
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
        auto stage = [&](int c0, int buf) {  // positions [c0, min(c0 + 64, j1)) -> chunk buffer `buf`
            const int n = min(kMBChunk, j1 - c0);
            float *ks = kvb + buf * kBuf, *vs = ks + kMBChunk * kMBKvStride;
            for (int i = tid; i < n * 16; i += kMBWorkers) {
                const int j = i >> 4, sg = i & 15;
                cp_async16(ks + j * kMBKvStride + sg * 4, kb + (size_t)(c0 + j) * kHd + sg * 4);
                cp_async16(vs + j * kMBKvStride + sg * 4, vb + (size_t)(c0 + j) * kHd + sg * 4);
            }
            cp_async_commit();
        };
        wsync();  // both buffers are free (previous item / previous user of the region)
        stage(j0, 0);
        int buf = 0;
        for (int c0 = j0; c0 < j1; c0 += kMBChunk, buf ^= 1) {
            const int n = min(kMBChunk, j1 - c0);
            // the next chunk streams in while this one is consumed
            if (c0 + kMBChunk < j1) {
                stage(c0 + kMBChunk, buf ^ 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                cp_async_wait_all();
            }
            wsync();
            const float *ks = kvb + buf * kBuf, *vs = ks + kMBChunk * kMBKvStride;
            for (int jb = 0; jb < n; jb += 8) {  // warp-uniform trip count (the shuffles need all lanes)
                const int j = jb + g;
                const bool valid = j < n;
                const int jc = valid ? j : 0;
                const float *kr = ks + jc * kMBKvStride + sub * 4, *vr = vs + jc * kMBKvStride + sub * 4;
                float dot = 0.f;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float4 kk = *reinterpret_cast<const float4 *>(kr + jj * 16);
                    dot = fmaf(qv[jj].x, kk.x, dot);
                    dot = fmaf(qv[jj].y, kk.y, dot);
                    dot = fmaf(qv[jj].z, kk.z, dot);
                    dot = fmaf(qv[jj].w, kk.w, dot);
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
            wsync();  // this buffer is overwritten by the chunk after next
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
            int total = 0;
            for (int b = 0; b < p.nb; ++b) total += act_s[b] ? pos_s[b] + 1 : 0;
            const int G = (int)gridDim.x;
            int pit = (total * kKV * 10 + G * 9 - 1) / (G * 9);
            pit = max(kMBChunk, (pit + kMBChunk - 1) / kMBChunk * kMBChunk);
            int base = 0;
            for (int b = 0; b < p.nb; ++b) {
                const int len = pos_s[b] + 1;
                int ns = min(kMBMaxSplit, (len + pit - 1) / pit);
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
            att_range(kcl, vcl, cache_len, b, kvh, j0, j1, m, l, o);
            const int h = kvh * kRep + warp;
            if (ns == 1) {
                if (g == 0) {
                    float *out = e.att + (size_t)b * kD + h * kHd + sub * 4;
                    const float inv = 1.0f / l;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        *reinterpret_cast<float4 *>(out + jj * 16) =
                            make_float4(o[jj * 4 + 0] * inv, o[jj * 4 + 1] * inv, o[jj * 4 + 2] * inv, o[jj * 4 + 3] * inv);
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
                float M = -INFINITY;
                for (int i = 0; i < ns; ++i) M = fmaxf(M, __ldcg(pp + i * (kHd + 4) + kHd));
                float L = 0.f, a0 = 0.f, a1 = 0.f;
                for (int i = 0; i < ns; ++i) {
                    const float mi = __ldcg(pp + i * (kHd + 4) + kHd), li = __ldcg(pp + i * (kHd + 4) + kHd + 1);
                    const float w = (mi == -INFINITY) ? 0.f : expf(mi - M);
                    const float2 ov = __ldcg(reinterpret_cast<const float2 *>(pp + i * (kHd + 4) + lane * 2));
                    L = fmaf(li, w, L);
                    a0 = fmaf(ov.x, w, a0);
                    a1 = fmaf(ov.y, w, a1);
                }
                *reinterpret_cast<float2 *>(e.att + (size_t)b * kD + h * kHd + lane * 2) = make_float2(a0 / L, a1 / L);
            }
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
        if (!eos) {
            RepPenState *rp = s_rep + cb;
            if (frame > 0) {
                if (tid == 0) rep_pen_update(rp, s_prev[1 + cb]);
                wsync();
            }
            const float *lg = p.logits + (size_t)b * p.ldl;
            for (int i = tid; i < n; i += kMBWorkers) {
                float v = __ldcg(lg + i);
                if (frame > 0 && ((rp->seen[i >> 5] >> (i & 31)) & 1u)) v = __fdiv_rn(v, st.sp.penalty);
                vals[i] = v;
            }
            wsync();
            const float u = philox_uniform(st.sp.seed, (uint64_t)frame * C1 + cb + 1, (uint32_t)(p.row0 + b));
            const int a = block_sample_sel<MBSync>(vals, scratch, sred, n, st.sp, u);
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
            if (cont) {
                // next frame's slow input: DualARTransformer::embed (dual_ar.rs:532-567) on this frame's codes
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
                finish_row(acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w, e.ssq_x + (size_t)b * kMBSsq);
            }
        }
        wsync();
    }

    // ------------------------------------------------------------ frame loop (workers)
    __device__ __forceinline__ void run() {
#pragma unroll
        for (int i = 0; i < G_COUNT; ++i) epoch_reg[i] = 0;
        Step cur = first_step();
        const bool samples = (int)blockIdx.x < p.nb;
        if (samples) load_sampler_state();
        // sum of squares of the prefilled hidden rows (the launch starts at the slow head of frame 0)
        if (samples) {
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(p.x + (size_t)blockIdx.x * kD) + tid);
            finish_row(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w, e.ssq_x + (size_t)blockIdx.x * kMBSsq);
        }
        frame_prep();
        grid_arrive();
        prep_step(cur);
        grid_wait();
        const bool timed = p.dbg != nullptr && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
        unsigned long long *dbg = p.dbg + (blockIdx.x == 0 ? 0 : 32);
        while (cur.kind != K_END) {
            unsigned long long t0 = 0, t1 = 0, t2 = 0;
            if (timed) t0 = clock64();
            if (cur.kind == K_SAMPLE) {
                if (samples) {
                    if (cur.pass == 0) sample_slow(cur.frame);
                    else sample_fast(cur.pass - 1);
                }
            } else if (cur.kind == K_ATT) {
                attn_phase(cur);
            } else {
                gemm_phase(cur);
            }
            const Step nxt = advance(cur);
            if (timed) t1 = clock64();
            grid_arrive();
            if (nxt.kind == K_SAMPLE) prep_step(advance(nxt));
            else if (cur.kind != K_SAMPLE) prep_step(nxt);
            if (timed) t2 = clock64();
            grid_wait();
            if (timed) {
                const unsigned long long t3 = clock64();
                dbg[cur.kind * 4 + 0] += t1 - t0;
                dbg[cur.kind * 4 + 1] += t2 - t1;
                dbg[cur.kind * 4 + 2] += t3 - t2;
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
            m1_mbar_arrive(x_ready);
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
        m1_mbar_init(m.x_ready, 1);
        for (int i = 0; i < 2; ++i) {
            m1_mbar_init(m.acc_full + i, 1);
            m1_mbar_init(m.acc_empty + i, kMBWorkers);
        }
        *m.go_frames = 1;
        *m.done_flag = 0;
        *m.cmd = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        // norm-weight pointers of every layer (chased by prep_step every phase)
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
