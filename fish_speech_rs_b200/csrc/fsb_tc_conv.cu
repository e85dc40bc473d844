// HiFi-GAN ResBlock convolutions on tcgen05 + TMEM (replaces the FP32-FMA resconv_tma_kernel for C = 256 / 128 / 64).
//
//   y[t, co] = bias[co] + sum_{tap, ci} W[co, ci, tap] * silu(x)[t - (K - 1 - tap) * dil, ci]        (hifi_gan.rs:60-86)
//
// is a GEMM per tap with M = time (128 rows per MMA), N = couts (= C), K = cins, all taps accumulating into the same TMEM
// tile.  The activations live in HBM as fp16 "images": [2 terms][C / 8 chunks][rows][8 channels], i.e. for every group of 8
// channels a time-major array of 16-byte rows.  That is exactly the no-swizzle K-major UMMA operand layout (core matrix =
// 8 rows x 16 bytes, rows 16 bytes apart), so
//   * a tile's input is ONE contiguous cp.async.bulk per (term, chunk) -- no tensor map, no transposition;
//   * the tap shift is a 16-byte * (tap * dil) offset of the descriptor start address: every tap reads the SAME staged tile;
//   * the causal left padding is kTcvPadL physical zero rows in front of every chunk.
// Weights are re-laid out at load time as [kb][tap][term][chunk][co][8 ci]: one contiguous bulk copy per (kb, tap) stage.
//
// Precision: an fp32 value v is split as hi = fp16(v), lo = fp16(v - hi) after scaling by a power of two (activations x 16,
// weights so that max |w| lands in [2^13, 2^14)): 22 mantissa bits per operand.  fp16 x fp16 products are exact in the fp32
// accumulator; the kernel issues hi*hi + hi*lo + lo*hi (the dropped lo*lo term is 2^-22 relative), so the result matches the
// FP32 FMA path to ~1e-6 relative -- the PCM parity bar (1e-4) is unchanged -- at 1/3 of the tensor pipe's fp16 rate.
//
// Structure (one CTA = MT x 128 time steps x all C couts, 320 threads, accumulators = MT * C TMEM columns):
//   warp 8 / lane 0 : producer   cp.async.bulk -> 2 activation stages (one per 16-channel block) + ring of weight stages
//   warp 9          : MMA issuer tcgen05.mma.cta_group::1.kind::f16 (fp16 in, fp32 accumulate), tcgen05.commit -> mbarrier
//   warps 0..7      : epilogue   tcgen05.ld -> + bias (+ residual) -> chunked f32 / next image (silu, split) / mean buffer
#include "fsb_tc_conv.cuh"

namespace fsb {

namespace {

constexpr int kTcvThreads = 320;
constexpr int kTcvBarBytes = 512;

__device__ __forceinline__ uint32_t tv_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tv_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tv_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tv_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tv_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tv_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(tv_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tv_bulk(void *smem, const void *gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tv_smem_u32(smem)),
                 "l"(gmem), "r"(bytes), "r"(tv_smem_u32(bar))
                 : "memory");
}
// K-major, no swizzle: core matrices of 8 rows x 16 bytes; `lbo` = bytes between the two core matrices along K,
// 128 bytes between 8-row groups (rows are uniformly 16 bytes apart)
__device__ __forceinline__ uint64_t tv_desc_hi(uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((128 >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100); layout type 0 = no swizzle
    return d;
}
__device__ __forceinline__ void tv_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tv_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tv_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tv_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
        "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tv_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <int CB>
__device__ __forceinline__ void tv_ld(uint32_t taddr, uint32_t (&r)[CB]) {
    if constexpr (CB == 32) tv_ld32(taddr, r);
    else tv_ld16(taddr, r);
}
// silu for the epilogue: x * 1 / (1 + 2^(-x log2 e)) on ex2.approx / rcp.approx (2^-22 relative, the precision of the image)
__device__ __forceinline__ float tv_silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// two scaled values -> packed fp16 hi pair and lo pair
__device__ __forceinline__ void tv_split2(float v0, float v1, uint32_t &hi, uint32_t &lo) {
    v0 = fminf(fmaxf(v0, -60000.f), 60000.f);
    v1 = fminf(fmaxf(v1, -60000.f), 60000.f);
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - f.x, v1 - f.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
__device__ __forceinline__ bool tv_elect() {
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

// v (already scaled) -> hi, lo fp16 with hi + lo == v to 2^-22 relative; saturates instead of overflowing to inf
__device__ __forceinline__ void tv_split(float v, __half &hi, __half &lo) {
    v = fminf(fmaxf(v, -60000.f), 60000.f);
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}
__device__ __forceinline__ uint32_t tv_pack(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// N = couts per tile (the MMA N); a.C = channels (cins = all couts).  Persistent: CTA b walks tiles b, b + gridDim.x, ...
// (tile = (time tile, cout tile)); the producer and the MMA warp run ahead into the next tile while the epilogue warps
// drain the previous one (two TMEM accumulator buffers).
template <int N>
__global__ void __launch_bounds__(kTcvThreads, 1) tcconv_kernel(const TcConvArgs a) {
    extern __shared__ uint8_t tv_smem_raw[];
    uint8_t *smem = tv_smem_raw + ((128u - (tv_smem_u32(tv_smem_raw) & 127u)) & 127u);
    const int MT = a.MT, K = a.K, dil = a.dil, S = a.wstages;
    const int halo = (K - 1) * dil;
    const int rows_ld = 128 * MT + halo;                  // staged rows per (term, chunk)
    const uint32_t x_plane = (uint32_t)rows_ld * 16u;     // bytes of one (term, chunk) plane
    const uint32_t x_stage = 4u * x_plane;
    constexpr uint32_t w_plane = (uint32_t)N * 16u, w_tap = 4u * w_plane;   // one (kb, tap): [2 chunks][hi | lo][N][16 B]
    // N <= 64: the hi and lo weight rows of a chunk form ONE B operand of 2 N rows, so x_hi * [w_hi ; w_lo] is a single MMA
    // into an accumulator tile of 2 N columns (hi*hi | hi*lo) and x_lo * w_hi a second one into its first N columns: the
    // activation tile -- the bulk of the shared-memory operand traffic that bounds this kernel -- is read twice, not 3 times
    constexpr bool STK = N <= 64;
    constexpr int AW = STK ? 2 * N : N;  // accumulator columns per M tile
    const int tps = a.tps;                                                   // taps per weight stage (1, or K for small N)
    const uint32_t w_stage = (uint32_t)tps * w_tap;
    const int NCH = a.C / 8, NKB = a.C / 16, NCT = a.ncols / N;  // input chunks / 16-channel blocks, column tiles
    uint64_t *xfull = reinterpret_cast<uint64_t *>(smem);  // [2]
    uint64_t *xempty = xfull + 2;                          // [2]
    uint64_t *accfull = xempty + 2;                        // [2]
    uint64_t *accempty = accfull + 2;                      // [2]
    uint64_t *wfull = accempty + 2;                        // [S]
    uint64_t *wempty = wfull + S;                          // [S]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(wempty + S);
    const uint32_t x_stage_al = (x_stage + 127u) & ~127u;
    uint8_t *xs = smem + kTcvBarBytes;
    uint8_t *ws = xs + 2 * (size_t)x_stage_al;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = a.ntiles;
    const size_t Lp = tcv_image_rows(a.L);
    // output geometry: column n of the GEMM is (phase r, cout) = (n / cout, n % cout), written at row t * ostride + r
    // (ostride = 1, r = 0 for a plain conv; the transposed convs interleave their `ostride` phases)
    const int nch_out = a.cout / 8, Lout = a.L * a.ostride;
    const size_t Lp_out = tcv_image_rows(Lout);
    const uint32_t acc_cols = (uint32_t)(MT * AW);
    const int kb_rot = a.rot ? (int)(blockIdx.x % (unsigned)NKB) : 0;  // experiment: per-CTA cyclic order of the channel blocks

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            tv_mbar_init(xfull + i, 1);
            tv_mbar_init(xempty + i, 1);
            tv_mbar_init(accfull + i, 1);
            tv_mbar_init(accempty + i, 8);
        }
        for (int i = 0; i < S; ++i) {
            tv_mbar_init(wfull + i, 1);
            tv_mbar_init(wempty + i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // two accumulator buffers of MT tiles x N fp32 columns, allocation rounded up to a power of two
    uint32_t ncols = 32;
    while (ncols < 2 * acc_cols) ncols <<= 1;
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tv_smem_u32(tmem_slot)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            // ---- producer
            int wi = 0, xi = 0;
            uint32_t wph = 0, xph = 0;
            for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
                const int t0 = (ti / NCT) * 128 * MT, ct = ti % NCT;
                const __half *xbase = a.ximg + ((size_t)kTcvPadL + t0 - halo) * 8;
                const __half *wsrc = a.wimg + (size_t)ct * NKB * K * (w_tap / 2);
                for (int kbi = 0; kbi < NKB; ++kbi) {
                    const int kb = (kbi + kb_rot) % NKB;
                    tv_wait(xempty + xi, xph ^ 1);
                    tv_expect_tx(xfull + xi, x_stage);
                    uint8_t *xd = xs + (size_t)xi * x_stage_al;
#pragma unroll
                    for (int term = 0; term < 2; ++term)
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch)
                            tv_bulk(xd + (term * 2 + ch) * x_plane, xbase + ((size_t)(term * NCH + 2 * kb + ch) * Lp) * 8, x_plane,
                                    xfull + xi);
                    if (++xi == 2) {
                        xi = 0;
                        xph ^= 1;
                    }
                    for (int tap = 0; tap < K; tap += tps) {
                        const uint32_t bytes = (uint32_t)min(tps, K - tap) * w_tap;  // the last stage of a block may be short
                        tv_wait(wempty + wi, wph ^ 1);
                        tv_expect_tx(wfull + wi, bytes);
                        tv_bulk(ws + (size_t)wi * w_stage, wsrc + (size_t)(kb * K + tap) * (w_tap / 2), bytes, wfull + wi);
                        if (++wi == S) {
                            wi = 0;
                            wph ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ---- MMA issuer (whole warp converged, one elected lane issues)
        // D[t, co] += X_term[t + tap * dil, 16 ci] * W_term[co, 16 ci]:  hi*hi, hi*lo, lo*hi
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // f16 x f16 -> f32
        constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t a_hi = tv_desc_hi(x_plane), b_hi = tv_desc_hi(2 * w_plane);
        int wi = 0, xi = 0, it = 0;
        uint32_t wph = 0, xph = 0;
        long long tw_acc = 0, tw_x = 0, tw_w = 0, tw_issue = 0, tq = 0, n_st = 0;
        const bool dbg = a.dbg != nullptr && blockIdx.x == 0;
        const long long t_begin = dbg ? clock64() : 0;
        for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x, ++it) {
            const int buf = it & 1;
            if (dbg) tq = clock64();
            tv_wait(accempty + buf, ((it >> 1) & 1) ^ 1);  // the epilogue warps are done with this buffer (tile it - 2)
            if (dbg) tw_acc += clock64() - tq;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t dbase = tmem_base + (uint32_t)buf * acc_cols;
            for (int kbi = 0; kbi < NKB; ++kbi) {
                if (dbg) tq = clock64();
                tv_wait(xfull + xi, xph);
                if (dbg) tw_x += clock64() - tq;
                const uint32_t xaddr = tv_smem_u32(xs + (size_t)xi * x_stage_al);
                for (int tap0 = 0; tap0 < K; tap0 += tps) {
                    const int ntp = min(tps, K - tap0);
                    if (dbg) tq = clock64();
                    tv_wait(wfull + wi, wph);
                    if (dbg) { const long long t1 = clock64(); tw_w += t1 - tq; tq = t1; ++n_st; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (tv_elect()) {
                        uint32_t waddr = tv_smem_u32(ws + (size_t)wi * w_stage);
                        uint32_t xrow = xaddr + (uint32_t)(tap0 * dil) * 16u;
                        for (int tp = 0; tp < ntp; ++tp, waddr += w_tap, xrow += (uint32_t)dil * 16u) {
                            const uint64_t b0 = b_hi | (uint64_t)((waddr >> 4) & 0x3FFF);
                            const uint64_t b1 = b_hi | (uint64_t)(((waddr + w_plane) >> 4) & 0x3FFF);
                            const uint32_t first = (kbi | tap0 | tp) != 0 ? 1u : 0u;
                            for (int mt = 0; mt < MT; ++mt) {
                                const uint32_t xa = xrow + (uint32_t)mt * 2048u;
                                const uint64_t a0 = a_hi | (uint64_t)((xa >> 4) & 0x3FFF);
                                const uint64_t a1 = a_hi | (uint64_t)(((xa + 2 * x_plane) >> 4) & 0x3FFF);
                                const uint32_t d = dbase + (uint32_t)(mt * AW);
                                if constexpr (STK) {
                                    tv_mma(d, a0, b0, idesc2, first);  // [hi*hi | hi*lo]
                                    tv_mma(d, a1, b0, idesc, 1u);      // lo*hi into the first N columns
                                } else {
                                    tv_mma(d, a0, b0, idesc, first);
                                    tv_mma(d, a0, b1, idesc, 1u);
                                    tv_mma(d, a1, b0, idesc, 1u);
                                }
                            }
                        }
                        tv_commit(wempty + wi);
                        if (tap0 + tps >= K) tv_commit(xempty + xi);
                        if (tap0 + tps >= K && kbi == NKB - 1) tv_commit(accfull + buf);
                    }
                    __syncwarp();
                    if (dbg) tw_issue += clock64() - tq;
                    if (++wi == S) {
                        wi = 0;
                        wph ^= 1;
                    }
                }
                if (++xi == 2) {
                    xi = 0;
                    xph ^= 1;
                }
            }
        }
        if (dbg && lane == 0) {
            a.dbg[0] = clock64() - t_begin; a.dbg[1] = tw_acc; a.dbg[2] = tw_x; a.dbg[3] = tw_w; a.dbg[4] = tw_issue; a.dbg[5] = n_st;
        }
    } else {
        // ---- epilogue: warp w reads TMEM lanes [32 (w & 3), +32) (= time rows); the (M tile, column block) pairs of a tile
        // are dealt alternately to the warp groups 0..3 and 4..7
        const int q = warp & 3, grp = warp >> 2;
        const int row = q * 32 + lane;
        constexpr int CB = N < 32 ? N : 32, NB = N / CB;  // column block
        const float inv_scale = a.inv_scale;
        const int nblk = MT * NB;
        int it = 0;
        for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x, ++it) {
            const int buf = it & 1;
            const int t0 = (ti / NCT) * 128 * MT, co0 = (ti % NCT) * N;
            if (a.yimg && t0 == 0 && row < kTcvPadL && grp == 0) {
                // the consumer's causal left padding
                const uint4 z = make_uint4(0, 0, 0, 0);
                for (int n8 = co0 / 8; n8 < (co0 + N) / 8; ++n8)
#pragma unroll
                    for (int term = 0; term < 2; ++term)
                        *reinterpret_cast<uint4 *>(a.yimg + ((size_t)(term * nch_out + n8 % nch_out) * Lp_out + row) * 8) = z;
            }
            // residual / mean-accumulator operands of the first block are requested before the accumulators are ready
            float4 r[CB / 4];
            auto load_res = [&](int blk) {
                const int mt = blk / NB, cb = (blk % NB) * CB;
                const int t = t0 + mt * 128 + row;
                if (t >= a.L) return;
                if (a.res) {
#pragma unroll
                    for (int g = 0; g < CB / 8; ++g) {
                        const size_t ro = ((size_t)((co0 + cb) / 8 + g) * a.L + t) * 8;  // (residuals only with ostride == 1)
                        r[2 * g] = *reinterpret_cast<const float4 *>(a.res + ro);
                        r[2 * g + 1] = *reinterpret_cast<const float4 *>(a.res + ro + 4);
                    }
                }
            };
            int blk = grp;
            if (blk < nblk) load_res(blk);
            tv_wait(accfull + buf, (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (; blk < nblk; blk += 2) {
                const int mt = blk / NB, cb = (blk % NB) * CB;
                const int t = t0 + mt * 128 + row;
                uint32_t v[CB];
                __syncwarp();
                tv_ld<CB>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * acc_cols + (uint32_t)(mt * AW + cb), v);
                if constexpr (STK) {
                    uint32_t v2[CB];
                    tv_ld<CB>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * acc_cols + (uint32_t)(mt * AW + N + cb), v2);
#pragma unroll
                    for (int j = 0; j < CB; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
                }
                float o[CB];
                if (t < a.L) {
#pragma unroll
                    for (int g = 0; g < CB / 8; ++g) {
                        const int co = (co0 + cb + g * 8) % a.cout;
                        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(a.bias + co));
                        const float4 b1 = __ldg(reinterpret_cast<const float4 *>(a.bias + co + 4));
                        o[g * 8 + 0] = fmaf(__uint_as_float(v[g * 8 + 0]), inv_scale, b0.x);
                        o[g * 8 + 1] = fmaf(__uint_as_float(v[g * 8 + 1]), inv_scale, b0.y);
                        o[g * 8 + 2] = fmaf(__uint_as_float(v[g * 8 + 2]), inv_scale, b0.z);
                        o[g * 8 + 3] = fmaf(__uint_as_float(v[g * 8 + 3]), inv_scale, b0.w);
                        o[g * 8 + 4] = fmaf(__uint_as_float(v[g * 8 + 4]), inv_scale, b1.x);
                        o[g * 8 + 5] = fmaf(__uint_as_float(v[g * 8 + 5]), inv_scale, b1.y);
                        o[g * 8 + 6] = fmaf(__uint_as_float(v[g * 8 + 6]), inv_scale, b1.z);
                        o[g * 8 + 7] = fmaf(__uint_as_float(v[g * 8 + 7]), inv_scale, b1.w);
                        if (a.res) {
                            o[g * 8 + 0] = r[2 * g].x + o[g * 8 + 0]; o[g * 8 + 1] = r[2 * g].y + o[g * 8 + 1];
                            o[g * 8 + 2] = r[2 * g].z + o[g * 8 + 2]; o[g * 8 + 3] = r[2 * g].w + o[g * 8 + 3];
                            o[g * 8 + 4] = r[2 * g + 1].x + o[g * 8 + 4]; o[g * 8 + 5] = r[2 * g + 1].y + o[g * 8 + 5];
                            o[g * 8 + 6] = r[2 * g + 1].z + o[g * 8 + 6]; o[g * 8 + 7] = r[2 * g + 1].w + o[g * 8 + 7];
                        }
                    }
                }
                // the next block's residual is in flight while this one is stored (a.y may alias a.res: same thread, same
                // elements, already consumed above)
                if (blk + 2 < nblk) load_res(blk + 2);
                if (t < a.L) {
                    if (a.m) {
                        float *mp = a.m + (size_t)(co0 + cb) * a.L + t;
                        if (a.acc_mode != 0) {
                            float pm[CB];
#pragma unroll
                            for (int j = 0; j < CB; ++j) pm[j] = mp[(size_t)j * a.L];
#pragma unroll
                            for (int j = 0; j < CB; ++j) o[j] = a.acc_mode == 2 ? (pm[j] + o[j]) * a.scale : pm[j] + o[j];
                        }
#pragma unroll
                        for (int j = 0; j < CB; ++j) mp[(size_t)j * a.L] = o[j];
                    }
#pragma unroll
                    for (int g = 0; g < CB / 8; ++g) {
                        const int n8 = (co0 + cb) / 8 + g, r = n8 / nch_out, c = n8 - r * nch_out;
                        const int to = t * a.ostride + r;
                        const size_t ro = ((size_t)c * Lout + to) * 8;
                        if (a.y) {
                            *reinterpret_cast<float4 *>(a.y + ro) = make_float4(o[g * 8 + 0], o[g * 8 + 1], o[g * 8 + 2], o[g * 8 + 3]);
                            *reinterpret_cast<float4 *>(a.y + ro + 4) = make_float4(o[g * 8 + 4], o[g * 8 + 5], o[g * 8 + 6], o[g * 8 + 7]);
                        }
                        if (a.yimg) {
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                tv_split2(tv_silu_fast(o[g * 8 + 2 * j]) * kTcvXScale, tv_silu_fast(o[g * 8 + 2 * j + 1]) * kTcvXScale,
                                          hi[j], lo[j]);
                            const size_t io = ((size_t)c * Lp_out + kTcvPadL + to) * 8;
                            *reinterpret_cast<uint4 *>(a.yimg + io) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            *reinterpret_cast<uint4 *>(a.yimg + (size_t)nch_out * Lp_out * 8 + io) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                }
            }
            // hand the accumulator buffer back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tv_smem_u32(accempty + buf)) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 9) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// (Cout, Cin, K) f32 -> [co tile][kb][tap][chunk][term][NT co][8 ci] fp16, scaled (hi rows, then lo rows, per 8-channel chunk)
__global__ void tcv_weight_image_kernel(const float *__restrict__ raw, __half *__restrict__ img, int C, int K, int NT, float s_w) {
    const size_t n = (size_t)C * C * K;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int tap = (int)(i % K), ci = (int)((i / K) % C), co = (int)(i / ((size_t)C * K));
        const int kb = ci >> 4, ch = (ci >> 3) & 1, j = ci & 7;
        __half hi, lo;
        tv_split(raw[i] * s_w, hi, lo);
        const int ct = co / NT, col = co % NT;
        const size_t o = (((((size_t)(ct * (C / 16) + kb) * K + tap) * 2 + ch) * 2 + 0) * NT + col) * 8 + j;
        img[o] = hi;
        img[o + (size_t)NT * 8] = lo;
    }
}

// ConvTranspose1d weight (Cin, Cout, 2 s) -> image of the equivalent 2-tap causal conv with s * Cout output columns:
//   column n = r * Cout + co;  tap 0 (x[t - 1]) = W[ci, co, r + s],  tap 1 (x[t]) = W[ci, co, r]     (utils/mod.rs:110-122)
__global__ void tcv_weight_image_t_kernel(const float *__restrict__ raw, __half *__restrict__ img, int Cin, int Cout, int s, int NT,
                                          float s_w) {
    const size_t n = (size_t)Cin * Cout * 2 * s;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % (2 * s)), co = (int)((i / (2 * s)) % Cout), ci = (int)(i / ((size_t)Cout * 2 * s));
        const int tap = k >= s ? 0 : 1, r = k >= s ? k - s : k;
        const int col = r * Cout + co, ct = col / NT, cl = col % NT;
        const int kb = ci >> 4, ch = (ci >> 3) & 1, j = ci & 7;
        __half hi, lo;
        tv_split(raw[i] * s_w, hi, lo);
        const size_t o = (((((size_t)(ct * (Cin / 16) + kb) * 2 + tap) * 2 + ch) * 2 + 0) * NT + cl) * 8 + j;
        img[o] = hi;
        img[o + (size_t)NT * 8] = lo;
    }
}

__global__ void tcv_absmax_kernel(const float *__restrict__ x, size_t n, unsigned int *out) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(x[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// u (C, L) -> uc [C / 8][L][8] f32 and img = split(silu(u) * kTcvXScale); also writes the image's zero rows
template <bool SILU>
__global__ void tcv_chunk_kernel(const float *__restrict__ u, int C, int L, float *__restrict__ uc, __half *__restrict__ img) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y, NCH = C / 8;
    const size_t Lp = tcv_image_rows(L);
    if (t < kTcvPadL) {
        const uint4 z = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(img + ((size_t)c * Lp + t) * 8) = z;
        *reinterpret_cast<uint4 *>(img + ((size_t)(NCH + c) * Lp + t) * 8) = z;
    }
    if (t >= L) return;
    float o[8];
    __half hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = u[(size_t)(c * 8 + j) * L + t];
    if (uc) {
        const size_t ro = ((size_t)c * L + t) * 8;
        *reinterpret_cast<float4 *>(uc + ro) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4 *>(uc + ro + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) tv_split((SILU ? silu_f(o[j]) : o[j]) * kTcvXScale, hi[j], lo[j]);
    const size_t io = ((size_t)c * Lp + kTcvPadL + t) * 8;
    *reinterpret_cast<uint4 *>(img + io) =
        make_uint4(tv_pack(hi[0], hi[1]), tv_pack(hi[2], hi[3]), tv_pack(hi[4], hi[5]), tv_pack(hi[6], hi[7]));
    *reinterpret_cast<uint4 *>(img + (size_t)NCH * Lp * 8 + io) =
        make_uint4(tv_pack(lo[0], lo[1]), tv_pack(lo[2], lo[3]), tv_pack(lo[4], lo[5]), tv_pack(lo[6], lo[7]));
}

constexpr size_t kTcvSmemMax = 200 * 1024;

size_t tcv_smem_bytes(int N, int MT, int K, int dil, int S, int tps) {
    const size_t rows_ld = 128 * (size_t)MT + (size_t)(K - 1) * dil;
    const size_t x_stage = (4 * rows_ld * 16 + 127) & ~(size_t)127;
    return 128 + kTcvBarBytes + 2 * x_stage + (size_t)S * tps * 64 * N;
}

int g_tcv_sms = 0;

}  // namespace

int tcv_init() {
    static bool done = false;
    if (done) return FSB_OK;
    FSB_CUDA_OK(cudaFuncSetAttribute(tcconv_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcvSmemMax));
    FSB_CUDA_OK(cudaFuncSetAttribute(tcconv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcvSmemMax));
    FSB_CUDA_OK(cudaFuncSetAttribute(tcconv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcvSmemMax));
    FSB_CUDA_OK(cudaFuncSetAttribute(tcconv_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcvSmemMax));
    int dev = 0;
    FSB_CUDA_OK(cudaGetDevice(&dev));
    FSB_CUDA_OK(cudaDeviceGetAttribute(&g_tcv_sms, cudaDevAttrMultiProcessorCount, dev));
    done = true;
    return FSB_OK;
}

static int tcv_weight_scale(const float *raw_dev, size_t n, cudaStream_t st, float *s_w) {
    unsigned int *d_max = nullptr;
    FSB_CUDA_OK(cudaMalloc(&d_max, sizeof(unsigned int)));
    FSB_CUDA_OK(cudaMemsetAsync(d_max, 0, sizeof(unsigned int), st));
    tcv_absmax_kernel<<<64, 256, 0, st>>>(raw_dev, n, d_max);
    unsigned int bits = 0;
    cudaError_t e = cudaMemcpyAsync(&bits, d_max, sizeof(bits), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_max);
    FSB_CUDA_OK(e);
    float mx;
    memcpy(&mx, &bits, sizeof(mx));
    // s_w = 2^k with max |w| * s_w in [2^13, 2^14): hi terms far from the fp16 limit, lo terms of typical weights normal
    int ex = 0;
    *s_w = 1.0f;
    if (mx > 0.f && std::isfinite(mx)) {
        frexpf(mx, &ex);  // mx = f * 2^ex, f in [0.5, 1)
        *s_w = ldexpf(1.0f, 14 - ex);
    }
    return FSB_OK;
}

int tcv_prepare_weights(const float *raw_dev, int C, int K, TcConvW *out, cudaStream_t st) {
    FSB_REQUIRE(tcv_supported(C), FSB_ERR_UNSUPPORTED, "tcconv: C=%d unsupported", C);
    const size_t n = (size_t)C * C * K;
    float s_w = 1.f;
    FSB_TRY(tcv_weight_scale(raw_dev, n, st, &s_w));
    __half *img = nullptr;
    FSB_CUDA_OK(cudaMalloc(&img, 2 * n * sizeof(__half)));
    const int NT = std::min(C, 128);
    tcv_weight_image_kernel<<<256, 256, 0, st>>>(raw_dev, img, C, K, NT, s_w);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaFree(img);
        set_error("tcconv weight image: %s", cudaGetErrorString(e));
        return FSB_ERR_CUDA;
    }
    out->img = img;
    out->inv_scale = 1.0f / (s_w * kTcvXScale);
    out->C = C;
    out->K = K;
    out->NT = NT;
    out->ncols = C;
    out->cout = C;
    out->ostride = 1;
    return FSB_OK;
}

int tcv_prepare_weights_t(const float *raw_dev, int Cin, int Cout, int stride, TcConvW *out, cudaStream_t st) {
    const int ncols = stride * Cout;
    FSB_REQUIRE(Cin % 16 == 0 && Cout % 8 == 0 && (ncols % 128 == 0 || ncols == 64 || ncols == 32 || ncols == 16), FSB_ERR_UNSUPPORTED,
                "tcconv (transposed): Cin=%d Cout=%d stride=%d unsupported", Cin, Cout, stride);
    const size_t n = (size_t)Cin * Cout * 2 * stride;
    float s_w = 1.f;
    FSB_TRY(tcv_weight_scale(raw_dev, n, st, &s_w));
    __half *img = nullptr;
    FSB_CUDA_OK(cudaMalloc(&img, 2 * n * sizeof(__half)));
    const int NT = std::min(ncols, 128);
    tcv_weight_image_t_kernel<<<256, 256, 0, st>>>(raw_dev, img, Cin, Cout, stride, NT, s_w);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaFree(img);
        set_error("tcconv weight image: %s", cudaGetErrorString(e));
        return FSB_ERR_CUDA;
    }
    out->img = img;
    out->inv_scale = 1.0f / (s_w * kTcvXScale);
    out->C = Cin;
    out->K = 2;
    out->NT = NT;
    out->ncols = ncols;
    out->cout = Cout;
    out->ostride = stride;
    return FSB_OK;
}

int tcv_chunk(const float *u, int C, int L, float *uc, __half *img, cudaStream_t st, bool silu) {
    const int n = std::max(L, kTcvPadL);
    if (silu) tcv_chunk_kernel<true><<<dim3((n + 127) / 128, C / 8), 128, 0, st>>>(u, C, L, uc, img);
    else tcv_chunk_kernel<false><<<dim3((n + 127) / 128, C / 8), 128, 0, st>>>(u, C, L, uc, img);
    FSB_CUDA_OK(cudaGetLastError());
    return FSB_OK;
}

int tcv_conv(const TcConvW &w, const float *bias, const __half *ximg, int L, int dil, const float *res, float *y, __half *yimg,
             float *m, int acc_mode, float scale, cudaStream_t st) {
    FSB_TRY(tcv_init());
    const int N = w.NT, K = w.K;
    TcConvArgs a;
    a.ximg = ximg; a.wimg = w.img; a.bias = bias; a.res = res; a.y = y; a.yimg = yimg; a.m = m;
    a.C = w.C; a.L = L; a.K = K; a.ncols = w.ncols; a.cout = w.cout; a.ostride = w.ostride; a.dil = dil; a.acc_mode = acc_mode; a.scale = scale; a.inv_scale = w.inv_scale;
    // MT tiles of 128 time steps per tile, two accumulator buffers: 2 * MT * N <= 512 TMEM columns.  Weight traffic per MMA
    // cycle falls as 1 / MT; the grid is persistent (one CTA per SM), so pick the MT that minimises rounds * MT, larger MT on
    // ties.  Small channel counts keep all taps of a 16-channel block in one weight stage.
    const int n128 = (L + 127) / 128, sms = g_tcv_sms > 0 ? g_tcv_sms : 148;
    const int nct = w.ncols / N;
    // weight stage = tps taps of one 16-channel block (<= 48 KB), taps split evenly over the stages of a block: the
    // per-stage handshake (mbarrier wait, commit) costs the MMA warp ~500 cycles, so few large stages beat many small ones
    const int tps_cap = std::max(1, (48 * 1024) / (64 * N));
    const int nst = (K + tps_cap - 1) / tps_cap;
    int tps = (K + nst - 1) / nst;
    if (const char *s = getenv("FSB_TCV_TPS")) tps = std::max(1, std::min(K, atoi(s)));
    const size_t budget = kTcvSmemMax;
    // per-tile time ~ (MT + 1) units (MT tiles of MMAs + the MT-independent weight staging): minimise rounds * (MT + 1)
    int MT = 1;
    long best = -1;
    const int aw = N <= 64 ? 2 * N : N;  // accumulator columns per M tile (stacked hi | lo products for N <= 64)
    for (int mt = std::min(8, 256 / aw); mt >= 1; --mt) {
        if (tcv_smem_bytes(N, mt, K, dil, 2, tps) > budget) continue;
        const long tiles = (long)((n128 + mt - 1) / mt) * nct;
        const long cost = ((tiles + sms - 1) / sms) * (mt + 1);
        if (best < 0 || cost < best) {
            best = cost;
            MT = mt;
        }
    }
    if (const char *s = getenv("FSB_TCV_MT")) MT = std::max(1, std::min(256 / aw, atoi(s)));
    a.MT = MT;
    a.tps = tps;
    int S = tps == 1 ? 16 : 4;
    while (S > 2 && tcv_smem_bytes(N, MT, K, dil, S, tps) > budget) --S;
    a.wstages = S;
    const size_t smem = tcv_smem_bytes(N, MT, K, dil, S, tps);
    FSB_REQUIRE(smem <= kTcvSmemMax, FSB_ERR_UNSUPPORTED, "tcconv tile needs %zu B of smem", smem);
    a.ntiles = ((n128 + MT - 1) / MT) * nct;
    a.rot = getenv("FSB_TCV_ROT") != nullptr;
    static long long *d_dbg = nullptr;
    const bool dbg = getenv("FSB_TCV_DBG") != nullptr;
    if (dbg && !d_dbg) FSB_CUDA_OK(cudaMalloc(&d_dbg, 8 * sizeof(long long)));
    a.dbg = dbg ? d_dbg : nullptr;
    const dim3 grid(std::min(a.ntiles, sms));
    if (N == 128) tcconv_kernel<128><<<grid, kTcvThreads, smem, st>>>(a);
    else if (N == 64) tcconv_kernel<64><<<grid, kTcvThreads, smem, st>>>(a);
    else if (N == 32) tcconv_kernel<32><<<grid, kTcvThreads, smem, st>>>(a);
    else tcconv_kernel<16><<<grid, kTcvThreads, smem, st>>>(a);
    if (dbg) {
        long long h[8];
        FSB_CUDA_OK(cudaStreamSynchronize(st));
        FSB_CUDA_OK(cudaMemcpy(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[tcconv N=%d C=%d K=%d dil=%d MT=%d S=%d tps=%d tiles=%d] mma warp: total %lld cyc, wait acc %lld x %lld w %lld, issue %lld, stages %lld\n",
                N, w.C, K, dil, MT, S, tps, a.ntiles, h[0], h[1], h[2], h[3], h[4], h[5]);
    }
    FSB_CUDA_OK(cudaGetLastError());
    return FSB_OK;
}

}  // namespace fsb
