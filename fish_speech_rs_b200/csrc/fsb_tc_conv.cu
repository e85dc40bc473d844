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

// N = couts per CTA (the MMA N); a.C = channels (cins = all couts); blockIdx.y = cout tile
template <int N>
__global__ void __launch_bounds__(kTcvThreads, 2) tcconv_kernel(const TcConvArgs a) {
    extern __shared__ uint8_t tv_smem_raw[];
    uint8_t *smem = tv_smem_raw + ((128u - (tv_smem_u32(tv_smem_raw) & 127u)) & 127u);
    const int MT = a.MT, K = a.K, dil = a.dil, S = a.wstages;
    const int halo = (K - 1) * dil;
    const int rows_ld = 128 * MT + halo;                  // staged rows per (term, chunk)
    const uint32_t x_plane = (uint32_t)rows_ld * 16u;     // bytes of one (term, chunk) plane
    const uint32_t x_stage = 4u * x_plane;
    constexpr uint32_t w_plane = (uint32_t)N * 16u, w_tap = 4u * w_plane;   // one (kb, tap): [2 terms][2 chunks][N][16 B]
    const int tps = a.tps;                                                   // taps per weight stage (1, or K for small N)
    const uint32_t w_stage = (uint32_t)tps * w_tap;
    const int NCH = a.C / 8, NKB = a.C / 16, co0 = blockIdx.y * N;
    uint64_t *xfull = reinterpret_cast<uint64_t *>(smem);  // [2]
    uint64_t *xempty = xfull + 2;                          // [2]
    uint64_t *accfull = xempty + 2;
    uint64_t *wfull = accfull + 1;                         // [S]
    uint64_t *wempty = wfull + S;                          // [S]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(wempty + S);
    const uint32_t x_stage_al = (x_stage + 127u) & ~127u;
    uint8_t *xs = smem + kTcvBarBytes;
    uint8_t *ws = xs + 2 * (size_t)x_stage_al;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * 128 * MT;
    const size_t Lp = tcv_image_rows(a.L);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            tv_mbar_init(xfull + i, 1);
            tv_mbar_init(xempty + i, 1);
        }
        tv_mbar_init(accfull, 1);
        for (int i = 0; i < S; ++i) {
            tv_mbar_init(wfull + i, 1);
            tv_mbar_init(wempty + i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // accumulator columns: MT tiles of N fp32 columns, allocation rounded up to a power of two
    uint32_t ncols = 32;
    while ((int)ncols < MT * N) ncols <<= 1;
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
            int wi = 0;
            uint32_t wph = 0;
            const __half *xbase = a.ximg + ((size_t)kTcvPadL + t0 - halo) * 8;
            const __half *wsrc = a.wimg + (size_t)blockIdx.y * NKB * K * (w_tap / 2);
            for (int kb = 0; kb < NKB; ++kb) {
                const int xsl = kb & 1;
                tv_wait(xempty + xsl, ((kb >> 1) & 1) ^ 1);
                tv_expect_tx(xfull + xsl, x_stage);
                uint8_t *xd = xs + (size_t)xsl * x_stage_al;
#pragma unroll
                for (int term = 0; term < 2; ++term)
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch)
                        tv_bulk(xd + (term * 2 + ch) * x_plane, xbase + ((size_t)(term * NCH + 2 * kb + ch) * Lp) * 8, x_plane,
                                xfull + xsl);
                for (int tap = 0; tap < K; tap += tps) {
                    tv_wait(wempty + wi, wph ^ 1);
                    tv_expect_tx(wfull + wi, w_stage);
                    tv_bulk(ws + (size_t)wi * w_stage, wsrc + (size_t)(kb * K + tap) * (w_tap / 2), w_stage, wfull + wi);
                    if (++wi == S) {
                        wi = 0;
                        wph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ---- MMA issuer (whole warp converged, one elected lane issues)
        // D[t, co] += X_term[t + tap * dil, 16 ci] * W_term[co, 16 ci]:  hi*hi, hi*lo, lo*hi
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // f16 x f16 -> f32
        const uint64_t a_hi = tv_desc_hi(x_plane), b_hi = tv_desc_hi(w_plane);
        int wi = 0;
        uint32_t wph = 0;
        for (int kb = 0; kb < NKB; ++kb) {
            const int xsl = kb & 1;
            tv_wait(xfull + xsl, (kb >> 1) & 1);
            const uint32_t xaddr = tv_smem_u32(xs + (size_t)xsl * x_stage_al);
            for (int tap0 = 0; tap0 < K; tap0 += tps) {
                tv_wait(wfull + wi, wph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tv_elect()) {
                    uint32_t waddr = tv_smem_u32(ws + (size_t)wi * w_stage);
                    uint32_t xrow = xaddr + (uint32_t)(tap0 * dil) * 16u;
                    for (int tp = 0; tp < tps; ++tp, waddr += w_tap, xrow += (uint32_t)dil * 16u) {
                        const uint64_t b0 = b_hi | (uint64_t)((waddr >> 4) & 0x3FFF);
                        const uint64_t b1 = b_hi | (uint64_t)(((waddr + 2 * w_plane) >> 4) & 0x3FFF);
                        const uint32_t first = (kb | tap0 | tp) != 0 ? 1u : 0u;
                        for (int mt = 0; mt < MT; ++mt) {
                            const uint32_t xa = xrow + (uint32_t)mt * 2048u;
                            const uint64_t a0 = a_hi | (uint64_t)((xa >> 4) & 0x3FFF);
                            const uint64_t a1 = a_hi | (uint64_t)(((xa + 2 * x_plane) >> 4) & 0x3FFF);
                            const uint32_t d = tmem_base + (uint32_t)(mt * N);
                            tv_mma(d, a0, b0, idesc, first);
                            tv_mma(d, a0, b1, idesc, 1u);
                            tv_mma(d, a1, b0, idesc, 1u);
                        }
                    }
                    tv_commit(wempty + wi);
                    if (tap0 + tps >= K) tv_commit(xempty + xsl);
                    if (tap0 + tps >= K && kb == NKB - 1) tv_commit(accfull);
                }
                __syncwarp();
                if (++wi == S) {
                    wi = 0;
                    wph ^= 1;
                }
            }
        }
    } else {
        // ---- epilogue: warp w reads TMEM lanes [32 (w & 3), +32) (= time rows); the (tile, column block) pairs are dealt
        // alternately to the warp groups 0..3 and 4..7
        const int q = warp & 3, grp = warp >> 2;
        const int row = q * 32 + lane;
        constexpr int CB = N < 32 ? N : 32, NB = N / CB;  // column block
        if (a.yimg && t0 == 0 && row < kTcvPadL && grp == 0) {
            // the consumer's causal left padding
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (int c = co0 / 8; c < (co0 + N) / 8; ++c)
#pragma unroll
                for (int term = 0; term < 2; ++term)
                    *reinterpret_cast<uint4 *>(a.yimg + ((size_t)(term * NCH + c) * Lp + row) * 8) = z;
        }
        tv_wait(accfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float inv_scale = a.inv_scale;
#pragma unroll 1
        for (int blk = grp; blk < MT * NB; blk += 2) {
            const int mt = blk / NB, cb = (blk % NB) * CB;
            const int t = t0 + mt * 128 + row;
            uint32_t v[CB];
            __syncwarp();
            tv_ld<CB>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * N + cb), v);
            if (t < a.L) {
#pragma unroll
                for (int g = 0; g < CB / 8; ++g) {
                    const int co = co0 + cb + g * 8, c = co >> 3;
                    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(a.bias + co));
                    const float4 b1 = __ldg(reinterpret_cast<const float4 *>(a.bias + co + 4));
                    float o[8];
                    o[0] = fmaf(__uint_as_float(v[g * 8 + 0]), inv_scale, b0.x);
                    o[1] = fmaf(__uint_as_float(v[g * 8 + 1]), inv_scale, b0.y);
                    o[2] = fmaf(__uint_as_float(v[g * 8 + 2]), inv_scale, b0.z);
                    o[3] = fmaf(__uint_as_float(v[g * 8 + 3]), inv_scale, b0.w);
                    o[4] = fmaf(__uint_as_float(v[g * 8 + 4]), inv_scale, b1.x);
                    o[5] = fmaf(__uint_as_float(v[g * 8 + 5]), inv_scale, b1.y);
                    o[6] = fmaf(__uint_as_float(v[g * 8 + 6]), inv_scale, b1.z);
                    o[7] = fmaf(__uint_as_float(v[g * 8 + 7]), inv_scale, b1.w);
                    const size_t ro = ((size_t)c * a.L + t) * 8;
                    if (a.res) {
                        const float4 r0 = *reinterpret_cast<const float4 *>(a.res + ro);
                        const float4 r1 = *reinterpret_cast<const float4 *>(a.res + ro + 4);
                        o[0] = r0.x + o[0]; o[1] = r0.y + o[1]; o[2] = r0.z + o[2]; o[3] = r0.w + o[3];
                        o[4] = r1.x + o[4]; o[5] = r1.y + o[5]; o[6] = r1.z + o[6]; o[7] = r1.w + o[7];
                    }
                    if (a.y) {
                        *reinterpret_cast<float4 *>(a.y + ro) = make_float4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<float4 *>(a.y + ro + 4) = make_float4(o[4], o[5], o[6], o[7]);
                    }
                    if (a.yimg) {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            tv_split2(tv_silu_fast(o[2 * j]) * kTcvXScale, tv_silu_fast(o[2 * j + 1]) * kTcvXScale, hi[j], lo[j]);
                        const size_t io = ((size_t)c * Lp + kTcvPadL + t) * 8;
                        *reinterpret_cast<uint4 *>(a.yimg + io) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4 *>(a.yimg + (size_t)NCH * Lp * 8 + io) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                    if (a.m) {
                        // all loads first: the stores below may alias them as far as the compiler knows
                        float *mp = a.m + (size_t)co * a.L + t;
                        float pm[8];
                        if (a.acc_mode != 0) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) pm[j] = mp[(size_t)j * a.L];
#pragma unroll
                            for (int j = 0; j < 8; ++j) o[j] = a.acc_mode == 2 ? (pm[j] + o[j]) * a.scale : pm[j] + o[j];
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) mp[(size_t)j * a.L] = o[j];
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 9) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// (Cout, Cin, K) f32 -> [co tile][kb][tap][term][chunk][NT co][8 ci] fp16, scaled
__global__ void tcv_weight_image_kernel(const float *__restrict__ raw, __half *__restrict__ img, int C, int K, int NT, float s_w) {
    const size_t n = (size_t)C * C * K;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int tap = (int)(i % K), ci = (int)((i / K) % C), co = (int)(i / ((size_t)C * K));
        const int kb = ci >> 4, ch = (ci >> 3) & 1, j = ci & 7;
        __half hi, lo;
        tv_split(raw[i] * s_w, hi, lo);
        const int ct = co / NT, col = co % NT;
        const size_t o = (((((size_t)(ct * (C / 16) + kb) * K + tap) * 2 + 0) * 2 + ch) * NT + col) * 8 + j;
        img[o] = hi;
        img[o + (size_t)2 * NT * 8] = lo;
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
    const size_t ro = ((size_t)c * L + t) * 8;
    *reinterpret_cast<float4 *>(uc + ro) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4 *>(uc + ro + 4) = make_float4(o[4], o[5], o[6], o[7]);
#pragma unroll
    for (int j = 0; j < 8; ++j) tv_split(silu_f(o[j]) * kTcvXScale, hi[j], lo[j]);
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

int tcv_prepare_weights(const float *raw_dev, int C, int K, TcConvW *out, cudaStream_t st) {
    FSB_REQUIRE(tcv_supported(C), FSB_ERR_UNSUPPORTED, "tcconv: C=%d unsupported", C);
    const size_t n = (size_t)C * C * K;
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
    float s_w = 1.0f;
    if (mx > 0.f && std::isfinite(mx)) {
        frexpf(mx, &ex);  // mx = f * 2^ex, f in [0.5, 1)
        s_w = ldexpf(1.0f, 14 - ex);
    }
    __half *img = nullptr;
    FSB_CUDA_OK(cudaMalloc(&img, 2 * n * sizeof(__half)));
    const int NT = std::min(C, 128);
    tcv_weight_image_kernel<<<256, 256, 0, st>>>(raw_dev, img, C, K, NT, s_w);
    e = cudaGetLastError();
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
    return FSB_OK;
}

int tcv_chunk(const float *u, int C, int L, float *uc, __half *img, cudaStream_t st) {
    const int n = std::max(L, kTcvPadL);
    tcv_chunk_kernel<<<dim3((n + 127) / 128, C / 8), 128, 0, st>>>(u, C, L, uc, img);
    FSB_CUDA_OK(cudaGetLastError());
    return FSB_OK;
}

int tcv_conv(const TcConvW &w, const float *bias, const __half *ximg, int L, int dil, const float *res, float *y, __half *yimg,
             float *m, int acc_mode, float scale, cudaStream_t st) {
    FSB_TRY(tcv_init());
    const int N = w.NT, K = w.K;
    TcConvArgs a;
    a.ximg = ximg; a.wimg = w.img; a.bias = bias; a.res = res; a.y = y; a.yimg = yimg; a.m = m;
    a.C = w.C; a.L = L; a.K = K; a.dil = dil; a.acc_mode = acc_mode; a.scale = scale; a.inv_scale = w.inv_scale;
    // MT tiles of 128 time steps per CTA (MT * N TMEM columns).  Weight traffic per MMA cycle falls as 1 / MT; MT * N <= 256
    // and <= 110 KB of smem let two CTAs share an SM (one drains its accumulators while the other issues MMAs).
    // Small channel counts keep all taps of a 16-channel block in one weight stage.
    const int n128 = (L + 127) / 128, sms = g_tcv_sms > 0 ? g_tcv_sms : 148;
    const int tps = N <= 32 ? K : 1;
    const int nkb = w.C / 16;
    const size_t budget = 110 * 1024;
    int MT = std::min(4, 256 / N);
    while (MT > 1 && (((n128 + MT - 1) / MT) * (w.C / N) < 2 * sms || tcv_smem_bytes(N, MT, K, dil, 2, tps) > budget)) --MT;
    if (const char *s = getenv("FSB_TCV_MT")) MT = std::max(1, std::min(std::min(4, 512 / N), atoi(s)));
    a.MT = MT;
    a.tps = tps;
    int S = tps == 1 ? 16 : std::min(4, nkb);
    while (S > 2 && tcv_smem_bytes(N, MT, K, dil, S, tps) > budget) --S;
    a.wstages = S;
    const size_t smem = tcv_smem_bytes(N, MT, K, dil, S, tps);
    FSB_REQUIRE(smem <= kTcvSmemMax, FSB_ERR_UNSUPPORTED, "tcconv tile needs %zu B of smem", smem);
    const dim3 grid((n128 + MT - 1) / MT, w.C / N);
    if (N == 128) tcconv_kernel<128><<<grid, kTcvThreads, smem, st>>>(a);
    else if (N == 64) tcconv_kernel<64><<<grid, kTcvThreads, smem, st>>>(a);
    else if (N == 32) tcconv_kernel<32><<<grid, kTcvThreads, smem, st>>>(a);
    else tcconv_kernel<16><<<grid, kTcvThreads, smem, st>>>(a);
    FSB_CUDA_OK(cudaGetLastError());
    return FSB_OK;
}

}  // namespace fsb
