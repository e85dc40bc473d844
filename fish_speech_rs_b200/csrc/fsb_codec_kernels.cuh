// Firefly-GAN-VQ codec kernels (fp32, batch 1 per launch).
// Activations are channel-major (C, L) like the reference's (1, C, L) tensors.
// Reference call sites are cited per kernel (paths relative to the reference
// root, fish_speech_core/lib/codec/).
#pragma once
#include "fsb_common.cuh"

namespace fsb {

// gelu() of Candle == tanh approximation (Q10): 0.5 x (1 + tanh( sqrt(2/pi) x (1 + 0.044715 x^2) ))
__device__ __forceinline__ float gelu_tanh_f(float x) {
    const float k = 0.7978845608028654f;
    return 0.5f * x * (1.0f + tanhf(k * x * (1.0f + 0.044715f * x * x)));
}

// ------------------------------------------------------------------ FSQ lookup + project_out
// GroupedResidualFSQ::get_output_from_indices, grouped_residual_fsq.rs:95-114,175-185;
// FSQ::indices_to_codes, fsq.rs:132-159.  codes (G, T) u32 -> z (G*64, T).
// levels (8,5,5,5), basis (1,8,40,200); err[0] set to 1 if a code >= 1000 (Q11).
__global__ void fsq_decode_kernel(const uint32_t *__restrict__ codes, int T, int G,
                                  const float *__restrict__ proj_w,  // (G, 64, 4)
                                  const float *__restrict__ proj_b,  // (G, 64)
                                  float *__restrict__ z, int *err) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (t >= T) return;
    const uint32_t idx = codes[(size_t)g * T + t];
    if (idx >= 1000u) {
        *err = 1;
        return;
    }
    // floor(idx / basis) mod levels, then (d - half) / half in f32 like the reference
    const float fi = (float)idx;
    const float lv[4] = {8.f, 5.f, 5.f, 5.f}, bs[4] = {1.f, 8.f, 40.f, 200.f};
    float c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float q = floorf(fi / bs[i]);
        q = q - floorf(q / lv[i]) * lv[i];
        const float hw = floorf(lv[i] / 2.0f);
        c[i] = (q - hw) / hw;
    }
    const float *w = proj_w + (size_t)g * 64 * 4;
    const float *b = proj_b + (size_t)g * 64;
    for (int j = 0; j < 64; ++j) {
        // Linear(4 -> 64): sequential f32 dot + bias
        float acc = c[0] * w[j * 4 + 0];
        acc = fmaf(c[1], w[j * 4 + 1], acc);
        acc = fmaf(c[2], w[j * 4 + 2], acc);
        acc = fmaf(c[3], w[j * 4 + 3], acc);
        z[(size_t)(g * 64 + j) * T + t] = acc + b[j];
    }
}

// ------------------------------------------------------------------ dense causal conv as implicit GEMM on FP32 FMA
// FishConvNet::forward (utils/mod.rs:53-63): left pad (K-1)*dil + 1 - stride, Conv1d.
// FishTransConvNet::forward (utils/mod.rs:110-122) runs through the same kernel, one
// launch per output phase r (= t mod stride): taps {r, r + s}, x index j - tap.
//   y[co, t*ostride + ooff] = epi( bias[co] + sum_ci sum_k Wt[(ci*Kw + k0 + k*kstep)*Cout + co] * pre(x[ci, t*stride + k*dil - pad]) )
// Wt is the weight re-laid out at load time as (Cin, Kw, Cout) so tiles load coalesced.
struct ConvArgs {
    const float *x;     // (Cin, Lin)
    const float *wt;    // (Cin, Kw, Cout)
    const float *bias;  // (Cout) or null
    const float *res;   // (Cout, Lres) residual added to the conv result, or null
    float *y;           // (Cout, Ly)
    int Cin, Cout, Lin, Lout;  // Lout = number of t positions this launch computes
    int Ly;                    // row stride of y (and res)
    int K;                     // taps this launch walks
    int Kw, k0, kstep;         // tap k reads weight slot k0 + k*kstep of Kw
    int dil, pad, stride;      // x index = t*stride + k*dil - pad   (dil may be negative)
    int ostride, ooff;         // output column = t*ostride + ooff
    int z_k0, z_ooff;          // per blockIdx.z increments of k0 / ooff (transposed conv: one z slice per output phase)
    int pre_silu;              // silu on the input (hifi_gan.rs:76-79,211-213)
    int acc_mode;              // 0: y = v;  1: y = y + v;  2: y = (y + v) * scale   (stack + mean, hifi_gan.rs:113-118)
    float scale;
    int post_tanh;             // hifi_gan.rs:215
    float *y_act;              // optional second output: silu(result), same layout as y (input of a TMA-staged ResBlock conv)
};

constexpr int kConvCK = 16;  // input channels staged per step (fewer load-sync-compute rounds: the global loads of a round are exposed)

// KT: taps as a compile-time constant (0 = run-time a.K): the tap loop is then fully unrolled, its address
// arithmetic (k * BM, k * dil) disappears and the loads of consecutive taps overlap.
template <int BM, int TM, int TN, int KT = 0>
__global__ void __launch_bounds__(256) conv1d_kernel(ConvArgs a) {
    constexpr int NTX = 32, BN = NTX * TN;  // 32 lanes along t (stride-32 interleave -> conflict-free, coalesced)
    constexpr int NTY = BM / TM;               // 8 warps along co
    static_assert(NTY * NTX == 256, "256 threads");
    extern __shared__ float smem[];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int t0 = blockIdx.x * BN, co0 = blockIdx.y * BM;
    // span of x indices needed for the tile: t in [t0, t0+BN), k in [0, K)
    const int off_lo = min(0, (a.K - 1) * a.dil) - a.pad;
    const int off_hi = max(0, (a.K - 1) * a.dil) - a.pad;
    const int x_lo = t0 * a.stride + off_lo;
    const int span = (BN - 1) * a.stride + off_hi - off_lo + 1;
    float *ws = smem;                       // [kConvCK][K][BM]
    float *xs = smem + kConvCK * a.K * BM;  // [kConvCK][span]
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int k0 = a.k0 + (int)blockIdx.z * a.z_k0, ooff = a.ooff + (int)blockIdx.z * a.z_ooff;
    const bool full_m = co0 + BM <= a.Cout && (a.Cout & 3) == 0;
    for (int c0 = 0; c0 < a.Cin; c0 += kConvCK) {
        const int nc = min(kConvCK, a.Cin - c0);
        // stage weights: rows (ci, k) of BM contiguous output channels; one division per 16-byte piece
        if (full_m) {
            constexpr int PR = BM / 4;  // float4 pieces per row
            for (int i = tid; i < nc * a.K * PR; i += 256) {
                const int row = i / PR, q = i - row * PR;
                const int ci = row / a.K, k = row - ci * a.K;
                const float4 w4 = *reinterpret_cast<const float4 *>(
                    a.wt + ((size_t)(c0 + ci) * a.Kw + k0 + k * a.kstep) * a.Cout + co0 + q * 4);
                *reinterpret_cast<float4 *>(ws + (size_t)row * BM + q * 4) = w4;
            }
        } else {
            for (int i = tid; i < nc * a.K * BM; i += 256) {
                const int co = i % BM, k = (i / BM) % a.K, ci = i / (BM * a.K);
                float w = 0.f;
                if (co0 + co < a.Cout) w = a.wt[((size_t)(c0 + ci) * a.Kw + k0 + k * a.kstep) * a.Cout + co0 + co];
                ws[i] = w;
            }
        }
        // stage x with the causal zero padding: one input channel per warp pass, no divisions
        for (int ci = tid >> 5; ci < nc; ci += 8) {
            const float *xg = a.x + (size_t)(c0 + ci) * a.Lin;
            float *xd = xs + (size_t)ci * span;
            for (int pp = tid & 31; pp < span; pp += 32) {
                const int xi = x_lo + pp;
                float v = 0.f;
                if (xi >= 0 && xi < a.Lin) {
                    v = xg[xi];
                    if (a.pre_silu) v = silu_f(v);
                }
                xd[pp] = v;
            }
        }
        __syncthreads();
        for (int ci = 0; ci < nc; ++ci) {
            const float *wrow = ws + (size_t)ci * a.K * BM + ty * TM;
            const float *xrow = xs + (size_t)ci * span - off_lo - a.pad;  // xrow[t_local*stride + k*dil]
            const int ntap = KT > 0 ? KT : a.K;
#pragma unroll
            for (int k = 0; k < ntap; ++k) {
                float w[TM], xv[TN];
                if (TM % 4 == 0) {  // the warp's TM output channels are contiguous and 16-byte aligned: broadcast LDS.128
#pragma unroll
                    for (int i = 0; i < TM; i += 4) {
                        const float4 w4 = *reinterpret_cast<const float4 *>(wrow + k * BM + i);
                        w[i] = w4.x; w[i + 1] = w4.y; w[i + 2] = w4.z; w[i + 3] = w4.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < TM; ++i) w[i] = wrow[k * BM + i];
                }
#pragma unroll
                for (int j = 0; j < TN; ++j) xv[j] = xrow[(tx + j * NTX) * a.stride + k * a.dil];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(w[i], xv[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int co = co0 + ty * TM + i;
        if (co >= a.Cout) continue;
        const float bv = a.bias ? a.bias[co] : 0.f;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int t = t0 + tx + j * NTX;
            if (t >= a.Lout) continue;
            const size_t o = (size_t)co * a.Ly + (size_t)t * a.ostride + ooff;
            float v = acc[i][j] + bv;
            if (a.res) v = a.res[o] + v;
            if (a.acc_mode == 1) v = a.y[o] + v;
            else if (a.acc_mode == 2) v = (a.y[o] + v) * a.scale;
            if (a.post_tanh) v = tanhf(v);
            a.y[o] = v;
            if (a.y_act) a.y_act[o] = silu_f(v);
        }
    }
}

// ------------------------------------------------------------------ ResBlock1 convs with TMA-staged tiles
// ResBlock1::forward (hifi_gan.rs:73-86): x += conv2_d(silu(conv1_d(silu(x)))), both convs causal with dilation d, C -> C
// channels.  These 90 convs are 95 % of the vocoder's MACs.  Same FP32 FMA implicit GEMM as conv1d_kernel (BM x 128
// output tile, TM x 4 per thread), but the tiles are moved by the TMA engine into a two-stage shared-memory ring while the
// previous chunk is multiplied:
//   * weights are re-laid out at load time as [co tile][ci chunk][ci][k][BM]: the (CK ci x K taps x BM co) block of a
//     chunk is ONE contiguous cp.async.bulk;
//   * the input tile (CK ci x span t) is ONE cp.async.bulk.tensor.2d on a (C, L) tensor map; columns left of t = 0 are
//     out of bounds and arrive as zeros -- the causal left padding of FishConvNet (utils/mod.rs:53-63) for free;
//   * the input is stored pre-activated by its producer (y_act of the previous conv), so nothing is computed at staging
//     time; the epilogue writes the raw result (residual stream / mean accumulator) and/or silu(result) for its consumer.
struct ResConvArgs {
    const float *wt;    // tiled weights [C / BM][C / CK][CK][K][BM]
    const float *bias;  // (C)
    const float *res;   // (C, L) residual (raw x), or null
    float *y;           // (C, L) raw result, or null (conv1: only silu(result) is ever read)
    float *y_act;       // (C, L) silu(result), or null
    int C, L;
    int acc_mode;       // 0: y = v;  1: y = y + v;  2: y = (y + v) * scale   (stack + mean, hifi_gan.rs:113-118)
    float scale;
};
constexpr int kResCK = 16;    // input channels per chunk
constexpr int kResBN = 128;   // time positions per tile

__device__ __forceinline__ uint32_t rc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// The TMA box must START on a 16-byte boundary (measured with tools/tma_probe.cu: a start column of -2 traps with
// "illegal instruction", -4 works), so the left halo (K - 1) * d is rounded up to a multiple of 4 columns.
__host__ __device__ constexpr int res_halo(int KT, int DIL) { return ((KT - 1) * DIL + 3) & ~3; }
__host__ __device__ constexpr int res_span(int KT, int DIL) { return kResBN + res_halo(KT, DIL); }
__host__ __device__ constexpr int res_stage_floats(int BM, int KT, int DIL) {
    return (kResCK * KT * BM + kResCK * res_span(KT, DIL) + 31) & ~31;  // 128-byte aligned stages
}

template <int BM, int KT, int DIL>
__global__ void __launch_bounds__(256) resconv_tma_kernel(const __grid_constant__ CUtensorMap tmx, ResConvArgs a) {
    constexpr int TM = BM / 8, TN = 4, NTX = 32;
    constexpr int SPAN = res_span(KT, DIL);
    constexpr int W_FLOATS = kResCK * KT * BM, X_FLOATS = kResCK * SPAN, STAGE = res_stage_floats(BM, KT, DIL);
    extern __shared__ float rc_smem_raw[];
    // TMA destinations need 128-byte alignment (dynamic shared memory only guarantees 16)
    float *rc_smem = rc_smem_raw + (((128u - (rc_smem_u32(rc_smem_raw) & 127u)) & 127u) >> 2);
    __shared__ __align__(8) unsigned long long full[2];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int t0 = blockIdx.x * kResBN, co0 = blockIdx.y * BM;
    const int nchunk = a.C / kResCK;
    const float *wsrc = a.wt + (size_t)blockIdx.y * nchunk * W_FLOATS;
    const CUtensorMap *pmap = &tmx;  // address of the __grid_constant__ parameter itself (never a local copy)
    const uint32_t bar0 = rc_smem_u32(&full[0]);
    const int xcol0 = t0 - res_halo(KT, DIL);
    auto issue = [=](int c) {  // chunk c -> stage c & 1 (one thread)
        float *st = rc_smem + (c & 1) * STAGE;
        const uint32_t bar = bar0 + (c & 1) * 8;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((W_FLOATS + X_FLOATS) * 4) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rc_smem_u32(st)),
                     "l"(wsrc + (size_t)c * W_FLOATS), "r"(W_FLOATS * 4), "r"(bar)
                     : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                         rc_smem_u32(st + W_FLOATS)),
                     "l"(pmap), "r"(bar), "r"(xcol0), "r"(c * kResCK)
                     : "memory");
    };
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rc_smem_u32(&full[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rc_smem_u32(&full[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        issue(0);
    }
    __syncthreads();
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int c = 0; c < nchunk; ++c) {
        if (tid == 0 && c + 1 < nchunk) issue(c + 1);  // stage (c + 1) & 1 was released by the barrier below
        {
            const uint32_t bar = rc_smem_u32(&full[c & 1]), parity = (uint32_t)(c >> 1) & 1u;
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "WAIT_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra DONE_%=;\n\t"
                "bra WAIT_%=;\n\t"
                "DONE_%=:\n\t"
                "}\n" ::"r"(bar),
                "r"(parity)
                : "memory");
        }
        const float *ws = rc_smem + (c & 1) * STAGE, *xs = ws + W_FLOATS;
#pragma unroll 2
        for (int ci = 0; ci < kResCK; ++ci) {
            const float *wrow = ws + ci * KT * BM + ty * TM;
            // xrow[j * 32 + k * DIL]: tile column 0 == t0 - halo (rounded up), tap 0 of output t reads column t - (KT - 1) * DIL
            const float *xrow = xs + ci * SPAN + tx + (res_halo(KT, DIL) - (KT - 1) * DIL);
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                float w[TM], xv[TN];
                if (TM % 4 == 0) {
#pragma unroll
                    for (int i = 0; i < TM; i += 4) {
                        const float4 w4 = *reinterpret_cast<const float4 *>(wrow + k * BM + i);
                        w[i] = w4.x; w[i + 1] = w4.y; w[i + 2] = w4.z; w[i + 3] = w4.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < TM; ++i) w[i] = wrow[k * BM + i];
                }
#pragma unroll
                for (int j = 0; j < TN; ++j) xv[j] = xrow[j * NTX + k * DIL];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(w[i], xv[j], acc[i][j]);
            }
        }
        __syncthreads();  // everybody is done with stage c & 1: it may be refilled with chunk c + 2
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int co = co0 + ty * TM + i;
        const float bv = a.bias[co];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int t = t0 + tx + j * NTX;
            if (t >= a.L) continue;
            const size_t o = (size_t)co * a.L + t;
            float v = acc[i][j] + bv;
            if (a.res) v = a.res[o] + v;
            if (a.acc_mode == 1) v = a.y[o] + v;
            else if (a.acc_mode == 2) v = (a.y[o] + v) * a.scale;
            if (a.y) a.y[o] = v;
            if (a.y_act) a.y_act[o] = silu_f(v);
        }
    }
}

// ------------------------------------------------------------------ conv_post (hifi_gan.rs:213-215): C -> 1, K taps, causal
//   y[t] = tanh( bias + sum_ci sum_k W[ci, k] * silu(x[ci, t + k - (K - 1)]) )
// One output channel: an implicit-GEMM tile would idle 15 of 16 rows, so every thread owns outputs instead (kPostTT per
// block), the silu'd input tile and the C x K weights sit in shared memory.
constexpr int kPostTT = 512, kPostThreads = 256, kPostMaxCK = 16 * 13;
__global__ void __launch_bounds__(kPostThreads) conv_post_kernel(const float *__restrict__ x, const float *__restrict__ wt /* (C, K, 1) */,
                                                                 const float *__restrict__ bias, float *__restrict__ y, int C, int K,
                                                                 int L) {
    extern __shared__ float cp_smem[];
    const int span = kPostTT + K - 1;
    float *xs = cp_smem;             // [C][span]
    float *ws = cp_smem + C * span;  // [C * K]
    const int t0 = blockIdx.x * kPostTT;
    for (int i = threadIdx.x; i < C * K; i += kPostThreads) ws[i] = wt[i];
    for (int i = threadIdx.x; i < C * span; i += kPostThreads) {
        const int ci = i / span, j = i - ci * span, t = t0 + j - (K - 1);
        xs[i] = (t >= 0 && t < L) ? silu_f(x[(size_t)ci * L + t]) : 0.f;
    }
    __syncthreads();
    const float bv = bias[0];
#pragma unroll
    for (int r = 0; r < kPostTT / kPostThreads; ++r) {
        const int j = threadIdx.x + r * kPostThreads, t = t0 + j;
        if (t >= L) continue;
        float acc = 0.f;
        for (int ci = 0; ci < C; ++ci) {
            const float *xr = xs + ci * span + j, *wr = ws + ci * K;
            for (int k = 0; k < K; ++k) acc = fmaf(wr[k], xr[k], acc);
        }
        y[t] = tanhf(acc + bv);
    }
}

// ------------------------------------------------------------------ ConvNeXt block pieces (convnext.rs:109-127)
// depthwise causal conv k=7 (pad 6) + LayerNorm over channels (biased variance, eps) -> h (L, C) time-major.
// One warp per time step.
__global__ void dwconv_ln_kernel(const float *__restrict__ x, int C, int L, const float *__restrict__ dw_w,  // (C, 7)
                                 const float *__restrict__ dw_b, const float *__restrict__ ln_w,
                                 const float *__restrict__ ln_b, float eps, float *__restrict__ h) {
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= L) return;
    extern __shared__ float sm[];
    float *row = sm + (size_t)(threadIdx.x >> 5) * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int xi = t + k - 6;
            if (xi >= 0) acc = fmaf(dw_w[c * 7 + k], x[(size_t)c * L + xi], acc);
        }
        acc += dw_b[c];
        row[c] = acc;
        s += acc;
    }
    s = warp_sum(s);
    const float mean = s / (float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float d = row[c] - mean;
        v = fmaf(d, d, v);
    }
    v = warp_sum(v);
    const float rstd = 1.0f / sqrtf(v / (float)C + eps);
    for (int c = lane; c < C; c += 32) h[(size_t)t * C + c] = (row[c] - mean) * rstd * ln_w[c] + ln_b[c];
}

// C[M, N] = A[M, K] . W[N, K]^T + bias, epilogues for the two pointwise convs:
//   EPI 0: gelu_tanh(v)                               -> Cm (M, N) row-major
//   EPI 1: res[n, m] + gamma[n] * v, stored transposed -> Cm (N, M)  (back to channel-major + residual)
template <int EPI>
__global__ void __launch_bounds__(256) pw_gemm_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                      const float *__restrict__ bias, const float *__restrict__ gamma,
                                                      const float *res, float *Cm, int M,
                                                      int N, int K) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = tid % 16, ty = tid / 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid / 4, lk = (tid % 4) * 4;
    for (int k0 = 0; k0 < K; k0 += BK) {
        float v[4] = {0.f, 0.f, 0.f, 0.f}, w[4] = {0.f, 0.f, 0.f, 0.f};
        if (m0 + lr < M) {
            const float4 t = *reinterpret_cast<const float4 *>(A + (size_t)(m0 + lr) * K + k0 + lk);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        }
        if (n0 + lr < N) {
            const float4 t = *reinterpret_cast<const float4 *>(W + (size_t)(n0 + lr) * K + k0 + lk);
            w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            As[lk + i][lr] = v[i];
            Ws[lk + i][lr] = w[i];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + bias[n];
            if (EPI == 0) {
                Cm[(size_t)m * N + n] = gelu_tanh_f(v);
            } else {
                if (gamma) v = gamma[n] * v;
                Cm[(size_t)n * M + m] = res[(size_t)n * M + m] + v;
            }
        }
    }
}

// ------------------------------------------------------------------ encoder-only pieces
// LayerNormChannelsFirst (convnext.rs:144-154): normalise over channels at each t.  One warp per t.
__global__ void ln_channels_first_kernel(const float *__restrict__ x, int C, int L, const float *__restrict__ w,
                                         const float *__restrict__ b, float eps, float *__restrict__ y) {
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= L) return;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += x[(size_t)c * L + t];
    s = warp_sum(s);
    const float u = s / (float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float d = x[(size_t)c * L + t] - u;
        v = fmaf(d, d, v);
    }
    v = warp_sum(v);
    const float den = sqrtf(v / (float)C + eps);
    for (int c = lane; c < C; c += 32) y[(size_t)c * L + t] = (x[(size_t)c * L + t] - u) / den * w[c] + b[c];
}

// FSQ encode of one group at one position (grouped_residual_fsq.rs:75-93, fsq.rs:68-130):
// project_in (64 -> 4), bound twice, round half away from zero, index = sum digit * basis.
// z (512, L) channel-major -> idx (G, L) int64.
__global__ void fsq_encode_kernel(const float *__restrict__ z, int L, int G, const float *__restrict__ pin_w,  // (G,4,64)
                                  const float *__restrict__ pin_b,                                               // (G,4)
                                  long long *__restrict__ idx) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (t >= L) return;
    const float lv[4] = {8.f, 5.f, 5.f, 5.f}, bs[4] = {1.f, 8.f, 40.f, 200.f};
    float out = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float acc = 0.f;
        for (int d = 0; d < 64; ++d) acc = fmaf(z[(size_t)(g * 64 + d) * L + t], pin_w[((size_t)g * 4 + i) * 64 + d], acc);
        acc += pin_b[g * 4 + i];
        const float half_l = (lv[i] - 1.0f) * 1.001f / 2.0f;
        const float offset = (lv[i] - floorf(lv[i] / 2.0f) * 2.0f == 0.f) ? 0.5f : 0.f;
        const float r = offset / half_l;
        const float shift = logf((1.0f + r) / (1.0f - r)) * 0.5f;  // atanh
        float v = tanhf(acc + shift) * half_l - offset;             // implicit first step
        v = tanhf(v + shift) * half_l - offset;                     // FSQ::forward bounds again
        const float q = copysignf(floorf(fabsf(v) + 0.5f), v);      // round half away from zero
        const float hw = floorf(lv[i] / 2.0f);
        const float zhat = (q / hw) * hw + hw;
        out += zhat * bs[i];
    }
    idx[(size_t)g * L + t] = (long long)out;
}

// ------------------------------------------------------------------ log-mel front-end (voice-clone path)
// LogMelSpectrogram::forward, audio/spectrogram.rs:29-83,141-158 on the streaming STFT of audio/stft.rs:52-90:
// reflect pad 768 (edge sample repeated), frame f = padded[512 f, 512 f + 2048) x periodic Hann, 2048-point DFT in
// f64 (the reference: rustfft in f64), magnitude of the first 1025 bins -> f32, + 1e-6, x mel table, clamp, ln.
constexpr int kFft = 2048, kHop = 512, kBins = 1025, kMels = 160;

// tw[m] = cos(2 pi m / 2048), tw[2048 + m] = sin(2 pi m / 2048), correctly rounded (cospi / sinpi)
__global__ void stft_twiddle_kernel(double *tw) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < kFft) {
        tw[m] = cospi((double)m / 1024.0);
        tw[kFft + m] = sinpi((double)m / 1024.0);
    }
}

__global__ void reflect_pad_kernel(const float *__restrict__ x, int N, int pad, float *__restrict__ xp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N + 2 * pad) return;
    int j;
    if (i < pad) j = pad - 1 - i;                 // spectrogram.rs:18: signal[0..pad] reversed
    else if (i < pad + N) j = i - pad;
    else j = N - 1 - (i - pad - N);               // :24: signal[N-pad..] reversed
    xp[i] = x[j];
}

// one CTA per frame; direct DFT with a shared twiddle table: bin k = sum_n xw[n] * (cos, -sin)(2 pi k n / 2048).
// 2.1 M double FMAs per frame -- a few ms for a 13 s clip; exact to ~1e-13 like an f64 FFT.
__global__ void __launch_bounds__(256) stft_mag_kernel(const float *__restrict__ xp, int Lp, const double *__restrict__ tw,
                                                       float *__restrict__ mag) {
    extern __shared__ double stft_sm[];
    double *xw = stft_sm, *cs = stft_sm + kFft, *sn = cs + kFft;
    const int f = blockIdx.x;
    for (int n = threadIdx.x; n < kFft; n += 256) {
        const int i = f * kHop + n;
        const double c = tw[n], s = tw[kFft + n];
        cs[n] = c;
        sn[n] = s;
        // periodic Hann (stft.rs:33-35); a final partial chunk is zero padded (:61-64)
        xw[n] = i < Lp ? (double)xp[i] * (0.5 * (1.0 - c)) : 0.0;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kBins; k += 256) {
        double re = 0.0, im = 0.0;
        int idx = 0;
        for (int n = 0; n < kFft; ++n) {
            re = fma(xw[n], cs[idx], re);
            im = fma(xw[n], sn[idx], im);
            idx = (idx + k) & (kFft - 1);
        }
        mag[(size_t)f * kBins + k] = (float)sqrt(re * re + im * im) + 1e-6f;  // spectrogram.rs:10,82
    }
}

// mel[m, f] = ln(clamp(sum_k mag[f, k] * fb[k, m], 1e-5, 100)); one CTA per frame, thread = mel bin
__global__ void __launch_bounds__(kMels) mel_log_kernel(const float *__restrict__ mag, const float *__restrict__ fb,
                                                        int nframes, float *__restrict__ mel) {
    __shared__ float row[kBins];
    const int f = blockIdx.x, m = threadIdx.x;
    for (int k = m; k < kBins; k += kMels) row[k] = mag[(size_t)f * kBins + k];
    __syncthreads();
    float acc = 0.f;
    for (int k = 0; k < kBins; ++k) acc = fmaf(row[k], fb[k * kMels + m], acc);
    mel[(size_t)m * nframes + f] = logf(fminf(fmaxf(acc, 1e-5f), 100.0f));
}

// ------------------------------------------------------------------ output stage (streaming path)
// audio/functional.rs:3-37 `resample`: linear interpolation, index arithmetic in f64, weights in f32, two products
// and one sum (no contraction: bit-exact with the reference's separate mul / add tensors).
__global__ void resample_linear_kernel(const float *__restrict__ x, long long n, double ratio, long long n_out,
                                       float *__restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const double idx = (double)i / ratio;
    const double fl = floor(idx);
    const long long i0 = (long long)fl;
    long long i1 = (long long)ceil(idx);
    if (i1 > n - 1) i1 = n - 1;
    const float t = (float)(idx - fl);
    const float omt = __fsub_rn(1.0f, t);
    y[i] = __fadd_rn(__fmul_rn(x[i0], omt), __fmul_rn(x[i1], t));
}

// audio/wav.rs:9-13 `Sample for f32`: (x.clamp(-1, 1) * 32767.0) as i16 (truncation toward zero; NaN -> 0)
__global__ void pcm_f32_to_s16_kernel(const float *__restrict__ x, long long n, short *__restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = x[i];
    v = v != v ? 0.f : fminf(fmaxf(v, -1.0f), 1.0f);
    y[i] = (short)(int)__fmul_rn(v, 32767.0f);
}

}  // namespace fsb
