// tcgen05 / TMEM implicit-GEMM for the HiFi-GAN ResBlock convolutions (fsb_tc_conv.cu).
#pragma once
#include <cuda_fp16.h>

#include "fsb_common.cuh"

namespace fsb {

constexpr int kTcvPadL = 64;          // zero rows in front of every image chunk: the causal left padding, >= (K-1)*dil = 50
constexpr int kTcvTailRows = 640;     // rows past round_up(L, 128) a tile may read (never contribute to a stored output)
constexpr float kTcvXScale = 16.0f;   // activations are multiplied by this before the fp16 split (range 4094, 2^-22 rel.)

// rows of one (term, chunk) plane of an activation image for L time steps
__host__ __device__ inline size_t tcv_image_rows(int L) { return (size_t)kTcvPadL + (size_t)((L + 127) / 128) * 128 + kTcvTailRows; }
// __half elements of the image of a (C, L) activation: [2 terms][C / 8 chunks][rows][8 channels]
inline size_t tcv_image_halves(int C, int L) { return (size_t)2 * C * tcv_image_rows(L); }

struct TcConvW {
    __half *img = nullptr;   // [C / NT cout tiles][C / 16 kb][K taps][2 terms][2 chunks][NT couts][8 cins], scaled by s_w (a power of two)
    float inv_scale = 0.f;   // 1 / (s_w * kTcvXScale)
    int C = 0, K = 0;
    int NT = 0;              // GEMM columns per tile (= MMA N): min(ncols, 128)
    int ncols = 0, cout = 0, ostride = 1;  // plain conv: ncols = cout = C; transposed conv (stride s): ncols = s * cout, K = 2
};

struct TcConvArgs {
    const __half *ximg;   // input image: silu(x) * kTcvXScale split into hi + lo fp16
    const __half *wimg;
    const float *bias;    // (C)
    const float *res;     // chunked f32 residual [C / 8][L][8], or null
    float *y;             // chunked f32 raw result, or null (may alias res)
    __half *yimg;         // image of silu(result) for the next conv, or null
    float *m;             // (C, L) channel-major mean accumulator, or null
    int C, L, K, dil, MT, wstages, tps, ntiles, rot;   // C = input channels, L = input time steps
    int ncols, cout, ostride;
    long long *dbg;
    int acc_mode;         // for m: 0: m = v;  1: m = m + v;  2: m = (m + v) * scale   (hifi_gan.rs:113-118)
    float scale, inv_scale;
};

// true when the ResBlock convs of a stage with C channels can run on tcconv_kernel
inline bool tcv_supported(int C) { return C == 512 || C == 256 || C == 128 || C == 64 || C == 32 || C == 16; }

// (Cout, Cin, K) f32 weights -> TcConvW (allocates the image; the caller owns it)
int tcv_prepare_weights(const float *raw_dev, int C, int K, TcConvW *out, cudaStream_t st);
// ConvTranspose1d (Cin, Cout, 2 * stride) f32 weights -> TcConvW of the equivalent 2-tap conv with stride * Cout columns
int tcv_prepare_weights_t(const float *raw_dev, int Cin, int Cout, int stride, TcConvW *out, cudaStream_t st);
// u (C, L) f32 channel-major -> uc chunked f32 [C / 8][L][8] and img = split(act(u) * kTcvXScale), act = silu or identity
int tcv_chunk(const float *u, int C, int L, float *uc /* or null */, __half *img, cudaStream_t st, bool silu = true);
// one ResBlock conv (C -> C, causal, dilation dil) on the 5th-gen tensor cores
int tcv_conv(const TcConvW &w, const float *bias, const __half *ximg, int L, int dil, const float *res, float *y, __half *yimg,
             float *m, int acc_mode, float scale, cudaStream_t st);
int tcv_init();

}  // namespace fsb
