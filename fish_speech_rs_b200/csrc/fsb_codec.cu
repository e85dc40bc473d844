// FireflyCodec behind the C ABI: host-side mirror of fish_speech_core/lib/codec/
// {firefly,decoder,encoder,quantizer,hifi_gan,convnext}.rs for Fish >= 1.4.
#include <algorithm>
#include <memory>

#include <cuda.h>

#include <tuple>

#include "fsb_codec_kernels.cuh"
#include "fsb_tc_conv.cuh"
#include "fsb_tc_gemm.cuh"

namespace fsb {

struct ConvW {
    float *wt = nullptr;  // (Cin, K, Cout)
    float *bias = nullptr;
    int Cin = 0, Cout = 0, K = 0;
    float *wt_tiled = nullptr;  // ResBlock convs: [C / BM][C / 16][16][K][BM] (one contiguous block per TMA bulk copy)
    int tile_bm = 0;
    TcConvW tc;  // ResBlock convs with 64 / 128 / 256 channels: fp16 hi + lo weight image for tcconv_kernel
};

struct ConvNeXtW {
    float *dw_w = nullptr, *dw_b = nullptr, *ln_w = nullptr, *ln_b = nullptr;
    float *pw1_w = nullptr, *pw1_b = nullptr, *pw2_w = nullptr, *pw2_b = nullptr, *gamma = nullptr;
    int dim = 0;
};

constexpr int kUpRates[5] = {8, 8, 2, 2, 2};     // codec/config.rs presets (hifi_gan.rs:145-167)
constexpr int kUpKernels[5] = {16, 16, 4, 4, 4};
constexpr int kResKernels[3] = {3, 7, 11};       // hifi_gan.rs:100-107
constexpr int kResDilations[3] = {1, 3, 5};
constexpr int kGroups = 8, kDim = 512;
constexpr int kEncDims[4] = {128, 256, 384, 512};
constexpr int kEncDepths[4] = {3, 3, 9, 3};

}  // namespace fsb
using namespace fsb;

struct fsb_codec {
    fsb_codec_options opt;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::vector<void *> owned;
    // quantizer
    float *proj_out_w = nullptr, *proj_out_b = nullptr;  // (G,64,4), (G,64)
    float *proj_in_w = nullptr, *proj_in_b = nullptr;    // (G,4,64), (G,4)
    ConvW up_conv[2], down_conv[2];
    ConvNeXtW up_block[2], down_block[2];
    // head
    ConvW conv_pre, conv_post, ups[5];
    ConvW res_c1[5][3][3], res_c2[5][3][3];
    // encoder
    ConvW stem;
    float *stem_ln_w = nullptr, *stem_ln_b = nullptr;
    float *mid_ln_w[4] = {}, *mid_ln_b[4] = {};
    ConvW mid_conv[4];
    std::vector<ConvNeXtW> enc_blocks[4];
    float *enc_norm_w = nullptr, *enc_norm_b = nullptr;
    // scratch
    int max_frames = 0;
    uint32_t *d_codes = nullptr;
    int *d_err = nullptr;
    float *buf[7] = {};  // activation buffers of 32768 * max_frames floats (+ slack for the image padding rows): 4 raw + 2
                         // pre-activated copies for the TMA convs + the chunked residual stream of the tcgen05 convs
    bool res_tma = false;  // ResBlock convs run on resconv_tma_kernel
    bool res_tc = false;   // ... and those with >= 64 channels on tcconv_kernel (tcgen05)
    std::map<std::tuple<const void *, int, int, int>, TcMap> xmaps;  // (buffer, C, L, span) -> tensor map
    float *cn_h = nullptr, *cn_g = nullptr;
    long long *d_idx = nullptr;
    double *d_tw = nullptr;   // STFT twiddles (cos | sin), encoder only
    float *d_melfb = nullptr; // mel table (1025, 160), encoder only
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evd0 = nullptr, evd1 = nullptr;
    double device_ms_acc = 0;
    fsb_codec_stats stats;
    uint64_t launches = 0;
};

namespace fsb {

#define CLAUNCH_CHECK(c)                                                                          \
    do {                                                                                          \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) {                                                                  \
            set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return FSB_ERR_CUDA;                                                                  \
        }                                                                                         \
        (c)->launches++;                                                                          \
    } while (0)

template <typename T>
static int calloc_dev(fsb_codec *c, T **p, size_t n) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? FSB_ERR_OOM : FSB_ERR_CUDA;
    }
    c->owned.push_back(q);
    *p = reinterpret_cast<T *>(q);
    return FSB_OK;
}

// dst[(i1*d2 + i2)*d0 + i0] = src[(i0*d1 + i1)*d2 + i2]   (Cout,Cin,K) -> (Cin,K,Cout)
__global__ void relayout_021_to_120(const float *__restrict__ src, float *__restrict__ dst, int d0, int d1, int d2) {
    const size_t n = (size_t)d0 * d1 * d2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int i2 = (int)(i % d2), i1 = (int)((i / d2) % d1), i0 = (int)(i / ((size_t)d1 * d2));
        dst[((size_t)i1 * d2 + i2) * d0 + i0] = src[i];
    }
}
// dst[(i0*d2 + i2)*d1 + i1] = src[(i0*d1 + i1)*d2 + i2]   (Cin,Cout,K) -> (Cin,K,Cout)
__global__ void relayout_012_to_021(const float *__restrict__ src, float *__restrict__ dst, int d0, int d1, int d2) {
    const size_t n = (size_t)d0 * d1 * d2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int i2 = (int)(i % d2), i1 = (int)((i / d2) % d1), i0 = (int)(i / ((size_t)d1 * d2));
        dst[((size_t)i0 * d2 + i2) * d1 + i1] = src[i];
    }
}

// (Cout, Cin, K) -> [Cout / BM][Cin / 16][16][K][BM]
__global__ void relayout_res_tiled(const float *__restrict__ src, float *__restrict__ dst, int C, int K, int BM) {
    const size_t n = (size_t)C * C * K;
    const int nch = C / kResCK;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % K), ci = (int)((i / K) % C), co = (int)(i / ((size_t)C * K));
        const int ct = co / BM, col = co % BM, ch = ci / kResCK, cil = ci % kResCK;
        dst[((((size_t)ct * nch + ch) * kResCK + cil) * K + k) * BM + col] = src[i];
    }
}

static int load_vec(fsb_codec *c, const fsb_tensor *w, size_t n, const std::string &name, std::vector<int64_t> shape,
                    float **out) {
    DevTensor t;
    FSB_TRY(upload_tensor(w, n, name, shape, FSB_F32, c->stream, &t, &c->owned));
    *out = (float *)t.ptr;
    return FSB_OK;
}

// Conv1d weight (Cout, Cin, K) / ConvTranspose1d weight (Cin, Cout, K) -> ConvW
static int load_conv(fsb_codec *c, const fsb_tensor *w, size_t n, const std::string &prefix, int Cin, int Cout, int K,
                     bool transposed, ConvW *out, bool res_tiled = false, bool tc_only = false) {
    DevTensor raw;
    std::vector<void *> tmp_owned;
    std::vector<int64_t> shape = transposed ? std::vector<int64_t>{Cin, Cout, K} : std::vector<int64_t>{Cout, Cin, K};
    int st = upload_tensor(w, n, prefix + ".weight", shape, FSB_F32, c->stream, &raw, &tmp_owned);
    if (st != FSB_OK) return st;
    float *dst = nullptr;
    st = calloc_dev(c, &dst, (size_t)Cin * Cout * K);
    if (st == FSB_OK) {
        if (transposed) relayout_012_to_021<<<256, 256, 0, c->stream>>>((const float *)raw.ptr, dst, Cin, Cout, K);
        else relayout_021_to_120<<<256, 256, 0, c->stream>>>((const float *)raw.ptr, dst, Cout, Cin, K);
        if (res_tiled && !transposed && Cin == Cout && Cin % kResCK == 0) {
            // second copy for resconv_tma_kernel: [C / BM][C / 16][16][K][BM], BM = min(64, C)
            float *tl = nullptr;
            st = calloc_dev(c, &tl, (size_t)Cin * Cout * K);
            if (st == FSB_OK) {
                out->tile_bm = std::min(64, Cout);
                relayout_res_tiled<<<256, 256, 0, c->stream>>>((const float *)raw.ptr, tl, Cin, K, out->tile_bm);
                out->wt_tiled = tl;
            }
            if (st == FSB_OK && c->res_tc && tcv_supported(Cin)) {
                st = tcv_prepare_weights((const float *)raw.ptr, Cin, K, &out->tc, c->stream);
                if (st == FSB_OK) c->owned.push_back(out->tc.img);
            }
        }
        if (st == FSB_OK && tc_only && !transposed && Cin == Cout && c->res_tc && tcv_supported(Cin)) {
            st = tcv_prepare_weights((const float *)raw.ptr, Cin, K, &out->tc, c->stream);
            if (st == FSB_OK) c->owned.push_back(out->tc.img);
        }
        if (st == FSB_OK && tc_only && transposed && c->res_tc && K % 2 == 0 && Cin % 16 == 0) {
            // HiFi-GAN upsampling: ConvTranspose1d with kernel 2 * stride as a 2-tap conv over stride * Cout columns
            st = tcv_prepare_weights_t((const float *)raw.ptr, Cin, Cout, K / 2, &out->tc, c->stream);
            if (st == FSB_OK) c->owned.push_back(out->tc.img);
        }
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            set_error("weight relayout: %s", cudaGetErrorString(e));
            st = FSB_ERR_CUDA;
        }
    }
    for (void *p : tmp_owned) cudaFree(p);
    FSB_TRY(st);
    out->wt = dst;
    out->Cin = Cin;
    out->Cout = Cout;
    out->K = K;
    return load_vec(c, w, n, prefix + ".bias", {Cout}, &out->bias);
}

static int load_convnext(fsb_codec *c, const fsb_tensor *w, size_t n, const std::string &p, int dim, ConvNeXtW *o) {
    o->dim = dim;
    FSB_TRY(load_vec(c, w, n, p + "dwconv.conv.weight", {dim, 1, 7}, &o->dw_w));
    FSB_TRY(load_vec(c, w, n, p + "dwconv.conv.bias", {dim}, &o->dw_b));
    FSB_TRY(load_vec(c, w, n, p + "norm.weight", {dim}, &o->ln_w));
    FSB_TRY(load_vec(c, w, n, p + "norm.bias", {dim}, &o->ln_b));
    FSB_TRY(load_vec(c, w, n, p + "pwconv1.weight", {4 * dim, dim}, &o->pw1_w));
    FSB_TRY(load_vec(c, w, n, p + "pwconv1.bias", {4 * dim}, &o->pw1_b));
    FSB_TRY(load_vec(c, w, n, p + "pwconv2.weight", {dim, 4 * dim}, &o->pw2_w));
    FSB_TRY(load_vec(c, w, n, p + "pwconv2.bias", {dim}, &o->pw2_b));
    FSB_TRY(load_vec(c, w, n, p + "gamma", {dim}, &o->gamma));  // layer_scale_init_value > 0 in every preset
    return FSB_OK;
}

// ---------------------------------------------------------------- log-mel front-end
static double hz_to_mel_slaney(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_hz / f_sp + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz_slaney(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}
// The table the reference embeds as audio/melfilters160.bytes (spectrogram.rs:85-96): triangular filters on the
// Slaney mel scale with area normalisation, 1025 x 160, f_max = 22050.  Rebuilt here in double precision; equal to
// the reference's bytes to 1.8e-7 (tests/golden/mel_golden.json).
static std::vector<float> mel_filterbank() {
    const double nyq = 22050.0;
    std::vector<double> f_pts(kMels + 2);
    const double m_lo = hz_to_mel_slaney(0.0), m_hi = hz_to_mel_slaney(nyq);
    for (int i = 0; i < kMels + 2; ++i) f_pts[i] = mel_to_hz_slaney(m_lo + (m_hi - m_lo) * i / (kMels + 1));
    std::vector<float> fb((size_t)kBins * kMels);
    for (int k = 0; k < kBins; ++k) {
        const double fr = nyq * k / (kBins - 1);
        for (int m = 0; m < kMels; ++m) {
            const double down = (fr - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
            const double up = (f_pts[m + 2] - fr) / (f_pts[m + 2] - f_pts[m + 1]);
            const double v = std::max(0.0, std::min(down, up)) * (2.0 / (f_pts[m + 2] - f_pts[m]));
            fb[(size_t)k * kMels + m] = (float)v;
        }
    }
    return fb;
}

// frames the streaming STFT emits for n samples (stft.rs:52-90 driven by spectrogram.rs:44-66)
static int n_mel_frames_of(long long n) {
    const long long lp = n + (kFft - kHop);
    const long long full = lp / kHop, rem = lp % kHop;
    long long nf = std::max<long long>(full - (kFft / kHop - 1), 0);
    if (rem > 0 && lp >= kFft) nf += 1;
    return (int)nf;
}

// pcm (host, n samples) -> log-mel (160, Lm) in buf[2]
static int log_mel_to_device(fsb_codec *c, const float *pcm, long long n, int *Lm_out) {
    FSB_REQUIRE(c->opt.with_encoder && c->d_tw && c->d_melfb, FSB_ERR_STATE, "codec was created without the encoder");
    const int pad = (kFft - kHop) / 2;
    FSB_REQUIRE(n > pad, FSB_ERR_INVALID, "log_mel: %lld samples, need more than the reflect pad (%d)", n, pad);
    const int Lm = n_mel_frames_of(n);
    FSB_REQUIRE(Lm >= 4 && Lm <= 4 * c->max_frames, FSB_ERR_INVALID, "log_mel: %d mel frames outside [4, %d]", Lm,
                4 * c->max_frames);
    cudaStream_t st = c->stream;
    const int Lp = (int)n + 2 * pad;
    float *raw = c->buf[1], *xp = c->buf[3], *mag = c->buf[0];
    FSB_CUDA_OK(cudaMemcpyAsync(raw, pcm, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
    reflect_pad_kernel<<<(Lp + 255) / 256, 256, 0, st>>>(raw, (int)n, pad, xp);
    CLAUNCH_CHECK(c);
    const size_t smem = 3 * kFft * sizeof(double);
    stft_mag_kernel<<<Lm, 256, smem, st>>>(xp, Lp, c->d_tw, mag);
    CLAUNCH_CHECK(c);
    mel_log_kernel<<<Lm, kMels, 0, st>>>(mag, c->d_melfb, Lm, c->buf[2]);
    CLAUNCH_CHECK(c);
    *Lm_out = Lm;
    return FSB_OK;
}

// one ResBlock conv (C -> C, causal, dilation d) on pre-activated input `xact`
static int res_conv_tma(fsb_codec *c, const ConvW &w, const float *xact, int L, int dil, const float *res, float *y,
                        float *y_act, int acc_mode, float scale) {
    const int C = w.Cin, K = w.K, BM = w.tile_bm;
    const int span = res_span(K, dil);
    auto key = std::make_tuple((const void *)xact, C, L, span);
    auto it = c->xmaps.find(key);
    if (it == c->xmaps.end()) {
        TcMap m;
        FSB_TRY(tc_make_map_f32_2d(&m, xact, (uint64_t)C, (uint64_t)L, (uint32_t)span, (uint32_t)kResCK));
        it = c->xmaps.emplace(key, m).first;
    }
    ResConvArgs a;
    a.wt = w.wt_tiled; a.bias = w.bias; a.res = res; a.y = y; a.y_act = y_act;
    a.C = C; a.L = L; a.acc_mode = acc_mode; a.scale = scale;
    const dim3 grid((L + kResBN - 1) / kResBN, C / BM);
#define RES_LAUNCH(BMv, KTv, DILv)                                                                                   \
    {                                                                                                                \
        const size_t smem = (size_t)2 * res_stage_floats(BMv, KTv, DILv) * sizeof(float) + 128;                      \
        static bool attr_done = false;                                                                               \
        if (!attr_done && smem > 48 * 1024) {                                                                        \
            FSB_CUDA_OK(cudaFuncSetAttribute(resconv_tma_kernel<BMv, KTv, DILv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_done = true;                                                                                        \
        }                                                                                                            \
        resconv_tma_kernel<BMv, KTv, DILv><<<grid, 256, smem, c->stream>>>(*reinterpret_cast<const CUtensorMap *>(&it->second), a); \
    }
#define RES_DIL(BMv, KTv)                                        \
    switch (dil) {                                               \
        case 1: RES_LAUNCH(BMv, KTv, 1) break;                   \
        case 3: RES_LAUNCH(BMv, KTv, 3) break;                   \
        default: RES_LAUNCH(BMv, KTv, 5) break;                  \
    }
#define RES_K(BMv)                                               \
    switch (K) {                                                 \
        case 3: RES_DIL(BMv, 3) break;                           \
        case 7: RES_DIL(BMv, 7) break;                           \
        default: RES_DIL(BMv, 11) break;                         \
    }
    if (BM == 64) RES_K(64)
    else if (BM == 32) RES_K(32)
    else RES_K(16)
#undef RES_K
#undef RES_DIL
#undef RES_LAUNCH
    CLAUNCH_CHECK(c);
    return FSB_OK;
}

// ---------------------------------------------------------------- launches
static int launch_conv(fsb_codec *c, ConvArgs a, int nz = 1) {
    const int off_lo = std::min(0, (a.K - 1) * a.dil) - a.pad, off_hi = std::max(0, (a.K - 1) * a.dil) - a.pad;
    if (a.Lout <= 0) return FSB_OK;
#define CONV_LAUNCH(BM, TM, TN, KT)                                                                \
    {                                                                                              \
        const int BN = 32 * TN;                                                                    \
        const int span = (BN - 1) * a.stride + off_hi - off_lo + 1;                                \
        const int gx = (a.Lout + BN - 1) / BN;                                                     \
        const size_t smem = ((size_t)kConvCK * a.K * BM + (size_t)kConvCK * span) * sizeof(float); \
        FSB_REQUIRE(smem <= 200 * 1024, FSB_ERR_UNSUPPORTED, "conv tile needs %zu B of smem", smem); \
        if (smem > 48 * 1024)                                                                      \
            FSB_CUDA_OK(cudaFuncSetAttribute(conv1d_kernel<BM, TM, TN, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        conv1d_kernel<BM, TM, TN, KT><<<dim3(gx, (a.Cout + BM - 1) / BM, nz), 256, smem, c->stream>>>(a); \
    }
    // taps of the ResBlock convs (3 / 7 / 11), of the transposed-conv phases (2) and of conv_pre / conv_post (13) are
    // compile-time constants of the kernel; anything else runs the generic tap loop
#define CONV_CASE(BM, TM, TN)                                                                      \
    switch (a.K) {                                                                                 \
        case 2: CONV_LAUNCH(BM, TM, TN, 2) break;                                                  \
        case 3: CONV_LAUNCH(BM, TM, TN, 3) break;                                                  \
        case 7: CONV_LAUNCH(BM, TM, TN, 7) break;                                                  \
        case 11: CONV_LAUNCH(BM, TM, TN, 11) break;                                                \
        case 13: CONV_LAUNCH(BM, TM, TN, 13) break;                                                \
        default: CONV_LAUNCH(BM, TM, TN, 0) break;                                                 \
    }
    // 64 x 128 output tiles; launches that would leave most SMs idle (short sequences at the top of the
    // stack: conv_pre, the first transposed convs) fall back to 32 x 64 tiles = 4x the CTAs
    const long long ctas64 = (long long)((a.Lout + 127) / 128) * ((a.Cout + 63) / 64) * nz;
    if (a.Cout >= 64 && ctas64 >= 120) CONV_CASE(64, 8, 4)
    else if (a.Cout >= 32 && a.Cout < 64 && (long long)((a.Lout + 127) / 128) * ((a.Cout + 31) / 32) * nz >= 120) CONV_CASE(32, 4, 4)
    else if (a.Cout >= 32) CONV_CASE(32, 4, 2)
    else CONV_CASE(16, 2, 4)
#undef CONV_LAUNCH
#undef CONV_CASE
    CLAUNCH_CHECK(c);
    return FSB_OK;
}

// FishConvNet::forward: y (Cout, Lout) = conv(x (Cin, Lin)), Lout = (Lin + pad - (K-1)*dil - 1)/stride + 1 == Lin/stride
static int conv_fwd(fsb_codec *c, const ConvW &w, const float *x, int Lin, float *y, int dil, int stride,
                    bool pre_silu, const float *res, int acc_mode, float scale, bool post_tanh, int *Lout_out,
                    float *y_act = nullptr) {
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    const int pad = (w.K - 1) * dil + 1 - stride;  // utils/mod.rs:43,55
    const int Lout = (Lin + pad - (w.K - 1) * dil - 1) / stride + 1;
    a.x = x; a.wt = w.wt; a.bias = w.bias; a.res = res; a.y = y;
    a.Cin = w.Cin; a.Cout = w.Cout; a.Lin = Lin; a.Lout = Lout; a.Ly = Lout;
    a.K = w.K; a.Kw = w.K; a.k0 = 0; a.kstep = 1;
    a.dil = dil; a.pad = pad; a.stride = stride; a.ostride = 1; a.ooff = 0;
    a.pre_silu = pre_silu; a.acc_mode = acc_mode; a.scale = scale; a.post_tanh = post_tanh;
    a.y_act = y_act;
    if (Lout_out) *Lout_out = Lout;
    return launch_conv(c, a);
}

// FishTransConvNet::forward: y (Cout, Lin*stride); kernel K in {stride, 2*stride}
static int convT_fwd(fsb_codec *c, const ConvW &w, const float *x, int Lin, float *y, int stride, bool pre_silu,
                     float *y_act = nullptr) {
    FSB_REQUIRE(w.K == stride || w.K == 2 * stride, FSB_ERR_UNSUPPORTED, "ConvTranspose1d k=%d s=%d unsupported", w.K,
                stride);
    // one launch, one grid.z slice per output phase r (= t mod stride): taps {r, r + stride}
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.wt = w.wt; a.bias = w.bias; a.res = nullptr; a.y = y;
    a.Cin = w.Cin; a.Cout = w.Cout; a.Lin = Lin; a.Lout = Lin; a.Ly = Lin * stride;
    a.K = w.K / stride; a.Kw = w.K; a.k0 = 0; a.kstep = stride;
    a.dil = -1; a.pad = 0; a.stride = 1; a.ostride = stride; a.ooff = 0;
    a.z_k0 = 1; a.z_ooff = 1;
    a.pre_silu = pre_silu;
    a.y_act = y_act;
    return launch_conv(c, a, stride);
}

// ConvNeXtBlock::forward in place on x (C, L)
static int convnext_fwd(fsb_codec *c, const ConvNeXtW &w, float *x, int L) {
    const int C = w.dim;
    const int warps = 4;
    dwconv_ln_kernel<<<(L + warps - 1) / warps, warps * 32, warps * C * sizeof(float), c->stream>>>(
        x, C, L, w.dw_w, w.dw_b, w.ln_w, w.ln_b, 1e-6f, c->cn_h);
    CLAUNCH_CHECK(c);
    pw_gemm_kernel<0><<<dim3((4 * C + 63) / 64, (L + 63) / 64), 256, 0, c->stream>>>(c->cn_h, w.pw1_w, w.pw1_b, nullptr,
                                                                                     nullptr, c->cn_g, L, 4 * C, C);
    CLAUNCH_CHECK(c);
    pw_gemm_kernel<1><<<dim3((C + 63) / 64, (L + 63) / 64), 256, 0, c->stream>>>(c->cn_g, w.pw2_w, w.pw2_b, w.gamma, x,
                                                                                 x, L, C, 4 * C);
    CLAUNCH_CHECK(c);
    return FSB_OK;
}

// FireflyDecoder::decode for one utterance; codes already on the device at c->d_codes (8, T)
static int decode_device(fsb_codec *c, int T, float *pcm_dev_out /* device (2048*T) */) {
    float *A = c->buf[0], *B = c->buf[1], *R = c->buf[2], *X1 = c->buf[3];
    cudaStream_t st = c->stream;
    FSB_CUDA_OK(cudaMemsetAsync(c->d_err, 0, sizeof(int), st));
    fsq_decode_kernel<<<dim3((T + 127) / 128, kGroups), 128, 0, st>>>(c->d_codes, T, kGroups, c->proj_out_w,
                                                                      c->proj_out_b, A, c->d_err);
    CLAUNCH_CHECK(c);
    // quantizer.upsample: upsample.0 then upsample.1 (quantizer.rs:126-133)
    int L = T;
    float *cur = A, *nxt = B;
    for (int i = 0; i < 2; ++i) {
        FSB_TRY(convT_fwd(c, c->up_conv[i], cur, L, nxt, 2, false));
        L *= 2;
        FSB_TRY(convnext_fwd(c, c->up_block[i], nxt, L));
        std::swap(cur, nxt);
    }
    // HiFiGAN::forward (hifi_gan.rs:207-216)
    if (c->res_tc && c->conv_pre.tc.img) {
        // conv_pre on the tensor cores: image of the raw input (no activation), result straight into the (C, L) buffer
        __half *img = reinterpret_cast<__half *>(c->buf[4]);
        FSB_TRY(tcv_chunk(cur, c->conv_pre.Cin, L, nullptr, img, st, false));
        FSB_TRY(tcv_conv(c->conv_pre.tc, c->conv_pre.bias, img, L, 1, nullptr, nullptr, nullptr, nxt, 0, 0.f, st));
        c->launches += 2;
    } else {
        FSB_TRY(conv_fwd(c, c->conv_pre, cur, L, nxt, 1, 1, false, nullptr, 0, 0.f, false, nullptr));
    }
    std::swap(cur, nxt);  // cur = M (stage input), nxt = U
    const float third = (float)(1.0 / 3.0);
    for (int i = 0; i < 5; ++i) {
        float *M = cur, *U = nxt;
        if (c->res_tc && c->res_c1[i][0][0].tc.img) {
            // tcgen05 ResBlocks: activations as fp16 hi + lo images (time-major rows of 8 channels), residual stream chunked
            float *Uc = c->buf[6], *Rc = c->buf[2];
            __half *SUi = reinterpret_cast<__half *>(c->buf[4]), *SRi = reinterpret_cast<__half *>(c->buf[5]);
            __half *X1i = reinterpret_cast<__half *>(c->buf[3]);
            if (c->ups[i].tc.img) {
                // upsampling on the tensor cores too: image of silu(M), then the phases of the transposed conv as GEMM
                // columns; the epilogue writes the chunked residual stream and the image of silu(U) directly
                FSB_TRY(tcv_chunk(M, c->ups[i].Cin, L, nullptr, SRi, st, true));
                FSB_TRY(tcv_conv(c->ups[i].tc, c->ups[i].bias, SRi, L, 1, nullptr, Uc, SUi, nullptr, 0, 0.f, st));
                L *= kUpRates[i];
                c->launches += 2;
            } else {
                FSB_TRY(convT_fwd(c, c->ups[i], M, L, U, kUpRates[i], true));
                L *= kUpRates[i];
                FSB_TRY(tcv_chunk(U, c->ups[i].Cout, L, Uc, SUi, st));
                c->launches++;
            }
            for (int j = 0; j < 3; ++j) {
                const float *xin = Uc;
                const __half *xin_img = SUi;
                for (int m = 0; m < 3; ++m) {
                    const int d = kResDilations[m];
                    const ConvW &w1 = c->res_c1[i][j][m], &w2 = c->res_c2[i][j][m];
                    FSB_TRY(tcv_conv(w1.tc, w1.bias, xin_img, L, d, nullptr, nullptr, X1i, nullptr, 0, 0.f, st));
                    if (m < 2) {
                        FSB_TRY(tcv_conv(w2.tc, w2.bias, X1i, L, d, xin, Rc, SRi, nullptr, 0, 0.f, st));
                        xin = Rc;
                        xin_img = SRi;
                    } else {
                        FSB_TRY(tcv_conv(w2.tc, w2.bias, X1i, L, d, xin, nullptr, nullptr, M, j == 0 ? 0 : (j == 1 ? 1 : 2), third, st));
                    }
                    c->launches += 2;
                }
            }
            continue;
        }
        if (c->res_tma) {
            // TMA-staged ResBlock convs read pre-activated inputs: every producer also writes silu(result) where its
            // consumer applies silu (hifi_gan.rs:76-79): SU = silu(U), SR = silu(R), X1 holds silu(conv1 output) only
            float *SU = c->buf[4], *SR = c->buf[5];
            FSB_TRY(convT_fwd(c, c->ups[i], M, L, U, kUpRates[i], true, SU));
            L *= kUpRates[i];
            for (int j = 0; j < 3; ++j) {
                const float *xin = U, *xin_act = SU;
                for (int m = 0; m < 3; ++m) {
                    const int d = kResDilations[m];
                    FSB_TRY(res_conv_tma(c, c->res_c1[i][j][m], xin_act, L, d, nullptr, nullptr, X1, 0, 0.f));
                    if (m < 2) {
                        FSB_TRY(res_conv_tma(c, c->res_c2[i][j][m], X1, L, d, xin, R, SR, 0, 0.f));
                        xin = R;
                        xin_act = SR;
                    } else {
                        FSB_TRY(res_conv_tma(c, c->res_c2[i][j][m], X1, L, d, xin, M, nullptr, j == 0 ? 0 : (j == 1 ? 1 : 2), third));
                    }
                }
            }
            continue;
        }
        FSB_TRY(convT_fwd(c, c->ups[i], M, L, U, kUpRates[i], true));
        L *= kUpRates[i];
        for (int j = 0; j < 3; ++j) {
            const float *xin = U;
            for (int m = 0; m < 3; ++m) {
                const int d = kResDilations[m];
                FSB_TRY(conv_fwd(c, c->res_c1[i][j][m], xin, L, X1, d, 1, true, nullptr, 0, 0.f, false, nullptr));
                if (m < 2) {
                    FSB_TRY(conv_fwd(c, c->res_c2[i][j][m], X1, L, R, d, 1, true, xin, 0, 0.f, false, nullptr));
                    xin = R;
                } else {
                    // last conv of the block feeds only stack+mean (hifi_gan.rs:113-118)
                    FSB_TRY(conv_fwd(c, c->res_c2[i][j][m], X1, L, M, d, 1, true, xin, j == 0 ? 0 : (j == 1 ? 1 : 2),
                                     third, false, nullptr));
                }
            }
        }
        // M holds the stage output; U is free
    }
    if (c->conv_post.Cout == 1 && c->conv_post.Cin * c->conv_post.K <= kPostMaxCK && !getenv("FSB_CODEC_OLD_POST")) {
        const int C = c->conv_post.Cin, K = c->conv_post.K;
        const size_t smem = ((size_t)C * (kPostTT + K - 1) + (size_t)C * K) * sizeof(float);
        conv_post_kernel<<<(L + kPostTT - 1) / kPostTT, kPostThreads, smem, st>>>(cur, c->conv_post.wt, c->conv_post.bias, pcm_dev_out,
                                                                                 C, K, L);
        CLAUNCH_CHECK(c);
        return FSB_OK;
    }
    FSB_TRY(conv_fwd(c, c->conv_post, cur, L, pcm_dev_out, 1, 1, true, nullptr, 0, 0.f, true, nullptr));
    return FSB_OK;
}

static int codec_create_impl(fsb_codec *c, const fsb_tensor *w, size_t n) {
    FSB_TRY(select_device(c->opt.device));
    if (c->opt.stream) c->stream = (cudaStream_t)c->opt.stream;
    else {
        FSB_CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    FSB_CUDA_OK(cudaEventCreate(&c->ev0));
    FSB_CUDA_OK(cudaEventCreate(&c->ev1));
    FSB_CUDA_OK(cudaEventCreate(&c->evd0));
    FSB_CUDA_OK(cudaEventCreate(&c->evd1));
    c->res_tma = getenv("FSB_CODEC_NO_TMA") == nullptr;
    c->res_tc = c->res_tma && getenv("FSB_CODEC_NO_TC") == nullptr;
    if (c->res_tc) FSB_TRY(tcv_init());
    // FSQ projections, gathered per group into one array each
    FSB_TRY(calloc_dev(c, &c->proj_out_w, (size_t)kGroups * 64 * 4));
    FSB_TRY(calloc_dev(c, &c->proj_out_b, (size_t)kGroups * 64));
    FSB_TRY(calloc_dev(c, &c->proj_in_w, (size_t)kGroups * 4 * 64));
    FSB_TRY(calloc_dev(c, &c->proj_in_b, (size_t)kGroups * 4));
    for (int g = 0; g < kGroups; ++g) {
        const std::string p = "quantizer.residual_fsq.rvqs." + std::to_string(g) + ".";
        float *t = nullptr;
        FSB_TRY(load_vec(c, w, n, p + "project_out.weight", {64, 4}, &t));
        FSB_CUDA_OK(cudaMemcpyAsync(c->proj_out_w + (size_t)g * 256, t, 256 * 4, cudaMemcpyDeviceToDevice, c->stream));
        FSB_TRY(load_vec(c, w, n, p + "project_out.bias", {64}, &t));
        FSB_CUDA_OK(cudaMemcpyAsync(c->proj_out_b + (size_t)g * 64, t, 64 * 4, cudaMemcpyDeviceToDevice, c->stream));
        if (c->opt.with_encoder) {
            FSB_TRY(load_vec(c, w, n, p + "project_in.weight", {4, 64}, &t));
            FSB_CUDA_OK(cudaMemcpyAsync(c->proj_in_w + (size_t)g * 256, t, 256 * 4, cudaMemcpyDeviceToDevice, c->stream));
            FSB_TRY(load_vec(c, w, n, p + "project_in.bias", {4}, &t));
            FSB_CUDA_OK(cudaMemcpyAsync(c->proj_in_b + (size_t)g * 4, t, 4 * 4, cudaMemcpyDeviceToDevice, c->stream));
        }
    }
    for (int i = 0; i < 2; ++i) {
        const std::string p = "quantizer.upsample." + std::to_string(i) + ".";
        FSB_TRY(load_conv(c, w, n, p + "0.conv", kDim, kDim, 2, true, &c->up_conv[i]));
        FSB_TRY(load_convnext(c, w, n, p + "1.", kDim, &c->up_block[i]));
    }
    // (conv_pre stays on the FP32-FMA kernel unless FSB_CODEC_TC_PRE is set: 22-bit operands at the very first conv move the
    // PCM by 6e-5, too close to the 1e-4 parity bar; the ResBlock / upsampling convs together stay at 1-2e-5)
    FSB_TRY(load_conv(c, w, n, "head.conv_pre.conv", kDim, kDim, 13, false, &c->conv_pre, false, getenv("FSB_CODEC_TC_PRE") != nullptr));
    for (int i = 0; i < 5; ++i) {
        const int cin = kDim >> i, cout = kDim >> (i + 1);
        FSB_TRY(load_conv(c, w, n, "head.ups." + std::to_string(i) + ".conv", cin, cout, kUpKernels[i], true, &c->ups[i], false,
                          kUpKernels[i] == 2 * kUpRates[i] && getenv("FSB_CODEC_NO_TC_UPS") == nullptr));
        for (int j = 0; j < 3; ++j)
            for (int m = 0; m < 3; ++m) {
                const std::string p = "head.resblocks." + std::to_string(i) + ".blocks." + std::to_string(j) + ".";
                FSB_TRY(load_conv(c, w, n, p + "convs1." + std::to_string(m) + ".conv", cout, cout, kResKernels[j], false,
                                  &c->res_c1[i][j][m], true));
                FSB_TRY(load_conv(c, w, n, p + "convs2." + std::to_string(m) + ".conv", cout, cout, kResKernels[j], false,
                                  &c->res_c2[i][j][m], true));
            }
    }
    FSB_TRY(load_conv(c, w, n, "head.conv_post.conv", 16, 1, 13, false, &c->conv_post));
    if (c->opt.with_encoder) {
        // log-mel front-end tables: STFT twiddles (f64) and the Slaney mel table
        FSB_TRY(calloc_dev(c, &c->d_tw, (size_t)2 * kFft));
        stft_twiddle_kernel<<<kFft / 256, 256, 0, c->stream>>>(c->d_tw);
        FSB_CUDA_OK(cudaGetLastError());
        FSB_CUDA_OK(cudaFuncSetAttribute(stft_mag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * kFft * sizeof(double))));
        {
            const std::vector<float> fb = mel_filterbank();
            FSB_TRY(calloc_dev(c, &c->d_melfb, fb.size()));
            FSB_CUDA_OK(cudaMemcpyAsync(c->d_melfb, fb.data(), fb.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
            FSB_CUDA_OK(cudaStreamSynchronize(c->stream));  // fb is a local
        }
        for (int i = 0; i < 2; ++i) {
            const std::string p = "quantizer.downsample." + std::to_string(i) + ".";
            FSB_TRY(load_conv(c, w, n, p + "0.conv", kDim, kDim, 2, false, &c->down_conv[i]));
            FSB_TRY(load_convnext(c, w, n, p + "1.", kDim, &c->down_block[i]));
        }
        const std::string d = "backbone.downsample_layers.";
        FSB_TRY(load_conv(c, w, n, d + "0.0.conv", 160, kEncDims[0], 7, false, &c->stem));
        FSB_TRY(load_vec(c, w, n, d + "0.1.weight", {kEncDims[0]}, &c->stem_ln_w));
        FSB_TRY(load_vec(c, w, n, d + "0.1.bias", {kEncDims[0]}, &c->stem_ln_b));
        for (int i = 1; i < 4; ++i) {
            const std::string p = d + std::to_string(i) + ".";
            FSB_TRY(load_vec(c, w, n, p + "0.weight", {kEncDims[i - 1]}, &c->mid_ln_w[i]));
            FSB_TRY(load_vec(c, w, n, p + "0.bias", {kEncDims[i - 1]}, &c->mid_ln_b[i]));
            FSB_TRY(load_conv(c, w, n, p + "1", kEncDims[i - 1], kEncDims[i], 1, false, &c->mid_conv[i]));
        }
        for (int i = 0; i < 4; ++i) {
            c->enc_blocks[i].resize(kEncDepths[i]);
            for (int j = 0; j < kEncDepths[i]; ++j)
                FSB_TRY(load_convnext(c, w, n, "backbone.stages." + std::to_string(i) + "." + std::to_string(j) + ".",
                                      kEncDims[i], &c->enc_blocks[i][j]));
        }
        FSB_TRY(load_vec(c, w, n, "backbone.norm.weight", {kDim}, &c->enc_norm_w));
        FSB_TRY(load_vec(c, w, n, "backbone.norm.bias", {kDim}, &c->enc_norm_b));
    }
    // scratch
    const size_t T = (size_t)c->max_frames;
    FSB_TRY(calloc_dev(c, &c->d_codes, (size_t)kGroups * T));
    FSB_TRY(calloc_dev(c, &c->d_err, 1));
    for (int i = 0; i < (c->res_tc ? 7 : (c->res_tma ? 6 : 4)); ++i) FSB_TRY(calloc_dev(c, &c->buf[i], 32768 * T + 512 * 1024));
    FSB_TRY(calloc_dev(c, &c->cn_h, 4 * T * kDim));
    FSB_TRY(calloc_dev(c, &c->cn_g, 4 * T * kDim * 4));
    FSB_TRY(calloc_dev(c, &c->d_idx, (size_t)kGroups * T));
    FSB_CUDA_OK(cudaStreamSynchronize(c->stream));
    return FSB_OK;
}

static void codec_free(fsb_codec *c) {
    if (!c) return;
    for (void *p : c->owned) cudaFree(p);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->evd0) cudaEventDestroy(c->evd0);
    if (c->evd1) cudaEventDestroy(c->evd1);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int decode_one(fsb_codec *c, const uint32_t *codes, int T, float *pcm) {
    FSB_REQUIRE(codes && pcm, FSB_ERR_INVALID, "decode: null pointer");
    FSB_REQUIRE(T >= 1 && T <= c->max_frames, FSB_ERR_INVALID, "decode: %d frames outside [1, max_frames=%d]", T,
                c->max_frames);
    cudaStream_t st = c->stream;
    FSB_CUDA_OK(cudaMemcpyAsync(c->d_codes, codes, (size_t)kGroups * T * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    // final PCM lands in buf[3] (X1 is free by then)
    float *pcm_dev = c->buf[3];
    FSB_CUDA_OK(cudaEventRecord(c->evd0, st));
    FSB_TRY(decode_device(c, T, pcm_dev));
    FSB_CUDA_OK(cudaEventRecord(c->evd1, st));
    int err = 0;
    FSB_CUDA_OK(cudaMemcpyAsync(pcm, pcm_dev, (size_t)T * 2048 * sizeof(float), cudaMemcpyDeviceToHost, st));
    FSB_CUDA_OK(cudaMemcpyAsync(&err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    FSB_CUDA_OK(cudaStreamSynchronize(st));
    FSB_REQUIRE(err == 0, FSB_ERR_INVALID, "decode: a code is >= 1000 (outside the FSQ implicit codebook, Q11)");
    float dms = 0.f;
    FSB_CUDA_OK(cudaEventElapsedTime(&dms, c->evd0, c->evd1));
    c->device_ms_acc += dms;
    return FSB_OK;
}

// Left context (code frames) after which a causal decode no longer depends on earlier codes: ConvNeXt k7 at 2x and
// 4x (3 + 1.5 frames), conv_pre k13 at 4x (3), first transposed conv (0.25), ResBlock1 chains k=11, d=1,3,5 at
// 32x .. 2048x (180 samples each: 5.6 + 0.7 + 0.35 + 0.18 + 0.09) = 14.7 frames.
constexpr int kHaloFrames = 16;

// frames [t0, t1) of an utterance whose codes (8, T_total) sit on the host -> device PCM at buf[3] + offset
static int decode_block_device(fsb_codec *c, const uint32_t *codes, int T_total, int t0, int t1, float **pcm_dev,
                               long long *n_samples) {
    FSB_REQUIRE(codes, FSB_ERR_INVALID, "decode: null pointer");
    FSB_REQUIRE(0 <= t0 && t0 < t1 && t1 <= T_total, FSB_ERR_INVALID, "decode_block: bad frame range [%d, %d) of %d", t0,
                t1, T_total);
    const int h = std::min(t0, kHaloFrames), Tb = t1 - t0 + h;
    FSB_REQUIRE(Tb <= c->max_frames, FSB_ERR_INVALID, "decode_block: %d frames (with halo) exceed max_frames=%d", Tb,
                c->max_frames);
    cudaStream_t st = c->stream;
    // gather the (8, Tb) window out of the row-major (8, T_total) host array
    FSB_CUDA_OK(cudaMemcpy2DAsync(c->d_codes, (size_t)Tb * sizeof(uint32_t), codes + (t0 - h),
                                  (size_t)T_total * sizeof(uint32_t), (size_t)Tb * sizeof(uint32_t), kGroups,
                                  cudaMemcpyHostToDevice, st));
    FSB_CUDA_OK(cudaEventRecord(c->evd0, st));
    FSB_TRY(decode_device(c, Tb, c->buf[3]));
    FSB_CUDA_OK(cudaEventRecord(c->evd1, st));
    *pcm_dev = c->buf[3] + (size_t)h * 2048;
    *n_samples = (long long)(t1 - t0) * 2048;
    return FSB_OK;
}

static int finish_decode(fsb_codec *c, const char *what) {
    int err = 0;
    FSB_CUDA_OK(cudaMemcpyAsync(&err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    FSB_CUDA_OK(cudaStreamSynchronize(c->stream));
    FSB_REQUIRE(err == 0, FSB_ERR_INVALID, "%s: a code is >= 1000 (outside the FSQ implicit codebook, Q11)", what);
    float dms = 0.f;
    FSB_CUDA_OK(cudaEventElapsedTime(&dms, c->evd0, c->evd1));
    c->stats.device_ms = dms;
    return FSB_OK;
}

}  // namespace fsb

extern "C" {

int fsb_codec_decode_block(fsb_codec *c, const uint32_t *codes, int32_t n_frames_total, int32_t t0, int32_t t1, float *pcm) {
    FSB_REQUIRE(c && pcm, FSB_ERR_INVALID, "decode_block: null argument");
    FSB_CUDA_OK(cudaSetDevice(c->opt.device));
    float *dev = nullptr;
    long long n = 0;
    FSB_TRY(decode_block_device(c, codes, n_frames_total, t0, t1, &dev, &n));
    FSB_CUDA_OK(cudaMemcpyAsync(pcm, dev, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    return finish_decode(c, "decode_block");
}

int fsb_codec_decode_block_s16(fsb_codec *c, const uint32_t *codes, int32_t n_frames_total, int32_t t0, int32_t t1,
                               uint32_t to_rate, int16_t *out, size_t cap, size_t *out_len) {
    FSB_REQUIRE(c && out && out_len, FSB_ERR_INVALID, "decode_block_s16: null argument");
    FSB_CUDA_OK(cudaSetDevice(c->opt.device));
    float *dev = nullptr;
    long long n = 0;
    FSB_TRY(decode_block_device(c, codes, n_frames_total, t0, t1, &dev, &n));
    cudaStream_t st = c->stream;
    const uint32_t from_rate = 44100;
    long long n_out = n;
    float *src = dev;
    if (to_rate != 0 && to_rate != from_rate) {
        // audio/functional.rs:8-9: ratio and length in f64
        const double ratio = (double)to_rate / (double)from_rate;
        n_out = (long long)std::ceil((double)n * ratio);
        FSB_REQUIRE(n_out >= 1 && (size_t)n_out <= (size_t)32768 * c->max_frames, FSB_ERR_INVALID, "decode_block_s16: %lld output samples",
                    n_out);
        resample_linear_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(dev, n, ratio, n_out, c->buf[0]);
        CLAUNCH_CHECK(c);
        src = c->buf[0];
    }
    FSB_REQUIRE((size_t)n_out <= cap, FSB_ERR_INVALID, "decode_block_s16: capacity %zu < %lld samples", cap, n_out);
    short *s16 = reinterpret_cast<short *>(c->buf[1]);
    pcm_f32_to_s16_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(src, n_out, s16);
    CLAUNCH_CHECK(c);
    FSB_CUDA_OK(cudaMemcpyAsync(out, s16, (size_t)n_out * sizeof(short), cudaMemcpyDeviceToHost, st));
    *out_len = (size_t)n_out;
    return finish_decode(c, "decode_block_s16");
}

int fsb_codec_create(const fsb_tensor *weights, size_t n_weights, const fsb_codec_options *opts, fsb_codec **out) {
    FSB_REQUIRE(weights && opts && out, FSB_ERR_INVALID, "fsb_codec_create: null argument");
    *out = nullptr;
    FSB_REQUIRE(opts->fish_version == FSB_FISH_1_4 || opts->fish_version == FSB_FISH_1_5, FSB_ERR_UNSUPPORTED,
                "only the causal (Fish >= 1.4) codec is implemented");
    FSB_REQUIRE(opts->max_frames >= 1, FSB_ERR_INVALID, "max_frames must be >= 1");
    std::unique_ptr<fsb_codec> c(new fsb_codec());
    c->opt = *opts;
    c->max_frames = opts->max_frames;
    memset(&c->stats, 0, sizeof(c->stats));
    int st = codec_create_impl(c.get(), weights, n_weights);
    if (st != FSB_OK) {
        codec_free(c.release());
        (void)cudaGetLastError();
        return st;
    }
    *out = c.release();
    return FSB_OK;
}

int fsb_codec_destroy(fsb_codec *c) {
    if (!c) return FSB_OK;
    cudaSetDevice(c->opt.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    codec_free(c);
    (void)cudaGetLastError();
    return FSB_OK;
}

int fsb_codec_decode(fsb_codec *c, const uint32_t *codes, int32_t n_frames, float *pcm) {
    FSB_REQUIRE(c, FSB_ERR_INVALID, "null codec handle");
    FSB_CUDA_OK(cudaSetDevice(c->opt.device));
    const uint64_t l0 = c->launches;
    c->device_ms_acc = 0;
    FSB_CUDA_OK(cudaEventRecord(c->ev0, c->stream));
    FSB_TRY(decode_one(c, codes, n_frames, pcm));
    FSB_CUDA_OK(cudaEventRecord(c->ev1, c->stream));
    FSB_CUDA_OK(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    FSB_CUDA_OK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.decode_ms = ms;
    c->stats.device_ms = c->device_ms_acc;
    c->stats.kernel_launches = c->launches - l0;
    return FSB_OK;
}

int fsb_codec_decode_batch(fsb_codec *c, const uint32_t *const *codes, const int32_t *n_frames, int32_t n,
                           float *const *pcm) {
    FSB_REQUIRE(c && codes && n_frames && pcm && n >= 1, FSB_ERR_INVALID, "decode_batch: bad arguments");
    FSB_CUDA_OK(cudaSetDevice(c->opt.device));
    const uint64_t l0 = c->launches;
    c->device_ms_acc = 0;
    FSB_CUDA_OK(cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < n; ++i) FSB_TRY(decode_one(c, codes[i], n_frames[i], pcm[i]));
    FSB_CUDA_OK(cudaEventRecord(c->ev1, c->stream));
    FSB_CUDA_OK(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    FSB_CUDA_OK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.decode_ms = ms;
    c->stats.device_ms = c->device_ms_acc;
    c->stats.kernel_launches = c->launches - l0;
    return FSB_OK;
}

static int encode_from_device_mel(fsb_codec *c, int Lm, int64_t *codes, size_t cap, size_t *out_len);

int fsb_codec_log_mel(fsb_codec *c, const float *pcm, int64_t n_samples, float *mel, size_t cap_frames, size_t *out_frames) {
    FSB_REQUIRE(c && pcm && mel && out_frames, FSB_ERR_INVALID, "log_mel: null argument");
    FSB_CUDA_OK(cudaSetDevice(c->opt.device));
    int Lm = 0;
    FSB_TRY(log_mel_to_device(c, pcm, n_samples, &Lm));
    FSB_REQUIRE((size_t)Lm <= cap_frames, FSB_ERR_INVALID, "log_mel: capacity %zu < %d frames", cap_frames, Lm);
    FSB_CUDA_OK(cudaMemcpyAsync(mel, c->buf[2], (size_t)kMels * Lm * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    FSB_CUDA_OK(cudaStreamSynchronize(c->stream));
    *out_frames = (size_t)Lm;
    return FSB_OK;
}

int fsb_codec_encode(fsb_codec *c, const float *pcm, int64_t n_samples, int64_t *codes, size_t cap, size_t *out_len) {
    FSB_REQUIRE(c && pcm && codes && out_len, FSB_ERR_INVALID, "encode: null argument");
    FSB_CUDA_OK(cudaSetDevice(c->opt.device));
    int Lm = 0;
    FSB_TRY(log_mel_to_device(c, pcm, n_samples, &Lm));  // the mel stays on the device (buf[2])
    return encode_from_device_mel(c, Lm, codes, cap, out_len);
}

int fsb_codec_encode_mel(fsb_codec *c, const float *mel, int32_t n_mel_frames, int64_t *codes, size_t cap,
                         size_t *out_len) {
    FSB_REQUIRE(c && mel && codes && out_len, FSB_ERR_INVALID, "encode_mel: null argument");
    FSB_REQUIRE(c->opt.with_encoder, FSB_ERR_STATE, "codec was created without the encoder");
    FSB_CUDA_OK(cudaSetDevice(c->opt.device));
    const int Lm = n_mel_frames;
    FSB_REQUIRE(Lm >= 4 && Lm <= 4 * c->max_frames, FSB_ERR_INVALID, "encode_mel: %d mel frames outside [4, %d]", Lm,
                4 * c->max_frames);
    // mel (160, Lm) staged in buf[2]
    FSB_CUDA_OK(cudaMemcpyAsync(c->buf[2], mel, (size_t)160 * Lm * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    return encode_from_device_mel(c, Lm, codes, cap, out_len);
}

static int encode_from_device_mel(fsb_codec *c, int Lm, int64_t *codes, size_t cap, size_t *out_len) {
    cudaStream_t st = c->stream;
    float *A = c->buf[0], *B = c->buf[1];
    // ConvNeXtEncoder::forward (convnext.rs:325-334)
    FSB_TRY(conv_fwd(c, c->stem, c->buf[2], Lm, A, 1, 1, false, nullptr, 0, 0.f, false, nullptr));
    ln_channels_first_kernel<<<(Lm + 3) / 4, 128, 0, st>>>(A, kEncDims[0], Lm, c->stem_ln_w, c->stem_ln_b, 1e-6f, B);
    CLAUNCH_CHECK(c);
    float *cur = B, *oth = A;
    for (int i = 0; i < 4; ++i) {
        if (i > 0) {
            ln_channels_first_kernel<<<(Lm + 3) / 4, 128, 0, st>>>(cur, kEncDims[i - 1], Lm, c->mid_ln_w[i],
                                                                   c->mid_ln_b[i], 1e-6f, oth);
            CLAUNCH_CHECK(c);
            // plain Conv1d 1x1 (convnext.rs:249-257): no causal padding is needed for k == 1
            FSB_TRY(conv_fwd(c, c->mid_conv[i], oth, Lm, cur, 1, 1, false, nullptr, 0, 0.f, false, nullptr));
        }
        for (auto &blk : c->enc_blocks[i]) FSB_TRY(convnext_fwd(c, blk, cur, Lm));
    }
    ln_channels_first_kernel<<<(Lm + 3) / 4, 128, 0, st>>>(cur, kDim, Lm, c->enc_norm_w, c->enc_norm_b, 1e-6f, oth);
    CLAUNCH_CHECK(c);
    std::swap(cur, oth);
    // DownsampleFiniteScalarQuantizer::encode (quantizer.rs:104-124)
    int L = Lm;
    for (int i = 0; i < 2; ++i) {
        int Lo = 0;
        FSB_TRY(conv_fwd(c, c->down_conv[i], cur, L, oth, 1, 2, false, nullptr, 0, 0.f, false, &Lo));
        L = Lo;
        FSB_TRY(convnext_fwd(c, c->down_block[i], oth, L));
        std::swap(cur, oth);
    }
    FSB_REQUIRE((size_t)L <= cap, FSB_ERR_INVALID, "encode_mel: capacity %zu < %d code frames", cap, L);
    fsq_encode_kernel<<<dim3((L + 127) / 128, kGroups), 128, 0, st>>>(cur, L, kGroups, c->proj_in_w, c->proj_in_b,
                                                                      c->d_idx);
    CLAUNCH_CHECK(c);
    std::vector<long long> host((size_t)kGroups * L);
    FSB_CUDA_OK(cudaMemcpyAsync(host.data(), c->d_idx, host.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
    FSB_CUDA_OK(cudaStreamSynchronize(st));
    for (int g = 0; g < kGroups; ++g)
        for (int t = 0; t < L; ++t) codes[(size_t)g * cap + t] = host[(size_t)g * L + t];
    *out_len = (size_t)L;
    return FSB_OK;
}

int fsb_codec_get_stats(fsb_codec *c, fsb_codec_stats *out) {
    FSB_REQUIRE(c && out, FSB_ERR_INVALID, "null argument");
    *out = c->stats;
    return FSB_OK;
}

int32_t fsb_codec_sample_rate(const fsb_codec *c) {
    (void)c;
    return 44100;  // codec/config.rs: every Fish >= 1.4 preset
}

}  // extern "C"
