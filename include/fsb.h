/*
 * fsb.h -- C ABI of the B200-native fish-speech hot path ("fsb").
 *
 * Drop-in boundary for fish-speech.rs (EndlessReform/fish-speech.rs @ e5a172a):
 * every entry point below replaces one method/function the reference's Rust
 * callers use today on `fish_speech_core::lm` / `fish_speech_core::codec`; the
 * reference location is cited per declaration (paths relative to the reference
 * root).  Plain C types only: pointers, sizes, PODs.  No torch / candle types.
 *
 * Conventions
 *   - every function returns an `fsb_status` (0 == ok, negative == error) and
 *     never throws or aborts across the ABI; `fsb_last_error()` holds the
 *     message of the last failure on the calling thread (reference: `bail!` /
 *     `candle_core::Error`, server/lib/handlers/error.rs:17-34).
 *   - a handle is NOT thread-safe (the reference serialises every generation
 *     behind `Arc<tokio::sync::Mutex<DualARTransformer>>`, server/lib/state.rs:13);
 *     distinct handles (one per GPU) are independent.
 *   - the library owns weights (copied at create), the KV arena and scratch;
 *     the caller owns every in/out buffer.  Nothing is allocated per frame.
 *   - pointers are HOST pointers unless the parameter is named `*_dev` or the
 *     call takes FSB_FLAG_DEVICE_PTRS.
 *   - there is NO CPU fallback: every compute entry point fails with
 *     FSB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef FSB_H
#define FSB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSB_ABI_VERSION 1

typedef enum fsb_status {
    FSB_OK = 0,
    FSB_ERR_INVALID = -1,        /* bad argument (null, out of range) */
    FSB_ERR_CUDA = -2,           /* CUDA runtime/driver error; sticky ones poison the handle */
    FSB_ERR_MISSING_WEIGHT = -3, /* a tensor the loader needs is not in the table */
    FSB_ERR_SHAPE = -4,          /* a tensor has the wrong shape / dtype */
    FSB_ERR_STATE = -5,          /* call not valid in the handle's state (e.g. KV overflow) */
    FSB_ERR_UNSUPPORTED = -6,    /* configuration outside what the kernels implement */
    FSB_ERR_OOM = -7
} fsb_status;

typedef enum fsb_dtype { FSB_F32 = 0, FSB_BF16 = 1, FSB_F16 = 2, FSB_U32 = 3, FSB_I64 = 4, FSB_U8 = 5, FSB_F64 = 6 } fsb_dtype;

/* fish_speech_core/lib/config.rs:3-8 (WhichFishVersion) */
typedef enum fsb_fish_version { FSB_FISH_1_2 = 12, FSB_FISH_1_4 = 14, FSB_FISH_1_5 = 15 } fsb_fish_version;

/* One named tensor of a checkpoint (what `VarBuilder::get(shape, name)` yields,
 * dual_ar.rs:125-156,466-511; codec/utils/mod.rs:25-40).  `data` is a host
 * pointer unless `on_device` != 0.  Row-major, contiguous. */
typedef struct fsb_tensor {
    const char *name;
    const void *data;
    int32_t dtype; /* fsb_dtype: FSB_F32 or FSB_BF16 */
    int32_t ndim;
    int64_t shape[4];
    int32_t on_device;
} fsb_tensor;

/* BaseModelArgs, dual_ar.rs:56-81 (serde field names kept). */
typedef struct fsb_model_args {
    int32_t attention_qkv_bias; /* must be 0 (no Fish checkpoint sets it) */
    int32_t codebook_size;
    int32_t dim;
    int32_t head_dim;
    int32_t intermediate_size; /* 0 == dim * 4 (Option::None) */
    int32_t max_seq_len;
    int32_t n_fast_layer;
    int32_t n_head;
    int32_t n_layer;
    int32_t n_local_heads;
    int32_t num_codebooks;
    int32_t vocab_size;
    int32_t tie_word_embeddings;
    float norm_eps;
    float rope_base;
} fsb_model_args;

/* TokenConfig, dual_ar.rs:17-23. */
typedef struct fsb_token_config {
    uint32_t im_end_id;
    uint32_t pad_id;
    uint32_t semantic_start_id;
    uint32_t semantic_end_id;
    int32_t has_semantic_end; /* Option<u32>::is_some(): 1 for Fish 1.5 / DualAR */
} fsb_token_config;

/* SamplingArgs, sampling/mod.rs:28-34, plus the Philox seed that replaces
 * `rand::random::<u64>()` (single_batch.rs:46) / 42 (static_batch.rs:63). */
typedef struct fsb_sampling_args {
    double temp; /* <= 1e-7 == argmax */
    double top_p;
    uint32_t top_k;
    float repetition_penalty;
    uint64_t seed;
} fsb_sampling_args;

typedef struct fsb_lm_options {
    int32_t device;        /* CUDA ordinal (reference: Device::cuda_if_available(0), server/src/main.rs:25) */
    void *stream;          /* cudaStream_t to run on; NULL == the handle creates its own */
    int32_t weight_dtype;  /* FSB_F32 (parity mode) or FSB_BF16 (weights stored bf16; activations, KV cache and all math stay fp32) */
    int32_t max_batch;     /* rows the KV arena is sized for (>= 1) */
    int32_t max_seq_len;   /* positions per row in the KV arena; 0 == model max_seq_len */
    int32_t fish_version;  /* fsb_fish_version */
    int32_t decode_mode;   /* 0 == auto, 1 == per-op kernels + CUDA graph, 2 == persistent megakernel */
} fsb_lm_options;

/* flags for the generate calls */
#define FSB_GEN_FIXED_LEN 0x1u   /* <|im_end|> not eligible; emit exactly `fixed_len` frames (bench harness) */
#define FSB_GEN_KEEP_SLOW_KV 0x2u /* do not clear the slow KV before prefill: prefix reuse (speech.rs:40) */

typedef struct fsb_lm_stats {
    double prefill_ms;     /* device time of the last generate call's prefill (CUDA events on the handle's stream) */
    double decode_ms;      /* device time of its frame loop */
    uint64_t frames;       /* frames produced by the last generate call, summed over rows */
    uint64_t kernel_launches; /* kernels this library launched in the last generate call */
    double dominant_kernel_ms; /* summed device time of the weight-streaming kernel(s), when FSB_PROFILE=1 */
    uint64_t dominant_kernel_launches;
    uint64_t weight_bytes_per_frame; /* algorithmic bytes one decode frame must stream */
    uint64_t dominant_kernel_bytes;  /* algorithmic weight bytes the profiled launches streamed */
} fsb_lm_stats;

typedef struct fsb_lm fsb_lm;
typedef struct fsb_codec fsb_codec;

/* ---- library ------------------------------------------------------------ */
int fsb_abi_version(void);
const char *fsb_last_error(void);
/* number of usable sm_100 devices (0 when none: every compute call will fail loudly) */
int fsb_device_count(void);

/* ---- DualARTransformer (fish_speech_core/lib/lm/dual_ar.rs) ------------- */

/* DualARTransformer::load, dual_ar.rs:460-529.  Copies (and converts to
 * opts->weight_dtype) every tensor it needs out of `weights`. */
int fsb_lm_create(const fsb_model_args *args, const fsb_token_config *tok, const fsb_tensor *weights,
                  size_t n_weights, const fsb_lm_options *opts, fsb_lm **out);
int fsb_lm_destroy(fsb_lm *lm);

/* forward_generate, dual_ar.rs:574-635.  inp: u32 (B, C+1, S).  All rows share
 * `input_pos` (as in the reference).  logits: f32 (B, 1, vocab) or NULL;
 * hidden: f32 (B, 1, dim), the PRE-norm last-position state (Q1), or NULL. */
int fsb_lm_forward_generate(fsb_lm *lm, const uint32_t *inp, int32_t bsz, int32_t seq_len, size_t input_pos,
                            float *logits, float *hidden);

/* forward_generate_fast, dual_ar.rs:638-673.  x: f32 (B, 1, dim) -> logits f32 (B, 1, codebook_size). */
int fsb_lm_forward_generate_fast(fsb_lm *lm, const float *x, int32_t bsz, size_t input_pos, float *logits);

/* `model.fast_embeddings.forward(ids)` (pub field, dual_ar.rs:447; single_batch.rs:176-182). out f32 (n, dim). */
int fsb_lm_fast_embeddings(fsb_lm *lm, const uint32_t *ids, int32_t n, float *out);

int fsb_lm_clear_fast_layer_caches(fsb_lm *lm);                 /* dual_ar.rs:675-679 */
int fsb_lm_clear_slow_layer_caches(fsb_lm *lm);                 /* dual_ar.rs:681-685 */
int fsb_lm_clear_slow_caches_until(fsb_lm *lm, size_t pos);     /* dual_ar.rs:687-693 (O(1) here: a length, not a copy) */
int fsb_lm_curr_kv_size(fsb_lm *lm, size_t *out);               /* dual_ar.rs:695-700 (row 0) */

/* generate_blocking, single_batch.rs:308-324 (== generate_blocking_with_hidden
 * without hidden states).  prompt: u32 (C+1, P).  out_codes: u32 (C, cap)
 * row-major with row stride `cap`; *out_len = T frames written (T <= cap).
 * The whole frame loop (slow step, constrained sampling, 8 fast steps, rep-pen,
 * fast_embeddings gather) runs on the device without host round trips. */
int fsb_lm_generate_blocking(fsb_lm *lm, const uint32_t *prompt, int32_t prompt_len, size_t max_new_tokens,
                             const fsb_sampling_args *sampling, uint32_t flags, int32_t fixed_len,
                             uint32_t *out_codes, size_t cap, size_t *out_len);

/* generate_blocking_with_hidden(collect_hidden_states = true), single_batch.rs:217-306 (the server's hidden-state
 * export, server/lib/handlers/speech.rs:24-48): same generation, plus the PRE-norm slow hidden state handed to the fast
 * stack (Q1) of EVERY yielded frame, <|im_end|> frames included (single_batch.rs:251,268-270).
 * hidden: f32 (hidden_cap, dim) row-major; *n_hidden = frames written (== fsb_lm_last_frames count). */
int fsb_lm_generate_blocking_with_hidden(fsb_lm *lm, const uint32_t *prompt, int32_t prompt_len, size_t max_new_tokens,
                                         const fsb_sampling_args *sampling, uint32_t flags, int32_t fixed_len,
                                         uint32_t *out_codes, size_t cap, size_t *out_len, float *hidden,
                                         size_t hidden_cap, size_t *n_hidden);

/* generate_static_batch, static_batch.rs:282-390, with "independent utterances"
 * semantics: row i is exactly fsb_lm_generate_blocking on prompts[i] with Philox
 * row index i (per-row positions and KV lengths; the reference's unmasked left
 * padding, SURVEY Q7, is NOT reproduced).  out_codes[i]: u32 (C, cap). */
int fsb_lm_generate_static_batch(fsb_lm *lm, const uint32_t *const *prompts, const int32_t *prompt_lens,
                                 int32_t bsz, size_t max_new_tokens, const fsb_sampling_args *sampling,
                                 uint32_t flags, int32_t fixed_len, uint32_t *const *out_codes, size_t cap,
                                 size_t *out_lens);

/* Every frame the last generate call emitted for batch row `row`, as `SingleBatchGenerator::next` yields them
 * (single_batch.rs:76-214: a (C+1, 1) column per frame: the slow/semantic token, then the C codes), INCLUDING
 * <|im_end|> frames that generate_blocking drops (single_batch.rs:262-266).  out: u32 (C+1, cap) row-major with row
 * stride `cap`; *out_len = frames written.  Lets a caller (and the parity tests) replay a generation step by step. */
int fsb_lm_last_frames(fsb_lm *lm, int32_t row, uint32_t *out, size_t cap, size_t *out_len);

/* ---- per-voice conditioning KV ------------------------------------------------------------------
 * The server keeps the system + voice KV between the chunks of ONE request (`clear_slow_caches_until(n_cond)`,
 * server/lib/handlers/speech.rs:40) and re-prefills it whenever the voice changes.  A snapshot holds the first
 * `n_positions` cached positions of a row (every slow block, K and V) on the device and can be restored into any row:
 * prefill the voice prompt once, snapshot, and start every later utterance of that voice with a restore +
 * FSB_GEN_KEEP_SLOW_KV generation of the remaining text columns. */
typedef struct fsb_kv_snapshot fsb_kv_snapshot;
int fsb_lm_kv_snapshot_save(fsb_lm *lm, int32_t row, size_t n_positions, fsb_kv_snapshot **out);
int fsb_lm_kv_snapshot_restore(fsb_lm *lm, const fsb_kv_snapshot *snap, int32_t row); /* row's KV length := n_positions */
int fsb_lm_kv_snapshot_free(fsb_lm *lm, fsb_kv_snapshot *snap);

/* ---- continuous batching ---------------------------------------------------------------------------
 * The reference serialises whole generations behind one Mutex (server/lib/state.rs:13) and its static batch waits for
 * the slowest row (static_batch.rs:160-173).  A session turns the `max_batch` rows of the handle into slots: an utterance
 * is admitted into a free slot between two runs, decodes together with whatever else is in flight, and is collected as
 * soon as it ends (<|im_end|> / budget) while the other slots keep going.  Row semantics are those of
 * fsb_lm_generate_blocking with Philox row index == slot.  Needs the wide-batch kernel (bf16 weights, Fish shapes,
 * max_batch >= 9); returns FSB_ERR_UNSUPPORTED otherwise.
 *   begin   : sampling parameters + flags (FSB_GEN_FIXED_LEN) of every utterance of the session; all slots free
 *   admit   : prefill `prompt` into `slot` and emit its first frame (the frame produced from the prompt itself)
 *   run     : up to `max_frames` more frames for every live slot in ONE launch; active[slot] = 1 while a slot is still
 *             generating (NULL allowed); *n_active = live slots left
 *   collect : codes u32 (C, cap) of a finished slot (same filtering as generate_blocking) and frees the slot */
int fsb_lm_session_begin(fsb_lm *lm, const fsb_sampling_args *sampling, uint32_t flags);
int fsb_lm_session_admit(fsb_lm *lm, int32_t slot, const uint32_t *prompt, int32_t prompt_len, size_t max_new_tokens,
                         int32_t fixed_len);
int fsb_lm_session_run(fsb_lm *lm, int32_t max_frames, int32_t *active, int32_t *n_active);
int fsb_lm_session_collect(fsb_lm *lm, int32_t slot, uint32_t *out_codes, size_t cap, size_t *out_len);

int fsb_lm_get_stats(fsb_lm *lm, fsb_lm_stats *out);
/* Profiling switch (the reference's only instrumentation is Instant + println!, single_batch.rs:233-303).
 * on != 0: generate calls run the frame loop eagerly and bracket every launch of the dominant
 * (weight-streaming) kernel with CUDA events on the handle's stream, for the first frames of the call;
 * the sums land in fsb_lm_stats.dominant_kernel_{ms,launches,bytes}.  Off by default (graphs). */
int fsb_lm_set_profile(fsb_lm *lm, int on);

/* ---- FireflyCodec (fish_speech_core/lib/codec/firefly.rs) --------------- */

typedef struct fsb_codec_options {
    int32_t device;
    void *stream;
    int32_t fish_version; /* FSB_FISH_1_4 / FSB_FISH_1_5 (causal convs); 1.2 unsupported */
    int32_t max_frames;   /* largest T one decode call may carry (scratch sizing) */
    int32_t with_encoder; /* load the ConvNeXt encoder + downsample path too */
} fsb_codec_options;

typedef struct fsb_codec_stats {
    double decode_ms;          /* last decode call, host copies included */
    double device_ms;          /* same call, kernels only (CUDA events on the handle's stream) */
    uint64_t kernel_launches;
    double dominant_kernel_ms; /* ResBlock conv kernels, when FSB_PROFILE=1 */
    uint64_t dominant_kernel_launches;
} fsb_codec_stats;

/* FireflyCodec::load, firefly.rs:20-34 (weights F32, load.rs:161-164). */
int fsb_codec_create(const fsb_tensor *weights, size_t n_weights, const fsb_codec_options *opts, fsb_codec **out);
int fsb_codec_destroy(fsb_codec *codec);
/* FireflyCodec::decode, firefly.rs:42-48.  codes: u32 (1, 8, T) -> pcm f32 (1, 1, 2048*T), 44.1 kHz.
 * Codes >= 1000 are rejected (FSB_ERR_INVALID), as the reference's gather would (Q11). */
int fsb_codec_decode(fsb_codec *codec, const uint32_t *codes, int32_t n_frames, float *pcm);
/* n independent batch-1 decodes (Q9: the reference decode is only valid for b == 1;
 * the server concatenates along time instead, speech.rs:90-91). */
int fsb_codec_decode_batch(fsb_codec *codec, const uint32_t *const *codes, const int32_t *n_frames, int32_t n,
                           float *const *pcm);
/* Streaming output stage (server/src/handlers/speech.rs:180-236 vocodes and ships audio block by block).
 * Frames [t0, t1) of an utterance whose codes (8, n_frames_total) are given in full: every convolution of the decoder
 * is causal and the receptive field is 14.7 code frames, so the block is decoded with a 16-frame left halo and the
 * result is bit-identical to the same samples of a whole-utterance fsb_codec_decode. */
int fsb_codec_decode_block(fsb_codec *codec, const uint32_t *codes, int32_t n_frames_total, int32_t t0, int32_t t1,
                           float *pcm);
/* Same block, finished on the device for the wire: optional linear-interpolation resample 44.1 kHz -> to_rate
 * (audio/functional.rs:3-37; 0 or 44100 = none; the server uses 24 kHz for Opus, opus.rs:12-93) and the f32 -> s16
 * conversion of audio/wav.rs:9-13; half the bytes of the f32 PCM cross PCIe. */
int fsb_codec_decode_block_s16(fsb_codec *codec, const uint32_t *codes, int32_t n_frames_total, int32_t t0, int32_t t1,
                               uint32_t to_rate, int16_t *out, size_t cap, size_t *out_len);
/* FireflyEncoder::encode on a log-mel input, encoder.rs:38-42: mel f32 (1, 160, Lm) -> i64 (1, 8, L). */
int fsb_codec_encode_mel(fsb_codec *codec, const float *mel, int32_t n_mel_frames, int64_t *codes, size_t cap,
                         size_t *out_len);
/* LogMelSpectrogram::forward, audio/spectrogram.rs:141-158 (on the streaming STFT of audio/stft.rs:52-90): mono
 * 44.1 kHz pcm f32 (n_samples) -> log-mel f32 (1, 160, Lm), Lm = fsb-internal frame count of the reference's
 * chunked STFT (1099 for the 562 265 samples of tests/resources/sky.wav).  Needs with_encoder. */
int fsb_codec_log_mel(fsb_codec *codec, const float *pcm, int64_t n_samples, float *mel, size_t cap_frames,
                      size_t *out_frames);
/* FireflyCodec::encode, firefly.rs:36-39: pcm f32 (1, 1, n_samples) -> i64 (1, 8, L); the mel never leaves the device. */
int fsb_codec_encode(fsb_codec *codec, const float *pcm, int64_t n_samples, int64_t *codes, size_t cap, size_t *out_len);
int fsb_codec_get_stats(fsb_codec *codec, fsb_codec_stats *out);
int32_t fsb_codec_sample_rate(const fsb_codec *codec); /* FireflyCodec.sample_rate, firefly.rs:13 */

/* ---- operator level (candle_core::CustomOp1 precedent) ------------------- */

/* RepeatKV::cuda_fwd, lm/ops/repeat_kv.rs:30-101 + candle-gqa-kernels/src/unary.cu:8-58:
 * dst[(h*n_rep + r), s, d] = src[h, s, d].  Contiguous device buffers, bsz == 1
 * like the reference (repeat_kv.rs:81-83).  Kept for callers that still want the
 * materialised form; the fused attention below makes it unnecessary. */
int fsb_op_repeat_kv(const void *src_dev, void *dst_dev, int32_t dtype, int32_t n_local_heads, int32_t n_rep,
                     int32_t seqlen, int32_t head_dim, void *stream);

/* Fused replacement of rope_i + Tensor::cat + repeat_kv + matmul + softmax + matmul
 * (dual_ar.rs:239-249,316-376) for one decode step of `bsz` rows.
 *   qkv_dev   f32 (bsz, (n_head + 2*n_local_heads) * head_dim)  output of wqkv
 *   k/v cache f32 (bsz, n_local_heads, max_len, head_dim), appended in place at pos[b]
 *   cos/sin   f32 (max_pos, head_dim/2)
 *   pos_dev   i32 (bsz): rows already cached == position of the new token
 *   out_dev   f32 (bsz, n_head * head_dim)
 */
int fsb_op_gqa_decode_attn(const float *qkv_dev, float *kcache_dev, float *vcache_dev, const float *cos_dev,
                           const float *sin_dev, const int32_t *pos_dev, int32_t bsz, int32_t n_head,
                           int32_t n_local_heads, int32_t head_dim, int32_t max_len, float *out_dev,
                           void *scratch_dev, size_t scratch_bytes, void *stream);
size_t fsb_op_gqa_decode_attn_scratch_bytes(int32_t bsz, int32_t n_head, int32_t head_dim);

#ifdef __cplusplus
}
#endif
#endif /* FSB_H */
