//! Raw FFI declarations for libfsb.so (include/fsb.h, ABI version 1).
//!
//! SOURCE ONLY: there is no Rust toolchain in the build image, so this file is not compiled or
//! tested here; it is the binding a fish-speech.rs maintainer would add.
//! It follows the reference's own precedent for native code behind Candle:
//! `fish_speech_core/lib/lm/ops/repeat_kv.rs` (CustomOp1 over a cudarc launch).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const FSB_OK: c_int = 0;
pub const FSB_F32: i32 = 0;
pub const FSB_BF16: i32 = 1;
pub const FSB_FISH_1_4: i32 = 14;
pub const FSB_FISH_1_5: i32 = 15;
pub const FSB_GEN_FIXED_LEN: u32 = 0x1;
pub const FSB_GEN_KEEP_SLOW_KV: u32 = 0x2;

#[repr(C)]
pub struct fsb_tensor {
    pub name: *const c_char,
    pub data: *const c_void,
    pub dtype: i32,
    pub ndim: i32,
    pub shape: [i64; 4],
    pub on_device: i32,
}

/// BaseModelArgs, fish_speech_core/lib/lm/dual_ar.rs:56-81
#[repr(C)]
pub struct fsb_model_args {
    pub attention_qkv_bias: i32,
    pub codebook_size: i32,
    pub dim: i32,
    pub head_dim: i32,
    pub intermediate_size: i32,
    pub max_seq_len: i32,
    pub n_fast_layer: i32,
    pub n_head: i32,
    pub n_layer: i32,
    pub n_local_heads: i32,
    pub num_codebooks: i32,
    pub vocab_size: i32,
    pub tie_word_embeddings: i32,
    pub norm_eps: f32,
    pub rope_base: f32,
}

/// TokenConfig, dual_ar.rs:17-23
#[repr(C)]
pub struct fsb_token_config {
    pub im_end_id: u32,
    pub pad_id: u32,
    pub semantic_start_id: u32,
    pub semantic_end_id: u32,
    pub has_semantic_end: i32,
}

/// SamplingArgs, sampling/mod.rs:28-34 (+ Philox seed)
#[repr(C)]
pub struct fsb_sampling_args {
    pub temp: f64,
    pub top_p: f64,
    pub top_k: u32,
    pub repetition_penalty: f32,
    pub seed: u64,
}

#[repr(C)]
pub struct fsb_lm_options {
    pub device: i32,
    pub stream: *mut c_void,
    pub weight_dtype: i32,
    pub max_batch: i32,
    pub max_seq_len: i32,
    pub fish_version: i32,
    pub decode_mode: i32,
}

#[repr(C)]
pub struct fsb_codec_options {
    pub device: i32,
    pub stream: *mut c_void,
    pub fish_version: i32,
    pub max_frames: i32,
    pub with_encoder: i32,
}

pub enum fsb_lm {}
pub enum fsb_codec {}
pub enum fsb_kv_snapshot {}

/// fsb_lm_stats (include/fsb.h), same field order
#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct fsb_lm_stats {
    pub prefill_ms: f64,
    pub decode_ms: f64,
    pub frames: u64,
    pub kernel_launches: u64,
    pub dominant_kernel_ms: f64,
    pub dominant_kernel_launches: u64,
    pub weight_bytes_per_frame: u64,
    pub dominant_kernel_bytes: u64,
}

/// fsb_codec_stats (include/fsb.h), same field order
#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct fsb_codec_stats {
    pub decode_ms: f64,
    pub device_ms: f64,
    pub kernel_launches: u64,
    pub dominant_kernel_ms: f64,
    pub dominant_kernel_launches: u64,
}

#[link(name = "fsb")]
extern "C" {
    pub fn fsb_abi_version() -> c_int;
    pub fn fsb_last_error() -> *const c_char;
    pub fn fsb_device_count() -> c_int;

    pub fn fsb_lm_create(
        args: *const fsb_model_args, tok: *const fsb_token_config, weights: *const fsb_tensor, n_weights: usize,
        opts: *const fsb_lm_options, out: *mut *mut fsb_lm,
    ) -> c_int;
    pub fn fsb_lm_destroy(lm: *mut fsb_lm) -> c_int;
    pub fn fsb_lm_forward_generate(
        lm: *mut fsb_lm, inp: *const u32, bsz: i32, seq_len: i32, input_pos: usize, logits: *mut f32, hidden: *mut f32,
    ) -> c_int;
    pub fn fsb_lm_forward_generate_fast(lm: *mut fsb_lm, x: *const f32, bsz: i32, input_pos: usize, logits: *mut f32) -> c_int;
    pub fn fsb_lm_fast_embeddings(lm: *mut fsb_lm, ids: *const u32, n: i32, out: *mut f32) -> c_int;
    pub fn fsb_lm_clear_fast_layer_caches(lm: *mut fsb_lm) -> c_int;
    pub fn fsb_lm_clear_slow_layer_caches(lm: *mut fsb_lm) -> c_int;
    pub fn fsb_lm_clear_slow_caches_until(lm: *mut fsb_lm, pos: usize) -> c_int;
    pub fn fsb_lm_curr_kv_size(lm: *mut fsb_lm, out: *mut usize) -> c_int;
    pub fn fsb_lm_generate_blocking(
        lm: *mut fsb_lm, prompt: *const u32, prompt_len: i32, max_new_tokens: usize, sampling: *const fsb_sampling_args,
        flags: u32, fixed_len: i32, out_codes: *mut u32, cap: usize, out_len: *mut usize,
    ) -> c_int;
    pub fn fsb_lm_generate_blocking_with_hidden(
        lm: *mut fsb_lm, prompt: *const u32, prompt_len: i32, max_new_tokens: usize, sampling: *const fsb_sampling_args,
        flags: u32, fixed_len: i32, out_codes: *mut u32, cap: usize, out_len: *mut usize, hidden: *mut f32,
        hidden_cap: usize, n_hidden: *mut usize,
    ) -> c_int;
    pub fn fsb_lm_last_frames(lm: *mut fsb_lm, row: i32, out: *mut u32, cap: usize, out_len: *mut usize) -> c_int;
    pub fn fsb_lm_kv_snapshot_save(lm: *mut fsb_lm, row: i32, n_positions: usize, out: *mut *mut fsb_kv_snapshot) -> c_int;
    pub fn fsb_lm_kv_snapshot_restore(lm: *mut fsb_lm, snap: *const fsb_kv_snapshot, row: i32) -> c_int;
    pub fn fsb_lm_kv_snapshot_free(lm: *mut fsb_lm, snap: *mut fsb_kv_snapshot) -> c_int;
    pub fn fsb_lm_session_begin(lm: *mut fsb_lm, sampling: *const fsb_sampling_args, flags: u32) -> c_int;
    pub fn fsb_lm_session_admit(
        lm: *mut fsb_lm, slot: i32, prompt: *const u32, prompt_len: i32, max_new_tokens: usize, fixed_len: i32,
    ) -> c_int;
    pub fn fsb_lm_session_run(lm: *mut fsb_lm, max_frames: i32, active: *mut i32, n_active: *mut i32) -> c_int;
    pub fn fsb_lm_session_collect(lm: *mut fsb_lm, slot: i32, out_codes: *mut u32, cap: usize, out_len: *mut usize) -> c_int;
    pub fn fsb_lm_get_stats(lm: *mut fsb_lm, out: *mut fsb_lm_stats) -> c_int;
    pub fn fsb_lm_set_profile(lm: *mut fsb_lm, on: c_int) -> c_int;
    pub fn fsb_lm_generate_static_batch(
        lm: *mut fsb_lm, prompts: *const *const u32, prompt_lens: *const i32, bsz: i32, max_new_tokens: usize,
        sampling: *const fsb_sampling_args, flags: u32, fixed_len: i32, out_codes: *const *mut u32, cap: usize,
        out_lens: *mut usize,
    ) -> c_int;

    pub fn fsb_codec_create(weights: *const fsb_tensor, n_weights: usize, opts: *const fsb_codec_options, out: *mut *mut fsb_codec) -> c_int;
    pub fn fsb_codec_destroy(codec: *mut fsb_codec) -> c_int;
    pub fn fsb_codec_decode(codec: *mut fsb_codec, codes: *const u32, n_frames: i32, pcm: *mut f32) -> c_int;
    pub fn fsb_codec_decode_batch(
        codec: *mut fsb_codec, codes: *const *const u32, n_frames: *const i32, n: i32, pcm: *const *mut f32,
    ) -> c_int;
    pub fn fsb_codec_get_stats(codec: *mut fsb_codec, out: *mut fsb_codec_stats) -> c_int;
    pub fn fsb_codec_decode_block(codec: *mut fsb_codec, codes: *const u32, n_frames_total: i32, t0: i32, t1: i32, pcm: *mut f32) -> c_int;
    pub fn fsb_codec_decode_block_s16(codec: *mut fsb_codec, codes: *const u32, n_frames_total: i32, t0: i32, t1: i32, to_rate: u32, out: *mut i16, cap: usize, out_len: *mut usize) -> c_int;
    pub fn fsb_codec_log_mel(codec: *mut fsb_codec, pcm: *const f32, n_samples: i64, mel: *mut f32, cap_frames: usize, out_frames: *mut usize) -> c_int;
    pub fn fsb_codec_encode(codec: *mut fsb_codec, pcm: *const f32, n_samples: i64, codes: *mut i64, cap: usize, out_len: *mut usize) -> c_int;
    pub fn fsb_codec_encode_mel(codec: *mut fsb_codec, mel: *const f32, n_mel_frames: i32, codes: *mut i64, cap: usize, out_len: *mut usize) -> c_int;
    pub fn fsb_codec_sample_rate(codec: *const fsb_codec) -> i32;

    pub fn fsb_op_repeat_kv(
        src_dev: *const c_void, dst_dev: *mut c_void, dtype: i32, n_local_heads: i32, n_rep: i32, seqlen: i32,
        head_dim: i32, stream: *mut c_void,
    ) -> c_int;
    pub fn fsb_op_gqa_decode_attn(
        qkv_dev: *const f32, kcache_dev: *mut f32, vcache_dev: *mut f32, cos_dev: *const f32, sin_dev: *const f32,
        pos_dev: *const i32, bsz: i32, n_head: i32, n_local_heads: i32, head_dim: i32, max_len: i32, out_dev: *mut f32,
        scratch_dev: *mut c_void, scratch_bytes: usize, stream: *mut c_void,
    ) -> c_int;
    pub fn fsb_op_gqa_decode_attn_scratch_bytes(bsz: i32, n_head: i32, head_dim: i32) -> usize;
}
