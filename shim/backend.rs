//! Drop-in backend for `fish_speech_core::lm` / `fish_speech_core::codec` on top of libfsb.so.
//!
//! SOURCE ONLY (no cargo in the build image).  It keeps the reference's public names so the server
//! (`server/lib/handlers/speech.rs:19-68`), the CLIs and the PyO3 crate compile unchanged once
//! `fish_speech_core::lm::generate::generate_blocking` / `FireflyCodec::decode` are re-exported from here
//! behind a cargo feature (`b200`), next to the existing `cuda` / `metal` / `flash-attn` features
//! (`fish_speech_core/Cargo.toml:10-18`).
use super::fsb_sys::*;
use candle_core::{DType, Device, Result, Tensor};
use std::ffi::{CStr, CString};

fn check(status: i32) -> Result<()> {
    if status == FSB_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(fsb_last_error()) }.to_string_lossy().into_owned();
    candle_core::bail!("libfsb error {status}: {msg}")
}

/// Same fields the reference exposes (`dual_ar.rs:443-458`); the weights live in the library.
pub struct DualARTransformer {
    handle: *mut fsb_lm,
    pub cfg: crate::lm::dual_ar::BaseModelArgs,
    pub token_config: crate::lm::dual_ar::TokenConfig,
    pub model_type: crate::config::WhichLM,
}
unsafe impl Send for DualARTransformer {} // one generation in flight, behind the server's tokio Mutex (state.rs:13)

impl DualARTransformer {
    /// `DualARTransformer::load(&vb, &cfg, &token_config, model_type)` (dual_ar.rs:460): the caller hands
    /// over the mmapped safetensors views instead of a VarBuilder.
    pub fn load_from_views(
        names: &[String], views: &[(&[u8], DType, Vec<usize>)], cfg: &crate::lm::dual_ar::BaseModelArgs,
        token_config: &crate::lm::dual_ar::TokenConfig, model_type: crate::config::WhichLM, device_ordinal: i32,
        bf16: bool, max_batch: i32,
    ) -> Result<Self> {
        let cnames: Vec<CString> = names.iter().map(|n| CString::new(n.as_str()).unwrap()).collect();
        let table: Vec<fsb_tensor> = views
            .iter()
            .zip(&cnames)
            .map(|((bytes, dt, shape), n)| {
                let mut s = [0i64; 4];
                for (i, d) in shape.iter().enumerate() {
                    s[i] = *d as i64;
                }
                fsb_tensor {
                    name: n.as_ptr(),
                    data: bytes.as_ptr() as *const _,
                    dtype: if *dt == DType::BF16 { FSB_BF16 } else { FSB_F32 },
                    ndim: shape.len() as i32,
                    shape: s,
                    on_device: 0,
                }
            })
            .collect();
        let args = fsb_model_args {
            attention_qkv_bias: cfg.attention_qkv_bias as i32,
            codebook_size: cfg.codebook_size as i32,
            dim: cfg.dim as i32,
            head_dim: cfg.head_dim as i32,
            intermediate_size: cfg.intermediate_size.unwrap_or(0) as i32,
            max_seq_len: cfg.max_seq_len as i32,
            n_fast_layer: cfg.n_fast_layer as i32,
            n_head: cfg.n_head as i32,
            n_layer: cfg.n_layer as i32,
            n_local_heads: cfg.n_local_heads as i32,
            num_codebooks: cfg.num_codebooks as i32,
            vocab_size: cfg.vocab_size as i32,
            tie_word_embeddings: cfg.tie_word_embeddings as i32,
            norm_eps: cfg.norm_eps as f32,
            rope_base: cfg.rope_base as f32,
        };
        let tok = fsb_token_config {
            im_end_id: token_config.im_end_id,
            pad_id: token_config.pad_id,
            semantic_start_id: token_config.semantic_start_id,
            semantic_end_id: token_config.semantic_end_id.unwrap_or(0),
            has_semantic_end: token_config.semantic_end_id.is_some() as i32,
        };
        let opts = fsb_lm_options {
            device: device_ordinal,
            stream: std::ptr::null_mut(),
            weight_dtype: if bf16 { FSB_BF16 } else { FSB_F32 },
            max_batch,
            max_seq_len: 0,
            fish_version: match model_type {
                crate::config::WhichLM::Fish(crate::config::WhichFishVersion::Fish1_5) => FSB_FISH_1_5,
                _ => FSB_FISH_1_4,
            },
            decode_mode: 0,
        };
        let mut handle = std::ptr::null_mut();
        check(unsafe { fsb_lm_create(&args, &tok, table.as_ptr(), table.len(), &opts, &mut handle) })?;
        Ok(Self { handle, cfg: cfg.clone(), token_config: token_config.clone(), model_type })
    }

    pub fn clear_slow_layer_caches(&mut self) {
        unsafe { fsb_lm_clear_slow_layer_caches(self.handle) };
    }
    pub fn clear_fast_layer_caches(&mut self) {
        unsafe { fsb_lm_clear_fast_layer_caches(self.handle) };
    }
    pub fn clear_slow_caches_until(&mut self, pos: usize) -> Result<()> {
        check(unsafe { fsb_lm_clear_slow_caches_until(self.handle, pos) })
    }
    pub fn curr_kv_size(&self) -> Result<usize> {
        let mut n = 0usize;
        check(unsafe { fsb_lm_curr_kv_size(self.handle, &mut n) })?;
        Ok(n)
    }
}

impl Drop for DualARTransformer {
    fn drop(&mut self) {
        unsafe { fsb_lm_destroy(self.handle) };
    }
}

/// `generate_blocking(&mut model, &prompt, max_new_tokens, &sampling_args, show_progress)`
/// (lm/generate/single_batch.rs:308-324): prompt u32 (C+1, P) -> codes u32 (C, T).
pub fn generate_blocking(
    model: &mut DualARTransformer, prompt: &Tensor, max_new_tokens: usize,
    sampling_args: &crate::lm::sampling::SamplingArgs, _show_progress: bool,
) -> Result<Tensor> {
    let (rows, p) = prompt.dims2()?;
    let host: Vec<u32> = prompt.to_dtype(DType::U32)?.flatten_all()?.to_vec1()?;
    let c = rows - 1;
    let cap = max_new_tokens.saturating_sub(p) + 2;
    let mut out = vec![0u32; c * cap];
    let mut n = 0usize;
    let sa = fsb_sampling_args {
        temp: sampling_args.temp,
        top_p: sampling_args.top_p,
        top_k: sampling_args.top_k as u32,
        repetition_penalty: sampling_args.repetition_penalty,
        seed: rand::random::<u64>(), // single_batch.rs:46
    };
    check(unsafe {
        fsb_lm_generate_blocking(model.handle, host.as_ptr(), p as i32, max_new_tokens, &sa, 0, 0, out.as_mut_ptr(), cap, &mut n)
    })?;
    let mut codes = Vec::with_capacity(c * n);
    for r in 0..c {
        codes.extend_from_slice(&out[r * cap..r * cap + n]);
    }
    Tensor::from_vec(codes, (c, n), &Device::Cpu)
}

/// `FireflyCodec` (codec/firefly.rs:10-49)
pub struct FireflyCodec {
    handle: *mut fsb_codec,
    pub sample_rate: u32,
}
unsafe impl Send for FireflyCodec {}

impl FireflyCodec {
    /// `decode(&Tensor u32 (1, 8, T)) -> Tensor f32 (1, 1, 2048 T)` (firefly.rs:42-48)
    pub fn decode(&self, tokens: &Tensor) -> Result<Tensor> {
        let (_b, _g, t) = tokens.dims3()?;
        let host: Vec<u32> = tokens.to_dtype(DType::U32)?.flatten_all()?.to_vec1()?;
        let mut pcm = vec![0f32; 2048 * t];
        check(unsafe { fsb_codec_decode(self.handle, host.as_ptr(), t as i32, pcm.as_mut_ptr()) })?;
        Tensor::from_vec(pcm, (1, 1, 2048 * t), &Device::Cpu)
    }

    /// `encode(&Tensor f32 (1, 1, N)) -> Tensor i64 (1, 8, L)` (firefly.rs:36-39): log-mel, encoder and FSQ run on the GPU
    pub fn encode(&self, audio: &Tensor) -> Result<Tensor> {
        let host: Vec<f32> = audio.to_dtype(DType::F32)?.flatten_all()?.to_vec1()?;
        let cap = host.len() / 2048 + 8;
        let mut codes = vec![0i64; 8 * cap];
        let mut n = 0usize;
        check(unsafe { fsb_codec_encode(self.handle, host.as_ptr(), host.len() as i64, codes.as_mut_ptr(), cap, &mut n) })?;
        let rows: Vec<i64> = (0..8).flat_map(|g| codes[g * cap..g * cap + n].to_vec()).collect();
        Tensor::from_vec(rows, (1, 8, n), &Device::Cpu)
    }
}

impl Drop for FireflyCodec {
    fn drop(&mut self) {
        unsafe { fsb_codec_destroy(self.handle) };
    }
}
