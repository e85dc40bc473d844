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

/// Name + shape of every tensor `DualARTransformer::load` binds (dual_ar.rs:466-511, per block :125-156,219-236):
/// the 203 entries of `docs/llama-weight-dict.txt` for Fish 1.2, the same pattern for 1.4 / 1.5.
fn lm_tensor_specs(cfg: &crate::lm::dual_ar::BaseModelArgs) -> Vec<(String, Vec<usize>)> {
    let (d, i) = (cfg.dim, cfg.intermediate_size.unwrap_or(4 * cfg.dim));
    let qkv = (cfg.n_head + 2 * cfg.n_local_heads) * cfg.head_dim;
    let mut v = vec![
        ("embeddings.weight".to_string(), vec![cfg.vocab_size, d]),
        ("codebook_embeddings.weight".to_string(), vec![cfg.num_codebooks * cfg.codebook_size, d]),
        ("norm.weight".to_string(), vec![d]),
        ("fast_embeddings.weight".to_string(), vec![cfg.codebook_size, d]),
        ("fast_norm.weight".to_string(), vec![d]),
        ("fast_output.weight".to_string(), vec![cfg.codebook_size, d]),
    ];
    if !cfg.tie_word_embeddings {
        v.push(("output.weight".to_string(), vec![cfg.vocab_size, d])); // dual_ar.rs:486-490
    }
    for (prefix, n) in [("layers", cfg.n_layer), ("fast_layers", cfg.n_fast_layer)] {
        for l in 0..n {
            let p = format!("{prefix}.{l}.");
            v.push((format!("{p}attention.wqkv.weight"), vec![qkv, d]));
            v.push((format!("{p}attention.wo.weight"), vec![d, cfg.n_head * cfg.head_dim]));
            v.push((format!("{p}feed_forward.w1.weight"), vec![i, d]));
            v.push((format!("{p}feed_forward.w2.weight"), vec![d, i]));
            v.push((format!("{p}feed_forward.w3.weight"), vec![i, d]));
            v.push((format!("{p}ffn_norm.weight"), vec![d]));
            v.push((format!("{p}attention_norm.weight"), vec![d]));
        }
    }
    v
}

/// Pulls every tensor of `specs` out of the VarBuilder as host bytes (`Tensor::to_vec1` of the flattened tensor, bf16
/// kept as bf16) and describes it to the library.  The returned vectors own the storage the table points into.
fn tensor_table(
    vb: &candle_nn::VarBuilder, specs: &[(String, Vec<usize>)],
) -> Result<(Vec<CString>, Vec<Vec<u8>>, Vec<fsb_tensor>)> {
    let mut names = Vec::with_capacity(specs.len());
    let mut blobs: Vec<Vec<u8>> = Vec::with_capacity(specs.len());
    let mut table = Vec::with_capacity(specs.len());
    for (name, shape) in specs {
        let t = vb.get(shape.as_slice(), name)?.to_device(&Device::Cpu)?.flatten_all()?;
        let (bytes, dtype): (Vec<u8>, i32) = match t.dtype() {
            DType::BF16 => (t.to_vec1::<half::bf16>()?.iter().flat_map(|x| x.to_bits().to_le_bytes()).collect(), FSB_BF16),
            _ => (t.to_dtype(DType::F32)?.to_vec1::<f32>()?.iter().flat_map(|x| x.to_le_bytes()).collect(), FSB_F32),
        };
        let mut s = [0i64; 4];
        for (k, dim) in shape.iter().enumerate() {
            s[k] = *dim as i64;
        }
        names.push(CString::new(name.as_str()).unwrap());
        blobs.push(bytes);
        table.push(fsb_tensor {
            name: names.last().unwrap().as_ptr(),
            data: blobs.last().unwrap().as_ptr() as *const _,
            dtype,
            ndim: shape.len() as i32,
            shape: s,
            on_device: 0,
        });
    }
    Ok((names, blobs, table))
}

fn device_ordinal(dev: &Device) -> i32 {
    match dev {
        Device::Cuda(c) => c.ordinal() as i32, // candle_core::CudaDevice
        _ => 0,
    }
}

impl DualARTransformer {
    /// `DualARTransformer::load(&vb, &cfg, &token_config, model_type)`, dual_ar.rs:460-529 -- the reference's signature.
    /// The VarBuilder's dtype picks `weight_dtype` (bf16 on CUDA, `server/src/main.rs:39-42`).
    pub fn load(
        vb: &candle_nn::VarBuilder, cfg: &crate::lm::dual_ar::BaseModelArgs,
        token_config: &crate::lm::dual_ar::TokenConfig, model_type: crate::config::WhichLM,
    ) -> Result<Self> {
        let specs = lm_tensor_specs(cfg);
        let (_names, blobs, table) = tensor_table(vb, &specs)?;
        let names: Vec<String> = specs.iter().map(|(n, _)| n.clone()).collect();
        let views: Vec<(&[u8], DType, Vec<usize>)> = blobs
            .iter()
            .zip(&table)
            .zip(&specs)
            .map(|((b, t), (_, shape))| (b.as_slice(), if t.dtype == FSB_BF16 { DType::BF16 } else { DType::F32 }, shape.clone()))
            .collect();
        // max_batch 32: one handle serves the bs=1 server path, static batches and sessions
        Self::load_from_views(&names, &views, cfg, token_config, model_type, device_ordinal(vb.device()), vb.dtype() == DType::BF16, 32)
    }

    /// Same, from mmapped safetensors views (what `server/lib/utils/load.rs:62-188` holds) without a VarBuilder.
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

    /// `forward_generate(&inp u32 (B, C+1, S), input_pos, pad_mask) -> (logits (B, 1, V), hidden (B, 1, D))`,
    /// dual_ar.rs:574-635.  `pad_mask` must be None: the library keeps per-row lengths instead of left padding.
    pub fn forward_generate(&mut self, inp: &Tensor, input_pos: usize, pad_mask: Option<Tensor>) -> Result<(Tensor, Tensor)> {
        if pad_mask.is_some() {
            candle_core::bail!("b200 backend: pad_mask is not supported (rows carry their own lengths)");
        }
        let (b, _c1, s) = inp.dims3()?;
        let host: Vec<u32> = inp.to_dtype(DType::U32)?.flatten_all()?.to_vec1()?;
        let (v, d) = (self.cfg.vocab_size, self.cfg.dim);
        let mut logits = vec![0f32; b * v];
        let mut hidden = vec![0f32; b * d];
        check(unsafe {
            fsb_lm_forward_generate(self.handle, host.as_ptr(), b as i32, s as i32, input_pos, logits.as_mut_ptr(), hidden.as_mut_ptr())
        })?;
        Ok((Tensor::from_vec(logits, (b, 1, v), &Device::Cpu)?, Tensor::from_vec(hidden, (b, 1, d), &Device::Cpu)?))
    }

    /// `forward_generate_fast(&x f32 (B, 1, D), input_pos) -> logits (B, 1, codebook_size)`, dual_ar.rs:638-673
    pub fn forward_generate_fast(&mut self, x: &Tensor, input_pos: usize) -> Result<Tensor> {
        let (b, _s, _d) = x.dims3()?;
        let host: Vec<f32> = x.to_dtype(DType::F32)?.flatten_all()?.to_vec1()?;
        let cs = self.cfg.codebook_size;
        let mut logits = vec![0f32; b * cs];
        check(unsafe { fsb_lm_forward_generate_fast(self.handle, host.as_ptr(), b as i32, input_pos, logits.as_mut_ptr()) })?;
        Tensor::from_vec(logits, (b, 1, cs), &Device::Cpu)
    }

    /// `model.fast_embeddings.forward(&codes)` as the generators call it (single_batch.rs:176-182): u32 (n) -> f32 (n, D)
    pub fn fast_embeddings(&self, ids: &Tensor) -> Result<Tensor> {
        let host: Vec<u32> = ids.to_dtype(DType::U32)?.flatten_all()?.to_vec1()?;
        let d = self.cfg.dim;
        let mut out = vec![0f32; host.len() * d];
        check(unsafe { fsb_lm_fast_embeddings(self.handle, host.as_ptr(), host.len() as i32, out.as_mut_ptr()) })?;
        Tensor::from_vec(out, (host.len(), d), &Device::Cpu)
    }

    // ---- beyond the reference: the serving primitives of SURVEY 8f ----

    /// Snapshot of the first `n_positions` cached positions of `row` (the voice / system conditioning):
    /// what `clear_slow_caches_until(n_conditioning_tokens)` keeps within one request (speech.rs:40), kept across requests.
    pub fn kv_snapshot_save(&mut self, row: i32, n_positions: usize) -> Result<KvSnapshot> {
        let mut p = std::ptr::null_mut();
        check(unsafe { fsb_lm_kv_snapshot_save(self.handle, row, n_positions, &mut p) })?;
        Ok(KvSnapshot { lm: self.handle, ptr: p })
    }
    pub fn kv_snapshot_restore(&mut self, snap: &KvSnapshot, row: i32) -> Result<()> {
        check(unsafe { fsb_lm_kv_snapshot_restore(self.handle, snap.ptr, row) })
    }

    /// Continuous batching: slots instead of whole generations behind the Mutex (state.rs:13).
    pub fn session_begin(&mut self, sampling_args: &crate::lm::sampling::SamplingArgs, seed: u64) -> Result<()> {
        check(unsafe { fsb_lm_session_begin(self.handle, &to_ffi_sampling(sampling_args, seed), 0) })
    }
    pub fn session_admit(&mut self, slot: i32, prompt: &Tensor, max_new_tokens: usize) -> Result<()> {
        let (_rows, p) = prompt.dims2()?;
        let host: Vec<u32> = prompt.to_dtype(DType::U32)?.flatten_all()?.to_vec1()?;
        check(unsafe { fsb_lm_session_admit(self.handle, slot, host.as_ptr(), p as i32, max_new_tokens, 0) })
    }
    /// up to `max_frames` more frames for every live slot; returns the slots still generating
    pub fn session_run(&mut self, max_frames: i32) -> Result<Vec<bool>> {
        let mut active = vec![0i32; 32];
        let mut n = 0i32;
        check(unsafe { fsb_lm_session_run(self.handle, max_frames, active.as_mut_ptr(), &mut n) })?;
        Ok(active.iter().map(|a| *a != 0).collect())
    }
    /// codes u32 (C, T) of a finished slot; frees the slot
    pub fn session_collect(&mut self, slot: i32, max_frames: usize) -> Result<Tensor> {
        let c = self.cfg.num_codebooks;
        let mut out = vec![0u32; c * max_frames];
        let mut n = 0usize;
        check(unsafe { fsb_lm_session_collect(self.handle, slot, out.as_mut_ptr(), max_frames, &mut n) })?;
        let codes: Vec<u32> = (0..c).flat_map(|r| out[r * max_frames..r * max_frames + n].to_vec()).collect();
        Tensor::from_vec(codes, (c, n), &Device::Cpu)
    }
}

/// Device-resident conditioning KV of one voice (fsb_kv_snapshot)
pub struct KvSnapshot {
    lm: *mut fsb_lm,
    ptr: *mut fsb_kv_snapshot,
}
unsafe impl Send for KvSnapshot {}
impl Drop for KvSnapshot {
    fn drop(&mut self) {
        unsafe { fsb_lm_kv_snapshot_free(self.lm, self.ptr) };
    }
}

fn to_ffi_sampling(a: &crate::lm::sampling::SamplingArgs, seed: u64) -> fsb_sampling_args {
    fsb_sampling_args { temp: a.temp, top_p: a.top_p, top_k: a.top_k as u32, repetition_penalty: a.repetition_penalty, seed }
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

/// `generate_blocking_with_hidden(&mut model, &prompt, max_new_tokens, &sampling_args, collect_hidden_states, show_progress)
/// -> (codes (C, T), Some(hidden (T_all, D)))`, single_batch.rs:217-306: hidden rows are the pre-norm slow states of EVERY
/// yielded frame, im_end frames included (:251,268-270), as the server's hidden-state export wants (speech.rs:24-48).
pub fn generate_blocking_with_hidden(
    model: &mut DualARTransformer, prompt: &Tensor, max_new_tokens: usize,
    sampling_args: &crate::lm::sampling::SamplingArgs, collect_hidden_states: bool, show_progress: bool,
) -> Result<(Tensor, Option<Tensor>)> {
    if !collect_hidden_states {
        return Ok((generate_blocking(model, prompt, max_new_tokens, sampling_args, show_progress)?, None));
    }
    let (rows, p) = prompt.dims2()?;
    let host: Vec<u32> = prompt.to_dtype(DType::U32)?.flatten_all()?.to_vec1()?;
    let (c, d) = (rows - 1, model.cfg.dim);
    let cap = max_new_tokens.saturating_sub(p) + 2;
    let mut out = vec![0u32; c * cap];
    let mut hidden = vec![0f32; cap * d];
    let (mut n, mut nh) = (0usize, 0usize);
    let sa = to_ffi_sampling(sampling_args, rand::random::<u64>());
    check(unsafe {
        fsb_lm_generate_blocking_with_hidden(
            model.handle, host.as_ptr(), p as i32, max_new_tokens, &sa, 0, 0, out.as_mut_ptr(), cap, &mut n,
            hidden.as_mut_ptr(), cap, &mut nh,
        )
    })?;
    let codes: Vec<u32> = (0..c).flat_map(|r| out[r * cap..r * cap + n].to_vec()).collect();
    hidden.truncate(nh * d);
    Ok((Tensor::from_vec(codes, (c, n), &Device::Cpu)?, Some(Tensor::from_vec(hidden, (nh, d), &Device::Cpu)?)))
}

/// `generate_static_batch(&mut model, &prompts, max_new_tokens, audio_only, sampling_args) -> (Vec<codes>, Vec<Vec<bool>>)`,
/// static_batch.rs:282-390.  Row i is exactly `generate_blocking` on `prompts[i]` (independent-utterance semantics: the
/// reference's unmasked left padding, SURVEY Q7, is not reproduced); the second value is each row's per-frame "still
/// running" mask, all true up to the row's own length.
pub fn generate_static_batch(
    model: &mut DualARTransformer, prompts: &[Tensor], max_new_tokens: usize, _audio_only: bool,
    sampling_args: crate::lm::sampling::SamplingArgs,
) -> Result<(Vec<Tensor>, Vec<Vec<bool>>)> {
    let c = model.cfg.num_codebooks;
    let hosts: Vec<Vec<u32>> =
        prompts.iter().map(|t| t.to_dtype(DType::U32)?.flatten_all()?.to_vec1()).collect::<Result<_>>()?;
    let lens: Vec<i32> = prompts.iter().map(|t| t.dims2().map(|(_, p)| p as i32)).collect::<Result<_>>()?;
    let cap = max_new_tokens + 2;
    let mut outs: Vec<Vec<u32>> = prompts.iter().map(|_| vec![0u32; c * cap]).collect();
    let in_ptrs: Vec<*const u32> = hosts.iter().map(|h| h.as_ptr()).collect();
    let out_ptrs: Vec<*mut u32> = outs.iter_mut().map(|o| o.as_mut_ptr()).collect();
    let mut out_lens = vec![0usize; prompts.len()];
    let sa = to_ffi_sampling(&sampling_args, rand::random::<u64>());
    check(unsafe {
        fsb_lm_generate_static_batch(
            model.handle, in_ptrs.as_ptr(), lens.as_ptr(), prompts.len() as i32, max_new_tokens, &sa, 0, 0, out_ptrs.as_ptr(),
            cap, out_lens.as_mut_ptr(),
        )
    })?;
    let mut codes = Vec::with_capacity(prompts.len());
    let mut masks = Vec::with_capacity(prompts.len());
    for (o, n) in outs.iter().zip(&out_lens) {
        let v: Vec<u32> = (0..c).flat_map(|r| o[r * cap..r * cap + n].to_vec()).collect();
        codes.push(Tensor::from_vec(v, (c, *n), &Device::Cpu)?);
        masks.push(vec![true; *n]);
    }
    Ok((codes, masks))
}

/// Name + shape of every tensor `FireflyCodec::load` binds for Fish >= 1.4 (codec/firefly.rs:20-34 -> decoder.rs / encoder.rs /
/// quantizer.rs / hifi_gan.rs / convnext.rs loaders; the list is `tests/golden/codec_weight_dims_fish12.txt` after the
/// documented 1.2 -> 1.4 renaming, plus the second down/up-sampling stage of codec/config.rs:146-163).
fn codec_tensor_specs(with_encoder: bool) -> Vec<(String, Vec<usize>)> {
    let mut v: Vec<(String, Vec<usize>)> = Vec::new();
    let mut conv = |v: &mut Vec<(String, Vec<usize>)>, p: String, shape: Vec<usize>, bias: usize| {
        v.push((format!("{p}.weight"), shape));
        v.push((format!("{p}.bias"), vec![bias]));
    };
    let convnext = |v: &mut Vec<(String, Vec<usize>)>, p: String, dim: usize| {
        v.push((format!("{p}dwconv.conv.weight"), vec![dim, 1, 7]));
        v.push((format!("{p}dwconv.conv.bias"), vec![dim]));
        v.push((format!("{p}norm.weight"), vec![dim]));
        v.push((format!("{p}norm.bias"), vec![dim]));
        v.push((format!("{p}pwconv1.weight"), vec![4 * dim, dim]));
        v.push((format!("{p}pwconv1.bias"), vec![4 * dim]));
        v.push((format!("{p}pwconv2.weight"), vec![dim, 4 * dim]));
        v.push((format!("{p}pwconv2.bias"), vec![dim]));
        v.push((format!("{p}gamma"), vec![dim]));
    };
    for g in 0..8 {
        let p = format!("quantizer.residual_fsq.rvqs.{g}.");
        v.push((format!("{p}project_out.weight"), vec![64, 4]));
        v.push((format!("{p}project_out.bias"), vec![64]));
        if with_encoder {
            v.push((format!("{p}project_in.weight"), vec![4, 64]));
            v.push((format!("{p}project_in.bias"), vec![4]));
        }
    }
    for i in 0..2 {
        conv(&mut v, format!("quantizer.upsample.{i}.0.conv"), vec![512, 512, 2], 512);
        convnext(&mut v, format!("quantizer.upsample.{i}.1."), 512);
        if with_encoder {
            conv(&mut v, format!("quantizer.downsample.{i}.0.conv"), vec![512, 512, 2], 512);
            convnext(&mut v, format!("quantizer.downsample.{i}.1."), 512);
        }
    }
    conv(&mut v, "head.conv_pre.conv".to_string(), vec![512, 512, 13], 512);
    let (rates, kernels) = ([8usize, 8, 2, 2, 2], [16usize, 16, 4, 4, 4]);
    for i in 0..5 {
        let (cin, cout) = (512 >> i, 512 >> (i + 1));
        let _ = rates[i];
        conv(&mut v, format!("head.ups.{i}.conv"), vec![cin, cout, kernels[i]], cout);
        for (j, k) in [3usize, 7, 11].iter().enumerate() {
            for m in 0..3 {
                for which in ["convs1", "convs2"] {
                    conv(&mut v, format!("head.resblocks.{i}.blocks.{j}.{which}.{m}.conv"), vec![cout, cout, *k], cout);
                }
            }
        }
    }
    conv(&mut v, "head.conv_post.conv".to_string(), vec![1, 16, 13], 1);
    if with_encoder {
        let dims = [128usize, 256, 384, 512];
        let depths = [3usize, 3, 9, 3];
        conv(&mut v, "backbone.downsample_layers.0.0.conv".to_string(), vec![dims[0], 160, 7], dims[0]);
        v.push(("backbone.downsample_layers.0.1.weight".to_string(), vec![dims[0]]));
        v.push(("backbone.downsample_layers.0.1.bias".to_string(), vec![dims[0]]));
        for i in 1..4 {
            v.push((format!("backbone.downsample_layers.{i}.0.weight"), vec![dims[i - 1]]));
            v.push((format!("backbone.downsample_layers.{i}.0.bias"), vec![dims[i - 1]]));
            conv(&mut v, format!("backbone.downsample_layers.{i}.1"), vec![dims[i], dims[i - 1], 1], dims[i]);
        }
        for i in 0..4 {
            for j in 0..depths[i] {
                convnext(&mut v, format!("backbone.stages.{i}.{j}."), dims[i]);
            }
        }
        v.push(("backbone.norm.weight".to_string(), vec![512]));
        v.push(("backbone.norm.bias".to_string(), vec![512]));
    }
    v
}

/// `FireflyCodec` (codec/firefly.rs:10-49)
pub struct FireflyCodec {
    handle: *mut fsb_codec,
    pub sample_rate: u32,
}
unsafe impl Send for FireflyCodec {}

impl FireflyCodec {
    /// `FireflyCodec::load(cfg, vb, version)`, firefly.rs:20-34 -- the reference's signature.  Fish 1.2 checkpoints
    /// (un-merged weight norm, non-causal convs) are not supported by the library (FSB_ERR_UNSUPPORTED).
    pub fn load(_cfg: crate::codec::FireflyConfig, vb: candle_nn::VarBuilder, version: crate::config::WhichFishVersion) -> Result<Self> {
        let specs = codec_tensor_specs(true);
        let (_names, _blobs, table) = tensor_table(&vb, &specs)?;
        let opts = fsb_codec_options {
            device: device_ordinal(vb.device()),
            stream: std::ptr::null_mut(),
            fish_version: match version {
                crate::config::WhichFishVersion::Fish1_5 => FSB_FISH_1_5,
                crate::config::WhichFishVersion::Fish1_4 => FSB_FISH_1_4,
                _ => 12,
            },
            max_frames: 2048, // 95 s of audio per call; the server vocodes block-wise (speech.rs:180-236)
            with_encoder: 1,
        };
        let mut handle = std::ptr::null_mut();
        check(unsafe { fsb_codec_create(table.as_ptr(), table.len(), &opts, &mut handle) })?;
        let sample_rate = unsafe { fsb_codec_sample_rate(handle) } as u32;
        Ok(Self { handle, sample_rate })
    }

    /// Block-wise vocoding for the streaming handler (speech.rs:180-236): frames [t0, t1) of `tokens`, resampled to
    /// `to_rate` and converted to s16 on the device -- the same samples a whole decode would give, bit for bit.
    pub fn decode_block_s16(&self, tokens: &Tensor, t0: usize, t1: usize, to_rate: u32) -> Result<Vec<i16>> {
        let (_b, _g, t) = tokens.dims3()?;
        let host: Vec<u32> = tokens.to_dtype(DType::U32)?.flatten_all()?.to_vec1()?;
        let cap = (t1 - t0) * 2048 * (to_rate as usize).max(self.sample_rate as usize) / self.sample_rate as usize + 16;
        let mut out = vec![0i16; cap];
        let mut n = 0usize;
        check(unsafe {
            fsb_codec_decode_block_s16(self.handle, host.as_ptr(), t as i32, t0 as i32, t1 as i32, to_rate, out.as_mut_ptr(), cap, &mut n)
        })?;
        out.truncate(n);
        Ok(out)
    }

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
