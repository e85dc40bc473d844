//! `CustomOp1` wrappers over the operator-level entry points, in the shape of the reference's
//! `RepeatKV` (fish_speech_core/lib/lm/ops/repeat_kv.rs:8-107).  SOURCE ONLY (no cargo here).
use super::fsb_sys::*;
use candle_core::{backend::BackendStorage, cuda_backend::cudarc::driver::DevicePtr, CpuStorage, CudaStorage, CustomOp1, Layout, Result, Shape};

pub struct RepeatKV {
    pub n_rep: usize,
}

impl CustomOp1 for RepeatKV {
    fn name(&self) -> &'static str {
        "repeat-kv"
    }
    fn cpu_fwd(&self, _s: &CpuStorage, _l: &Layout) -> Result<(CpuStorage, Shape)> {
        candle_core::bail!("repeat_kv: no CPU path (same as the reference, repeat_kv.rs:13-20)")
    }
    fn cuda_fwd(&self, s: &CudaStorage, l: &Layout) -> Result<(CudaStorage, Shape)> {
        if !l.is_contiguous() {
            candle_core::bail!("repeat_kv: input must be contiguous") // repeat_kv.rs:52-55
        }
        let (bsz, n_local_heads, seqlen, head_dim) = l.shape().dims4()?;
        if bsz != 1 {
            candle_core::bail!("repeat_kv: bsz must be 1") // repeat_kv.rs:81-83
        }
        let dev = s.device().clone();
        let src = s.as_cuda_slice::<f32>()?;
        let n_out = n_local_heads * self.n_rep * seqlen * head_dim;
        let dst = unsafe { dev.alloc::<f32>(n_out) }?; // repeat_kv.rs:69
        let st = unsafe {
            fsb_op_repeat_kv(
                *src.device_ptr() as *const _, *dst.device_ptr() as *mut _, FSB_F32, n_local_heads as i32, self.n_rep as i32,
                seqlen as i32, head_dim as i32, dev.cu_stream() as *mut _,
            )
        };
        if st != FSB_OK {
            candle_core::bail!("fsb_op_repeat_kv failed: {st}")
        }
        let out = CudaStorage::wrap_cuda_slice(dst, dev);
        Ok((out, Shape::from((bsz, n_local_heads * self.n_rep, seqlen, head_dim))))
    }
}
