"""ctypes binding of libfsb.so (include/fsb.h).  The library is the product; this
module only marshals arguments.  There is no Python/CPU fallback: if the shared
library is missing the import fails loudly, and every compute entry point
returns FSB_ERR_CUDA when no sm_100 device is present."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FSB_LIB: same-box A/B runs of two builds of the library (tools/gpu_ab2.sh); the default is the in-tree build
LIB_PATH = os.environ.get("FSB_LIB") or os.path.join(_HERE, "libfsb.so")

FSB_F32, FSB_BF16, FSB_F16, FSB_U32, FSB_I64, FSB_U8, FSB_F64 = range(7)
FSB_FISH_1_2, FSB_FISH_1_4, FSB_FISH_1_5 = 12, 14, 15
FSB_GEN_FIXED_LEN, FSB_GEN_KEEP_SLOW_KV = 0x1, 0x2
STATUS = {0: "FSB_OK", -1: "FSB_ERR_INVALID", -2: "FSB_ERR_CUDA", -3: "FSB_ERR_MISSING_WEIGHT",
          -4: "FSB_ERR_SHAPE", -5: "FSB_ERR_STATE", -6: "FSB_ERR_UNSUPPORTED", -7: "FSB_ERR_OOM"}


class FsbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"{STATUS.get(status, status)}: {msg}")
        self.status = status


class fsb_tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 4), ("on_device", C.c_int32)]


class fsb_model_args(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "attention_qkv_bias", "codebook_size", "dim", "head_dim", "intermediate_size", "max_seq_len", "n_fast_layer",
        "n_head", "n_layer", "n_local_heads", "num_codebooks", "vocab_size", "tie_word_embeddings")] + [
        ("norm_eps", C.c_float), ("rope_base", C.c_float)]


class fsb_token_config(C.Structure):
    _fields_ = [("im_end_id", C.c_uint32), ("pad_id", C.c_uint32), ("semantic_start_id", C.c_uint32),
                ("semantic_end_id", C.c_uint32), ("has_semantic_end", C.c_int32)]


class fsb_sampling_args(C.Structure):
    _fields_ = [("temp", C.c_double), ("top_p", C.c_double), ("top_k", C.c_uint32),
                ("repetition_penalty", C.c_float), ("seed", C.c_uint64)]


class fsb_lm_options(C.Structure):
    _fields_ = [("device", C.c_int32), ("stream", C.c_void_p), ("weight_dtype", C.c_int32), ("max_batch", C.c_int32),
                ("max_seq_len", C.c_int32), ("fish_version", C.c_int32), ("decode_mode", C.c_int32)]


class fsb_lm_stats(C.Structure):
    _fields_ = [("prefill_ms", C.c_double), ("decode_ms", C.c_double), ("frames", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("dominant_kernel_ms", C.c_double),
                ("dominant_kernel_launches", C.c_uint64), ("weight_bytes_per_frame", C.c_uint64),
                ("dominant_kernel_bytes", C.c_uint64)]


class fsb_codec_options(C.Structure):
    _fields_ = [("device", C.c_int32), ("stream", C.c_void_p), ("fish_version", C.c_int32),
                ("max_frames", C.c_int32), ("with_encoder", C.c_int32)]


class fsb_codec_stats(C.Structure):
    _fields_ = [("decode_ms", C.c_double), ("device_ms", C.c_double), ("kernel_launches", C.c_uint64),
                ("dominant_kernel_ms", C.c_double),
                ("dominant_kernel_launches", C.c_uint64)]


# every symbol include/fsb.h declares: name -> (restype, argtypes)
P = C.POINTER
SYMBOLS = {
    "fsb_abi_version": (C.c_int, []),
    "fsb_last_error": (C.c_char_p, []),
    "fsb_device_count": (C.c_int, []),
    "fsb_lm_create": (C.c_int, [P(fsb_model_args), P(fsb_token_config), P(fsb_tensor), C.c_size_t,
                                P(fsb_lm_options), P(C.c_void_p)]),
    "fsb_lm_destroy": (C.c_int, [C.c_void_p]),
    "fsb_lm_forward_generate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_size_t, C.c_void_p,
                                          C.c_void_p]),
    "fsb_lm_forward_generate_fast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_size_t, C.c_void_p]),
    "fsb_lm_fast_embeddings": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "fsb_lm_clear_fast_layer_caches": (C.c_int, [C.c_void_p]),
    "fsb_lm_clear_slow_layer_caches": (C.c_int, [C.c_void_p]),
    "fsb_lm_clear_slow_caches_until": (C.c_int, [C.c_void_p, C.c_size_t]),
    "fsb_lm_curr_kv_size": (C.c_int, [C.c_void_p, P(C.c_size_t)]),
    "fsb_lm_generate_blocking": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_size_t, P(fsb_sampling_args),
                                           C.c_uint32, C.c_int32, C.c_void_p, C.c_size_t, P(C.c_size_t)]),
    "fsb_lm_generate_blocking_with_hidden": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_size_t, P(fsb_sampling_args),
                                                       C.c_uint32, C.c_int32, C.c_void_p, C.c_size_t, P(C.c_size_t),
                                                       C.c_void_p, C.c_size_t, P(C.c_size_t)]),
    "fsb_lm_generate_static_batch": (C.c_int, [C.c_void_p, P(C.c_void_p), P(C.c_int32), C.c_int32, C.c_size_t,
                                               P(fsb_sampling_args), C.c_uint32, C.c_int32, P(C.c_void_p),
                                               C.c_size_t, P(C.c_size_t)]),
    "fsb_lm_last_frames": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, P(C.c_size_t)]),
    "fsb_lm_kv_snapshot_save": (C.c_int, [C.c_void_p, C.c_int32, C.c_size_t, P(C.c_void_p)]),
    "fsb_lm_kv_snapshot_restore": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "fsb_lm_kv_snapshot_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fsb_lm_session_begin": (C.c_int, [C.c_void_p, P(fsb_sampling_args), C.c_uint32]),
    "fsb_lm_session_admit": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_size_t, C.c_int32]),
    "fsb_lm_session_run": (C.c_int, [C.c_void_p, C.c_int32, P(C.c_int32), P(C.c_int32)]),
    "fsb_lm_session_collect": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, P(C.c_size_t)]),
    "fsb_lm_get_stats": (C.c_int, [C.c_void_p, P(fsb_lm_stats)]),
    "fsb_lm_set_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "fsb_codec_create": (C.c_int, [P(fsb_tensor), C.c_size_t, P(fsb_codec_options), P(C.c_void_p)]),
    "fsb_codec_destroy": (C.c_int, [C.c_void_p]),
    "fsb_codec_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "fsb_codec_decode_batch": (C.c_int, [C.c_void_p, P(C.c_void_p), P(C.c_int32), C.c_int32, P(C.c_void_p)]),
    "fsb_codec_encode_mel": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, P(C.c_size_t)]),
    "fsb_codec_decode_block": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "fsb_codec_decode_block_s16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_void_p,
                                            C.c_size_t, P(C.c_size_t)]),
    "fsb_codec_log_mel": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, P(C.c_size_t)]),
    "fsb_codec_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, P(C.c_size_t)]),
    "fsb_codec_get_stats": (C.c_int, [C.c_void_p, P(fsb_codec_stats)]),
    "fsb_codec_sample_rate": (C.c_int32, [C.c_void_p]),
    "fsb_op_repeat_kv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p]),
    "fsb_op_gqa_decode_attn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    "fsb_op_gqa_decode_attn_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
}

_lib = None


def lib():
    """The loaded C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()); "
                              "there is no fallback implementation")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status):
    if status != 0:
        raise FsbError(status, lib().fsb_last_error().decode("utf-8", "replace"))


def tensor_table(weights):
    """dict name -> torch.Tensor (f32 / bf16, CPU or CUDA) -> (fsb_tensor array, keep-alive list)."""
    import torch
    arr = (fsb_tensor * len(weights))()
    keep = []
    for i, (name, t) in enumerate(weights.items()):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t)
        if t.dtype not in (torch.float32, torch.bfloat16):
            t = t.to(torch.float32)
        t = t.contiguous()
        nb = name.encode()
        keep.append((t, nb))
        arr[i].name = nb
        arr[i].data = t.data_ptr()
        arr[i].dtype = FSB_F32 if t.dtype == torch.float32 else FSB_BF16
        arr[i].ndim = t.dim()
        for d in range(t.dim()):
            arr[i].shape[d] = t.shape[d]
        arr[i].on_device = 1 if t.is_cuda else 0
    return arr, keep
