"""CPU tests: the C-ABI library loads, exports every symbol include/fsb.h declares,
and refuses to compute without a GPU (no fallback path)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from fish_speech_rs_b200 import _ffi as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fsb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fsb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = C.CDLL(F.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fsb.h but not exported by libfsb.so"
    # the Python binding covers the same set
    assert sorted(F.SYMBOLS) == names


def test_rust_binding_declares_the_same_symbols():
    """shim/fsb_sys.rs is shipped as source (no Rust toolchain in the image): its `extern "C"` block must name exactly
    the functions of include/fsb.h, and every one of them must be used by the safe wrappers or re-exported."""
    rs = open(os.path.join(ROOT, "shim", "fsb_sys.rs")).read()
    rs = re.sub(r"//.*", "", rs)
    assert sorted(set(re.findall(r"\bpub fn (fsb_[a-z0-9_]+)\s*\(", rs))) == declared_symbols()
    # INTEGRATION.md shows the reference-side binding: every entry point it names exists
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for n in set(re.findall(r"\b(fsb_[a-z0-9_]+)\s*\(", doc)):
        assert n in declared_symbols() or n.startswith("fsb_sys"), f"INTEGRATION.md mentions {n}, not in include/fsb.h"


def test_abi_version_and_error_string():
    lib = F.lib()
    assert lib.fsb_abi_version() == 1
    assert isinstance(lib.fsb_last_error(), bytes)


def test_struct_layouts_match_header():
    # sizes computed by hand from include/fsb.h (natural alignment)
    assert C.sizeof(F.fsb_tensor) == 8 + 8 + 4 + 4 + 32 + 4 + 4
    assert C.sizeof(F.fsb_model_args) == 15 * 4
    assert C.sizeof(F.fsb_token_config) == 20
    assert C.sizeof(F.fsb_sampling_args) == 8 + 8 + 4 + 4 + 8
    assert C.sizeof(F.fsb_lm_options) == 4 + 4 + 8 + 4 * 5 + 4
    assert C.sizeof(F.fsb_lm_stats) == 8 * 8
    assert C.sizeof(F.fsb_codec_options) == 4 + 4 + 8 + 12 + 4


def test_invalid_arguments_are_reported_not_crashed():
    lib = F.lib()
    assert lib.fsb_lm_create(None, None, None, 0, None, None) == -1
    assert b"null" in lib.fsb_last_error()
    assert lib.fsb_codec_create(None, 0, None, None) == -1
    assert lib.fsb_lm_destroy(None) == 0
    assert lib.fsb_codec_destroy(None) == 0


@pytest.mark.skipif(F.lib().fsb_device_count() > 0, reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(tiny_lm):
    """Without an sm_100 device every compute entry point must fail with FSB_ERR_CUDA."""
    from fish_speech_rs_b200 import DualARTransformer, FireflyCodec
    cfg, tok, w = tiny_lm
    with pytest.raises(F.FsbError) as e:
        DualARTransformer(w, cfg, tok)
    assert e.value.status == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(F.FsbError) as e:
        FireflyCodec({"x": np.zeros(1, np.float32)})
    assert e.value.status == -2
