"""fish-speech.rs hot path, B200-native (sm_100a): dual-AR token loop + Firefly codec.

The product is `libfsb.so` (C ABI in include/fsb.h); these modules mirror the
reference's `fish_speech_core::lm` / `fish_speech_core::codec` surface on top of it.
"""
from . import _ffi
from .codec import FireflyCodec
from .lm import (DualARTransformer, SamplingArgs, generate_blocking, generate_blocking_with_hidden,
                 generate_static_batch)

__all__ = ["DualARTransformer", "FireflyCodec", "SamplingArgs", "generate_blocking", "generate_blocking_with_hidden",
           "generate_static_batch", "_ffi"]
