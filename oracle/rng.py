"""Counter-based uniform generator shared by the oracle and the CUDA sampler.

The reference seeds ``rand::rngs::StdRng`` (ChaCha12) from OS entropy for the
bs=1 path (single_batch.rs:46) and from the constant 42 for the batched one
(static_batch.rs:63), so its bs=1 draws are not reproducible by construction.
This build replaces the stream with Philox4x32-10 keyed by the user's seed and
indexed by (draw, row); one 24-bit uniform per sampling call.  The CUDA side
(csrc/fsb_sample.cuh: philox_uniform) implements the identical function, which
is what makes sampled-token parity testable at all.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: 4 uint32, key: 2 uint32 -> 4 uint32 (Salmon et al. 2011)."""
    c = [np.uint64(x & 0xFFFFFFFF) for x in counter]
    k0, k1 = key[0] & 0xFFFFFFFF, key[1] & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c = [
            (hi1 ^ c[1] ^ np.uint64(k0)) & _MASK,
            lo1,
            (hi0 ^ c[3] ^ np.uint64(k1)) & _MASK,
            lo0,
        ]
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return [int(x) for x in c]


def philox_uniform(seed: int, draw: int, row: int = 0) -> np.float32:
    """Uniform in [0, 1) with 24 bits: word0 >> 8 scaled by 2^-24."""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    ctr = (draw & 0xFFFFFFFF, (draw >> 32) & 0xFFFFFFFF, row & 0xFFFFFFFF, 0)
    w = philox4x32_10(ctr, key)[0]
    return np.float32((w >> 8) * (1.0 / 16777216.0))
