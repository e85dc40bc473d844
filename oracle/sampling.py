"""Sampling and repetition penalty -- oracle (test infrastructure only).

Follows
  fish_speech_core/lib/lm/sampling/mod.rs:29-132   (SamplingArgs, top-k -> top-p -> multinomial)
  fish_speech_core/lib/lm/sampling/rep_pen.rs:4-72 (SingleBatchedRepPenProcessor)
  fish_speech_core/lib/lm/generate/single_batch.rs:38-46 (candle_transformers
      LogitsProcessor::from_sampling(ArgMax | TopKThenTopP); candle-transformers
      0.8.3 is un-vendored, its algorithm is restated from its published source)

Deliberate, documented deviations (DESIGN.md "sampler semantics"):
  * RNG: Philox4x32-10 (oracle/rng.py) instead of ChaCha12 seeded from entropy.
  * multinomial draw order: candidates are walked in descending-probability
    order (ties: lower index first).  The reference walks them in the order
    `select_nth_unstable_by` leaves them, which Rust documents as unspecified;
    the distribution is identical.
  * argmax tie-break: first maximal index (ties do not occur with
    continuous synthetic weights, tests assert that).
"""
from dataclasses import dataclass

import numpy as np
import torch

from .rng import philox_uniform


@dataclass
class SamplingArgs:
    """sampling/mod.rs:28-34."""
    temp: float = 0.7
    top_p: float = 0.8
    top_k: int = 256
    repetition_penalty: float = 1.4
    seed: int = 0  # ours: Philox key (reference: rand::random / 42)


def softmax_probs(logits: torch.Tensor, temp: float) -> np.ndarray:
    """`softmax_last_dim(logits / temp)` in f32 (candle affine: x * f32(1/temp))."""
    inv = np.float32(1.0 / temp)
    x = logits.to(torch.float32).flatten() * float(inv)
    return torch.softmax(x, dim=-1).numpy().astype(np.float32)


def _sorted_desc(probs: np.ndarray) -> np.ndarray:
    # descending probability, ties -> lower index first (stable sort on -p)
    return np.argsort(-probs, kind="stable")


def sample_from_probs(probs: np.ndarray, top_k: int, top_p: float, u: np.float32) -> int:
    """top-k -> top-p -> multinomial (mod.rs:51-75,113-132; candle LogitsProcessor
    sample_topk_topp).  `u` is the uniform in [0,1) consumed by this draw."""
    n = probs.shape[0]
    order = _sorted_desc(probs)
    k = n if top_k >= n else top_k
    cand = order[:k]
    w = probs[cand].astype(np.float32).copy()
    # sum_p over the kept k (f32 sequential, mod.rs:66)
    sum_p = np.float32(0.0)
    for v in w:
        sum_p = np.float32(sum_p + v)
    top_p32 = np.float32(top_p)
    # the gate is evaluated in f64 (mod.rs:67: `top_p <= 0.0 || top_p >= sum_p as f64`), the cut below in f32 (`top_p as f32`)
    if not (top_p <= 0.0 or float(top_p) >= float(sum_p)) or top_k >= n:
        # sample_topp: zero everything once the running sum has reached top_p (mod.rs:119-129)
        cumsum = np.float32(0.0)
        for i in range(k):
            if cumsum >= top_p32:
                w[i] = np.float32(0.0)
            cumsum = np.float32(cumsum + w[i])
    # WeightedIndex (rand 0.8.5): chosen = u * total; first i with cum[i] > chosen
    total = np.float32(0.0)
    for v in w:
        total = np.float32(total + v)
    chosen = np.float32(u * total)
    cum = np.float32(0.0)
    last_kept = 0
    for i in range(k):
        if w[i] > 0:
            last_kept = i
        cum = np.float32(cum + w[i])
        if cum > chosen and w[i] > 0:
            return int(cand[i])
    return int(cand[last_kept])


def sample(logits: torch.Tensor, args: SamplingArgs, draw: int, row: int = 0) -> int:
    """LogitsProcessor::sample (ArgMax when temp == 0, else TopKThenTopP)."""
    flat = logits.to(torch.float32).flatten()
    if args.temp <= 1e-7:
        return int(torch.argmax(flat).item())  # first maximal index
    probs = softmax_probs(flat, args.temp)
    u = philox_uniform(args.seed, draw, row)
    return sample_from_probs(probs, args.top_k, args.top_p, u)


class RepPenProcessor:
    """SingleBatchedRepPenProcessor, rep_pen.rs:4-72, bug-for-bug.

    `tokens_seen.entry(t).or_insert(1)` never increments an existing count, so a
    token is un-penalised as soon as ANY occurrence of it leaves the window, even
    if a younger occurrence is still inside (rep_pen.rs:43-61).
    """

    def __init__(self, vocab_size: int, max_ctxt_size: int, penalty: float):
        self.mask = torch.ones(vocab_size, dtype=torch.float32)
        self.penalty = float(np.float32(penalty))
        self.context = []  # index 0 == front
        self.tokens_seen = {}
        self.max_ctxt_size = max_ctxt_size
        self.vocab_size = vocab_size

    def apply(self, logits: torch.Tensor, last_token: int) -> torch.Tensor:
        if last_token >= self.vocab_size:
            raise ValueError("Token must be within vocab size")
        count = self.tokens_seen.setdefault(last_token, 1)
        if count == 1:
            self.mask[last_token] = self.penalty
        self.context.insert(0, last_token)
        if len(self.context) > self.max_ctxt_size:
            dropped = self.context.pop()
            if dropped in self.tokens_seen:
                self.tokens_seen[dropped] -= 1
                if self.tokens_seen[dropped] == 0:
                    del self.tokens_seen[dropped]
                    self.mask[dropped] = 1.0
        return logits / self.mask  # sign-agnostic divide (rep_pen.rs:64)
