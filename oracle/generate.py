"""Generation loops -- oracle (test infrastructure only).

Restates
  fish_speech_core/lib/lm/generate/single_batch.rs:19-324 (SingleBatchGenerator,
      generate_blocking_with_hidden)
  fish_speech_core/lib/lm/generate/utils.rs:6-56 (constrain / rescale)
including quirks Q1-Q6 of SURVEY.md section 8a.  The batched entry point follows
this build's "independent utterances" semantics: each row is the bs=1 path
(SURVEY Q7), not the reference's unmasked left-padded static batch.
"""
from typing import List, Optional, Tuple

import torch

from .dual_ar import DualARTransformer
from .sampling import RepPenProcessor, SamplingArgs, sample

REP_PEN_WINDOW = 16  # single_batch.rs:51


def constrain_probs_to_audio(logits: torch.Tensor, model: DualARTransformer) -> torch.Tensor:
    """utils.rs:6-33."""
    tc = model.token_config
    if model.model_type != "1.5":
        return logits
    if tc.im_end_id == tc.semantic_start_id - 1:
        return logits[..., tc.im_end_id:]
    return torch.cat([logits[..., tc.im_end_id: tc.im_end_id + 1], logits[..., tc.semantic_start_id:]], dim=-1)


def rescale_semantic_token(token: int, model: DualARTransformer) -> int:
    """utils.rs:36-56."""
    tc = model.token_config
    if model.model_type != "1.5":
        return token
    if tc.im_end_id == tc.semantic_start_id - 1:
        return token + tc.im_end_id
    if token == 0:
        return tc.im_end_id
    return token - 1 + tc.semantic_start_id


class SingleBatchGenerator:
    """single_batch.rs:19-215.  `force_slow` replaces the non-reproducible
    thread_rng PAD/EOS draw of Fish <=1.4 (Q8): frame i takes force_slow[i]."""

    def __init__(self, model: DualARTransformer, prompt: torch.Tensor, max_new_tokens: int,
                 args: SamplingArgs, audio_only: bool = True, row: int = 0,
                 fixed_len: Optional[int] = None, force_slow: Optional[List[int]] = None):
        self.model = model
        self.args = args
        self.row = row
        self.rep_pen = [RepPenProcessor(model.cfg.codebook_size, REP_PEN_WINDOW, args.repetition_penalty)
                        for _ in range(model.cfg.num_codebooks)]
        self.input_pos = model.curr_kv_size()
        self.max_new_tokens = max_new_tokens + model.curr_kv_size()
        self.prompt: Optional[torch.Tensor] = prompt.clone()
        self.previous_codes: Optional[List[int]] = None
        self.audio_only = audio_only
        self.frame = 0
        self.fixed_len = fixed_len
        self.force_slow = force_slow
        self.last_hidden = None
        self.last_slow_logits = None
        self.last_fast_logits = []

    def _draw(self, slot: int) -> int:
        return self.frame * (self.model.cfg.num_codebooks + 1) + slot

    def next(self) -> Optional[List[int]]:
        m = self.model
        C = m.cfg.num_codebooks
        if self.input_pos > self.max_new_tokens:
            return None
        if self.prompt is None:
            return None
        x = self.prompt
        prompt_length = x.shape[-1]
        if x.dim() == 2:
            x = x.unsqueeze(0)
        logits, hidden = m.forward_generate(x, self.input_pos)
        self.last_hidden = hidden
        im_end = m.token_config.im_end_id
        if self.force_slow is not None:
            semantic_token = self.force_slow[min(self.frame, len(self.force_slow) - 1)]
        elif self.audio_only and m.model_type != "1.5":
            raise NotImplementedError("Fish <=1.4 slow token is a thread_rng draw (Q8); pass force_slow")
        elif self.audio_only:
            slow = constrain_probs_to_audio(logits, m).flatten()
            if self.fixed_len is not None:
                slow = slow.clone()
                # fixed_len harness flag: <|im_end|> is not eligible
                slow[0] = float("-inf")
            self.last_slow_logits = slow
            semantic_token = rescale_semantic_token(sample(slow, self.args, self._draw(0), self.row), m)
        else:
            semantic_token = sample(logits.flatten(), self.args, self._draw(0), self.row)
        codebooks = [semantic_token]
        m.clear_fast_layer_caches()
        self.last_fast_logits = []
        xh = hidden.clone()
        for cb in range(C):
            if self.audio_only and semantic_token == im_end:
                codebooks.append(0)
                continue
            fl = m.forward_generate_fast(xh, cb).flatten()
            if self.previous_codes is None:
                adj = fl
            else:
                adj = self.rep_pen[cb].apply(fl, self.previous_codes[cb + 1])
            self.last_fast_logits.append(adj)
            a = sample(adj, self.args, self._draw(cb + 1), self.row)
            if cb != C - 1:
                xh = m.fast_embeddings[a].reshape(1, 1, -1)
            codebooks.append(a)
        if self.previous_codes is None:
            self.input_pos += prompt_length
        else:
            self.input_pos += 1
        self.previous_codes = list(codebooks)
        self.frame += 1
        if self.audio_only and semantic_token == im_end:
            self.prompt = None
        else:
            self.prompt = torch.tensor(codebooks, dtype=torch.int64).unsqueeze(-1)
        return codebooks


def generate_blocking(model: DualARTransformer, prompt: torch.Tensor, max_new_tokens: int,
                      args: SamplingArgs, row: int = 0, fixed_len: Optional[int] = None,
                      force_slow: Optional[List[int]] = None) -> torch.Tensor:
    """single_batch.rs:217-324 -> codes (C, T): frames whose slow token is
    <|im_end|> are dropped except the first (Q4), row 0 removed (:280-282).
    With `fixed_len`, exactly that many frames are produced (EOS disabled)."""
    gen = SingleBatchGenerator(model, prompt, max_new_tokens, args, True, row, fixed_len, force_slow)
    im_end = model.token_config.im_end_id
    first = gen.next()
    if first is None:
        raise RuntimeError("Prefill mistakenly thought generation ended. Please check max tokens")
    frames = [first]
    while True:
        if fixed_len is not None and len(frames) >= fixed_len:
            break
        tok = gen.next()
        if tok is None:
            break
        if tok[0] != im_end:
            frames.append(tok)
    full = torch.tensor(frames, dtype=torch.int64).T  # (C+1, T)
    return full[1:, :]


def generate_independent_batch(model: DualARTransformer, prompts: List[torch.Tensor], max_new_tokens: int,
                               args: SamplingArgs, fixed_len: Optional[int] = None) -> List[torch.Tensor]:
    """This build's batched semantics: row i == bs=1 path on prompt i,
    Philox row index i, slow KV cleared before each row."""
    outs = []
    for i, p in enumerate(prompts):
        model.clear_slow_layer_caches()
        outs.append(generate_blocking(model, p, max_new_tokens, args, row=i, fixed_len=fixed_len))
    model.clear_slow_layer_caches()
    return outs


# ---------------------------------------------------------------------------------------------------------------
# Teacher-forced replay: verifies EVERY sampling decision of a GPU generation against the oracle's logits.
#
# A bf16 tensor-core path accumulates in a different order (and with the tensor pipe's fp32 rounding) than the CPU
# reference, so over hundreds of decisions on flat synthetic distributions a few near-ties flip and the two
# autoregressive runs then diverge for good.  Comparing final token arrays cannot tell such a flip from a bug.
# The replay feeds the GPU's OWN frames to the oracle (so the two never diverge), and classifies each decision:
#   exact     the oracle, given the same history, picks the same token;
#   near_tie  it picks another one, but its own margin is below the stated fp tolerance (greedy: logit gap;
#             sampled: the uniform draw lies within `tol_cdf` of the CDF boundary between the two candidates, or
#             the two candidates' probabilities are within `tol_cdf` and adjacent in the ordering);
#   violation anything else -- a real disagreement.
# Tolerances are the repo-wide ones: logits atol 1e-3 (tests/e2e/backbone-allclose.py:82) -> tol_logit = 2e-3
# (both operands of a gap), tol_cdf = 2e-3 / temp relative to the kept mass.
def _classify(logits: torch.Tensor, args: SamplingArgs, draw: int, row: int, tok: int, tol_logit: float):
    flat = logits.to(torch.float32).flatten()
    if args.temp <= 1e-7:
        best = int(torch.argmax(flat).item())
        if best == tok:
            return "exact"
        return "near_tie" if float(flat[best] - flat[tok]) <= tol_logit else "violation"
    pick = sample(flat, args, draw, row)
    if pick == tok:
        return "exact"
    import numpy as np
    from .rng import philox_uniform
    from .sampling import softmax_probs, _sorted_desc
    probs = softmax_probs(flat, args.temp).astype(np.float64)
    order = _sorted_desc(probs.astype(np.float32))
    n = probs.shape[0]
    k = n if args.top_k >= n else args.top_k
    cand = [int(c) for c in order[:k]]
    tol = tol_logit / args.temp  # relative tolerance on probabilities / cumulative sums
    if tok not in cand:
        # just outside the oracle's top-k: acceptable only if it ties with the k-th candidate and the draw falls on it
        tie = abs(probs[tok] - probs[cand[-1]]) <= tol * probs[cand[-1]]
        return "near_tie" if tie and pick == cand[-1] else "violation"
    w = probs[cand]
    cum = np.cumsum(w)
    before = cum - w  # running sum BEFORE each candidate (the top-p rule looks at it, mod.rs:119-129)
    use_topp = (not (args.top_p <= 0.0 or args.top_p >= cum[-1])) or args.top_k >= n
    kept0 = int((before < args.top_p).sum()) if use_topp else k
    alts = {kept0}
    if use_topp:
        for kk in (kept0 - 1, kept0 + 1):  # the cut moves by one candidate if a running sum sits on top_p
            if 1 <= kk <= k and abs(before[min(kk, kept0)] - args.top_p) <= tol:
                alts.add(kk)
    u = float(philox_uniform(args.seed, draw, row))
    i_g = cand.index(tok)
    for kept in sorted(alts):
        total = cum[kept - 1]
        chosen = u * total
        i = int(np.searchsorted(cum[:kept], chosen, side="right"))
        i = min(i, kept - 1)
        if i == i_g:
            return "near_tie"  # same rule, top-p cut one candidate over
        if abs(i - i_g) == 1 and i_g < kept:
            # the draw lands within tolerance of the boundary between the two neighbours ...
            if abs(chosen - cum[min(i, i_g)]) <= tol * total:
                return "near_tie"
            # ... or the two neighbours have (numerically) equal probability and swap places in the ordering
            if abs(w[i] - w[i_g]) <= tol * max(w[i], w[i_g]):
                return "near_tie"
    return "violation"


def replay_frames(model: DualARTransformer, prompt: torch.Tensor, frames, args: SamplingArgs, row: int = 0,
                  fixed_len: Optional[int] = None, tol_logit: float = 2e-3, force_slow: bool = False,
                  keep_slow_kv: bool = False):
    """frames: int array (C+1, T) as `fsb_lm_last_frames` returns them.  Returns a dict of counts and the list of
    (frame, slot, oracle_pick_class) for every non-exact decision.  The oracle model's slow KV is cleared first."""
    import numpy as np
    m = model
    C = m.cfg.num_codebooks
    frames = np.asarray(frames).astype(np.int64)
    T = frames.shape[1]
    if not keep_slow_kv:  # keep_slow_kv: the prompt continues on top of a kept prefix (speech.rs:40)
        m.clear_slow_layer_caches()
    rep = [RepPenProcessor(m.cfg.codebook_size, REP_PEN_WINDOW, args.repetition_penalty) for _ in range(C)]
    stats = {"decisions": 0, "exact": 0, "near_tie": 0, "violation": 0, "events": []}
    im_end = m.token_config.im_end_id
    pos = m.curr_kv_size()
    x = prompt.clone()
    prev = None
    with torch.no_grad():
        for f in range(T):
            xin = x.unsqueeze(0) if x.dim() == 2 else x
            logits, hidden = m.forward_generate(xin, pos)
            pos += xin.shape[-1]
            tok0 = int(frames[0, f])
            if not force_slow and m.model_type == "1.5":
                slow = constrain_probs_to_audio(logits, m).flatten()
                if fixed_len is not None:
                    slow = slow.clone()
                    slow[0] = float("-inf")
                # GPU token back to its index in the constrained vector (inverse of rescale_semantic_token)
                tc = m.token_config
                if tc.im_end_id == tc.semantic_start_id - 1:
                    idx = tok0 - tc.im_end_id
                else:
                    idx = 0 if tok0 == tc.im_end_id else tok0 - tc.semantic_start_id + 1
                cls = _classify(slow, args, f * (C + 1), row, idx, tol_logit) if 0 <= idx < slow.numel() else "violation"
                stats["decisions"] += 1
                stats[cls] += 1
                if cls != "exact":
                    stats["events"].append((f, 0, cls))
            m.clear_fast_layer_caches()
            xh = hidden.clone()
            if tok0 == im_end:
                if any(int(frames[1 + c, f]) != 0 for c in range(C)):
                    stats["violation"] += 1
                    stats["events"].append((f, -1, "eos frame with non-zero codes"))
                break
            for cb in range(C):
                fl = m.forward_generate_fast(xh, cb).flatten()
                adj = fl if prev is None else rep[cb].apply(fl, int(prev[cb + 1]))
                a = int(frames[1 + cb, f])
                cls = _classify(adj, args, f * (C + 1) + cb + 1, row, a, tol_logit) if 0 <= a < adj.numel() else "violation"
                stats["decisions"] += 1
                stats[cls] += 1
                if cls != "exact":
                    stats["events"].append((f, cb + 1, cls))
                if cb != C - 1:
                    xh = m.fast_embeddings[a].reshape(1, 1, -1)
            prev = [int(v) for v in frames[:, f]]
            x = torch.tensor(prev, dtype=torch.int64).unsqueeze(-1)
    if not keep_slow_kv:
        m.clear_slow_layer_caches()
    return stats
