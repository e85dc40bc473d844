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
