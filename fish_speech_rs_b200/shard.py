"""Multi-GPU sharding of independent utterances: one process per GPU, every GPU holds a
full replica of the LM + codec, utterance i goes to exactly one rank, and there is NO collective on
the decode path.  The only exchange is the (optional) gather of results on rank 0 after the timed
region.  The reference has nothing to mirror here (single device, server/src/main.rs:25)."""
from typing import Callable, Dict, List, Optional, Sequence, Tuple


def assign(costs: Sequence[int], world_size: int) -> List[List[int]]:
    """Length-balanced static partition: longest-processing-time greedy bin packing on
    cost_i = P_i + N_i (prompt + frames).  Returns, per rank, the utterance indices it owns (each in
    ascending order).  Deterministic: every rank computes the same table without communicating."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    load = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(costs[i])
    return [sorted(x) for x in out]


def my_shard(costs: Sequence[int], rank: int, world_size: int) -> List[int]:
    return assign(costs, world_size)[rank]


def gather_results(local: dict, world_size: int, group=None) -> dict:
    """Rank 0 receives {utterance index: payload} from every rank (torch.distributed all_gather_object;
    gloo on CPU, nccl on GPU boxes).  Outside any timed region."""
    import torch.distributed as dist
    if world_size == 1 or not dist.is_initialized():
        return dict(local)
    parts = [None] * world_size
    dist.all_gather_object(parts, local, group=group)
    merged: dict = {}
    for part in parts:
        for k, v in part.items():
            if k in merged:
                raise RuntimeError(f"utterance {k} was produced by two ranks")
            merged[k] = v
    return merged


MAX_ROWS = 32  # rows of one wide-batch launch (csrc/fsb_lm_megab.cuh); more rows per GPU run as several launches


def plan(costs: Sequence[int], rank: int, world_size: int, max_rows: int = MAX_ROWS) -> List[List[int]]:
    """The launches of one rank: its utterances (see `assign`) sorted by cost, longest first, and cut into batches of
    at most `max_rows` rows.  A static batch runs until its longest row is done, so rows of similar length share a
    launch; ties keep index order.  Deterministic and communication-free like `assign`."""
    if max_rows < 1:
        raise ValueError("max_rows must be >= 1")
    mine = sorted(my_shard(costs, rank, world_size), key=lambda i: (-int(costs[i]), i))
    return [mine[k:k + max_rows] for k in range(0, len(mine), max_rows)]


class ShardedSynthesizer:
    """Data-parallel driver above the C ABI: one instance per process / GPU (rank r of world_size), every instance owns
    a full replica (an LM handle and a codec handle).  `synthesize` runs THIS rank's share of a list of utterances --
    token loop, then vocoder, batch by batch -- and touches no other rank: there is no collective on the data path.
    `gather=True` adds the one exchange SURVEY 8e allows, after the work: rank 0 receives everybody's results.

    `generate(prompts, max_new_tokens, fixed_len) -> [codes (C, T_i)]` and `vocode([codes]) -> [pcm]` default to the
    library (`generate_static_batch` on `lm`, `FireflyCodec.decode_batch` on `codec`); tests inject stand-ins, because
    the host logic is what the CPU suite can check (tests/test_shard_gloo.py)."""

    def __init__(self, lm=None, codec=None, rank: int = 0, world_size: int = 1, sampling_args=None,
                 generate: Optional[Callable] = None, vocode: Optional[Callable] = None, max_rows: int = MAX_ROWS):
        if not 0 <= rank < world_size:
            raise ValueError("rank must be in [0, world_size)")
        self.lm, self.codec, self.rank, self.world_size, self.max_rows = lm, codec, rank, world_size, max_rows
        self.sampling_args = sampling_args
        self._generate = generate or self._lib_generate
        self._vocode = vocode or self._lib_vocode
        self.launches: List[List[int]] = []  # the batches of the last call (utterance indices), for inspection

    def _lib_generate(self, prompts, max_new_tokens, fixed_len):
        from .lm import generate_static_batch
        if self.lm is None or self.sampling_args is None:
            raise RuntimeError("ShardedSynthesizer needs an LM handle and sampling args (or an injected generate)")
        return generate_static_batch(self.lm, prompts, max_new_tokens, self.sampling_args, fixed_len=fixed_len)

    def _lib_vocode(self, codes):
        if self.codec is None:
            raise RuntimeError("ShardedSynthesizer needs a codec handle (or an injected vocode)")
        return self.codec.decode_batch(codes)

    def synthesize(self, prompts: Sequence, max_new_tokens: int, frames: Optional[Sequence[int]] = None,
                   fixed_len: Optional[int] = None, gather: bool = False) -> Dict[int, Tuple[object, object]]:
        """prompts: ALL utterances of the job, identical on every rank (each (C + 1, P_i)); frames: expected frames per
        utterance for the cost model (default: `fixed_len` or `max_new_tokens`).  Returns {utterance index: (codes,
        pcm)} for this rank's utterances, or for all of them on rank 0 when `gather` is set (other ranks: their own)."""
        n = len(prompts)
        if frames is None:
            frames = [int(fixed_len if fixed_len is not None else max_new_tokens)] * n
        if len(frames) != n:
            raise ValueError("frames must have one entry per prompt")
        costs = [int(p.shape[1]) + int(f) for p, f in zip(prompts, frames)]
        self.launches = plan(costs, self.rank, self.world_size, self.max_rows)
        out: Dict[int, Tuple[object, object]] = {}
        for batch in self.launches:
            codes = self._generate([prompts[i] for i in batch], max_new_tokens, fixed_len)
            pcm = self._vocode(codes)
            if len(codes) != len(batch) or len(pcm) != len(batch):
                raise RuntimeError("backend returned a different number of rows than it was given")
            for i, c, w in zip(batch, codes, pcm):
                out[i] = (c, w)
        if gather:
            merged = gather_results(out, self.world_size)
            if self.rank == 0:
                if sorted(merged) != list(range(n)):
                    raise RuntimeError("gathered results do not cover every utterance exactly once")
                return merged
        return out
