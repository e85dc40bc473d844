"""Multi-GPU sharding of independent utterances: one process per GPU, every GPU holds a
full replica of the LM + codec, utterance i goes to exactly one rank, and there is NO collective on
the decode path.  The only exchange is the (optional) gather of results on rank 0 after the timed
region.  The reference has nothing to mirror here (single device, server/src/main.rs:25)."""
from typing import List, Sequence


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
