"""CPU tests of the N>1 host logic: utterance sharding is a partition, balanced, identical on every
rank, and the result gather works over a world_size-2 gloo group (no data-path collective exists)."""
import os
import socket

import torch.distributed as dist
import torch.multiprocessing as mp

from fish_speech_rs_b200 import shard


def test_assign_is_a_balanced_partition():
    costs = [300 + 28 * i + 216 for i in range(16)]  # cfg3: mixed prompts + 216 frames
    for ws in (1, 2, 4, 8):
        parts = shard.assign(costs, ws)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(16))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs)
    assert shard.assign([], 4) == [[], [], [], []]
    assert shard.assign([5, 5, 5], 8)[3:] == [[]] * 5


def test_plan_cuts_a_rank_into_launches_of_similar_length():
    costs = [384 + 1292] * 40 + [300 + 28 * i + 216 for i in range(16)]  # 56 utterances, two length classes
    for ws in (1, 2, 4):
        seen = []
        for r in range(ws):
            batches = shard.plan(costs, r, ws)
            assert all(1 <= len(b) <= shard.MAX_ROWS for b in batches)
            flat = [i for b in batches for i in b]
            assert sorted(flat) == shard.my_shard(costs, r, ws)
            assert [costs[i] for i in flat] == sorted((costs[i] for i in flat), reverse=True)  # longest first
            seen += flat
        assert sorted(seen) == list(range(len(costs)))
    assert shard.plan(costs, 0, 1, max_rows=8)[0] == list(range(8))  # ties keep index order
    assert shard.plan([], 0, 2) == []


class _Arr:
    """stand-in for a (C + 1, P) prompt array"""

    def __init__(self, P):
        self.shape = (9, P)


def _fake_backend(log):
    def generate(prompts, max_new_tokens, fixed_len):
        log.append([p.shape[1] for p in prompts])
        return [("codes", p.shape[1], fixed_len) for p in prompts]

    def vocode(codes):
        return [("pcm", c[1]) for c in codes]
    return generate, vocode


def test_sharded_synthesizer_runs_only_its_share():
    prompts = [_Arr(100 + 3 * i) for i in range(70)]
    got = {}
    for r in range(2):
        log = []
        g, v = _fake_backend(log)
        syn = shard.ShardedSynthesizer(rank=r, world_size=2, generate=g, vocode=v)
        res = syn.synthesize(prompts, 4096, fixed_len=216)
        assert sorted(res) == shard.my_shard([p.shape[1] + 216 for p in prompts], r, 2)
        assert [len(b) for b in syn.launches] == [32, 3] and len(log) == 2  # 35 utterances: two launches
        assert all(res[i] == (("codes", prompts[i].shape[1], 216), ("pcm", prompts[i].shape[1])) for i in res)
        assert not set(res) & set(got)
        got.update(res)
    assert sorted(got) == list(range(70))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, costs):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.my_shard(costs, rank, world)
    # stand-in for "generate + vocode my utterances": payload derived from the index only
    local = {i: (i * 7 + 1, costs[i]) for i in mine}
    merged = shard.gather_results(local, world)
    assert sorted(merged) == list(range(len(costs)))
    assert all(merged[i] == (i * 7 + 1, costs[i]) for i in merged)
    # the driver: every rank runs its own launches, rank 0 ends up with everything, in utterance order
    prompts = [_Arr(c - 216) for c in costs]
    g, v = _fake_backend([])
    syn = shard.ShardedSynthesizer(rank=rank, world_size=world, generate=g, vocode=v, max_rows=2)
    res = syn.synthesize(prompts, 4096, frames=[216] * len(costs), fixed_len=216, gather=True)
    if rank == 0:
        assert sorted(res) == list(range(len(costs)))
        assert all(res[i][1] == ("pcm", prompts[i].shape[1]) for i in res)
    else:
        assert sorted(res) == sorted(mine)
    assert all(len(b) <= 2 for b in syn.launches)
    # every rank derived the same table without talking to the others
    tables = [None] * world
    dist.all_gather_object(tables, shard.assign(costs, world))
    assert all(t == tables[0] for t in tables)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    costs = [384 + 1292] * 5 + [300 + 216, 720 + 216, 512 + 216]
    mp.spawn(_worker, args=(2, _free_port(), costs), nprocs=2, join=True)
