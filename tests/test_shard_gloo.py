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
    # every rank derived the same table without talking to the others
    tables = [None] * world
    dist.all_gather_object(tables, shard.assign(costs, world))
    assert all(t == tables[0] for t in tables)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    costs = [384 + 1292] * 5 + [300 + 216, 720 + 216, 512 + 216]
    mp.spawn(_worker, args=(2, _free_port(), costs), nprocs=2, join=True)
