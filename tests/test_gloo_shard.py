"""world_size-2 gloo test of the multi-GPU host logic (sharding + the output-offset exchange)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minialign_b200 import shard


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_batches = 7
    mine = shard.batches_of(rank, world, n_batches)
    sizes = {b: 1000 + 37 * b for b in mine}                      # pretend SAM byte counts
    res = []
    for wave in range((n_batches + world - 1) // world):
        b = wave * world + rank
        ofs, total = shard.output_offsets(sizes.get(b, 0))
        res.append((b if b < n_batches else -1, ofs, total))
    q.put((rank, mine, res))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_partition_and_offsets():
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    out = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    assert out[0][1] == [0, 2, 4, 6] and out[1][1] == [1, 3, 5]
    # offsets of a wave are the exclusive prefix sum in rank (= batch id) order, totals agree on both ranks
    for w in range(4):
        (b0, o0, t0), (b1, o1, t1) = out[0][2][w], out[1][2][w]
        assert o0 == 0 and t0 == t1
        assert o1 == 1000 + 37 * b0
        assert t0 == (1000 + 37 * b0) + ((1000 + 37 * b1) if b1 >= 0 else 0)
    assert [b for b, _ in shard.merged_order(7, 2)] == list(range(7))


def test_single_process_is_identity():
    assert shard.output_offsets(123) == (0, 123)
    assert shard.batches_of(0, 1, 5) == [0, 1, 2, 3, 4]
