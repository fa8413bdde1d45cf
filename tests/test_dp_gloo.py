"""CPU tier: the N>1 host logic (sharding, rank seeds, bucketed gradient all-reduce) over gloo,
world_size 2, rendezvous on 127.0.0.1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from ups_b200.dp import GradAllReducer, init_from_env, rank_seed, shard_bounds
    r, _, w = init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = shard_bounds(11, rank, world)
    g = torch.Generator().manual_seed(rank_seed(7, rank))
    grads = torch.randn(10_000, generator=g)
    mine = grads.clone()
    red = GradAllReducer(flat_grads=grads, bucket_bytes=4096 * 4, world_size=world)
    assert red.impl == "gloo" and red.transport == "gloo" and red.grad_scale == 1.0
    assert len(red.buckets) == 3
    red.launch()
    red.wait()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    want = sum(gathered) / world
    q.put((rank, lo, hi, bool(torch.allclose(grads, want, rtol=1e-6, atol=1e-7))))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_allreduce_and_shards():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, 0, 6, True), (1, 6, 11, True)]


def test_shard_bounds_cover_batch():
    from ups_b200.dp import shard_bounds, rank_seed
    for gb in (1, 7, 8, 256, 1001):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(gb, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == gb
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    assert len({rank_seed(0, r) for r in range(8)}) == 8


def test_bucket_split_is_16_byte_aligned_and_covers():
    from ups_b200.dp import _split, two_buckets, TAIL_BUCKET_FLOATS
    for n in (4, 1000, 33_300_000, 33_300_004):
        for parts in (1, 2, 3, 7):
            b = _split(n, parts)
            assert b[0][0] == 0 and b[-1][0] + b[-1][1] == n
            assert all(o % 4 == 0 for o, _ in b)
            assert all(b[i][0] + b[i][1] == b[i + 1][0] for i in range(parts - 1))
    b = two_buckets(33_300_002)
    assert b[0][0] == 0 and b[0][1] + b[1][1] == 33_300_004 and b[1][1] == TAIL_BUCKET_FLOATS and b[1][0] % 4 == 0
